// sufsort.cu -- createSuffixArray (src/Data/BWT/Internal.hs:110-134) as a GPU suffix sort.
//
//   1. byte histogram -> alphabet; every symbol gets a code 1..sigma (0 = "past the end",
//      which is the sentinel: unique and smaller than every symbol, like the empty suffix
//      sorting first under Ord (Seq a)).  The text is re-packed at b = bits(sigma) bits per
//      symbol, so the key of suffix i (its first k = 64/b symbols: 21 for ACGT(N), 7 for
//      bytes) is two 64-bit loads and a funnel shift.
//   2. sort all suffixes by key.  Two paths:
//      MSD path (balanced inputs): counting sort on a <= 24-bit key prefix with global
//        atomics (not stable -- it need not be), then every bucket (<= 1024 suffixes) is sorted
//        by one warp with a bitonic network in shared memory.  ~3 sweeps over the data instead
//        of 8 LSD passes; the LSD passes are ALU-bound on ballot matching (profiles/README.md).
//      LSD path (fallback: skewed or very large inputs): 8-bit radix passes (radix.cu).
//   3. group heads (key != previous key) -> running max = group id = rank; suffixes in
//      singleton groups are final.
//   4. while unresolved suffixes remain: compact them, key2 = (group << 32 | rank[i + h]),
//      radix sort only those, write back, split groups, h *= 2.  In the first round rank[i + h]
//      is found by binary search in the sorted key array, so the inverse suffix array (a 4N-byte
//      random scatter) is only built if a second round is needed.
// Suffixes that reach the end of the text inside their first h symbols are always unique,
// so i + h <= n for every unresolved suffix.
#include <math.h>

#include <algorithm>
#include <utility>

#include "common.cuh"

namespace {
struct Code256 {
    uint16_t code[256]; // byte -> code 1..sigma (0 if the byte does not occur)
};

__global__ void __launch_bounds__(256) byte_hist_kernel(const uint8_t *__restrict__ t, uint64_t n,
                                                        uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t stride = (uint64_t)gridDim.x * 256 * 16;
    for (uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16; base < n; base += stride) {
        if (base + 16 <= n && ((reinterpret_cast<uintptr_t>(t + base) & 15) == 0)) {
            uint4 v = ld_stream_u4(t + base);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; k++) atomicAdd(&h[(w[k >> 2] >> ((k & 3) * 8)) & 0xff], 1u);
        } else {
            for (int k = 0; k < 16 && base + k < n; k++) atomicAdd(&h[t[base + k]], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// ---- packed text ---------------------------------------------------------------------
// Symbol j occupies bits [j*b, (j+1)*b) of an MSB-first bit stream of 64-bit words; symbols
// past the end are 0.  extract_key returns the first kb = k*b bits of suffix i, right-aligned.
__device__ __forceinline__ uint64_t extract_key(const uint64_t *__restrict__ pw, int b, int kb, uint64_t i) {
    uint64_t o = i * (uint64_t)b;
    uint64_t wi = o >> 6;
    int sh = (int)(o & 63);
    uint64_t hi = pw[wi], lo = pw[wi + 1];
    uint64_t v = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
    return v >> (64 - kb);
}

__global__ void __launch_bounds__(256)
    sa_pack_kernel(const uint8_t *__restrict__ t, uint64_t n, Code256 lut, int b, uint64_t nwords,
                   uint64_t *__restrict__ pw) {
    __shared__ uint16_t s_lut[256];
    s_lut[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    uint64_t w = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (w >= nwords) return;
    uint64_t bit0 = w * 64;
    uint64_t word = 0;
    for (uint64_t j = bit0 / b; j * b < bit0 + 64; j++) {
        uint64_t code = j < n ? s_lut[t[j]] : 0;
        int s = 64 - (int)((int64_t)(j * b) - (int64_t)bit0) - b; // left shift that puts the symbol in place
        word |= s >= 0 ? (code << s) : (code >> (-s));
    }
    pw[w] = word;
}

__global__ void sa_keys_from_packed_kernel(const uint64_t *__restrict__ pw, int b, int kb, uint64_t N,
                                           uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    keys[i] = extract_key(pw, b, kb, i);
    vals[i] = (uint32_t)i;
}

// ---- MSD path ------------------------------------------------------------------------
__global__ void msd_hist_kernel(const uint64_t *__restrict__ pw, int b, int kb, int bshift, uint64_t N,
                                uint32_t *__restrict__ cnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    atomicAdd(&cnt[extract_key(pw, b, kb, i) >> bshift], 1u);
}
__global__ void msd_max_kernel(const uint32_t *__restrict__ cnt, uint64_t nbk, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbk; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, cnt[i]);
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(TC_FULL, m, d));
    if (lane_id() == 0 && m) atomicMax(out, m);
}
// cursor[] = exclusive scan of the counts; after this kernel cursor[bkt] = end of bucket bkt
__global__ void msd_scatter_kernel(const uint64_t *__restrict__ pw, int b, int kb, int bshift, uint64_t N,
                                   uint32_t *__restrict__ cursor, uint32_t *__restrict__ sa) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t slot = atomicAdd(&cursor[extract_key(pw, b, kb, i) >> bshift], 1u);
    sa[slot] = (uint32_t)i;
}

constexpr int BS_WARPS = 8;
constexpr int BS_CAP = 512;   // largest bucket one warp sorts
constexpr int BS_PER = BS_CAP / 32;
struct BsWarp {
    uint64_t keys[BS_CAP];
    uint32_t vals[BS_CAP];
    uint32_t cnt[256 + 32];
};
// bitonic network over (key, suffix) pairs in shared memory: fallback for buckets whose next
// 8 bits are badly skewed
__device__ __forceinline__ void warp_bitonic(uint64_t *keys, uint32_t *vals, uint32_t bn) {
    const unsigned lane = lane_id();
    uint32_t P = 32;
    while (P < bn) P <<= 1;
    for (uint32_t j = bn + lane; j < P; j += 32) {
        keys[j] = ~0ull;
        vals[j] = 0xffffffffu;
    }
    __syncwarp();
    for (uint32_t k2 = 2; k2 <= P; k2 <<= 1) {
        for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (uint32_t idx = lane; idx < P / 2; idx += 32) {
                uint32_t t = ((idx & ~(j2 - 1)) << 1) | (idx & (j2 - 1));
                uint32_t p = t | j2;
                bool up = (t & k2) == 0;
                uint64_t kt = keys[t], kp = keys[p];
                uint32_t vt = vals[t], vp = vals[p];
                bool gt = kt > kp || (kt == kp && vt > vp);
                if (gt == up) {
                    keys[t] = kp;
                    keys[p] = kt;
                    vals[t] = vp;
                    vals[p] = vt;
                }
            }
            __syncwarp();
        }
    }
}

// One warp sorts one bucket [bs, bs+bn) of sa by (key, suffix).  The keys of a bucket share
// their prefix; the `fbits` key bits below it are sorted by 1 or 2 stable counting rounds of 8
// bits (LSD; ballot ranking into warp-private counters, as in the radix scatter).  What is
// left are short runs of equal field -- about one element for balanced text -- finished by an
// insertion sort run by the lane that owns the run head.  Badly skewed buckets fall back to a
// bitonic network.
__device__ __forceinline__ void warp_bucket_sort(BsWarp &W, uint32_t bs, uint32_t bn, int bshift, int nrounds,
                                                 const uint64_t *__restrict__ pw, int b, int kb,
                                                 uint32_t *__restrict__ sa, uint32_t *__restrict__ ties) {
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int fbits = bshift < 8 * nrounds ? bshift : 8 * nrounds;
    const int fshift = bshift - fbits;
    const int nr = (fbits + 7) / 8;
    uint64_t key[BS_PER];
    uint32_t val[BS_PER];
    uint16_t rnk[BS_PER];
#pragma unroll
    for (int i = 0; i < BS_PER; i++) {
        uint32_t j = lane + 32 * i;
        val[i] = j < bn ? sa[bs + j] : 0;
    }
#pragma unroll
    for (int i = 0; i < BS_PER; i++) {
        uint32_t j = lane + 32 * i;
        key[i] = j < bn ? extract_key(pw, b, kb, val[i]) : 0;
    }
    if (nr == 0) { // the prefix is the whole key: nothing to count on
        for (uint32_t j = lane; j < bn; j += 32) {
            W.keys[j] = key[j >> 5];
            W.vals[j] = val[j >> 5];
        }
    }
    for (int r = 0; r < nr; r++) {
        const int dsh = fshift + 8 * r;
        if (r > 0) { // re-read in the order the previous round produced
#pragma unroll
            for (int i = 0; i < BS_PER; i++) {
                uint32_t j = lane + 32 * i;
                if (j < bn) {
                    key[i] = W.keys[j];
                    val[i] = W.vals[j];
                }
            }
        }
        for (int j = lane; j < 256 + 32; j += 32) W.cnt[j] = 0;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < BS_PER; i++) {
            if (32u * i < bn) { // warp-uniform
                bool valid = lane + 32 * i < bn;
                uint32_t d = (uint32_t)(key[i] >> dsh) & 255u;
                unsigned peers = match_bits<8>(d, valid);
                uint32_t pre = valid ? W.cnt[d] : 0;
                __syncwarp();
                if (valid && (peers & lt) == 0) W.cnt[d] = pre + __popc(peers);
                __syncwarp();
                rnk[i] = (uint16_t)(pre + __popc(peers & lt));
            }
        }
        // exclusive scan of the 256 counters (8 per lane); cnt[d] becomes the start of digit d
        uint32_t c[8], run = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            c[q] = W.cnt[lane * 8 + q];
            run += c[q];
        }
        uint32_t excl = warp_incl_sum(run) - run;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; q++) {
            W.cnt[lane * 8 + q] = excl;
            excl += c[q];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < BS_PER; i++) {
            if (lane + 32 * i < bn) {
                uint32_t d = (uint32_t)(key[i] >> dsh) & 255u;
                uint32_t p = W.cnt[d] + rnk[i];
                W.keys[p] = key[i];
                W.vals[p] = val[i];
            }
        }
        __syncwarp();
    }
    // Runs of equal field are short (about one element for balanced text).  Every element counts
    // the members of its run that are smaller than itself and moves straight to its final slot;
    // all lanes walk outwards together, so there is no divergence.  A run longer than 32 means
    // the bucket is badly skewed: fall back to the bitonic network.
    uint32_t newpos[BS_PER];
    bool fallback = false;
#pragma unroll
    for (int i = 0; i < BS_PER; i++) {
        if (32u * i < bn) { // warp-uniform
            const uint32_t j = lane + 32 * i;
            const bool valid = j < bn;
            const uint64_t kj = valid ? W.keys[j] : 0;
            const uint32_t vj = valid ? W.vals[j] : 0;
            const uint64_t fj = kj >> fshift;
            key[i] = kj;
            val[i] = vj;
            uint32_t nleft = 0, nless = 0;
            bool goL = valid, goR = valid;
            for (uint32_t d = 1;; d++) {
                bool l = goL && j >= d;
                uint64_t kl = l ? W.keys[j - d] : 0;
                l = l && (kl >> fshift) == fj;
                goL = l;
                bool r = goR && j + d < bn;
                uint64_t kr = r ? W.keys[j + d] : 0;
                r = r && (kr >> fshift) == fj;
                goR = r;
                if (!__any_sync(TC_FULL, l || r)) break;
                if (d > 32) {
                    fallback = true;
                    break;
                }
                if (l) {
                    nleft++;
                    nless += (kl < kj) || (kl == kj && W.vals[j - d] < vj);
                }
                if (r) nless += (kr < kj) || (kr == kj && W.vals[j + d] < vj);
            }
            newpos[i] = j - nleft + nless;
        }
    }
    fallback = __any_sync(TC_FULL, fallback);
    __syncwarp();
    if (!fallback) {
#pragma unroll
        for (int i = 0; i < BS_PER; i++) {
            if (lane + 32 * i < bn) {
                W.keys[newpos[i]] = key[i];
                W.vals[newpos[i]] = val[i];
            }
        }
        __syncwarp();
    } else {
        warp_bitonic(W.keys, W.vals, bn);
    }
    // suffixes whose key equals a neighbour's are unresolved (equal keys never span buckets)
    uint32_t tied = 0;
    for (uint32_t j = lane; j < bn; j += 32) {
        sa[bs + j] = W.vals[j];
        uint64_t kj = W.keys[j];
        tied += (j > 0 && W.keys[j - 1] == kj) || (j + 1 < bn && W.keys[j + 1] == kj);
    }
    for (int dlt = 16; dlt; dlt >>= 1) tied += __shfl_xor_sync(TC_FULL, tied, dlt);
    if (lane == 0 && tied) atomicAdd(ties, tied);
    __syncwarp();
}

// Warps take groups of 32 consecutive buckets from a global counter (bucket sizes differ and
// populated bucket ids come in runs, so a static split is badly unbalanced).  *ties counts the
// suffixes that still share their key with a neighbour.
__global__ void __launch_bounds__(BS_WARPS * 32)
    msd_bucket_sort_kernel(const uint32_t *__restrict__ ends, uint64_t nbk, int bshift, int nrounds,
                           const uint64_t *__restrict__ pw, int b, int kb, uint32_t *__restrict__ sa, uint32_t *__restrict__ ties,
                           unsigned long long *__restrict__ next_group) {
    extern __shared__ __align__(16) unsigned char bs_raw[];
    BsWarp &W = reinterpret_cast<BsWarp *>(bs_raw)[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    // work unit: a multiple of 32 bucket ids, sized so that there are ~64K units (one global
    // atomic per unit; populated ids come in runs, so units must stay much smaller than the table)
    const uint64_t usz = 32 * (nbk / (32ull * 65536) > 1 ? nbk / (32ull * 65536) : 1);
    const uint64_t nunits = (nbk + usz - 1) / usz;
    for (;;) {
        unsigned long long unit = 0;
        if (lane == 0) unit = atomicAdd(next_group, 1ull);
        unit = __shfl_sync(TC_FULL, unit, 0);
        if (unit >= nunits) break;
        for (uint64_t base = unit * usz; base < (unit + 1) * usz && base < nbk; base += 32) {
            uint64_t id = base + lane;
            uint32_t e = id < nbk ? ends[id] : 0;
            uint32_t s = __shfl_up_sync(TC_FULL, e, 1);
            if (lane == 0) s = base ? ends[base - 1] : 0;
            uint32_t size = id < nbk ? e - s : 0;
            unsigned todo = __ballot_sync(TC_FULL, size >= 2);
            while (todo) {
                int l = __ffs(todo) - 1;
                todo &= todo - 1;
                uint32_t bs = __shfl_sync(TC_FULL, s, l), bn = __shfl_sync(TC_FULL, size, l);
                warp_bucket_sort(W, bs, bn, bshift, nrounds, pw, b, kb, sa, ties);
            }
        }
    }
}

__global__ void sa_gather_keys_kernel(const uint64_t *__restrict__ pw, int b, int kb, const uint32_t *__restrict__ sa,
                                      uint64_t N, uint64_t *__restrict__ keys) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    keys[j] = extract_key(pw, b, kb, sa[j]);
}

// h[j] = j+1 if slot j starts a new group (key differs from its predecessor), else 0
__global__ void sa_heads_kernel(const uint64_t *__restrict__ keys, uint64_t N, uint32_t *__restrict__ h) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    h[j] = (j == 0 || keys[j] != keys[j - 1]) ? (uint32_t)j + 1 : 0u;
}

// g = running max of h = (group head slot + 1); ns[j] = 1 iff the group of slot j has more than one member
__global__ void sa_flags_kernel(const uint32_t *__restrict__ g, uint64_t N, uint32_t *__restrict__ ns) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    bool head = g[j] == (uint32_t)j + 1;
    bool next_head = (j + 1 == N) || (g[j + 1] == (uint32_t)j + 2);
    ns[j] = (head && next_head) ? 0u : 1u;
}
__global__ void sa_build_isa_kernel(const uint32_t *__restrict__ g, const uint32_t *__restrict__ sa, uint64_t N,
                                    uint32_t *__restrict__ isa) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    isa[sa[j]] = g[j] - 1;
}
// first doubling round: rank of suffix s+h = first slot holding its key (binary search in the
// sorted keys), so no inverse suffix array is needed yet
__global__ void sa_keys2_lookup_kernel(const uint32_t *__restrict__ cj, uint64_t U, const uint32_t *__restrict__ sa,
                                       const uint64_t *__restrict__ keys, uint64_t N, const uint32_t *__restrict__ g,
                                       uint64_t h, uint64_t n, const uint64_t *__restrict__ pw, int b, int kb,
                                       uint64_t *__restrict__ key2, uint32_t *__restrict__ val2) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= U) return;
    uint32_t j = cj[m];
    uint32_t s = sa[j];
    uint64_t x = (uint64_t)s + h;
    uint32_t r2 = 0;
    if (x <= n) {
        uint64_t kx = extract_key(pw, b, kb, x);
        uint64_t lo = 0, hi = N; // first slot with keys[slot] >= kx
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (keys[mid] < kx) lo = mid + 1; else hi = mid;
        }
        r2 = (uint32_t)lo;
    }
    key2[m] = ((uint64_t)(g[j] - 1) << 32) | r2;
    val2[m] = s;
}

__global__ void sa_compact_kernel(const uint32_t *__restrict__ ns, const uint32_t *__restrict__ cpos,
                                  const uint32_t *__restrict__ src /*nullable: identity*/, uint64_t count,
                                  uint32_t *__restrict__ dst) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    if (ns[j]) dst[cpos[j]] = src ? src[j] : (uint32_t)j;
}

__global__ void sa_keys2_kernel(const uint32_t *__restrict__ cj, uint64_t U, const uint32_t *__restrict__ sa,
                                const uint32_t *__restrict__ isa, const uint32_t *__restrict__ g, uint64_t h,
                                uint64_t n, uint64_t *__restrict__ key2, uint32_t *__restrict__ val2) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= U) return;
    uint32_t j = cj[m];
    uint32_t s = sa[j];
    uint64_t nx = (uint64_t)s + h;
    uint32_t r2 = nx <= n ? isa[nx] : 0u; // nx <= n always holds for unresolved suffixes
    key2[m] = ((uint64_t)(g[j] - 1) << 32) | r2;
    val2[m] = s;
}

__global__ void sa_update_kernel(const uint32_t *__restrict__ g2, const uint32_t *__restrict__ cj,
                                 const uint32_t *__restrict__ val2s, uint64_t U, uint32_t *__restrict__ sa,
                                 uint32_t *__restrict__ isa, uint32_t *__restrict__ g, uint32_t *__restrict__ ns2) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= U) return;
    uint32_t slot = cj[m];
    uint32_t gm = g2[m];
    uint32_t headslot = cj[gm - 1];
    uint32_t s = val2s[m];
    sa[slot] = s;
    if (isa) isa[s] = headslot;
    g[slot] = headslot + 1;
    bool head = gm == (uint32_t)m + 1;
    bool next_head = (m + 1 == U) || (g2[m + 1] == (uint32_t)m + 2);
    ns2[m] = (head && next_head) ? 0u : 1u;
}

inline int bits_for(uint64_t maxval) { // bits needed to represent values 0..maxval
    int b = 0;
    while (maxval) {
        b++;
        maxval >>= 1;
    }
    return b ? b : 1;
}
} // namespace

// Byte histogram -> host (256 counts).  Shared with the FM-index builder.
int tc_byte_hist_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *h_hist /*256, host*/) {
    uint32_t *d_hist;
    TC_TRY(ws_alloc(ctx, 256, &d_hist));
    TC_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint32_t), ctx->stream));
    if (n) {
        unsigned grid = (unsigned)std::min<uint64_t>(ceil_div_u64(n, 256 * 16), (uint64_t)ctx->sm_count * 8);
        TC_LAUNCH(ctx, byte_hist_kernel, grid, 256, 0, d_text, n, d_hist);
    }
    uint32_t *h = (uint32_t *)ctx->h_scal;
    TC_CUDA(cudaMemcpyAsync(h, d_hist, 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, h, 256 * sizeof(uint32_t));
    return TC_OK;
}

int tc_suffix_sort_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_sa) {
    const uint64_t N = n + 1;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    if (n == 0) {
        TC_CUDA(cudaMemsetAsync(d_sa, 0, sizeof(uint32_t), ctx->stream));
        return TC_OK;
    }
    WsMark mk = tc_ws_mark(ctx);
    uint32_t hist[256];
    TC_TRY(tc_byte_hist_dev(ctx, d_text, n, hist));
    Code256 lut;
    int sigma = 0;
    double entropy = 0;
    for (int c = 0; c < 256; c++) {
        lut.code[c] = hist[c] ? (uint16_t)(++sigma) : 0;
        if (hist[c]) {
            double p = (double)hist[c] / (double)n;
            entropy -= p * log2(p);
        }
    }
    const int b = bits_for((uint64_t)sigma); // codes 0..sigma; 9 bits when all 256 byte values occur
    const int k = 64 / b;
    const int kb = b * k;
    const unsigned gridN = (unsigned)ceil_div_u64(N, 256);

    // packed text (+ enough zero symbols behind the end for any key read)
    const uint64_t nwords = ((N + (uint64_t)k + 64) * (uint64_t)b + 63) / 64 + 2;
    uint64_t *pw, *keys = nullptr;
    uint32_t *g, *ns, *cpos, *d_U;
    TC_TRY(ws_alloc(ctx, nwords, &pw));
    TC_TRY(ws_alloc(ctx, 2, &d_U));
    TC_LAUNCH(ctx, sa_pack_kernel, (unsigned)ceil_div_u64(nwords, 256), 256, 0, d_text, n, lut, b, nwords, pw);
    uint32_t *hU = (uint32_t *)ctx->h_scal;

    // ---- sort by key: MSD path when the buckets of a <= 24-bit prefix stay small
    bool sorted = false;
    if (N >= 4096) {
        // prefix length: about N/128 populated buckets, judged by the symbol entropy
        double want_bits = log2((double)N / 128.0);
        int p = (int)ceil(want_bits / (entropy > 0.05 ? entropy : 0.05));
        int pmax = 24 / b;
        if (p > pmax) p = pmax;
        if (p > k) p = k;
        if (p < 1) p = 1;
        const int PB = p * b;
        const int bshift = kb - PB;
        const uint64_t nbk = 1ull << PB;
        uint32_t *cnt;
        TC_TRY(ws_alloc(ctx, nbk, &cnt));
        TC_CUDA(cudaMemsetAsync(cnt, 0, nbk * sizeof(uint32_t), ctx->stream));
        TC_CUDA(cudaMemsetAsync(d_U, 0, 2 * sizeof(uint32_t), ctx->stream));
        TC_LAUNCH(ctx, msd_hist_kernel, gridN, 256, 0, pw, b, kb, bshift, N, cnt);
        unsigned mgrid = (unsigned)std::min<uint64_t>(ceil_div_u64(nbk, 256), (uint64_t)ctx->sm_count * 8);
        TC_LAUNCH(ctx, msd_max_kernel, mgrid, 256, 0, cnt, nbk, d_U);
        TC_CUDA(cudaMemcpyAsync(hU, d_U, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (hU[0] <= (uint32_t)BS_CAP) {
            TC_TRY(tc_scan_exclusive_u32(ctx, cnt, cnt, nbk, (uint32_t *)nullptr));
            TC_LAUNCH(ctx, msd_scatter_kernel, gridN, 256, 0, pw, b, kb, bshift, N, cnt, d_sa);
            const size_t bs_smem = sizeof(BsWarp) * BS_WARPS;
            TC_CUDA(cudaFuncSetAttribute(msd_bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bs_smem));
            unsigned long long *d_next;
            TC_TRY(ws_alloc(ctx, 1, &d_next));
            TC_CUDA(cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), ctx->stream));
            unsigned sgrid = (unsigned)std::min<uint64_t>(ceil_div_u64(nbk, BS_WARPS * 32), (uint64_t)ctx->sm_count * 3);
            TC_CUDA(cudaMemsetAsync(d_U, 0, sizeof(uint32_t), ctx->stream));
            // key bits needed below the prefix to split a bucket into singletons, judged by the symbol
            // entropy per key bit: one counting round of 8 bits or two
            double ebit = (entropy > 0.05 ? entropy : 0.05) / (double)b;
            int nrounds = (log2((double)(hU[0] > 1 ? hU[0] : 2)) / ebit > 10.0) ? 2 : 1;
            TC_LAUNCH(ctx, msd_bucket_sort_kernel, sgrid, BS_WARPS * 32, bs_smem, cnt, nbk, bshift, nrounds, pw, b, kb,
                      d_sa, d_U, d_next);
            TC_CUDA(cudaMemcpyAsync(hU, d_U, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            TC_CUDA(cudaStreamSynchronize(ctx->stream));
            if (hU[0] == 0) { // every suffix is already distinguished by its key: done
                tc_ws_release(ctx, mk);
                return TC_OK;
            }
            TC_TRY(ws_alloc(ctx, N, &keys));
            TC_LAUNCH(ctx, sa_gather_keys_kernel, gridN, 256, 0, pw, b, kb, d_sa, N, keys);
            sorted = true;
        }
    }
    if (!sorted) { // LSD path
        TC_TRY(ws_alloc(ctx, N, &keys));
        int shifts[16];
        int npass = 0;
        for (int s = 0; s < kb; s += 8) shifts[npass++] = s;
        uint64_t *k1;
        uint32_t *vtmp;
        TC_TRY(ws_alloc(ctx, N, &k1));
        TC_TRY(ws_alloc(ctx, N, &vtmp));
        // arrange the ping-pong so the sorted values land in d_sa and the sorted keys in `keys`
        uint64_t *ka = (npass % 2 == 0) ? keys : k1, *kbuf = (npass % 2 == 0) ? k1 : keys;
        uint32_t *v0 = (npass % 2 == 0) ? d_sa : vtmp, *v1 = (npass % 2 == 0) ? vtmp : d_sa;
        TC_LAUNCH(ctx, sa_keys_from_packed_kernel, gridN, 256, 0, pw, b, kb, N, ka, v0);
        uint64_t *ks;
        uint32_t *vs;
        TC_TRY(tc_radix_sort_pairs(ctx, ka, v0, kbuf, v1, N, shifts, npass, &ks, &vs));
        if (vs != d_sa || ks != keys) return TC_E_ARG; // cannot happen
    }

    // ---- groups and unresolved suffixes
    TC_TRY(ws_alloc(ctx, N, &g));
    TC_TRY(ws_alloc(ctx, N, &ns));
    TC_TRY(ws_alloc(ctx, N, &cpos));
    TC_LAUNCH(ctx, sa_heads_kernel, gridN, 256, 0, keys, N, g);
    TC_TRY(tc_scan_inclusive_max_u32(ctx, g, g, N));
    TC_LAUNCH(ctx, sa_flags_kernel, gridN, 256, 0, g, N, ns);
    TC_TRY(tc_scan_exclusive_u32(ctx, ns, cpos, N, d_U));
    TC_CUDA(cudaMemcpyAsync(hU, d_U, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t U = hU[0];
    if (U > 0) {
        uint32_t *cjA, *cjB, *val2a, *val2b, *g2, *ns2, *cpos2, *isa = nullptr;
        uint64_t *key2a, *key2b;
        TC_TRY(ws_alloc(ctx, U, &cjA));
        TC_TRY(ws_alloc(ctx, U, &cjB));
        TC_TRY(ws_alloc(ctx, U, &val2a));
        TC_TRY(ws_alloc(ctx, U, &val2b));
        TC_TRY(ws_alloc(ctx, U, &g2));
        TC_TRY(ws_alloc(ctx, U, &ns2));
        TC_TRY(ws_alloc(ctx, U, &cpos2));
        TC_TRY(ws_alloc(ctx, U, &key2a));
        TC_TRY(ws_alloc(ctx, U, &key2b));
        TC_LAUNCH(ctx, sa_compact_kernel, gridN, 256, 0, ns, cpos, (const uint32_t *)nullptr, N, cjA);
        const int rb = bits_for(N - 1);
        int sh2[16];
        int np2 = 0;
        for (int s = 0; s < rb; s += 8) sh2[np2++] = s;
        for (int s = 0; s < rb; s += 8) sh2[np2++] = 32 + s;
        uint64_t h = (uint64_t)k;
        for (int round = 0; U > 0; round++) {
            if (round > 48) {
                snprintf(ctx->err, sizeof ctx->err, "suffix sort did not converge");
                return TC_E_CUDA;
            }
            const unsigned gridU = (unsigned)ceil_div_u64(U, 256);
            if (round == 0) {
                TC_LAUNCH(ctx, sa_keys2_lookup_kernel, gridU, 256, 0, cjA, U, d_sa, keys, N, g, h, n, pw, b, kb, key2a,
                          val2a);
            } else {
                if (!isa) { // second round: now the inverse suffix array pays for itself
                    TC_TRY(ws_alloc(ctx, N, &isa));
                    TC_LAUNCH(ctx, sa_build_isa_kernel, gridN, 256, 0, g, d_sa, N, isa);
                }
                TC_LAUNCH(ctx, sa_keys2_kernel, gridU, 256, 0, cjA, U, d_sa, isa, g, h, n, key2a, val2a);
            }
            uint64_t *k2s;
            uint32_t *v2s;
            TC_TRY(tc_radix_sort_pairs(ctx, key2a, val2a, key2b, val2b, U, sh2, np2, &k2s, &v2s));
            TC_LAUNCH(ctx, sa_heads_kernel, gridU, 256, 0, k2s, U, g2);
            TC_TRY(tc_scan_inclusive_max_u32(ctx, g2, g2, U));
            TC_LAUNCH(ctx, sa_update_kernel, gridU, 256, 0, g2, cjA, v2s, U, d_sa, isa, g, ns2);
            TC_TRY(tc_scan_exclusive_u32(ctx, ns2, cpos2, U, d_U));
            TC_LAUNCH(ctx, sa_compact_kernel, gridU, 256, 0, ns2, cpos2, (const uint32_t *)cjA, U, cjB);
            TC_CUDA(cudaMemcpyAsync(hU, d_U, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            TC_CUDA(cudaStreamSynchronize(ctx->stream));
            U = hU[0];
            std::swap(cjA, cjB);
            h *= 2;
        }
    }
    tc_ws_release(ctx, mk);
    return TC_OK;
}
