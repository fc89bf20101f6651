// sufsort.cu -- createSuffixArray (src/Data/BWT/Internal.hs:110-134) as a prefix-doubling
// GPU suffix sort.
//
//   1. byte histogram -> alphabet; every symbol gets a code 1..sigma (0 = "past the end",
//      which is the sentinel: unique and smaller than every symbol, like the empty suffix
//      sorting first under Ord (Seq a)).
//   2. initial key of suffix i = its first k symbols packed at b bits each into 64 bits
//      (b = bits(sigma), k = 64 / b: 21 symbols for ACGT(N), 7 for full bytes).
//   3. LSD radix sort of (key, i).
//   4. group heads (key != previous key) -> running max = group id = rank; suffixes in
//      singleton groups are final.
//   5. while unresolved suffixes remain: compact them, key2 = (group << 32 | rank[i + h]),
//      radix sort only those, write back, split groups, h *= 2.
// Suffixes that reach the end of the text inside their first h symbols are always unique,
// so i + h <= n for every unresolved suffix.
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

constexpr int IK_T = 256;
constexpr int IK_PER = 4;
constexpr int IK_TILE = IK_T * IK_PER;
__global__ void __launch_bounds__(IK_T)
    sa_init_keys_kernel(const uint8_t *__restrict__ t, uint64_t n, uint64_t N, Code256 lut, int b, int k,
                        uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    __shared__ uint16_t codes[IK_TILE + 64];
    __shared__ uint16_t s_lut[256];
    s_lut[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * IK_TILE;
    for (int j = threadIdx.x; j < IK_TILE + 64; j += IK_T) {
        uint64_t i = base + j;
        codes[j] = i < n ? s_lut[t[i]] : 0;
    }
    __syncthreads();
    const uint64_t mask = (b * k >= 64) ? ~0ull : ((1ull << (b * k)) - 1);
    int o = threadIdx.x * IK_PER;
    uint64_t key = 0;
    for (int j = 0; j < k; j++) key = (key << b) | codes[o + j];
#pragma unroll
    for (int q = 0; q < IK_PER; q++) {
        uint64_t p = base + o + q;
        if (p < N) {
            keys[p] = key;
            vals[p] = (uint32_t)p;
        }
        key = ((key << b) | codes[o + q + k]) & mask;
    }
}

// h[j] = j+1 if slot j starts a new group (key differs from its predecessor), else 0
__global__ void sa_heads_kernel(const uint64_t *__restrict__ keys, uint64_t N, uint32_t *__restrict__ h) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    h[j] = (j == 0 || keys[j] != keys[j - 1]) ? (uint32_t)j + 1 : 0u;
}

// g = running max of h = (group head slot + 1).  rank[sa[j]] = head slot; ns[j] = 1 iff the
// group of slot j has more than one member.
__global__ void sa_assign_kernel(const uint32_t *__restrict__ g, const uint32_t *__restrict__ sa, uint64_t N,
                                 uint32_t *__restrict__ isa, uint32_t *__restrict__ ns) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    uint32_t gj = g[j];
    isa[sa[j]] = gj - 1;
    bool head = gj == (uint32_t)j + 1;
    bool next_head = (j + 1 == N) || (g[j + 1] == (uint32_t)j + 2);
    ns[j] = (head && next_head) ? 0u : 1u;
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
    isa[s] = headslot;
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
    for (int c = 0; c < 256; c++) lut.code[c] = hist[c] ? (uint16_t)(++sigma) : 0;
    const int b = bits_for((uint64_t)sigma); // codes 0..sigma; 9 bits when all 256 byte values occur
    const int k = 64 / b;
    const int key_bits = b * k;
    int shifts[16];
    int npass = 0;
    for (int s = 0; s < key_bits; s += 8) shifts[npass++] = s;

    uint64_t *k0, *k1;
    uint32_t *vtmp, *isa, *g, *ns;
    TC_TRY(ws_alloc(ctx, N, &k0));
    TC_TRY(ws_alloc(ctx, N, &k1));
    TC_TRY(ws_alloc(ctx, N, &vtmp));
    TC_TRY(ws_alloc(ctx, N, &isa));
    TC_TRY(ws_alloc(ctx, N, &g));
    TC_TRY(ws_alloc(ctx, N, &ns));
    // arrange the ping-pong so the sorted values land in d_sa
    uint32_t *v0 = (npass % 2 == 0) ? d_sa : vtmp;
    uint32_t *v1 = (npass % 2 == 0) ? vtmp : d_sa;
    TC_LAUNCH(ctx, sa_init_keys_kernel, (unsigned)ceil_div_u64(N, IK_TILE), IK_T, 0, d_text, n, N, lut,
              b, k, k0, v0);
    uint64_t *ks;
    uint32_t *vs;
    TC_TRY(tc_radix_sort_pairs(ctx, k0, v0, k1, v1, N, shifts, npass, &ks, &vs));
    if (vs != d_sa) return TC_E_ARG; // cannot happen
    const unsigned gridN = (unsigned)ceil_div_u64(N, 256);
    TC_LAUNCH(ctx, sa_heads_kernel, gridN, 256, 0, ks, N, g);
    TC_TRY(tc_scan_inclusive_max_u32(ctx, g, g, N));
    TC_LAUNCH(ctx, sa_assign_kernel, gridN, 256, 0, g, d_sa, N, isa, ns);
    uint32_t *cpos = (uint32_t *)k0; // keys are dead now; reuse as scratch (N u32 fits in N u64)
    uint32_t *d_U;
    TC_TRY(ws_alloc(ctx, 1, &d_U));
    TC_TRY(tc_scan_exclusive_u32(ctx, ns, cpos, N, d_U));
    uint32_t *hU = (uint32_t *)ctx->h_scal;
    TC_CUDA(cudaMemcpyAsync(hU, d_U, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t U = hU[0];
    if (U > 0) {
        uint32_t *cjA, *cjB, *val2a, *val2b, *g2, *ns2, *cpos2;
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
            TC_LAUNCH(ctx, sa_keys2_kernel, gridU, 256, 0, cjA, U, d_sa, isa, g, h, n, key2a, val2a);
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
