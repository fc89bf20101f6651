// sufsort.cu -- createSuffixArray (src/Data/BWT/Internal.hs:110-134) as a GPU suffix sort.
//
//   1. byte histogram -> alphabet; every symbol gets a code 1..sigma (0 = "past the end",
//      which is the sentinel: unique and smaller than every symbol, like the empty suffix
//      sorting first under Ord (Seq a)).  The text is re-packed at b = bits(sigma) bits per
//      symbol, so the key of suffix i (its first k = 64/b symbols: 21 for ACGT(N), 7 for
//      bytes) is two 64-bit loads and a funnel shift.
//   2. sort all suffixes by key.  Two paths:
//      MSD path (4 Ki <= n <= 25 Mi, the block sizes of the compressor): the radix is taken
//        from a 32-bit *uniform key* -- the arithmetic code of the suffix's first symbols under
//        the text's own symbol frequencies -- so two 8-bit partition levels give 65,536 evenly
//        filled buckets for any memoryless text; one warp then sorts a bucket with a bitonic
//        network in registers and writes the suffix array and the BWT bytes.  Ties in the
//        32-bit key are settled with the packed 63-bit keys and, beyond them, by comparing the
//        suffixes k symbols at a time.  See "MSD path on uniform keys" below.
//      LSD path (tiny or huge inputs, and texts whose buckets overflow, i.e. strongly
//        correlated or repetitive text): 8-bit radix passes over the packed keys (radix.cu).
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

// One thread packs 64 symbols into exactly B words, so every shift is a compile-time constant.
template <int B>
__global__ void __launch_bounds__(256)
    sa_pack_kernel(const uint8_t *__restrict__ t, uint64_t n, Code256 lut, uint64_t nwords, uint64_t *__restrict__ pw) {
    __shared__ uint16_t s_lut[256];
    s_lut[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    const uint64_t g = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (g * B >= nwords) return;
    const uint64_t base = g * 64;
    uint32_t by[16]; // 64 text bytes
    if (base + 64 <= n && (reinterpret_cast<uintptr_t>(t + base) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint4 v = ld_stream_u4(t + base + 16 * q);
            by[4 * q] = v.x;
            by[4 * q + 1] = v.y;
            by[4 * q + 2] = v.z;
            by[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint64_t i = base + 4 * q + k;
                if (i < n) w |= (uint32_t)t[i] << (8 * k);
            }
            by[q] = w;
        }
    }
    uint64_t w[B];
#pragma unroll
    for (int k = 0; k < B; k++) w[k] = 0;
#pragma unroll
    for (int j = 0; j < 64; j++) {
        const uint64_t code = base + j < n ? (uint64_t)s_lut[(by[j >> 2] >> (8 * (j & 3))) & 0xffu] : 0ull;
        const int wi = (j * B) >> 6, o = (j * B) & 63;
        if (o + B <= 64) {
            w[wi] |= code << (64 - o - B);
        } else {
            w[wi] |= code >> (o + B - 64);
            w[wi + 1] |= code << (128 - o - B);
        }
    }
#pragma unroll
    for (int k = 0; k < B; k++)
        if (g * B + k < nwords) pw[g * B + k] = w[k];
}

__global__ void sa_keys_from_packed_kernel(const uint64_t *__restrict__ pw, int b, int kb, uint64_t N,
                                           uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    keys[i] = extract_key(pw, b, kb, i);
    vals[i] = (uint32_t)i;
}

// ---- MSD path on uniform keys ---------------------------------------------------------
// ukey(i): a 32-bit order-preserving code of suffix i's key -- the arithmetic code of its
// first symbols under the text's own symbol frequencies (looked up g symbols at a time).
// key(s) <= key(t) implies ukey(s) <= ukey(t), and for memoryless text ukey is uniform on
// [0, 2^32) whatever the alphabet or its skew, so plain bit fields of ukey are balanced
// radix digits: no sparse bucket ids (ACGTN in 3-bit codes fills 5/8 of each digit), no
// entropy guessing.  Ties in ukey are settled with the full 63-bit keys.
struct SymProb {
    float p[258]; // p[code], code 0 = past the end
};

// lut[e] = {cum16 << 16, freq16 << 16} over all g-grams e in numeric (= lexicographic) order
__global__ void __launch_bounds__(1024) uk_lut_kernel(SymProb sp, int sigma, int b, int g, uint2 *__restrict__ lut) {
    __shared__ uint32_t sh[1024 / 32 + 1];
    const uint32_t E = 1u << (g * b);
    float nvalid = 1.f;
    for (int j = 0; j < g; j++) nvalid *= (float)(sigma + 1);
    const float S = 65536.f - nvalid - 64.f;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < E; base += 1024) {
        uint32_t e = base + threadIdx.x;
        float pr = e < E ? 1.f : 0.f;
        for (int j = 0; j < g; j++) {
            uint32_t c = (e >> (b * (g - 1 - j))) & ((1u << b) - 1);
            pr *= c <= (uint32_t)sigma ? sp.p[c] : 0.f;
        }
        uint32_t f = 0;
        if (pr > 0.f) {
            f = (uint32_t)(pr * S);
            if (f < 1) f = 1;
        }
        uint32_t tot;
        uint32_t cum = carry + block_excl_sum<uint32_t, 1024>(f, sh, &tot);
        if (e < E) {
            if (cum > 65535u) {
                cum = 65535u;
                f = 0;
            }
            if (cum + f > 65536u) f = 65536u - cum;
            if (f > 65535u) f = 65535u;
            lut[e] = make_uint2(cum << 16, f << 16);
        }
        carry += tot;
    }
}

// Shared-memory slot of g-gram e.  Symbol codes use 1..sigma of the 2^b values of a field, so
// the raw indices of the g-grams that occur cluster on a few banks (ACGTN in 3-bit fields: 10
// of 16 bank pairs, 79 us of bank-conflict replays per SM in uk_keys); folding the upper index
// bits into the lower ones spreads them.
__device__ __forceinline__ uint32_t lut_slot(uint32_t e) { return e ^ (e >> 4); }

// G lookups of gb key bits each, from the top of the (right-aligned, kb-bit) key; G = 0: runtime Gr
template <int G>
__device__ __forceinline__ uint32_t ukey_of(uint64_t key, int kb, int gb, int Gr, const uint2 *lut /*shared*/) {
    uint32_t x = 0, r = 0xffffffffu;
    int sh = kb - gb;
    const uint32_t m = (1u << gb) - 1;
    if (G > 0) {
#pragma unroll
        for (int j = 0; j < G; j++, sh -= gb) {
            uint2 e = lut[lut_slot((uint32_t)(key >> sh) & m)];
            x += __umulhi(r, e.x);
            r = __umulhi(r, e.y);
        }
    } else {
        for (int j = 0; j < Gr; j++, sh -= gb) {
            uint2 e = lut[lut_slot((uint32_t)(key >> sh) & m)];
            x += __umulhi(r, e.x);
            r = __umulhi(r, e.y);
        }
    }
    return x;
}
// bytes of a uint4 summed (each field of the result holds at most 16 * 255)
__device__ __forceinline__ uint32_t byte_sum16(uint4 v) {
    uint32_t a = (v.x & 0x00ff00ffu) + ((v.x >> 8) & 0x00ff00ffu);
    a += (v.y & 0x00ff00ffu) + ((v.y >> 8) & 0x00ff00ffu);
    a += (v.z & 0x00ff00ffu) + ((v.z >> 8) & 0x00ff00ffu);
    a += (v.w & 0x00ff00ffu) + ((v.w >> 8) & 0x00ff00ffu);
    return (a & 0xffffu) + (a >> 16);
}

// Counting without atomics or matching: every lane owns a byte counter per digit
// (cnt[digit][lane], 8 KB per warp), bumped with a plain load/add/store; a lane sees at most
// PC_ITEMS keys per tile, so a byte is enough.  The warp then sums its 32 columns per digit.
constexpr int PC_T = 256;
constexpr int PC_WARPS = PC_T / 32;
constexpr int PC_ITEMS = 64;
constexpr int PC_TILE = PC_T * PC_ITEMS;
constexpr int PC_WCHUNK = 32 * PC_ITEMS; // keys per warp
constexpr size_t PC_SMEM = (size_t)PC_WARPS * 256 * 32;

__device__ __forceinline__ void pc_zero(uint8_t *wc, unsigned lane) {
    uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int q = 0; q < 16; q++) reinterpret_cast<uint4 *>(wc)[q * 32 + lane] = z;
}
// total of digit d = q*32 + lane over the warp's 32 lane columns
__device__ __forceinline__ uint32_t pc_total(const uint8_t *wc, int q, unsigned lane) {
    const uint4 *p = reinterpret_cast<const uint4 *>(wc + (q * 32 + lane) * 32);
    return byte_sum16(p[0]) + byte_sum16(p[1]);
}

// K_A: ukey of every suffix 0..n-1 + histogram of the level-1 digit
constexpr int UK_LUT_MAX = 512; // g-gram table entries (gb <= 9 bits), staged in shared memory
template <int G>
__global__ void __launch_bounds__(PC_T, 3)
    uk_keys_kernel(const uint64_t *__restrict__ pw, int b, int kb, int gb, int Gr, const uint2 *__restrict__ lut,
                   uint32_t n, int shift1, uint32_t *__restrict__ ukey, uint32_t *__restrict__ hist1) {
    extern __shared__ __align__(16) uint8_t pc_cnt[];
    __shared__ uint16_t wtot[PC_WARPS][256];
    __shared__ uint2 s_lut[UK_LUT_MAX];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    for (int j = threadIdx.x; j < (1 << gb); j += PC_T) s_lut[lut_slot(j)] = lut[j];
    uint8_t *wc = pc_cnt + (size_t)w * 8192;
    pc_zero(wc, lane);
    __syncthreads();
    const uint32_t tbase = blockIdx.x * PC_TILE + w * PC_WCHUNK;
    uint8_t *mine = wc + lane;
    // batches of 8 suffixes: all packed-text loads of a batch are issued before any is used
    for (int r0 = 0; r0 < PC_ITEMS; r0 += 8) {
        uint64_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t i = tbase + (r0 + q) * 32 + lane;
            const uint32_t o = (i < n ? i : 0u) * (uint32_t)b;
            hi[q] = pw[o >> 6];
            lo[q] = pw[(o >> 6) + 1];
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t i = tbase + (r0 + q) * 32 + lane;
            if (i < n) {
                const int sh = (int)((i * (uint32_t)b) & 63u);
                const uint64_t v = sh ? ((hi[q] << sh) | (lo[q] >> (64 - sh))) : hi[q];
                const uint32_t u = ukey_of<G>(v >> (64 - kb), kb, gb, Gr, s_lut);
                ukey[i] = u;
                mine[(u >> shift1) * 32]++;
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) wtot[w][q * 32 + lane] = (uint16_t)pc_total(wc, q, lane);
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int ww = 0; ww < PC_WARPS; ww++) s += wtot[ww][threadIdx.x];
    if (s) atomicAdd(&hist1[threadIdx.x], s);
}

// K_A for (nearly) equiprobable bytes -- already compressed, encrypted or random data: with all 256 byte values
// equally likely the arithmetic code of a suffix IS its leading bytes, so the uniform key is the big-endian word
// text[i .. i+3] (zeros past the end; ties fall to the packed keys as always) and no table walk is needed.
__global__ void __launch_bounds__(PC_T, 3)
    uk_keys_raw_kernel(const uint8_t *__restrict__ t, uint32_t n, int shift1, uint32_t *__restrict__ ukey,
                       uint32_t *__restrict__ hist1) {
    extern __shared__ __align__(16) uint8_t pc_cnt[];
    __shared__ uint16_t wtot[PC_WARPS][256];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    uint8_t *wc = pc_cnt + (size_t)w * 8192;
    pc_zero(wc, lane);
    __syncthreads();
    const uint32_t tbase = blockIdx.x * PC_TILE + w * PC_WCHUNK;
    uint8_t *mine = wc + lane;
    const bool aligned = (reinterpret_cast<uintptr_t>(t) & 3) == 0;
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(t);
    for (int r0 = 0; r0 < PC_ITEMS; r0 += 8) {
        uint32_t w0[8], w1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t i = tbase + (r0 + q) * 32 + lane;
            const bool fast = aligned && i + 8 <= n; // both words inside the text
            w0[q] = fast ? t32[i >> 2] : 0u;
            w1[q] = fast ? t32[(i >> 2) + 1] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t i = tbase + (r0 + q) * 32 + lane;
            if (i < n) {
                uint32_t u;
                if (aligned && i + 8 <= n) {
                    u = __byte_perm(w0[q], w1[q], 0x0123u + 0x1111u * (i & 3u));
                } else {
                    u = 0;
                    for (int k = 0; k < 4; k++) u = (u << 8) | (i + k < n ? (uint32_t)t[i + k] : 0u);
                }
                ukey[i] = u;
                mine[(u >> shift1) * 32]++;
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) wtot[w][q * 32 + lane] = (uint16_t)pc_total(wc, q, lane);
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int ww = 0; ww < PC_WARPS; ww++) s += wtot[ww][threadIdx.x];
    if (s) atomicAdd(&hist1[threadIdx.x], s);
}

// Partition tiles never straddle a segment (= a bucket of the previous level).  tilebase[s]
// = first tile of segment s, chunkbase[s] = first counting chunk of segment s.
constexpr int PT_T = 256;
constexpr int PT_ITEMS = 8;
constexpr int PT_TILE = PT_T * PT_ITEMS;
constexpr int PT_WARPS = PT_T / 32;

// one CTA: segment starts from the level-1 histogram, + tile and chunk tables for level 2.  hist1 == nullptr: the
// histogram comes by value (raw-byte keys: the level-1 digit is the suffix's first byte, so the level-1 histogram
// is the byte histogram the host already holds).
struct Hist256 {
    uint32_t c[256];
};
__global__ void __launch_bounds__(256)
    seg_tables_kernel(const uint32_t *__restrict__ hist1, Hist256 hv, int nseg, uint32_t *__restrict__ segstart /*nseg+1*/,
                      uint32_t *__restrict__ cursor1, uint32_t *__restrict__ tilebase /*nseg+1*/,
                      uint32_t *__restrict__ chunkbase /*nseg+1*/) {
    __shared__ uint32_t sh[256 / 32 + 1];
    const int s = threadIdx.x;
    uint32_t c = s < nseg ? (hist1 ? hist1[s] : hv.c[s]) : 0;
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t, 256>(c, sh, &tot);
    if (s < nseg) {
        segstart[s] = ex;
        cursor1[s] = ex;
    }
    if (s == 0) segstart[nseg] = tot;
    uint32_t t = (c + PT_TILE - 1) / PT_TILE;
    ex = block_excl_sum<uint32_t, 256>(t, sh, &tot);
    if (s < nseg) tilebase[s] = ex;
    if (s == 0) tilebase[nseg] = tot;
    uint32_t h = (c + PC_WCHUNK - 1) / PC_WCHUNK;
    ex = block_excl_sum<uint32_t, 256>(h, sh, &tot);
    if (s < nseg) chunkbase[s] = ex;
    if (s == 0) chunkbase[nseg] = tot;
}

// segment of unit t: base[s] <= t < base[s+1]   (base[nseg] = number of units)
__device__ __forceinline__ int seg_of(const uint32_t *__restrict__ base, int nseg, uint32_t t) {
    int lo = 0, hi = nseg;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (base[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

// level-2 histogram: one warp per chunk of a segment, hist2[seg * nb2 + digit]
__global__ void __launch_bounds__(PC_T)
    seg_hist_kernel(const uint2 *__restrict__ rec, const uint32_t *__restrict__ segstart,
                    const uint32_t *__restrict__ chunkbase, int nseg, int shift, uint32_t dmask,
                    uint32_t *__restrict__ hist2) {
    extern __shared__ __align__(16) uint8_t pc_cnt[];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const uint32_t c = blockIdx.x * PC_WARPS + w;
    if (c >= chunkbase[nseg]) return;
    const int s = seg_of(chunkbase, nseg, c);
    const uint32_t beg = segstart[s] + (c - chunkbase[s]) * PC_WCHUNK;
    const uint32_t end = min(segstart[s + 1], beg + PC_WCHUNK);
    uint8_t *wc = pc_cnt + (size_t)w * 8192;
    pc_zero(wc, lane);
    __syncwarp();
    // batches of 16 records: all loads of a batch are issued before any is used
    for (int r0 = 0; r0 < PC_ITEMS; r0 += 16) {
        uint32_t x[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            uint32_t i = beg + (r0 + q) * 32 + lane;
            x[q] = i < end ? rec[i].x : 0u;
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
            uint32_t i = beg + (r0 + q) * 32 + lane;
            if (i < end) wc[((x[q] >> shift) & dmask) * 32 + lane]++;
        }
    }
    __syncwarp();
    const uint32_t nb = dmask + 1;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        uint32_t d = q * 32 + lane;
        if (d < nb) {
            uint32_t t = pc_total(wc, q, lane);
            if (t) atomicAdd(&hist2[(uint32_t)s * nb + d], t);
        }
    }
}

// level-2 bucket starts: one CTA per segment scans the segment's digit counts
__global__ void __launch_bounds__(256)
    starts_kernel(const uint32_t *__restrict__ hist2, const uint32_t *__restrict__ segstart, uint32_t nb2,
                  uint32_t nbk, uint64_t n, uint32_t *__restrict__ starts, uint32_t *__restrict__ cursor,
                  uint32_t bucket_cap, uint32_t *__restrict__ overflow) {
    __shared__ uint32_t sh[256 / 32 + 1];
    const uint32_t s = blockIdx.x, d = threadIdx.x;
    uint32_t c = d < nb2 ? hist2[s * nb2 + d] : 0;
    // a bucket the final level cannot hold: the level-2 partition and the final sort return at once and the host
    // takes the other sort path (correlated text ends here 0.4 ms earlier than by overflowing in the final sort)
    if (c > bucket_cap) atomicOr(overflow, 1u);
    uint32_t ex = segstart[s] + block_excl_sum<uint32_t, 256>(c, sh, (uint32_t *)nullptr);
    if (d < nb2) {
        starts[s * nb2 + d] = ex;
        cursor[s * nb2 + d] = ex;
    }
    if (s == 0 && d == 0) starts[nbk] = (uint32_t)n;
}

struct PtSmem {
    uint2 stage[PT_TILE];
    uint16_t wcount[PT_WARPS][256];
    uint16_t dig_start[256];
    uint32_t gbase[256];
    uint32_t scan[PT_T / 32 + 1];
};

// One partition level.  In-tile ranking by warp match groups into warp-private counters (as
// in radix.cu), records staged in digit order, then written as runs; the tile's slice of every
// bucket is reserved with one atomicAdd on the bucket cursor (placement inside a bucket need
// not be deterministic -- the final sort orders it).
// FIRST: the input is ukey[i] -- or, with ukey_in == nullptr, the raw-byte key read straight from the text (big-endian
// text[i .. i+3], zeros past the end: uk_keys_raw_kernel's key without a pass of its own); the record becomes
// (ukey, i | T[i-1] << 24) when packprev.
__device__ __forceinline__ uint32_t raw_key_at(const uint8_t *__restrict__ t, uint32_t i, uint32_t n, bool aligned) {
    if (aligned && i + 8 <= n) { // both words inside the text
        const uint32_t *t32 = reinterpret_cast<const uint32_t *>(t);
        return __byte_perm(t32[i >> 2], t32[(i >> 2) + 1], 0x0123u + 0x1111u * (i & 3u));
    }
    uint32_t u = 0;
    for (int k = 0; k < 4; k++) u = (u << 8) | (i + k < n ? (uint32_t)t[i + k] : 0u);
    return u;
}
template <bool FIRST>
__global__ void __launch_bounds__(PT_T, 6)
    part_kernel(const uint32_t *__restrict__ ukey_in, const uint8_t *__restrict__ text, int packprev,
                const uint2 *__restrict__ rec_in, uint2 *__restrict__ rec_out, uint64_t n, int shift, uint32_t dmask,
                const uint32_t *__restrict__ segstart, const uint32_t *__restrict__ tilebase, int nseg,
                uint32_t *__restrict__ cursor, const uint32_t *__restrict__ stop) {
    extern __shared__ __align__(16) unsigned char pt_raw[];
    PtSmem &S = *reinterpret_cast<PtSmem *>(pt_raw);
    if (stop && *stop) return; // an oversized bucket was seen (starts_kernel)
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    uint32_t beg, end, seg = 0;
    if (FIRST) {
        beg = blockIdx.x * PT_TILE;
        end = (uint32_t)min((uint64_t)beg + PT_TILE, n);
    } else {
        if (blockIdx.x >= tilebase[nseg]) return;
        seg = (uint32_t)seg_of(tilebase, nseg, blockIdx.x);
        beg = segstart[seg] + (blockIdx.x - tilebase[seg]) * PT_TILE;
        end = min(segstart[seg + 1], beg + PT_TILE);
    }
    const uint32_t cnt = end - beg;
    for (int j = threadIdx.x; j < PT_WARPS * 256 / 2; j += PT_T) reinterpret_cast<uint32_t *>(&S.wcount[0][0])[j] = 0;
    __syncthreads();
    uint32_t key[PT_ITEMS];
    uint32_t rnk[PT_ITEMS / 2]; // two 16-bit ranks per register
#pragma unroll
    for (int r = 0; r < PT_ITEMS; r++) {
        uint32_t j = w * (32 * PT_ITEMS) + r * 32 + lane;
        key[r] = 0;
        if (j < cnt) {
            if (!FIRST) key[r] = rec_in[beg + j].x;
            else if (ukey_in) key[r] = ukey_in[beg + j];
            else key[r] = raw_key_at(text, beg + j, (uint32_t)n, (reinterpret_cast<uintptr_t>(text) & 3) == 0);
        }
    }
#pragma unroll
    for (int r = 0; r < PT_ITEMS; r++) {
        uint32_t j = w * (32 * PT_ITEMS) + r * 32 + lane;
        bool valid = j < cnt;
        uint32_t d = (key[r] >> shift) & dmask;
        unsigned peers = match_bits<8>(d, valid);
        uint32_t pre = valid ? S.wcount[w][d] : 0;
        __syncwarp();
        if (valid && (peers & lt) == 0) S.wcount[w][d] = (uint16_t)(pre + __popc(peers));
        __syncwarp();
        uint32_t rk = pre + __popc(peers & lt);
        rnk[r >> 1] = (r & 1) ? (rnk[r >> 1] | (rk << 16)) : rk;
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < PT_WARPS; ww++) {
            uint32_t t = S.wcount[ww][d];
            S.wcount[ww][d] = (uint16_t)run;
            run += t;
        }
        uint32_t total;
        uint32_t ex = block_excl_sum<uint32_t, PT_T>(run, S.scan, &total);
        S.dig_start[d] = (uint16_t)ex;
        uint32_t g = run ? atomicAdd(&cursor[seg * (dmask + 1) + d], run) : 0u;
        S.gbase[d] = g - ex; // global slot of local slot 0 of digit d
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PT_ITEMS; r++) {
        uint32_t j = w * (32 * PT_ITEMS) + r * 32 + lane;
        if (j < cnt) {
            // the second word is fetched only now (keeps the ranking phase's register footprint small)
            uint32_t i = beg + j, y;
            if (FIRST) y = (packprev && i) ? (i | ((uint32_t)text[i - 1] << 24)) : i;
            else y = rec_in[i].y;
            uint32_t d = (key[r] >> shift) & dmask;
            S.stage[(uint32_t)S.dig_start[d] + S.wcount[w][d] + ((rnk[r >> 1] >> (16 * (r & 1))) & 0xffffu)] =
                make_uint2(key[r], y);
        }
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < cnt; p += PT_T) {
        uint2 v = S.stage[p];
        rec_out[S.gbase[(v.x >> shift) & dmask] + p] = v;
    }
}

// Final level: one warp sorts one bucket (<= FS_CAP records) by the low `rb` bits of ukey.
// Elements are (field << 9 | local index) in one 32-bit word, so a compare-exchange is a min and
// a max: the bucket is sorted by a bitonic network held in registers (8 or 16 elements per
// lane; partner in the same lane for small strides, one shuffle otherwise) -- no shared-memory
// counters, no match groups, about 1/3 of the instructions of two 8-bit counting rounds.
// Records whose field ties with a neighbour are ordered by their full keys (rare: about
// N^2 / 2^33 pairs).
constexpr int FS_WARPS = 8;
constexpr int FS_CAP = 512;
struct FsWarp {
    uint32_t e[FS_CAP];
};
enum { FL_TIES = 0, FL_OVERFLOW = 1 };

// Ascending sort of 32 * PER words held as v[r] of lane l = element r * 32 + l, by the bitonic network of
// 32 * REG >= 32 * PER elements whose upper REG - PER registers are +infinity and therefore never materialise: a
// compare-exchange leaves the minimum at the lower index, so one whose upper side is +infinity changes nothing
// and is skipped at compile time.  Buckets hold 256 records on average, so half of them are a little larger than
// 256: they are sorted as 320 or 384 elements instead of 512.  "Flip" form of the network: the first step of
// every merge pairs e with e ^ (k-1), the others e with e ^ j -- no direction flags.  Strides below 32 are
// shuffles, strides of 32 and more stay inside the lane.
template <int PER, int REG>
__device__ __forceinline__ void warp_bitonic(uint32_t (&v)[PER], unsigned lane) {
#pragma unroll
    for (int k = 2; k <= 32 * REG; k <<= 1) {
        if (k <= 32) { // flip inside the lanes' dimension: partner lane = lane ^ (k - 1), same register
            const bool lower = (lane & (k >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < PER; r++) {
                const uint32_t o = __shfl_xor_sync(TC_FULL, v[r], k - 1);
                v[r] = lower ? min(v[r], o) : max(v[r], o);
            }
        } else { // partner = (r ^ (k/32 - 1), lane ^ 31)
#pragma unroll
            for (int r = 0; r < PER; r++) {
                const int p = r ^ (k / 32 - 1);
                if (p > r && p < PER) {
                    const uint32_t op = __shfl_xor_sync(TC_FULL, v[p], 31), orr = __shfl_xor_sync(TC_FULL, v[r], 31);
                    v[r] = min(v[r], op);
                    v[p] = max(v[p], orr);
                }
            }
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j < 32) {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < PER; r++) {
                    const uint32_t o = __shfl_xor_sync(TC_FULL, v[r], j);
                    v[r] = lower ? min(v[r], o) : max(v[r], o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < PER; r++) {
                    const int p = r ^ (j / 32);
                    if (p > r && p < PER) {
                        const uint32_t a = min(v[r], v[p]), c = max(v[r], v[p]);
                        v[r] = a;
                        v[p] = c;
                    }
                }
            }
        }
    }
}

// load the bucket (coalesced), sort it, leave the sorted words in we[0..M)
template <int PER, int REG>
__device__ __forceinline__ bool fs_sort_bucket(const uint2 *__restrict__ rec, uint32_t M, int fsh, uint32_t fmask,
                                               uint32_t *we, unsigned lane) {
    uint32_t v[PER];
#pragma unroll
    for (int r = 0; r < PER; r++) {
        uint32_t j = lane + 32 * r; // any assignment of records to registers will do
        v[r] = j < M ? ((((rec[j].x >> fsh) & fmask) << 9) | j) : 0xffffffffu;
    }
    warp_bitonic<PER, REG>(v, lane);
    // neighbours with equal field?  (padding words are all ones and never tie with a record)
    bool tie = false;
#pragma unroll
    for (int r = 0; r < PER; r++) {
        const uint32_t nx = __shfl_down_sync(TC_FULL, v[r], 1);
        const uint32_t n0 = r + 1 < PER ? __shfl_sync(TC_FULL, v[r + 1 < PER ? r + 1 : r], 0) : 0xffffffffu;
        const uint32_t next = lane == 31 ? n0 : nx;
        tie |= (v[r] >> 9) == (next >> 9) && r * 32 + lane + 1 < M;
        if (r * 32 + lane < M) we[r * 32 + lane] = v[r];
    }
    __syncwarp();
    return __any_sync(TC_FULL, tie);
}

// Order of two suffixes whose first k symbols agree: compare the following keys, k symbols at
// a time.  A suffix that runs past the end reads as zeros, which no symbol code equals, so two
// different suffixes always differ at some step; only very long repeats exhaust `steps`.
// returns -1 / +1, or 0 if still undecided
__device__ __noinline__ int suffix_cmp_deep(const uint64_t *__restrict__ pw, int b, int kb, int k, uint64_t n,
                                               uint64_t i, uint64_t j, int steps) {
    for (int s = 1; s <= steps; s++) {
        uint64_t pi = i + (uint64_t)s * k, pj = j + (uint64_t)s * k;
        uint64_t a = pi <= n ? extract_key(pw, b, kb, pi) : 0ull;
        uint64_t c = pj <= n ? extract_key(pw, b, kb, pj) : 0ull;
        if (a != c) return a < c ? -1 : 1;
    }
    return 0;
}

// Slow path of the final sort: the calling lane insertion-sorts, by full suffix comparison,
// every run of equal field whose head sits at one of its slots (lane, lane+32, ...).  Runs are
// disjoint, so lanes never touch each other's slots.  Returns undecided pairs | overflow << 16.
__device__ __noinline__ uint32_t fs_sort_runs(uint32_t *we, const uint2 *__restrict__ wv, uint32_t M, unsigned lane,
                                              const uint64_t *__restrict__ pw, int b, int kb, int k, uint64_t n,
                                              uint32_t idxmask) {
    uint32_t undecided = 0, overflow = 0;
    for (uint32_t a = lane; a + 1 < M; a += 32) {
        const uint32_t f = we[a] >> 9;
        if ((we[a + 1] >> 9) != f || (a > 0 && (we[a - 1] >> 9) == f)) continue;
        uint32_t c = a + 2;
        while (c < M && (we[c] >> 9) == f) c++;
        if (c - a > 64) { // a long repeat: leave it to the general path
            overflow = 1;
            continue;
        }
        for (uint32_t x = a + 1; x < c; x++) {
            const uint32_t ex = we[x], sx = wv[ex & 511u].y & idxmask;
            const uint64_t kx = extract_key(pw, b, kb, sx);
            uint32_t y = x;
            while (y > a) {
                const uint32_t ey = we[y - 1], sy = wv[ey & 511u].y & idxmask;
                const uint64_t ky = extract_key(pw, b, kb, sy);
                int cmp = ky < kx ? -1 : (ky > kx ? 1 : suffix_cmp_deep(pw, b, kb, k, n, sy, sx, 16));
                undecided += cmp == 0;
                if (cmp <= 0) break;
                we[y] = ey;
                y--;
            }
            we[y] = ex;
        }
    }
    return (undecided > 0xffffu ? 0xffffu : undecided) | (overflow << 16);
}

__global__ void __launch_bounds__(FS_WARPS * 32, 5)
    final_sort_kernel(const uint2 *__restrict__ rec, const uint32_t *__restrict__ starts, uint32_t nbuckets, int rb,
                      const uint64_t *__restrict__ pw, int b, int kb, int k, uint64_t n, int packprev,
                      const uint8_t *__restrict__ text, uint32_t *__restrict__ sa, uint8_t *__restrict__ bwt,
                      uint64_t *__restrict__ primary, uint32_t *__restrict__ flags) {
    __shared__ FsWarp fs[FS_WARPS];
    FsWarp &W = fs[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    if (blockIdx.x == 0 && threadIdx.x == 0) { // the empty suffix sorts first (row 0); its BWT symbol is the last text symbol
        sa[0] = (uint32_t)n;
        if (bwt) bwt[0] = text[n - 1];
    }
    const uint32_t bk = blockIdx.x * FS_WARPS + (threadIdx.x >> 5);
    if (bk >= nbuckets) return;
    const uint32_t s = starts[bk], M = starts[bk + 1] - s;
    if (M == 0) return;
    // once any bucket has overflowed the whole sort is redone by the general path: stop working
    if (*reinterpret_cast<volatile uint32_t *>(&flags[FL_OVERFLOW])) return;
    if (M > FS_CAP) {
        if (lane == 0) atomicOr(&flags[FL_OVERFLOW], 1u);
        return;
    }
    // field = the top fbits of the rb remaining bits (all of them unless rb > 23)
    const int fbits = rb < 23 ? rb : 23;
    const int fsh = rb - fbits;
    const uint32_t fmask = (fbits ? (0xffffffffu >> (32 - fbits)) : 0u);
    const bool any_tie = M <= 256   ? fs_sort_bucket<8, 8>(rec + s, M, fsh, fmask, W.e, lane)
                         : M <= 320 ? fs_sort_bucket<10, 16>(rec + s, M, fsh, fmask, W.e, lane)
                         : M <= 384 ? fs_sort_bucket<12, 16>(rec + s, M, fsh, fmask, W.e, lane)
                                    : fs_sort_bucket<16, 16>(rec + s, M, fsh, fmask, W.e, lane);
    // runs of equal field (rare): the lane holding the head of a run sorts it by full keys
    bool head = false;
    uint32_t eqpairs = 0;
    for (uint32_t j = lane; any_tie && j + 1 < M; j += 32) {
        const uint32_t f = W.e[j] >> 9;
        const bool eq = (W.e[j + 1] >> 9) == f;
        eqpairs += eq;
        head |= eq && (j == 0 || (W.e[j - 1] >> 9) != f);
    }
    for (int dlt = 16; dlt; dlt >>= 1) eqpairs += __shfl_xor_sync(TC_FULL, eqpairs, dlt);
    if (eqpairs > 48) { // memoryless text has one or two tied pairs per bucket: this is correlated text
        if (lane == 0) atomicOr(&flags[FL_OVERFLOW], 1u);
        return;
    }
    if (__any_sync(TC_FULL, head)) {
        uint32_t r = 0;
        if (head) r = fs_sort_runs(W.e, rec + s, M, lane, pw, b, kb, k, n, packprev ? 0x00ffffffu : 0xffffffffu);
        __syncwarp();
        uint32_t und = r & 0xffffu;
        for (int dlt = 16; dlt; dlt >>= 1) und += __shfl_xor_sync(TC_FULL, und, dlt);
        if (lane == 0 && und) atomicAdd(&flags[FL_TIES], und);
        if (r >> 16) atomicOr(&flags[FL_OVERFLOW], 1u);
    }
    // slot 0 of the suffix array belongs to the empty suffix, so bucket slot j is row 1 + s + j
    for (uint32_t j = lane; j < M; j += 32) {
        uint32_t v = rec[s + (W.e[j] & 511u)].y; // the bucket was just read: an L1/L2 hit
        uint32_t idx = packprev ? (v & 0x00ffffffu) : v;
        sa[1 + s + j] = idx;
        if (bwt) {
            uint8_t c = packprev ? (uint8_t)(v >> 24) : (idx ? text[idx - 1] : (uint8_t)0);
            bwt[1 + s + j] = c;
            if (idx == 0) *primary = 1ull + s + j;
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
                                       uint64_t h, uint64_t n, const uint64_t *__restrict__ pw, int b, int kb, int gsh,
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
    key2[m] = ((uint64_t)(g[j] - 1) << gsh) | r2; // group and rank packed tightly: fewer radix passes
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
                                uint64_t n, int gsh, uint64_t *__restrict__ key2, uint32_t *__restrict__ val2) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= U) return;
    uint32_t j = cj[m];
    uint32_t s = sa[j];
    uint64_t nx = (uint64_t)s + h;
    uint32_t r2 = nx <= n ? isa[nx] : 0u; // nx <= n always holds for unresolved suffixes
    key2[m] = ((uint64_t)(g[j] - 1) << gsh) | r2; // group and rank packed tightly: fewer radix passes
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
    TC_TRY(tc_d2h_small(ctx, h, d_hist, 256 * sizeof(uint32_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, h, 256 * sizeof(uint32_t));
    return TC_OK;
}

int tc_suffix_sort_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_sa) {
    return tc_suffix_sort_bwt_dev(ctx, d_text, n, d_sa, nullptr, nullptr, nullptr);
}

// Suffix array + (optionally) the BWT in the same sweep: when d_bwt is given and the fast path
// resolves every suffix, the last level writes the BWT bytes and the primary index itself and
// *bwt_done is set; otherwise the caller emits the BWT from d_sa.
int tc_suffix_sort_bwt_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_sa, uint8_t *d_bwt,
                           uint64_t *primary /*host*/, bool *bwt_done) {
    const uint64_t N = n + 1;
    if (bwt_done) *bwt_done = false;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    if (n == 0) {
        TC_CUDA(cudaMemsetAsync(d_sa, 0, sizeof(uint32_t), ctx->stream));
        return TC_OK;
    }
    WsMark mk = tc_ws_mark(ctx);
    uint32_t hist[256];
    TC_TRY(tc_byte_hist_dev(ctx, d_text, n, hist));
    memcpy(ctx->text_hist, hist, sizeof hist);
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
    {
        void (*kpack)(const uint8_t *, uint64_t, Code256, uint64_t, uint64_t *) =
            b == 1 ? sa_pack_kernel<1> : b == 2 ? sa_pack_kernel<2> : b == 3 ? sa_pack_kernel<3>
            : b == 4 ? sa_pack_kernel<4> : b == 5 ? sa_pack_kernel<5> : b == 6 ? sa_pack_kernel<6>
            : b == 7 ? sa_pack_kernel<7> : b == 8 ? sa_pack_kernel<8> : sa_pack_kernel<9>;
        ctx->prof_bytes_next = n + 8 * nwords;
        TC_LAUNCH_AS(ctx, "sa_pack_kernel", kpack, (unsigned)ceil_div_u64(ceil_div_u64(nwords, b), 256), 256, 0, d_text, n, lut,
                     nwords, pw);
    }
    uint32_t *hU = (uint32_t *)ctx->h_scal;

    // ---- sort by key: MSD path on uniform keys (two partition levels + per-warp final sort)
    bool sorted = false;
    if (bwt_done) *bwt_done = false;
    if (n >= 4096 && n <= (25ull << 20) && !ctx->no_msd) {
        int PB = 0;
        while (PB < 16 && (n >> PB) > 256) PB++;
        const int B1 = PB < 8 ? PB : 8, B2 = PB - B1;
        const int nb1 = 1 << B1, nb2 = 1 << B2;
        const uint32_t nbk = 1u << PB;
        const int packprev = (d_bwt && n <= (1ull << 24)) ? 1 : 0;
        // lookups of g symbols each; enough of them to spend the 32 bits of ukey
        int gs = 9 / b;
        if (gs < 1) gs = 1;
        if (gs > k) gs = k;
        const int gb = gs * b;
        int G = (int)ceil(34.0 / ((double)gs * (entropy > 0.02 ? entropy : 0.02)));
        if (G > k / gs) G = k / gs;
        if (G < 1) G = 1;
        SymProb sp;
        memset(&sp, 0, sizeof sp);
        sp.p[0] = 1.0f / (float)N;
        for (int c = 0; c < 256; c++)
            if (hist[c]) sp.p[lut.code[c]] = (float)((double)hist[c] / (double)N);
        uint2 *uk_lut;
        uint32_t *ukey, *hist1, *hist2, *segstart, *cursor1, *tilebase, *chunkbase, *starts2, *cursor2, *flags;
        uint2 *recA, *recB = nullptr;
        uint64_t *d_primary = nullptr;
        TC_TRY(ws_alloc(ctx, (size_t)1 << gb, &uk_lut));
        TC_TRY(ws_alloc(ctx, n, &ukey));
        TC_TRY(ws_alloc(ctx, 256 + (size_t)nbk + 8, &hist1)); // hist1 | hist2 | flags: one memset
        hist2 = hist1 + 256;
        flags = hist2 + ((nbk + 1) & ~(size_t)1);
        d_primary = reinterpret_cast<uint64_t *>(flags + 2); // next to the flags: one small copy brings both back
        TC_TRY(ws_alloc(ctx, 257, &segstart));
        TC_TRY(ws_alloc(ctx, 256, &cursor1));
        TC_TRY(ws_alloc(ctx, 257, &tilebase));
        TC_TRY(ws_alloc(ctx, 257, &chunkbase));
        TC_TRY(ws_alloc(ctx, (size_t)nbk + 1, &starts2));
        TC_TRY(ws_alloc(ctx, (size_t)nbk, &cursor2));
        TC_TRY(ws_alloc(ctx, n, &recA));
        if (B2) TC_TRY(ws_alloc(ctx, n, &recB));
        TC_CUDA(cudaMemsetAsync(hist1, 0, (256 + (size_t)nbk + 8) * sizeof(uint32_t), ctx->stream));
        void (*kkeys)(const uint64_t *, int, int, int, int, const uint2 *, uint32_t, int, uint32_t *, uint32_t *) =
            G == 3 ? uk_keys_kernel<3> : G == 4 ? uk_keys_kernel<4> : G == 5 ? uk_keys_kernel<5>
            : G == 6 ? uk_keys_kernel<6> : uk_keys_kernel<0>;
        TC_CUDA(cudaFuncSetAttribute(kkeys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
        TC_CUDA(cudaFuncSetAttribute(seg_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
        TC_CUDA(cudaFuncSetAttribute(part_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtSmem)));
        TC_CUDA(cudaFuncSetAttribute(part_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtSmem)));
        // equiprobable bytes (every value within 5 % of n / 256): the leading bytes are the uniform key
        bool flat = sigma == 256 && !ctx->no_rawkey;
        for (int c = 0; c < 256 && flat; c++) flat = fabs((double)hist[c] * 256.0 / (double)n - 1.0) < 0.05;
        ctx->prof_bytes_next = n + 4 * n;
        const bool raw_direct = flat && B1 == 8 && nb1 == 256; // level-1 digit = first byte: no key pass at all
        Hist256 hv;
        memcpy(hv.c, hist, sizeof hv.c);
        if (raw_direct) {
            ukey = nullptr;
            ctx->prof_bytes_next = 0;
        } else if (flat) {
            TC_CUDA(cudaFuncSetAttribute(uk_keys_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
            TC_LAUNCH(ctx, uk_keys_raw_kernel, (unsigned)ceil_div_u64(n, PC_TILE), PC_T, PC_SMEM, d_text, (uint32_t)n,
                      B1 ? 32 - B1 : 31, ukey, hist1);
        } else {
            TC_LAUNCH(ctx, uk_lut_kernel, 1, 1024, 0, sp, sigma, b, gs, uk_lut);
            ctx->prof_bytes_next = n + 4 * n;
            TC_LAUNCH_AS(ctx, "uk_keys_kernel", kkeys, (unsigned)ceil_div_u64(n, PC_TILE), PC_T, PC_SMEM, pw, b, kb, gb, G, uk_lut,
                         (uint32_t)n, B1 ? 32 - B1 : 31, ukey, hist1);
        }
        TC_LAUNCH(ctx, seg_tables_kernel, 1, 256, 0, raw_direct ? (const uint32_t *)nullptr : (const uint32_t *)hist1, hv, nb1,
                  segstart, cursor1, tilebase, chunkbase);
        ctx->prof_bytes_next = 4 * n + (packprev ? n : 0) + 8 * n;
        TC_LAUNCH(ctx, part_kernel<true>, (unsigned)ceil_div_u64(n, PT_TILE), PT_T, sizeof(PtSmem), ukey, d_text, packprev,
                  (const uint2 *)nullptr, recA, n, 32 - B1, (uint32_t)(nb1 - 1), (const uint32_t *)nullptr,
                  (const uint32_t *)nullptr, 1, cursor1, (const uint32_t *)nullptr);
        const uint2 *recF = recA;
        const uint32_t *startsF = segstart;
        if (B2) {
            ctx->prof_bytes_next = 8 * n;
            TC_LAUNCH(ctx, seg_hist_kernel, (unsigned)ceil_div_u64(ceil_div_u64(n, PC_WCHUNK) + nb1, PC_WARPS), PC_T,
                      PC_SMEM, recA, segstart, chunkbase, nb1, 32 - PB, (uint32_t)(nb2 - 1), hist2);
            TC_LAUNCH(ctx, starts_kernel, nb1, 256, 0, hist2, segstart, (uint32_t)nb2, nbk, n, starts2, cursor2,
                      (uint32_t)FS_CAP, flags + FL_OVERFLOW);
            ctx->prof_bytes_next = 16 * n;
            TC_LAUNCH(ctx, part_kernel<false>, (unsigned)(ceil_div_u64(n, PT_TILE) + nb1), PT_T, sizeof(PtSmem),
                      (const uint32_t *)nullptr, (const uint8_t *)nullptr, 0, (const uint2 *)recA, recB, n, 32 - PB,
                      (uint32_t)(nb2 - 1), segstart, tilebase, nb1, cursor2, (const uint32_t *)(flags + FL_OVERFLOW));
            recF = recB;
            startsF = starts2;
        }
        ctx->prof_bytes_next = 8 * n + 4 * n + (d_bwt ? n : 0);
        TC_LAUNCH(ctx, final_sort_kernel, (unsigned)ceil_div_u64(nbk, FS_WARPS), FS_WARPS * 32, 0, recF, startsF, nbk,
                  32 - PB, pw, b, kb, k, n, packprev, d_text, d_sa, d_bwt, d_primary, flags);
        TC_TRY(tc_d2h_small(ctx, hU, flags, 4 * sizeof(uint32_t)));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (hU[FL_OVERFLOW] == 0) {
            sorted = true;
            if (hU[FL_TIES] == 0) { // every suffix is already distinguished by its key: done
                if (bwt_done && d_bwt) {
                    *bwt_done = true;
                    if (primary) *primary = ctx->h_scal[1];
                }
                tc_ws_release(ctx, mk);
                return TC_OK;
            }
            TC_TRY(ws_alloc(ctx, N, &keys));
            TC_LAUNCH(ctx, sa_gather_keys_kernel, gridN, 256, 0, pw, b, kb, d_sa, N, keys);
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
    TC_TRY(tc_d2h_small(ctx, hU, d_U, sizeof(uint32_t)));
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
        for (int s = 0; s < 2 * rb; s += 8) sh2[np2++] = s; // key = group << rb | rank: 2*rb bits
        uint64_t h = (uint64_t)k;
        for (int round = 0; U > 0; round++) {
            if (round > 48) {
                snprintf(ctx->err, sizeof ctx->err, "suffix sort did not converge");
                return TC_E_CUDA;
            }
            const unsigned gridU = (unsigned)ceil_div_u64(U, 256);
            if (ctx->diag & 4) fprintf(stderr, "[tc_b200] doubling round %d: h = %llu, unresolved %llu of %llu\n", round, (unsigned long long)h, (unsigned long long)U, (unsigned long long)N);
            if (round == 0) {
                TC_LAUNCH(ctx, sa_keys2_lookup_kernel, gridU, 256, 0, cjA, U, d_sa, keys, N, g, h, n, pw, b, kb, rb, key2a,
                          val2a);
            } else {
                if (!isa) { // second round: now the inverse suffix array pays for itself
                    TC_TRY(ws_alloc(ctx, N, &isa));
                    TC_LAUNCH(ctx, sa_build_isa_kernel, gridN, 256, 0, g, d_sa, N, isa);
                }
                TC_LAUNCH(ctx, sa_keys2_kernel, gridU, 256, 0, cjA, U, d_sa, isa, g, h, n, rb, key2a, val2a);
            }
            uint64_t *k2s;
            uint32_t *v2s;
            TC_TRY(tc_radix_sort_pairs(ctx, key2a, val2a, key2b, val2b, U, sh2, np2, &k2s, &v2s));
            TC_LAUNCH(ctx, sa_heads_kernel, gridU, 256, 0, k2s, U, g2);
            TC_TRY(tc_scan_inclusive_max_u32(ctx, g2, g2, U));
            TC_LAUNCH(ctx, sa_update_kernel, gridU, 256, 0, g2, cjA, v2s, U, d_sa, isa, g, ns2);
            TC_TRY(tc_scan_exclusive_u32(ctx, ns2, cpos2, U, d_U));
            TC_LAUNCH(ctx, sa_compact_kernel, gridU, 256, 0, ns2, cpos2, (const uint32_t *)cjA, U, cjB);
            TC_TRY(tc_d2h_small(ctx, hU, d_U, sizeof(uint32_t)));
            TC_CUDA(cudaStreamSynchronize(ctx->stream));
            U = hU[0];
            std::swap(cjA, cjB);
            h *= 2;
        }
    }
    tc_ws_release(ctx, mk);
    return TC_OK;
}
