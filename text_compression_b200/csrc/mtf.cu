// mtf.cu -- seqToMTF / seqFromMTF (src/Data/MTF/Internal.hs:128-232) as a chunked
// parallel scan over composable move-to-front summaries.
//
// Symbols are first mapped to alphabet ranks 0..sigma-1 (nubSeq', :79-99: sorted,
// Nothing first), so the initial list L0 is the identity permutation.
//
// Encode, generic alphabet (sigma <= 257):
//   K1  thread per chunk: recency list of the chunk (distinct ranks, most recent first)
//   K2a warp per tile of chunks: exclusive chain of   acc <- r ++ (acc \ r)   inside the tile
//   K2b one warp: exclusive chain over tiles starting from L0 (also yields the FINAL list,
//       the second component seqToMTF returns)
//   K3  thread per chunk: replay.  The list is never materialised: every symbol owns a
//       time slot (virtual slots for the incoming order, then one slot per position) and a
//       bitmap marks the slots that are "latest occurrence of their symbol".  The MTF index
//       of c is popcount(bitmap in (last[c], now)) -- O(gap/32) instead of O(sigma).
// Decode: same chunk/tile structure; a chunk's summary is the permutation it applies to
//   list positions, composition is a gather.
// Algorithmic bytes: N * (1 + w_idx) (u8 symbols in, u16 indices out here: 3N).
#include <algorithm>

#include "common.cuh"
#include "impl.cuh"

namespace {
constexpr int SIGMAX = 257;
constexpr int LISTPAD = 264;
constexpr int LASTW = (SIGMAX + 1) / 2; // u16 pairs per thread, as 32-bit words

struct Lut {
    uint16_t rank[SIGMAX]; // symbol+1 -> alphabet rank, 0xffff if absent
};

struct SrcU8 {
    const uint8_t *p;
    uint64_t primary;
    __device__ __forceinline__ int at(uint64_t i) const { return i == primary ? 0 : (int)p[i] + 1; }
    static constexpr int VEC = 16;
    __device__ __forceinline__ bool can_vec(uint64_t base) const {
        return (reinterpret_cast<uintptr_t>(p + base) & 15) == 0;
    }
    __device__ __forceinline__ void load_vec(uint64_t base, int *c) const {
        uint4 v = *reinterpret_cast<const uint4 *>(p + base);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) c[k] = (int)((w[k >> 2] >> ((k & 3) * 8)) & 0xff) + 1;
        if (primary >= base && primary < base + 16) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (base + k == primary) c[k] = 0;
        }
    }
};
struct SrcI16 {
    const int16_t *p;
    __device__ __forceinline__ int at(uint64_t i) const {
        int v = p[i];
        return v < 0 ? 0 : (v & 0xff) + 1;
    }
    static constexpr int VEC = 16;
    __device__ __forceinline__ bool can_vec(uint64_t base) const {
        return (reinterpret_cast<uintptr_t>(p + base) & 15) == 0;
    }
    __device__ __forceinline__ void load_vec(uint64_t base, int *c) const {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            uint4 v = *reinterpret_cast<const uint4 *>(p + base + q * 8);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int s = (int)(int16_t)((w[k >> 1] >> ((k & 1) * 16)) & 0xffff);
                c[q * 8 + k] = s < 0 ? 0 : (s & 0xff) + 1;
            }
        }
    }
};

// ---- presence of each of the 257 codes ---------------------------------------------
template <class Src>
__global__ void __launch_bounds__(256) mtf_presence_kernel(Src src, uint64_t N, uint32_t *__restrict__ present) {
    __shared__ uint32_t s[SIGMAX];
    for (int j = threadIdx.x; j < SIGMAX; j += 256) s[j] = 0;
    __syncthreads();
    uint64_t stride = (uint64_t)gridDim.x * 256 * 16;
    for (uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 16; base < N; base += stride) {
        int c[16];
        if (base + 16 <= N && src.can_vec(base)) {
            src.load_vec(base, c);
#pragma unroll
            for (int k = 0; k < 16; k++) s[c[k]] = 1;
        } else {
            for (int k = 0; k < 16 && base + k < N; k++) s[src.at(base + k)] = 1;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < SIGMAX; j += 256)
        if (s[j]) present[j] = 1;
}

// ---- per-thread state in shared memory, word-interleaved across the block -------------
// word w of thread t lives at sm[w * blockDim.x + t]: conflict-free for any per-thread index.
struct TState {
    uint32_t *sm;
    __device__ __forceinline__ uint32_t &word(int w) const { return sm[w * blockDim.x + threadIdx.x]; }
    __device__ __forceinline__ uint16_t get16(int i) const {
        return reinterpret_cast<const uint16_t *>(&sm[(i >> 1) * blockDim.x + threadIdx.x])[i & 1];
    }
    __device__ __forceinline__ void set16(int i, uint16_t v) const {
        reinterpret_cast<uint16_t *>(&sm[(i >> 1) * blockDim.x + threadIdx.x])[i & 1] = v;
    }
};

// ---- K1: recency list of each chunk -------------------------------------------------
template <class Src>
__global__ void mtf_recency_kernel(Src src, Lut lut, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t sigma,
                                   uint16_t *__restrict__ rec, uint32_t *__restrict__ reclen) {
    extern __shared__ uint32_t smem[];
    __shared__ uint16_t s_rank[SIGMAX];
    for (int j = threadIdx.x; j < SIGMAX; j += blockDim.x) s_rank[j] = lut.rank[j];
    __syncthreads();
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    TState seen{smem}; // 9 words per thread
#pragma unroll
    for (int w = 0; w < 9; w++) seen.word(w) = 0;
    uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    uint16_t *out = rec + k * sigma;
    uint32_t n = 0;
    // walk backwards in 16-symbol vectors
    uint64_t i = end;
    while (i > beg) {
        uint64_t vb = (i - 1) & ~uint64_t(15);
        if (vb < beg) vb = beg;
        int c[16];
        int cnt = (int)(i - vb);
        if (cnt == 16 && src.can_vec(vb)) {
            src.load_vec(vb, c);
        } else {
            for (int q = 0; q < cnt; q++) c[q] = src.at(vb + q);
        }
        for (int q = cnt - 1; q >= 0; q--) {
            int r = s_rank[c[q]];
            uint32_t w = seen.word(r >> 5);
            uint32_t bit = 1u << (r & 31);
            if (!(w & bit)) {
                seen.word(r >> 5) = w | bit;
                out[n++] = (uint16_t)r;
            }
        }
        i = vb;
        if (n == sigma) break;
    }
    reclen[k] = n;
}

// dst = r ++ (acc \ r) -- the list after applying a chunk with recency list r to list acc.
__device__ __forceinline__ int compose_lists(uint16_t *dst, const uint16_t *acc, int alen, const uint16_t *r, int rlen,
                                             uint32_t *bm) {
    const unsigned lane = lane_id();
    if (lane < 9) bm[lane] = 0;
    __syncwarp();
    for (int j = lane; j < rlen; j += 32) {
        uint32_t c = r[j];
        dst[j] = (uint16_t)c;
        atomicOr(&bm[c >> 5], 1u << (c & 31));
    }
    __syncwarp();
    int out = rlen;
    for (int b = 0; b < alen; b += 32) {
        int j = b + lane;
        bool valid = j < alen;
        uint32_t c = valid ? acc[j] : 0;
        bool keep = valid && !((bm[c >> 5] >> (c & 31)) & 1);
        unsigned m = __ballot_sync(TC_FULL, keep);
        if (keep) dst[out + __popc(m & lanemask_lt())] = (uint16_t)c;
        out += __popc(m);
    }
    __syncwarp();
    return out;
}

// ---- K2a: exclusive chain inside each tile of G chunks (one warp per tile) ---------------
__global__ void __launch_bounds__(128)
    mtf_tile_chain_kernel(const uint16_t *__restrict__ rec, const uint32_t *__restrict__ reclen, uint64_t nchunks,
                          uint32_t G, uint32_t sigma, uint16_t *__restrict__ part, uint32_t *__restrict__ partlen,
                          uint16_t *__restrict__ tilesum, uint32_t *__restrict__ tilesumlen, uint64_t ntiles) {
    __shared__ uint16_t bufs[4][2][LISTPAD];
    __shared__ uint32_t bms[4][9];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    uint64_t t = (uint64_t)blockIdx.x * 4 + w;
    if (t >= ntiles) return;
    uint16_t *A = bufs[w][0], *B = bufs[w][1];
    int alen = 0;
    uint64_t k0 = t * G, k1 = k0 + G < nchunks ? k0 + G : nchunks;
    for (uint64_t k = k0; k < k1; k++) {
        for (int j = lane; j < alen; j += 32) part[k * sigma + j] = A[j];
        if (lane == 0) partlen[k] = alen;
        alen = compose_lists(B, A, alen, rec + k * sigma, (int)reclen[k], bms[w]);
        uint16_t *tmp = A;
        A = B;
        B = tmp;
    }
    for (int j = lane; j < alen; j += 32) tilesum[t * sigma + j] = A[j];
    if (lane == 0) tilesumlen[t] = alen;
}

// ---- K2b: exclusive chain over tiles from L0 = identity (one warp) ------------------------
__global__ void __launch_bounds__(32)
    mtf_top_chain_kernel(const uint16_t *__restrict__ tilesum, const uint32_t *__restrict__ tilesumlen,
                         uint64_t ntiles, uint32_t sigma, uint16_t *__restrict__ tileprefix,
                         uint16_t *__restrict__ final_list) {
    __shared__ uint16_t bufs[2][LISTPAD];
    __shared__ uint32_t bm[9];
    const unsigned lane = lane_id();
    uint16_t *A = bufs[0], *B = bufs[1];
    for (int j = lane; j < (int)sigma; j += 32) A[j] = (uint16_t)j;
    __syncwarp();
    for (uint64_t t = 0; t < ntiles; t++) {
        for (int j = lane; j < (int)sigma; j += 32) tileprefix[t * sigma + j] = A[j];
        compose_lists(B, A, (int)sigma, tilesum + t * sigma, (int)tilesumlen[t], bm);
        uint16_t *tmp = A;
        A = B;
        B = tmp;
    }
    for (int j = lane; j < (int)sigma; j += 32) final_list[j] = A[j];
}

// ---- K3: replay ----------------------------------------------------------------------------
// per-thread shared state: LASTW words of last[] (u16 slot per rank) + BW bitmap words.
template <class Src>
__global__ void mtf_replay_kernel(Src src, Lut lut, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t G,
                                  uint32_t sigma, const uint16_t *__restrict__ part,
                                  const uint32_t *__restrict__ partlen, const uint16_t *__restrict__ tileprefix,
                                  uint16_t *__restrict__ idx_out) {
    extern __shared__ uint32_t smem[];
    __shared__ uint16_t s_rank[SIGMAX];
    for (int j = threadIdx.x; j < SIGMAX; j += blockDim.x) s_rank[j] = lut.rank[j];
    __syncthreads();
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    TState last{smem};
    TState bits{smem + LASTW * blockDim.x};
    const int BW = (int)((sigma + L + 31) / 32);
    for (int w = 0; w < BW; w++) bits.word(w) = 0;
    for (int w = 0; w < (int)sigma / 32; w++) bits.word(w) = 0xffffffffu;
    if (sigma & 31) bits.word(sigma / 32) = (1u << (sigma & 31)) - 1;
    for (int r = 0; r < (int)sigma; r++) last.set16(r, 0xffff);
    // incoming order: the tile-local partial list first, then the tile prefix for the rest.
    // list position j owns virtual slot sigma-1-j (front of the list == most recent).
    uint32_t pl = partlen[k];
    const uint16_t *pp = part + k * sigma;
    for (uint32_t j = 0; j < pl; j++) last.set16(pp[j], (uint16_t)(sigma - 1 - j));
    const uint16_t *tp = tileprefix + (k / G) * sigma;
    uint32_t j2 = pl;
    for (uint32_t j = 0; j < sigma && j2 < sigma; j++) {
        uint16_t r = tp[j];
        if (last.get16(r) == 0xffff) last.set16(r, (uint16_t)(sigma - 1 - j2++));
    }
    uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    const bool vec_out = (reinterpret_cast<uintptr_t>(idx_out + beg) & 15) == 0;
    for (uint64_t vb = beg; vb < end; vb += 16) {
        int c[16];
        int cnt = (int)(end - vb < 16 ? end - vb : 16);
        if (cnt == 16 && src.can_vec(vb)) {
            src.load_vec(vb, c);
        } else {
            for (int q = 0; q < cnt; q++) c[q] = src.at(vb + q);
        }
        uint32_t o[16];
        for (int q = 0; q < cnt; q++) {
            const int r = s_rank[c[q]];
            const uint32_t a = last.get16(r);
            const uint32_t now = sigma + (uint32_t)(vb - beg) + q;
            const int wa = a >> 5, wn = now >> 5;
            // popcount of set bits strictly between a and now (bit `now` is not set yet)
            uint32_t cntbits;
            uint32_t wa_bits = bits.word(wa);
            uint32_t hi_a = (wa_bits >> (a & 31)) >> 1; // bits above a inside its word
            if (wa == wn) {
                cntbits = __popc(hi_a & ((1u << ((now & 31) - (a & 31) - 1)) - 1));
            } else {
                cntbits = __popc(hi_a);
                for (int w = wa + 1; w < wn; w++) cntbits += __popc(bits.word(w));
                cntbits += __popc(bits.word(wn) & ((1u << (now & 31)) - 1));
            }
            o[q] = cntbits;
            bits.word(wa) = wa_bits & ~(1u << (a & 31));
            bits.word(wn) |= 1u << (now & 31);
            last.set16(r, (uint16_t)now);
        }
        if (cnt == 16 && vec_out) {
            uint4 w0, w1;
            w0.x = o[0] | (o[1] << 16);
            w0.y = o[2] | (o[3] << 16);
            w0.z = o[4] | (o[5] << 16);
            w0.w = o[6] | (o[7] << 16);
            w1.x = o[8] | (o[9] << 16);
            w1.y = o[10] | (o[11] << 16);
            w1.z = o[12] | (o[13] << 16);
            w1.w = o[14] | (o[15] << 16);
            *reinterpret_cast<uint4 *>(idx_out + vb) = w0;
            *reinterpret_cast<uint4 *>(idx_out + vb + 8) = w1;
        } else {
            for (int q = 0; q < cnt; q++) idx_out[vb + q] = (uint16_t)o[q];
        }
    }
}

// ---- decode ----------------------------------------------------------------------------------
// D1: permutation each chunk applies to list positions (replay on the identity list).
__global__ void mtfd_perm_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t L, uint64_t nchunks,
                                 uint32_t sigma, uint16_t *__restrict__ perm, uint32_t *__restrict__ err) {
    extern __shared__ uint32_t smem[];
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    TState lst{smem};
    for (int j = 0; j < (int)sigma; j++) lst.set16(j, (uint16_t)j);
    uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    for (uint64_t i = beg; i < end; i++) {
        uint32_t r = idx[i];
        if (r >= sigma) {
            atomicMax(err, 1u);
            r = 0;
        }
        uint16_t h = lst.get16(r);
        for (int j = (int)r; j > 0; j--) lst.set16(j, lst.get16(j - 1));
        lst.set16(0, h);
    }
    for (int j = 0; j < (int)sigma; j++) perm[k * sigma + j] = lst.get16(j);
}

// D2a: exclusive chain of permutations inside a tile: acc'[j] = acc[perm[j]].
__global__ void __launch_bounds__(128)
    mtfd_tile_chain_kernel(const uint16_t *__restrict__ perm, uint64_t nchunks, uint32_t G, uint32_t sigma,
                           uint16_t *__restrict__ part, uint16_t *__restrict__ tilesum, uint64_t ntiles) {
    __shared__ uint16_t bufs[4][2][LISTPAD];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    uint64_t t = (uint64_t)blockIdx.x * 4 + w;
    if (t >= ntiles) return;
    uint16_t *A = bufs[w][0], *B = bufs[w][1];
    for (int j = lane; j < (int)sigma; j += 32) A[j] = (uint16_t)j;
    __syncwarp();
    uint64_t k0 = t * G, k1 = k0 + G < nchunks ? k0 + G : nchunks;
    for (uint64_t k = k0; k < k1; k++) {
        for (int j = lane; j < (int)sigma; j += 32) {
            part[k * sigma + j] = A[j];
            B[j] = A[perm[k * sigma + j]];
        }
        __syncwarp();
        uint16_t *tmp = A;
        A = B;
        B = tmp;
    }
    for (int j = lane; j < (int)sigma; j += 32) tilesum[t * sigma + j] = A[j];
}

// D2b: exclusive chain over tiles starting from L0 (actual symbols).
struct List0 {
    int16_t sym[SIGMAX];
};
__global__ void __launch_bounds__(32)
    mtfd_top_chain_kernel(const uint16_t *__restrict__ tilesum, uint64_t ntiles, uint32_t sigma, List0 l0,
                          int16_t *__restrict__ tileprefix) {
    __shared__ int16_t bufs[2][LISTPAD];
    const unsigned lane = lane_id();
    int16_t *A = bufs[0], *B = bufs[1];
    for (int j = lane; j < (int)sigma; j += 32) A[j] = l0.sym[j];
    __syncwarp();
    for (uint64_t t = 0; t < ntiles; t++) {
        for (int j = lane; j < (int)sigma; j += 32) {
            tileprefix[t * sigma + j] = A[j];
            B[j] = A[tilesum[t * sigma + j]];
        }
        __syncwarp();
        int16_t *tmp = A;
        A = B;
        B = tmp;
    }
}

// D3: replay with the real symbols.
__global__ void mtfd_replay_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t L, uint64_t nchunks,
                                   uint32_t G, uint32_t sigma, const uint16_t *__restrict__ part,
                                   const int16_t *__restrict__ tileprefix, int16_t *__restrict__ out) {
    extern __shared__ uint32_t smem[];
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    TState lst{smem};
    const int16_t *tp = tileprefix + (k / G) * sigma;
    const uint16_t *pp = part + k * sigma;
    for (int j = 0; j < (int)sigma; j++) lst.set16(j, (uint16_t)tp[pp[j]]);
    uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    for (uint64_t i = beg; i < end; i++) {
        uint32_t r = idx[i];
        if (r >= sigma) r = 0;
        uint16_t h = lst.get16(r);
        for (int j = (int)r; j > 0; j--) lst.set16(j, lst.get16(j - 1));
        lst.set16(0, h);
        out[i] = (int16_t)h;
    }
}

uint32_t pick_chunk_len(tc_ctx *ctx, uint64_t N, uint32_t lo, uint32_t hi) {
    uint64_t target_threads = (uint64_t)ctx->sm_count * 128;
    uint64_t L = ceil_div_u64(N, target_threads);
    L = (L + 15) / 16 * 16;
    if (L < lo) L = lo;
    if (L > hi) L = hi;
    return (uint32_t)L;
}

template <class Src>
int mtf_encode_impl(tc_ctx *ctx, Src src, uint64_t N, uint16_t *d_idx, int16_t *final_list, uint32_t *sigma_out) {
    *sigma_out = 0;
    if (N == 0) return TC_OK;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    WsMark mk = tc_ws_mark(ctx);
    // alphabet = nubSeq' (sorted, Nothing first)
    uint32_t *d_present;
    TC_TRY(ws_alloc(ctx, SIGMAX, &d_present));
    TC_CUDA(cudaMemsetAsync(d_present, 0, SIGMAX * sizeof(uint32_t), ctx->stream));
    unsigned pgrid = (unsigned)std::min<uint64_t>(ceil_div_u64(N, 256 * 16), (uint64_t)ctx->sm_count * 8);
    TC_LAUNCH(ctx, (mtf_presence_kernel<Src>), pgrid, 256, 0, src, N, d_present);
    uint32_t *h_present = (uint32_t *)ctx->h_scal;
    TC_CUDA(cudaMemcpyAsync(h_present, d_present, SIGMAX * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    Lut lut;
    int16_t alpha[SIGMAX];
    uint32_t sigma = 0;
    for (int c = 0; c < SIGMAX; c++) {
        if (h_present[c]) {
            lut.rank[c] = (uint16_t)sigma;
            alpha[sigma++] = (int16_t)(c - 1);
        } else {
            lut.rank[c] = 0;
        }
    }
    const uint32_t L = pick_chunk_len(ctx, N, 64, 1024);
    const uint64_t nchunks = ceil_div_u64(N, L);
    const uint32_t G = 128;
    const uint64_t ntiles = ceil_div_u64(nchunks, G);
    uint16_t *rec, *part, *tilesum, *tileprefix, *d_final;
    uint32_t *reclen, *partlen, *tilesumlen;
    TC_TRY(ws_alloc(ctx, nchunks * sigma, &rec));
    TC_TRY(ws_alloc(ctx, nchunks, &reclen));
    TC_TRY(ws_alloc(ctx, nchunks * sigma, &part));
    TC_TRY(ws_alloc(ctx, nchunks, &partlen));
    TC_TRY(ws_alloc(ctx, ntiles * sigma, &tilesum));
    TC_TRY(ws_alloc(ctx, ntiles, &tilesumlen));
    TC_TRY(ws_alloc(ctx, ntiles * sigma, &tileprefix));
    TC_TRY(ws_alloc(ctx, SIGMAX, &d_final));
    const int T = 64;
    unsigned cgrid = (unsigned)ceil_div_u64(nchunks, T);
    TC_LAUNCH(ctx, (mtf_recency_kernel<Src>), cgrid, T, 9 * T * sizeof(uint32_t), src, lut, N, L, nchunks, sigma, rec,
              reclen);
    TC_LAUNCH(ctx, mtf_tile_chain_kernel, (unsigned)ceil_div_u64(ntiles, 4), 128, 0, rec, reclen, nchunks, G, sigma,
              part, partlen, tilesum, tilesumlen, ntiles);
    TC_LAUNCH(ctx, mtf_top_chain_kernel, 1, 32, 0, tilesum, tilesumlen, ntiles, sigma, tileprefix, d_final);
    const int BW = (int)((sigma + L + 31) / 32);
    size_t smem = (size_t)(LASTW + BW) * T * sizeof(uint32_t);
    if (smem > 48 * 1024)
        TC_CUDA(cudaFuncSetAttribute(mtf_replay_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TC_LAUNCH(ctx, (mtf_replay_kernel<Src>), cgrid, T, smem, src, lut, N, L, nchunks, G, sigma, part, partlen,
              tileprefix, d_idx);
    uint16_t *h_final = (uint16_t *)ctx->h_scal;
    TC_CUDA(cudaMemcpyAsync(h_final, d_final, sigma * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t j = 0; j < sigma; j++) final_list[j] = alpha[h_final[j]];
    *sigma_out = sigma;
    tc_ws_release(ctx, mk);
    return TC_OK;
}
} // namespace

int mtf_encode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint16_t *d_idx,
                           int16_t *final_list, uint32_t *sigma) {
    if (N && primary >= N) primary = ~0ull;
    return mtf_encode_impl(ctx, SrcU8{d_bwt, primary}, N, d_idx, final_list, sigma);
}
int mtf_encode_i16_dev_impl(tc_ctx *ctx, const int16_t *d_sym, uint64_t N, uint16_t *d_idx, int16_t *final_list,
                               uint32_t *sigma) {
    return mtf_encode_impl(ctx, SrcI16{d_sym}, N, d_idx, final_list, sigma);
}

// seqFromMTF: initial list = nubSeq' (final list) (src/Data/MTF/Internal.hs:214)
int mtf_decode_dev_impl(tc_ctx *ctx, const uint16_t *d_idx, uint64_t N, const int16_t *final_list, uint32_t sigma_in,
                           int16_t *d_sym) {
    if (N == 0 || sigma_in == 0) return TC_OK; // empty guards (:202-209)
    if (sigma_in > SIGMAX) return TC_E_ARG;
    bool seen[SIGMAX] = {false};
    for (uint32_t j = 0; j < sigma_in; j++) {
        int v = final_list[j];
        seen[v < 0 ? 0 : (v & 0xff) + 1] = true;
    }
    List0 l0;
    uint32_t sigma = 0;
    for (int c = 0; c < SIGMAX; c++)
        if (seen[c]) l0.sym[sigma++] = (int16_t)(c - 1);
    WsMark mk = tc_ws_mark(ctx);
    const uint32_t L = pick_chunk_len(ctx, N, 64, 512);
    const uint64_t nchunks = ceil_div_u64(N, L);
    const uint32_t G = 128;
    const uint64_t ntiles = ceil_div_u64(nchunks, G);
    uint16_t *perm, *part, *tilesum;
    int16_t *tileprefix;
    uint32_t *d_err;
    TC_TRY(ws_alloc(ctx, nchunks * sigma, &perm));
    TC_TRY(ws_alloc(ctx, nchunks * sigma, &part));
    TC_TRY(ws_alloc(ctx, ntiles * sigma, &tilesum));
    TC_TRY(ws_alloc(ctx, ntiles * sigma, &tileprefix));
    TC_TRY(ws_alloc(ctx, 1, &d_err));
    TC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream));
    const int T = 64;
    unsigned cgrid = (unsigned)ceil_div_u64(nchunks, T);
    size_t smem = (size_t)LASTW * T * sizeof(uint32_t);
    TC_LAUNCH(ctx, mtfd_perm_kernel, cgrid, T, smem, d_idx, N, L, nchunks, sigma, perm, d_err);
    TC_LAUNCH(ctx, mtfd_tile_chain_kernel, (unsigned)ceil_div_u64(ntiles, 4), 128, 0, perm, nchunks, G, sigma, part,
              tilesum, ntiles);
    TC_LAUNCH(ctx, mtfd_top_chain_kernel, 1, 32, 0, tilesum, ntiles, sigma, l0, tileprefix);
    TC_LAUNCH(ctx, mtfd_replay_kernel, cgrid, T, smem, d_idx, N, L, nchunks, G, sigma, part, tileprefix, d_sym);
    uint32_t *h_err = (uint32_t *)ctx->h_scal;
    TC_CUDA(cudaMemcpyAsync(h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    tc_ws_release(ctx, mk);
    return h_err[0] ? TC_E_INDEX : TC_OK;
}
