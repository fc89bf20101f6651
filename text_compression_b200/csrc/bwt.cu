// bwt.cu -- saToBWT (src/Data/BWT/Internal.hs:98-106) and fromBWT
// (src/Data/BWT.hs:93-104 + sortTB / magicInverseBWT, src/Data/BWT/Internal.hs:144-200).
//
// Inverse: the reference sorts (symbol, index) pairs -- a stable counting
// sort, done here in one pass over the 257 symbol codes (cs_* kernels) -- which yields psi, then walks
//   f = psi[0]; while f != 0: emit F[f]; f = psi[f]
// The walk is a linked list over rows; it is parallelised by list ranking with splitters:
// every K-th row is a splitter, each splitter walks to the next splitter (sub-list length),
// the reduced list is ranked by pointer jumping, and a second walk writes the text.
#include <algorithm>
#include <utility>

#include <cooperative_groups.h>

#include "common.cuh"
#include "impl.cuh"

int tc_byte_hist_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *h_hist);

namespace cg = cooperative_groups;
namespace {
__global__ void bwt_emit_kernel(const uint8_t *__restrict__ t, const uint32_t *__restrict__ sa, uint64_t N,
                                uint8_t *__restrict__ bwt, uint64_t *__restrict__ d_primary,
                                uint32_t *__restrict__ sa1) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    uint32_t s = sa[j];
    if (s == 0) {
        bwt[j] = 0;
        *d_primary = j;
    } else {
        bwt[j] = t[s - 1];
    }
    if (sa1) sa1[j] = s + 1;
}

__global__ void sa_plus1_kernel(const uint32_t *__restrict__ sa, uint64_t N, uint32_t *__restrict__ sa1) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) sa1[j] = sa[j] + 1;
}

struct CodeU8 {
    const uint8_t *p;
    uint64_t primary;
    __device__ __forceinline__ uint32_t at(uint64_t i) const { return i == primary ? 0u : (uint32_t)p[i] + 1; }
};
struct CodeI16 {
    const int16_t *p;
    __device__ __forceinline__ uint32_t at(uint64_t i) const {
        int v = p[i];
        return v < 0 ? 0u : (uint32_t)(v & 0xff) + 1;
    }
};

// ---- psi by ONE stable counting pass over the 257 symbol codes ----------------------------------
// (the generic radix sort needs two 8-bit passes over (key, value) pairs for a 9-bit code)
//   cs_hist    per-tile code counts, code-major matrix hist[code * tiles + tile]
//   cs_rows    exclusive scan of every code's row over tiles + row totals
//   cs_base    exclusive scan of the 257 totals = first row of every code in F (the C table)
//   cs_scatter stable rank inside the tile (warp match groups, warp-private counters), then
//              psi[base[code] + offs[code][tile] + rank] = position
constexpr int CS_T = 256;
constexpr int CS_ITEMS = 16;
constexpr int CS_TILE = CS_T * CS_ITEMS;
constexpr int CS_WARPS = CS_T / 32;
constexpr int CS_BINS = 288; // 257 codes, padded

template <class Src>
__global__ void __launch_bounds__(CS_T)
    cs_hist_kernel(Src src, uint64_t N, uint32_t *__restrict__ hist, uint64_t tiles) {
    __shared__ uint32_t h[CS_WARPS][CS_BINS];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    for (int j = threadIdx.x; j < CS_WARPS * CS_BINS; j += CS_T) (&h[0][0])[j] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t)blockIdx.x * CS_TILE + (uint64_t)w * (32 * CS_ITEMS);
#pragma unroll 4
    for (int r = 0; r < CS_ITEMS; r++) {
        const uint64_t i = wbase + (uint64_t)r * 32 + lane;
        const bool valid = i < N;
        const uint32_t c = valid ? src.at(i) : 0u;
        const unsigned peers = match_bits<9>(c, valid);
        if (valid && (peers & lanemask_lt()) == 0) h[w][c] += __popc(peers); // warp-private: no atomics
        __syncwarp();
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 257; c += CS_T) {
        uint32_t s = 0;
#pragma unroll
        for (int ww = 0; ww < CS_WARPS; ww++) s += h[ww][c];
        hist[(uint64_t)c * tiles + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
    cs_rows_kernel(uint32_t *__restrict__ hist, uint64_t tiles, uint32_t *__restrict__ totals) {
    __shared__ uint32_t sh[1024 / 32 + 1];
    uint32_t *row = hist + (uint64_t)blockIdx.x * tiles;
    uint32_t carry = 0;
    for (uint64_t b0 = 0; b0 < tiles; b0 += 1024) {
        const uint64_t i = b0 + threadIdx.x;
        const uint32_t v = i < tiles ? row[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_sum<uint32_t, 1024>(v, sh, &tot);
        if (i < tiles) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// totals[0..257) -> exclusive scan in place (first F row per code) and a copy of the counts
__global__ void __launch_bounds__(512) cs_base_kernel(uint32_t *totals, uint32_t *__restrict__ counts) {
    __shared__ uint32_t sh[512 / 32 + 1];
    const uint32_t v = threadIdx.x < 257 ? totals[threadIdx.x] : 0;
    const uint32_t ex = block_excl_sum<uint32_t, 512>(v, sh, (uint32_t *)nullptr);
    if (threadIdx.x < 257) {
        totals[threadIdx.x] = ex;
        counts[threadIdx.x] = v;
    }
}

template <class Src>
__global__ void __launch_bounds__(CS_T)
    cs_scatter_kernel(Src src, uint64_t N, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ base,
                      uint64_t tiles, uint32_t *__restrict__ psi) {
    __shared__ uint16_t wcount[CS_WARPS][CS_BINS];
    __shared__ uint32_t gbase[CS_BINS];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    for (int j = threadIdx.x; j < CS_WARPS * CS_BINS / 2; j += CS_T) reinterpret_cast<uint32_t *>(&wcount[0][0])[j] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t)blockIdx.x * CS_TILE + (uint64_t)w * (32 * CS_ITEMS);
    uint32_t code[CS_ITEMS / 2]; // two 16-bit codes per register
    uint32_t rnk[CS_ITEMS / 2];  // two 16-bit ranks per register
#pragma unroll
    for (int r = 0; r < CS_ITEMS; r++) {
        const uint64_t i = wbase + (uint64_t)r * 32 + lane;
        const bool valid = i < N;
        const uint32_t c = valid ? src.at(i) : 0u;
        const unsigned peers = match_bits<9>(c, valid);
        const uint32_t pre = valid ? wcount[w][c] : 0;
        __syncwarp();
        if (valid && (peers & lt) == 0) wcount[w][c] = (uint16_t)(pre + __popc(peers));
        __syncwarp();
        const uint32_t rk = pre + __popc(peers & lt);
        code[r >> 1] = (r & 1) ? (code[r >> 1] | (c << 16)) : c;
        rnk[r >> 1] = (r & 1) ? (rnk[r >> 1] | (rk << 16)) : rk;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 257; c += CS_T) { // exclusive prefix over the tile's warps + global row start
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < CS_WARPS; ww++) {
            const uint32_t t = wcount[ww][c];
            wcount[ww][c] = (uint16_t)run;
            run += t;
        }
        gbase[c] = base[c] + offs[(uint64_t)c * tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CS_ITEMS; r++) {
        const uint64_t i = wbase + (uint64_t)r * 32 + lane;
        if (i < N) {
            const uint32_t c = (code[r >> 1] >> (16 * (r & 1))) & 0xffffu;
            psi[gbase[c] + wcount[w][c] + ((rnk[r >> 1] >> (16 * (r & 1))) & 0xffffu)] = (uint32_t)i;
        }
    }
}

struct CStart {
    uint32_t c[258]; // c[code] = first row of that code in F; c[257] = N
};

constexpr uint32_t NIL = 0xffffffffu;

// sub-list of splitter s: rows s*K, psi(s*K), ... up to (excluding) the next splitter row
// link[s] = (next splitter, hops to it), one 8-byte word per splitter: a jump round is one random access, not two
__global__ void inv_walk1_kernel(const uint32_t *__restrict__ psi, uint64_t S, uint32_t K, uint64_t N,
                                 uint2 *__restrict__ link) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    uint32_t cur = (uint32_t)(s * K);
    uint32_t cnt = 0;
    do {
        cur = psi[cur];
        cnt++;
    } while ((cur & (K - 1)) != 0 && cnt <= N); // psi is a permutation: the walk returns to a splitter within N steps
    uint32_t succ = cur / K;
    link[s] = make_uint2(succ == 0 ? NIL : succ, cnt); // the cycle through row 0 is cut just before row 0
}
__device__ __forceinline__ uint2 inv_jump_one(const uint2 *__restrict__ a, uint64_t s) {
    uint2 v = a[s];
    if (v.x != NIL) {
        const uint2 u = a[v.x];
        v.y += u.y;
        v.x = u.x;
    }
    return v;
}
__global__ void inv_jump_kernel(const uint2 *__restrict__ a, uint64_t S, uint2 *__restrict__ b) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) b[s] = inv_jump_one(a, s);
}
// All pointer-jumping rounds in one cooperative launch (a grid barrier between rounds instead of ~20 launches of
// ~10 us each).  The result of `rounds` rounds is in (rounds odd ? b : a).
constexpr int JP_PER = 8;  // elements per thread and round, their loads issued together
constexpr int JP_HOPS = 1; // dependent jumps per round (3 hops and half the rounds measured slower: the rounds are bound by L2 sector reads, not by the barrier)
__global__ void __launch_bounds__(512)
    inv_jump_all_kernel(uint2 *a, uint2 *b, uint64_t S, int rounds) {
    cg::grid_group grid = cg::this_grid();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < rounds; r++) {
        const uint2 *src = (r & 1) ? b : a;
        uint2 *dst = (r & 1) ? a : b;
        for (uint64_t s0 = t0; s0 < S; s0 += stride * JP_PER) {
            uint2 v[JP_PER], u[JP_PER];
#pragma unroll
            for (int q = 0; q < JP_PER; q++) { // written by other SMs in the round before: L2, not L1
                const uint64_t s = s0 + q * stride;
                v[q] = s < S ? __ldcg(&src[s]) : make_uint2(NIL, 0);
            }
#pragma unroll
            for (int hop = 0; hop < JP_HOPS; hop++) { // every hop reads the round's input array only: reach x (JP_HOPS + 1) per round
#pragma unroll
                for (int q = 0; q < JP_PER; q++) u[q] = v[q].x != NIL ? __ldcg(&src[v[q].x]) : make_uint2(NIL, 0);
#pragma unroll
                for (int q = 0; q < JP_PER; q++)
                    if (v[q].x != NIL) v[q] = make_uint2(u[q].x, v[q].y + u[q].y);
            }
#pragma unroll
            for (int q = 0; q < JP_PER; q++) {
                const uint64_t s = s0 + q * stride;
                if (s < S) dst[s] = v[q];
            }
        }
        grid.sync();
    }
}
// Second walk: every splitter on the cycle through row 0 walks its segment again and writes the text.  Walk lengths
// are geometric, so lanes idle while the longest walk of their warp finishes and every instruction in this loop costs
// ~4x: text[i] = F[row] (the code whose row range contains the row) is therefore a lookup in a bucket table over the
// row's high bits (built on the host with the C table) followed by a binary search between the bucket's two ends
// (usually zero or one step) instead of nine steps over the whole C table.
constexpr int IT_BUCKETS = 1024;
struct FTable {
    uint16_t bt[IT_BUCKETS + 2]; // bt[k] = largest code index whose first row is <= k << sh
    int sh;
};
__global__ void __launch_bounds__(128)
    inv_walk2_kernel(const uint32_t *__restrict__ psi, uint64_t S, uint32_t K, const uint2 *__restrict__ link,
                     CStart cs, FTable ft, uint8_t *__restrict__ text, uint64_t cap, uint64_t N,
                     uint32_t *__restrict__ err) {
    __shared__ uint32_t sc[258];
    __shared__ uint16_t bt[IT_BUCKETS + 2];
    for (int j = threadIdx.x; j < 258; j += blockDim.x) sc[j] = cs.c[j];
    for (int j = threadIdx.x; j < IT_BUCKETS + 2; j += blockDim.x) bt[j] = ft.bt[j];
    __syncthreads();
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint2 me = link[s];
    if (me.x != NIL) return; // not on the cycle through row 0
    const uint64_t total = link[0].y;
    uint64_t g = total - me.y;
    uint32_t cur = (uint32_t)(s * K);
    uint64_t guard = 0;
    bool bad = false;
    do {
        const uint32_t nxt = psi[cur]; // the next hop does not wait for this row's symbol
        if (g >= 1) {
            const uint32_t bkt = cur >> ft.sh;
            int lo = bt[bkt], hi = bt[bkt + 1] + 1; // sc[lo] <= cur < sc[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sc[mid] <= cur) lo = mid; else hi = mid;
            }
            if (lo == 0) bad = true; // fromJust Nothing (src/Data/BWT/Internal.hs:195)
            else if (g - 1 < cap) text[g - 1] = (uint8_t)(lo - 1);
        }
        cur = nxt;
        g++;
    } while ((cur & (K - 1)) != 0 && guard++ <= N);
    if (bad) atomicMax(err, 1u);
}

template <class Src>
int bwt_decode_impl(tc_ctx *ctx, Src src, uint64_t N, uint8_t *d_text, uint64_t cap, uint64_t *n_out) {
    *n_out = 0;
    if (N == 0) return TC_OK;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    WsMark mk = tc_ws_mark(ctx);
    // stable counting sort of the positions by symbol code == psi
    const uint64_t tiles = ceil_div_u64(N, CS_TILE);
    uint32_t *hist, *totals, *d_counts, *psi;
    TC_TRY(ws_alloc(ctx, 257 * tiles, &hist));
    TC_TRY(ws_alloc(ctx, 260, &totals));
    TC_TRY(ws_alloc(ctx, 260, &d_counts));
    TC_TRY(ws_alloc(ctx, N, &psi));
    TC_LAUNCH(ctx, (cs_hist_kernel<Src>), (unsigned)tiles, CS_T, 0, src, N, hist, tiles);
    TC_LAUNCH(ctx, cs_rows_kernel, 257, 1024, 0, hist, tiles, totals);
    TC_LAUNCH(ctx, cs_base_kernel, 1, 512, 0, totals, d_counts);
    uint32_t *h = (uint32_t *)ctx->h_scal;
    TC_TRY(tc_d2h_small(ctx, h, d_counts, 257 * sizeof(uint32_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h[0] == 0) { // no Nothing: empty result (src/Data/BWT/Internal.hs:174-175)
        tc_ws_release(ctx, mk);
        return TC_OK;
    }
    CStart cs;
    uint32_t acc = 0;
    for (int c = 0; c < 257; c++) {
        cs.c[c] = acc;
        acc += h[c];
    }
    cs.c[257] = acc;
    ctx->prof_bytes_next = N * (sizeof(*src.p) + 4);
    TC_LAUNCH(ctx, (cs_scatter_kernel<Src>), (unsigned)tiles, CS_T, 0, src, N, hist, totals, tiles, psi);
    // list ranking
    const uint32_t K = 16;
    const uint64_t S = ceil_div_u64(N, K);
    uint2 *linkA, *linkB;
    TC_TRY(ws_alloc(ctx, S, &linkA));
    TC_TRY(ws_alloc(ctx, S, &linkB));
    uint32_t *d_err;
    TC_TRY(ws_alloc(ctx, 1, &d_err));
    TC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream));
    const unsigned gridS = (unsigned)ceil_div_u64(S, 128);
    TC_LAUNCH(ctx, inv_walk1_kernel, gridS, 128, 0, psi, S, K, N, linkA);
    int rounds = 1;
    while ((1ull << rounds) < S) rounds++;
    if (ctx->coop_ok) {
        rounds = 1;
        for (uint64_t reach = JP_HOPS + 1; reach < S; reach *= JP_HOPS + 1) rounds++;
        int per_sm = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inv_jump_all_kernel, 512, 0));
        if (per_sm > 2) per_sm = 2;
        const unsigned cgrid = (unsigned)std::min<uint64_t>(ceil_div_u64(S, 512), (uint64_t)ctx->sm_count * std::max(per_sm, 1));
        uint64_t S_arg = S;
        void *args[] = {&linkA, &linkB, &S_arg, &rounds};
        if (ctx->prof_on) ctx->prof_begin("inv_jump_all_kernel");
        cudaError_t e = cudaLaunchCooperativeKernel((void *)inv_jump_all_kernel, dim3(cgrid), dim3(512), args, 0, ctx->stream);
        if (ctx->prof_on) ctx->prof_end();
        ctx->launches++;
        if (e != cudaSuccess) return ctx->fail(e, "inv_jump_all_kernel", __LINE__);
        if (rounds & 1) std::swap(linkA, linkB);
    } else {
        for (int r = 0; r < rounds; r++) {
            TC_LAUNCH(ctx, inv_jump_kernel, gridS, 128, 0, (const uint2 *)linkA, S, linkB);
            std::swap(linkA, linkB);
        }
    }
    FTable ft;
    ft.sh = 0;
    while (((N - 1) >> ft.sh) >= (uint64_t)IT_BUCKETS) ft.sh++;
    for (int k = 0, lo = 0; k < IT_BUCKETS + 2; k++) {
        const uint64_t x = (uint64_t)k << ft.sh;
        while (lo < 256 && (uint64_t)cs.c[lo + 1] <= x) lo++; // largest index with cs.c[index] <= x, at most 256
        ft.bt[k] = (uint16_t)lo;
    }
    ctx->prof_bytes_next = 5 * N;
    TC_LAUNCH(ctx, inv_walk2_kernel, gridS, 128, 0, psi, S, K, (const uint2 *)linkA, cs, ft, d_text, cap, N, d_err);
    TC_TRY(tc_d2h_small(ctx, h, &linkA[0].y, sizeof(uint32_t)));
    TC_TRY(tc_d2h_small(ctx, h + 1, d_err, sizeof(uint32_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t total = h[0];
    bool bad = h[1] != 0;
    tc_ws_release(ctx, mk);
    if (bad) return TC_E_FROMJUST;
    *n_out = total - 1;
    return (total - 1 > cap) ? TC_E_CAP : TC_OK;
}
} // namespace

// BWT bytes + primary from an existing suffix array (0-based starts, N entries)
int tc_bwt_emit_dev(tc_ctx *ctx, const uint8_t *d_text, const uint32_t *d_sa, uint64_t N, uint8_t *d_bwt,
                    uint64_t *primary) {
    uint64_t *d_primary;
    TC_TRY(ws_alloc(ctx, 1, &d_primary));
    TC_LAUNCH(ctx, bwt_emit_kernel, (unsigned)ceil_div_u64(N, 256), 256, 0, d_text, d_sa, N, d_bwt, d_primary,
              (uint32_t *)nullptr);
    TC_TRY(tc_d2h_small(ctx, ctx->h_scal, d_primary, sizeof(uint64_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    *primary = ctx->h_scal[0];
    return TC_OK;
}

int bwt_encode_dev_impl(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint8_t *d_bwt, uint64_t *primary,
                        uint32_t *d_sa_1based) {
    *primary = 0;
    if (n == 0) return TC_OK; // toBWT [] = BWT Empty (src/Data/BWT.hs:58)
    const uint64_t N = n + 1;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    WsMark mk = tc_ws_mark(ctx);
    uint32_t *d_sa;
    uint64_t *d_primary;
    TC_TRY(ws_alloc(ctx, N, &d_sa));
    TC_TRY(ws_alloc(ctx, 1, &d_primary));
    bool done = false;
    TC_TRY(tc_suffix_sort_bwt_dev(ctx, d_text, n, d_sa, d_bwt, primary, &done));
    if (done && d_sa_1based) {
        TC_LAUNCH(ctx, sa_plus1_kernel, (unsigned)ceil_div_u64(N, 256), 256, 0, d_sa, N, d_sa_1based);
    } else if (!done) {
        TC_LAUNCH(ctx, bwt_emit_kernel, (unsigned)ceil_div_u64(N, 256), 256, 0, d_text, d_sa, N, d_bwt, d_primary,
                  d_sa_1based);
        TC_TRY(tc_d2h_small(ctx, ctx->h_scal, d_primary, sizeof(uint64_t)));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        *primary = ctx->h_scal[0];
    }
    tc_ws_release(ctx, mk);
    return TC_OK;
}

int bwt_decode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint8_t *d_text,
                              uint64_t cap, uint64_t *n_out) {
    if (N && primary >= N) primary = ~0ull;
    return bwt_decode_impl(ctx, CodeU8{d_bwt, primary}, N, d_text, cap, n_out);
}
int bwt_decode_i16_dev_impl(tc_ctx *ctx, const int16_t *d_bwt, uint64_t N, uint8_t *d_text, uint64_t cap,
                               uint64_t *n_out) {
    return bwt_decode_impl(ctx, CodeI16{d_bwt}, N, d_text, cap, n_out);
}
