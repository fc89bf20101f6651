// radix.cu -- LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.
// Per pass:
//   (1) rs_hist_kernel    per-tile digit histogram.  Shared-memory atomics cost 2 cycles per
//                         lane on this part, so counting is done the way the ranking is: warp
//                         match groups (8 ballots, not MATCH.ANY), the group leader bumps a
//                         warp-private counter with a plain read-modify-write.  8 B/key.
//   (2) rs_rowscan_kernel exclusive scan of each digit's row of the digit-major histogram
//                         matrix + rs_base_kernel for the 256 row totals (two tiny launches).
//   (3) rs_scatter_kernel stable in-tile ranking with warp match/ballot, keys then values
//                         staged through shared memory in digit order and written out as
//                         coalesced runs.  Values are loaded only after the ranking so the
//                         kernel fits 4 CTAs per SM (latency hiding; the first version ran at
//                         16 warps/SM and stalled on long-scoreboard 44 % of the time).
// Traffic per pass and element: 8 B (histogram) + 12 B read + 12 B written.
#include <utility>

#include "common.cuh"

namespace {
constexpr int RS_T = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_T * RS_ITEMS;
constexpr int RS_WARPS = RS_T / 32;

__global__ void __launch_bounds__(RS_T)
    rs_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ hist,
                   uint64_t tiles) {
    __shared__ uint32_t h[RS_WARPS][256];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_T) (&h[0][0])[j] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)w * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        key[r] = i < n ? keys[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        bool valid = i < n;
        uint32_t d = (uint32_t)((key[r] >> shift) & 255);
        unsigned peers = match_bits<8>(d, valid);
        if (valid && (peers & lanemask_lt()) == 0) h[w][d] += __popc(peers); // warp-private: no atomics
        __syncwarp();
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) s += h[ww][threadIdx.x];
    hist[(uint64_t)threadIdx.x * tiles + blockIdx.x] = s;
}

// row d of the matrix (tiles entries): in-place exclusive scan, row total -> totals[d]
__global__ void __launch_bounds__(1024)
    rs_rowscan_kernel(uint32_t *__restrict__ hist, uint64_t tiles, uint32_t *__restrict__ totals) {
    __shared__ uint32_t sh[1024 / 32 + 1];
    uint32_t *row = hist + (uint64_t)blockIdx.x * tiles;
    uint32_t carry = 0;
    for (uint64_t b = 0; b < tiles; b += 1024) {
        uint64_t i = b + threadIdx.x;
        uint32_t v = i < tiles ? row[i] : 0;
        uint32_t tot;
        uint32_t ex = block_excl_sum<uint32_t, 1024>(v, sh, &tot);
        if (i < tiles) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(256) rs_base_kernel(uint32_t *totals) {
    __shared__ uint32_t sh[256 / 32 + 1];
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t, 256>(totals[threadIdx.x], sh, &tot);
    totals[threadIdx.x] = ex;
}

struct RsSmem {
    uint64_t keys[RS_TILE];
    uint32_t vals[RS_TILE];
    uint16_t wcount[RS_WARPS][256];
    uint16_t dig_start[256 + 1];
    uint32_t gbase[256];
    uint32_t scan[RS_WARPS + 1];
};

__global__ void __launch_bounds__(RS_T, 4)
    rs_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout,
                      uint32_t *__restrict__ vout, uint64_t n, int shift, const uint32_t *__restrict__ offs,
                      const uint32_t *__restrict__ base, uint64_t tiles) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem &S = *reinterpret_cast<RsSmem *>(rs_raw);
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    for (int j = threadIdx.x; j < RS_WARPS * 256 / 2; j += RS_T) reinterpret_cast<uint32_t *>(&S.wcount[0][0])[j] = 0;
    __syncthreads();
    const uint64_t tbase = (uint64_t)blockIdx.x * RS_TILE;
    const uint64_t wbase = tbase + (uint64_t)w * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint16_t rnk[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        key[r] = i < n ? kin[i] : ~0ull;
    }
    // stable rank of every key among the keys of its warp with the same digit
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        bool valid = i < n;
        uint32_t d = (uint32_t)((key[r] >> shift) & 255);
        unsigned peers = match_bits<8>(d, valid);
        uint32_t pre = valid ? S.wcount[w][d] : 0;
        __syncwarp();
        if (valid && (peers & lanemask_lt()) == 0) S.wcount[w][d] = (uint16_t)(pre + __popc(peers));
        __syncwarp();
        rnk[r] = (uint16_t)(pre + __popc(peers & lanemask_lt()));
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, then exclusive scan over digits
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            uint32_t t = S.wcount[ww][d];
            S.wcount[ww][d] = (uint16_t)run;
            run += t;
        }
        uint32_t total;
        uint32_t ex = block_excl_sum<uint32_t, RS_T>(run, S.scan, &total);
        S.dig_start[d] = (uint16_t)ex;
        S.gbase[d] = offs[(uint64_t)d * tiles + blockIdx.x] + base[d] - ex; // global slot of local slot 0 of digit d
    }
    __syncthreads();
    // final local slot of every key; keys go to shared memory in digit order
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)((key[r] >> shift) & 255);
            uint32_t p = (uint32_t)S.dig_start[d] + S.wcount[w][d] + rnk[r];
            rnk[r] = (uint16_t)p;
            S.keys[p] = key[r];
        }
    }
    // values: loaded only now (keeps the register footprint of the ranking phase small)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        if (i < n) S.vals[rnk[r]] = vin[i];
    }
    __syncthreads();
    const uint32_t cnt = (uint32_t)((n - tbase) < (uint64_t)RS_TILE ? (n - tbase) : (uint64_t)RS_TILE);
    for (uint32_t p = threadIdx.x; p < cnt; p += RS_T) {
        uint64_t k = S.keys[p];
        uint32_t d = (uint32_t)((k >> shift) & 255);
        uint32_t g = S.gbase[d] + p;
        kout[g] = k;
        vout[g] = S.vals[p];
    }
}
} // namespace

int tc_radix_sort_pairs(tc_ctx *ctx, uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, uint64_t n,
                        const int *shifts, int npass, uint64_t **out_keys, uint32_t **out_vals) {
    *out_keys = k0;
    *out_vals = v0;
    if (n == 0 || npass == 0) return TC_OK;
    // per-device attribute; cheap enough to set on every call (contexts may live on different GPUs)
    TC_CUDA(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
    uint64_t tiles = ceil_div_u64(n, RS_TILE);
    WsMark mk = tc_ws_mark(ctx);
    uint32_t *hist, *totals;
    TC_TRY(ws_alloc(ctx, 256 * tiles, &hist));
    TC_TRY(ws_alloc(ctx, 256, &totals));
    uint64_t *ki = k0, *ko = k1;
    uint32_t *vi = v0, *vo = v1;
    for (int p = 0; p < npass; p++) {
        ctx->prof_bytes_next = 8 * n; // keys read
        TC_LAUNCH(ctx, rs_hist_kernel, (unsigned)tiles, RS_T, 0, ki, n, shifts[p], hist, tiles);
        TC_LAUNCH(ctx, rs_rowscan_kernel, 256, 1024, 0, hist, tiles, totals);
        TC_LAUNCH(ctx, rs_base_kernel, 1, 256, 0, totals);
        ctx->prof_bytes_next = 24 * n; // (8 B key + 4 B value) read and written
        TC_LAUNCH(ctx, rs_scatter_kernel, (unsigned)tiles, RS_T, sizeof(RsSmem), ki, vi, ko, vo, n, shifts[p], hist,
                  totals, tiles);
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    *out_keys = ki;
    *out_vals = vi;
    tc_ws_release(ctx, mk);
    return TC_OK;
}
