// radix.cu -- LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.
// Per pass: (1) per-tile digit histogram with warp-aggregated (match) shared-memory
// atomics, (2) exclusive scan of the digit-major histogram matrix, (3) scatter: stable
// in-tile ranking with warp match/ballot, keys staged in shared memory in digit order and
// written out as coalesced runs.
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
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t tbase = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; k++) {
        uint64_t i = tbase + (uint64_t)k * RS_T + threadIdx.x;
        bool valid = i < n;
        uint32_t d = valid ? (uint32_t)((keys[i] >> shift) & 255) : (0x100u + lane_id());
        unsigned peers = __match_any_sync(TC_FULL, d);
        if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&h[d], __popc(peers));
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

struct RsSmem {
    uint64_t keys[RS_TILE];
    uint32_t vals[RS_TILE];
    uint32_t wcount[RS_WARPS][256];
    uint32_t dig_start[256];
    uint32_t gbase[256];
    uint32_t scan[RS_WARPS + 1];
};

__global__ void __launch_bounds__(RS_T)
    rs_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout,
                      uint32_t *__restrict__ vout, uint64_t n, int shift, const uint32_t *__restrict__ offs,
                      uint64_t tiles) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem &S = *reinterpret_cast<RsSmem *>(rs_raw);
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_T) (&S.wcount[0][0])[j] = 0;
    __syncthreads();
    const uint64_t tbase = (uint64_t)blockIdx.x * RS_TILE;
    const uint64_t wbase = tbase + (uint64_t)w * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    uint16_t rnk[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? kin[i] : ~0ull;
        val[r] = valid ? vin[i] : 0;
    }
    // stable rank of every key among the keys of its warp with the same digit
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? (uint32_t)((key[r] >> shift) & 255) : (0x100u + lane);
        unsigned peers = __match_any_sync(TC_FULL, d);
        uint32_t pre = valid ? S.wcount[w][d] : 0;
        __syncwarp();
        if (valid && (peers & lanemask_lt()) == 0) S.wcount[w][d] = pre + __popc(peers);
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
            S.wcount[ww][d] = run;
            run += t;
        }
        uint32_t total;
        uint32_t ex = block_excl_sum<uint32_t, RS_T>(run, S.scan, &total);
        S.dig_start[d] = ex;
        S.gbase[d] = offs[(uint64_t)d * tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)((key[r] >> shift) & 255);
            uint32_t p = S.dig_start[d] + S.wcount[w][d] + rnk[r];
            S.keys[p] = key[r];
            S.vals[p] = val[r];
        }
    }
    __syncthreads();
    const uint32_t cnt = (uint32_t)((n - tbase) < (uint64_t)RS_TILE ? (n - tbase) : (uint64_t)RS_TILE);
    for (uint32_t p = threadIdx.x; p < cnt; p += RS_T) {
        uint64_t k = S.keys[p];
        uint32_t d = (uint32_t)((k >> shift) & 255);
        uint32_t g = S.gbase[d] + (p - S.dig_start[d]);
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
    uint32_t *hist;
    TC_TRY(ws_alloc(ctx, 256 * tiles, &hist));
    uint64_t *ki = k0, *ko = k1;
    uint32_t *vi = v0, *vo = v1;
    for (int p = 0; p < npass; p++) {
        ctx->prof_bytes_next = 8 * n; // keys read
        TC_LAUNCH(ctx, rs_hist_kernel, (unsigned)tiles, RS_T, 0, ki, n, shifts[p], hist, tiles);
        TC_TRY(tc_scan_exclusive_u32(ctx, hist, hist, 256 * tiles, (uint32_t *)nullptr));
        ctx->prof_bytes_next = 24 * n; // (8 B key + 4 B value) read and written
        TC_LAUNCH(ctx, rs_scatter_kernel, (unsigned)tiles, RS_T, sizeof(RsSmem), ki, vi, ko, vo, n, shifts[p], hist,
                  tiles);
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
    *out_keys = ki;
    *out_vals = vi;
    tc_ws_release(ctx, mk);
    return TC_OK;
}
