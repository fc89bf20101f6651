// scan.cu -- device-wide prefix scans (reduce-then-scan, recursive).
// Used for radix-sort bucket offsets, RLE run offsets, compaction, group ids of the
// suffix sort (max-scan) and locate hit offsets.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kItems = 4;
constexpr int kTile = kThreads * kItems;

struct OpSum {
    template <typename T>
    __device__ __forceinline__ static T apply(T a, T b) { return a + b; }
};
struct OpMax {
    template <typename T>
    __device__ __forceinline__ static T apply(T a, T b) { return a > b ? a : b; }
};

template <typename T, class Op>
__device__ __forceinline__ T warp_incl(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(TC_FULL, v, d);
        if (lane_id() >= (unsigned)d) v = Op::apply(o, v);
    }
    return v;
}
// exclusive block scan with identity 0 (valid for sum, and for max over unsigned values)
template <typename T, class Op>
__device__ __forceinline__ T block_excl(T v, T *sh, T *total) {
    const int w = threadIdx.x >> 5;
    T inc = warp_incl<T, Op>(v);
    T prev = __shfl_up_sync(TC_FULL, inc, 1);
    if (lane_id() == 0) prev = T(0);
    if (lane_id() == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = (lane_id() < kThreads / 32) ? sh[lane_id()] : T(0);
        T xi = warp_incl<T, Op>(x);
        T xe = __shfl_up_sync(TC_FULL, xi, 1);
        if (lane_id() == 0) xe = T(0);
        if (lane_id() < kThreads / 32) sh[lane_id()] = xe;
        if (lane_id() == 31) sh[kThreads / 32] = xi;
    }
    __syncthreads();
    T r = Op::apply(sh[w], prev);
    *total = sh[kThreads / 32];
    __syncthreads();
    return r;
}

template <typename TIn, typename TOut, class Op>
__global__ void __launch_bounds__(kThreads) scan_reduce_kernel(const TIn *__restrict__ in, TOut *__restrict__ tile_sums,
                                                               uint64_t n) {
    __shared__ TOut sh[kThreads / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * kItems;
    TOut s = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++)
        if (base + k < n) s = Op::apply(s, (TOut)in[base + k]);
    TOut total;
    block_excl<TOut, Op>(s, sh, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename TIn, typename TOut, class Op, bool INCL>
__global__ void __launch_bounds__(kThreads)
    scan_apply_kernel(const TIn *in, TOut *out, const TOut *__restrict__ tile_off, uint64_t n, TOut *d_total) {
    __shared__ TOut sh[kThreads / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * kItems;
    TOut v[kItems];
    TOut s = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        v[k] = (base + k < n) ? (TOut)in[base + k] : TOut(0);
        s = Op::apply(s, v[k]);
    }
    TOut total;
    TOut ex = block_excl<TOut, Op>(s, sh, &total);
    TOut off = tile_off ? tile_off[blockIdx.x] : TOut(0);
    ex = Op::apply(off, ex);
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        TOut inc = Op::apply(ex, v[k]);
        if (base + k < n) out[base + k] = INCL ? inc : ex;
        ex = inc;
    }
    if (d_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *d_total = Op::apply(off, total);
}

template <typename TIn, typename TOut, class Op, bool INCL>
int scan_impl(tc_ctx *ctx, const TIn *in, TOut *out, uint64_t n, TOut *d_total) {
    if (n == 0) {
        if (d_total) TC_CUDA(cudaMemsetAsync(d_total, 0, sizeof(TOut), ctx->stream));
        return TC_OK;
    }
    uint64_t tiles = ceil_div_u64(n, kTile);
    if (tiles == 1) {
        TC_LAUNCH(ctx, (scan_apply_kernel<TIn, TOut, Op, INCL>), 1, kThreads, 0, in, out, (const TOut *)nullptr, n,
                  d_total);
        return TC_OK;
    }
    TOut *sums = nullptr;
    TC_TRY(ws_alloc(ctx, tiles, &sums));
    TC_LAUNCH(ctx, (scan_reduce_kernel<TIn, TOut, Op>), (unsigned)tiles, kThreads, 0, in, sums, n);
    TC_TRY((scan_impl<TOut, TOut, Op, false>(ctx, sums, sums, tiles, (TOut *)nullptr)));
    TC_LAUNCH(ctx, (scan_apply_kernel<TIn, TOut, Op, INCL>), (unsigned)tiles, kThreads, 0, in, out,
              (const TOut *)sums, n, d_total);
    return TC_OK;
}
} // namespace

int tc_scan_exclusive_u32_to_u64(tc_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n, uint64_t *d_total) {
    return scan_impl<uint32_t, uint64_t, OpSum, false>(ctx, in, out, n, d_total);
}
int tc_scan_exclusive_u32(tc_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *d_total) {
    return scan_impl<uint32_t, uint32_t, OpSum, false>(ctx, in, out, n, d_total);
}
int tc_scan_exclusive_u64(tc_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t n, uint64_t *d_total) {
    return scan_impl<uint64_t, uint64_t, OpSum, false>(ctx, in, out, n, d_total);
}
// inclusive running maximum over unsigned values (identity 0)
int tc_scan_inclusive_max_u32(tc_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n) {
    return scan_impl<uint32_t, uint32_t, OpMax, true>(ctx, in, out, n, (uint32_t *)nullptr);
}
