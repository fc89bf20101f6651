// common.cuh -- context, scratch arena and device helpers shared by all kernels.
// sm_100a only (B200): 148 SMs, 32-wide warps, 227 KB shared memory per CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/tc_b200.h"

#define TC_WARP 32
#define TC_FULL 0xffffffffu
#define TC_NONE32 0xffffffffu

struct tc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // copy streams + events of the batch entry point (created on first use)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    // Grow-only scratch arena: pointers handed out stay valid until the next ws_reset().
    struct Chunk {
        char *p;
        size_t cap;
    };
    std::vector<Chunk> chunks;
    size_t cur_chunk = 0; // cursor: chunk index and offset inside it
    size_t cur_off = 0;
    size_t used_total = 0;
    // pinned scalars for small device->host results
    uint64_t *h_scal = nullptr; // 1024 x u64, pinned
    uint64_t launches = 0;
    // MTF final list still in flight (composed helpers read it after the RLE stage's sync)
    struct {
        bool active = false;
        uint32_t sigma = 0;
        int16_t alpha[257];
        int16_t *final_list = nullptr;
        uint32_t h_off = 512; // where in h_scal (64-bit words) the device list lands
    } mtf_pending;
    uint32_t text_hist[256] = {0}; // byte histogram of the last text handed to the suffix sort
    // extra lanes of the batch entry points (tc_blocks_encode_packed / _dev): full contexts of their
    // own on the same device, each driven by a helper thread; created on first use
    static constexpr int MAX_LANES = 4;
    tc_ctx *child[MAX_LANES - 1] = {nullptr, nullptr, nullptr};
    int lanes = 3; // blocks in flight per call (TC_B200_LANES=1..4; measured 2: 18.8, 3: 19.6, 4: 19.8 GB/s on C2)
    bool no_msd = false; // TC_B200_NO_MSD=1: force the LSD suffix-sort path (tests exercise both)
    bool no_rawkey = false; // TC_B200_NO_RAWKEY=1: arithmetic-code uniform keys also for equiprobable bytes (tests compare both)
    bool mtf_v2 = false; // TC_B200_MTF_V2=1: warp-per-chunk MTF replay (the round-1 kernel) instead of thread-per-chunk
    bool mtfd_v1 = false; // TC_B200_MTFD_V1=1: list-shifting MTF decode kernels (cost grows with the index) instead of the select-based ones
    uint32_t mtf_L = 0;  // TC_B200_MTF_L: chunk length of the thread-per-chunk MTF replay (0 = one chunk per resident thread)
    void *mtf_auto[9] = {nullptr}; // per alphabet size: device tables of the MTF automata (mtf.cu: AutoTables)
    bool coop_ok = false; // the device supports cooperative launches (grid barriers inside a kernel)
    uint32_t diag = 0;   // TC_B200_DIAG (ctx.cu): timing experiments only, results are then incomplete
    uint32_t attr_done = 0; // kernels whose dynamic shared-memory limit has been raised on this context's device
    char err[512] = {0};
    // optional per-kernel timing (tc_ctx_profile): one event pair per launch
    struct ProfRec {
        const char *name;
        cudaEvent_t a, b;
        uint64_t bytes; // algorithmic bytes of this launch (0 if the caller did not say)
    };
    bool prof_on = false;
    uint64_t prof_bytes_next = 0;
    std::vector<ProfRec> prof;
    void prof_begin(const char *name) {
        ProfRec r{name, nullptr, nullptr, prof_bytes_next};
        prof_bytes_next = 0;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream);
        prof.push_back(r);
    }
    void prof_end() { cudaEventRecord(prof.back().b, stream); }

    int fail(cudaError_t e, const char *what, int line) {
        snprintf(err, sizeof err, "%s at line %d: %s", what, line, cudaGetErrorString(e));
        return TC_E_CUDA;
    }
};

#define TC_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return ctx->fail(e_, #call, __LINE__); \
    } while (0)
#define TC_TRY(call)               \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != TC_OK) return rc_; \
    } while (0)
// kernel launch + bookkeeping (TC_LAUNCH_AS: the profile name when `kernel` is a pointer)
#define TC_LAUNCH_AS(ctx, name, kernel, grid, block, smem, ...)               \
    do {                                                                      \
        if ((ctx)->prof_on) (ctx)->prof_begin(name);                          \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);      \
        if ((ctx)->prof_on) (ctx)->prof_end();                                \
        (ctx)->launches++;                                                    \
        cudaError_t e_ = cudaPeekAtLastError();                               \
        if (e_ != cudaSuccess) return (ctx)->fail(e_, name, __LINE__);        \
    } while (0)
#define TC_LAUNCH(ctx, kernel, grid, block, smem, ...) \
    TC_LAUNCH_AS(ctx, #kernel, kernel, grid, block, smem, __VA_ARGS__)

// ---- arena -------------------------------------------------------------------
// ws_reset(): called at the start of a top-level op; moves the cursor back to the first chunk
// (chunks are kept, see ctx.cu), so steady state does no cudaMalloc.
int tc_ws_reset(tc_ctx *ctx);
int tc_ws_alloc(tc_ctx *ctx, size_t bytes, void **out);
template <typename T>
static inline int ws_alloc(tc_ctx *ctx, size_t count, T **out) {
    void *p = nullptr;
    int rc = tc_ws_alloc(ctx, count * sizeof(T), &p);
    *out = (T *)p;
    return rc;
}
// nested ops use mark/release so a composite can reuse scratch between stages
struct WsMark {
    size_t nchunks, off, used; // nchunks = cursor chunk index at mark time
};
WsMark tc_ws_mark(tc_ctx *ctx);
void tc_ws_release(tc_ctx *ctx, WsMark m);

static inline uint64_t ceil_div_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// Small device -> host result (a few bytes into the context's pinned, device-mapped scalars):
// written by a one-block kernel instead of a memcpy, so it never queues behind a bulk D2H copy
// that another stream has on the copy engine (tc_blocks_encode overlaps those with compute).
int tc_d2h_small(tc_ctx *ctx, void *h_pinned_dst, const void *d_src, size_t bytes);

// ---- internal device-level entry points (all pointers are device pointers) ----
// radix sort of (u64 key, u32 val) pairs by the bit ranges listed (LSD, 8 bits per pass).
// On return *out_keys/*out_vals point at whichever buffer holds the result.
int tc_radix_sort_pairs(tc_ctx *ctx, uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, uint64_t n,
                        const int *shifts, int npass, uint64_t **out_keys, uint32_t **out_vals);
// exclusive prefix sums (out may alias in)
int tc_scan_exclusive_u32_to_u64(tc_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n, uint64_t *d_total);
int tc_scan_exclusive_u32(tc_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *d_total);
int tc_scan_exclusive_u64(tc_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t n, uint64_t *d_total);
int tc_scan_inclusive_max_u32(tc_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n);
// suffix array of text$ (0-based start positions, N = n+1 entries) in d_sa
int tc_suffix_sort_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_sa);
// same, and when d_bwt is given the fast path also emits the BWT bytes + primary (*bwt_done)
int tc_suffix_sort_bwt_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_sa, uint8_t *d_bwt,
                           uint64_t *primary, bool *bwt_done);

// ---- device helpers --------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// streaming 128-bit loads/stores that bypass L1 allocation (data touched once)
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// Lanes holding the same BITS-bit label, built from one ballot per label bit.  The MATCH.ANY
// instruction (__match_any_sync) measured ~10x slower than this on sm_100a (a radix histogram
// pass went from 113 us to memory-bound when it was replaced), so it is not used anywhere.
template <int BITS>
__device__ __forceinline__ unsigned match_bits(uint32_t label, bool valid) {
    unsigned peers = __ballot_sync(TC_FULL, valid);
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        bool bit = (label >> b) & 1;
        unsigned m = __ballot_sync(TC_FULL, bit);
        peers &= m ^ (bit ? 0u : 0xffffffffu); // one select + one 3-input logic op per bit
    }
    return peers;
}

template <typename T>
__device__ __forceinline__ T warp_incl_sum(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(TC_FULL, v, d);
        if (lane_id() >= (unsigned)d) v += o;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_incl_max(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(TC_FULL, v, d);
        if (lane_id() >= (unsigned)d) v = o > v ? o : v;
    }
    return v;
}
// Block-wide exclusive sum.  `sh` needs THREADS/32 + 1 elements.  All threads must call.
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_sum(T v, T *sh, T *total) {
    const int w = threadIdx.x >> 5;
    T inc = warp_incl_sum(v);
    if (lane_id() == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = (lane_id() < THREADS / 32) ? sh[lane_id()] : T(0);
        T xi = warp_incl_sum(x);
        if (lane_id() < THREADS / 32) sh[lane_id()] = xi - x;
        if (lane_id() == 31) sh[THREADS / 32] = xi;
    }
    __syncthreads();
    T r = sh[w] + inc - v;
    if (total) *total = sh[THREADS / 32];
    __syncthreads();
    return r;
}
// Block-wide exclusive max with identity `ident` (ident must be <= every value).
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_max(T v, T ident, T *sh, T *total) {
    const int w = threadIdx.x >> 5;
    T inc = warp_incl_max(v);
    T prev = __shfl_up_sync(TC_FULL, inc, 1);
    if (lane_id() == 0) prev = ident;
    if (lane_id() == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = (lane_id() < THREADS / 32) ? sh[lane_id()] : ident;
        T xi = warp_incl_max(x);
        T xe = __shfl_up_sync(TC_FULL, xi, 1);
        if (lane_id() == 0) xe = ident;
        if (lane_id() < THREADS / 32) sh[lane_id()] = xe;
        if (lane_id() == 31) sh[THREADS / 32] = xi;
    }
    __syncthreads();
    T base = sh[w];
    T r = base > prev ? base : prev;
    if (total) *total = sh[THREADS / 32];
    __syncthreads();
    return r;
}
// Run statistics of an MTF index stream (mtf.cu RunStat -> rle.cu rle_emit_tiled_kernel): tstat[T] = (run boundaries
// inside MTF tile T, 1 + position of the last one, first index, last index).  One CTA of THREADS threads: the
// boundary at the start of every tile is decided here (first index against the last index of the tile before),
// then exclusive sum of the runs and exclusive max of the heads over the tiles.  Called by the last CTA of the
// replay kernel (loads bypass L1: other SMs wrote the records), or as a kernel of its own.
template <int THREADS>
__device__ __noinline__ void runstat_scan(const uint4 *__restrict__ tstat, uint64_t ntiles, uint32_t tile_syms,
                                             uint64_t *__restrict__ toff, uint32_t *__restrict__ theadx) {
    constexpr int PER = 12; // records per thread and round: 16 MiB blocks (1,366 or 4,097 tiles) take one or two rounds
    __shared__ uint64_t rs_sh64[THREADS / 32 + 1];
    __shared__ uint32_t rs_sh32[THREADS / 32 + 1];
    uint64_t carry = 0;
    uint32_t ch = 0;
    for (uint64_t b = 0; b < ntiles; b += (uint64_t)THREADS * PER) {
        const uint64_t t0 = b + (uint64_t)threadIdx.x * PER;
        uint4 me[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) me[q] = t0 + q < ntiles ? __ldcg(&tstat[t0 + q]) : make_uint4(0, 0, 0, 0);
        uint32_t before = t0 > 0 && t0 - 1 < ntiles ? __ldcg(&tstat[t0 - 1]).w : 0u;
        uint64_t v = 0;
        uint32_t hm = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const uint64_t t = t0 + q;
            uint32_t p = 0, h = 0;
            if (t < ntiles) {
                const bool starts = t == 0 || me[q].z != before; // a run starts at the tile's first position
                p = me[q].x + (t > 0 && starts);
                h = max(me[q].y, starts ? (uint32_t)(t * tile_syms) + 1u : 0u);
                before = me[q].w;
            }
            me[q].x = p, me[q].y = h;
            v += p;
            hm = max(hm, h);
        }
        uint64_t tot;
        uint32_t th;
        uint64_t ex = carry + block_excl_sum<uint64_t, THREADS>(v, rs_sh64, &tot);
        uint32_t hx = max(ch, block_excl_max<uint32_t, THREADS>(hm, 0u, rs_sh32, &th));
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (t0 + q < ntiles) {
                toff[t0 + q] = ex;
                theadx[t0 + q] = hx;
            }
            ex += me[q].x;
            hx = max(hx, me[q].y);
        }
        carry += tot;
        ch = max(ch, th);
    }
}
// last-CTA election for kernels that finish with a small serial step: every thread of every CTA calls it after its
// global writes; true in exactly one CTA, and only after all the others' writes are visible
__device__ __forceinline__ bool last_cta_done(uint32_t *ticket, uint32_t nctas) {
    __shared__ uint32_t lc_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) lc_last = atomicAdd(ticket, 1u) == nctas - 1;
    __syncthreads();
    if (lc_last) __threadfence();
    return lc_last != 0;
}
#endif
