// fm.cu -- FM-index: C table (seqToCc, src/Data/FMIndex/Internal.hs:275-316), Occ
// (seqToOccCK, :195-259) and the suffix array (src/Data/FMIndex.hs:169-173), plus batched
// backward search: countFMIndex (:347-438) and locateFMIndex (:448-542) with the
// rank->position map of the wrappers (src/Data/FMIndex.hs:496,526,562,598).
//
// Device image (one contiguous allocation, so it can be replicated with one broadcast):
//   header | occurrence planes | raw BWT bytes | SA-mark plane | SA samples
// Occ is stored as one bit-plane per symbol in 32-byte rank blocks:
//   { u32 count of the symbol before the block ; 7 x u32 bits = 224 BWT positions }
// so Occ(c,k) costs exactly one 32-byte sector (the DRAM/L2 access granule) and <= 7 popcounts.
// The Nothing row is never queried (patterns are `Seq b`, matched against `Just a`).
// SA samples: rows whose suffix start is a multiple of `rate` are marked in a plane of the
// same layout; sample index = rank in that plane.  rate == 1 keeps the full SA.
#include <algorithm>

#include "common.cuh"
#include "impl.cuh"

namespace {
constexpr uint64_t FM_MAGIC = 0x30304d4642434254ull; // "TCBFM00"
constexpr uint32_t BLK = 224;

struct FmHeader {
    uint64_t magic, n, N, primary;
    uint32_t sigma, rate;
    int16_t alphabet[257];
    int64_t C[258];
    uint16_t code[256]; // byte -> alphabet index 1..sigma-1, 0 if absent
    uint64_t nblocks, n_samples, blob_bytes;
    uint64_t off_planes, off_bwt, off_mark, off_samples;
};
constexpr size_t HDR_BYTES = 8192;
static_assert(sizeof(FmHeader) <= HDR_BYTES, "header too large");

struct RankBlock {
    uint32_t cnt;
    uint32_t bits[7];
};
static_assert(sizeof(RankBlock) == 32, "rank block must be one sector");

struct FmDev { // what kernels need, passed by value
    const RankBlock *planes;
    const uint8_t *bwt;
    const RankBlock *mark;
    const uint32_t *samples;
    uint64_t nblocks, N;
    uint32_t sigma;
};
struct FmTables {
    uint32_t C[258];
    uint16_t code[256];
};

__device__ __forceinline__ uint32_t rank_in(const RankBlock *plane, uint32_t k) {
    // number of set bits at positions < k
    uint32_t b = k / BLK, r = k % BLK;
    const uint4 *p = reinterpret_cast<const uint4 *>(plane + b);
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t c = w[0];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int rem = (int)r - 32 * i;
        uint32_t m = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1));
        c += __popc(w[i + 1] & m);
    }
    return c;
}

// ---- build -----------------------------------------------------------------------
struct CodeLut {
    uint16_t code[256];
};
// one warp per rank block: ballots build the bit words of every plane
__global__ void __launch_bounds__(256)
    fm_planes_kernel(const uint8_t *__restrict__ bwt, uint64_t N, uint64_t primary, CodeLut lut, uint32_t nplanes,
                     uint64_t nblocks, RankBlock *__restrict__ planes) {
    __shared__ uint16_t s_code[256];
    s_code[threadIdx.x] = lut.code[threadIdx.x];
    __syncthreads();
    uint64_t blk = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= nblocks) return;
    const unsigned lane = lane_id();
    uint32_t c[7];
#pragma unroll
    for (int w = 0; w < 7; w++) {
        uint64_t pos = blk * BLK + w * 32 + lane;
        c[w] = (pos < N && pos != primary) ? s_code[bwt[pos]] : 0u;
    }
    for (uint32_t p = 0; p < nplanes; p++) {
        uint32_t words[7];
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < 7; w++) {
            words[w] = __ballot_sync(TC_FULL, c[w] == p + 1);
            total += __popc(words[w]);
        }
        RankBlock *rb = planes + (uint64_t)p * nblocks + blk;
        if (lane == 0) rb->cnt = total; // own count for now; fm_plane_scan_kernel turns it into a prefix
#pragma unroll
        for (int w = 0; w < 7; w++)
            if (lane == w + 1) rb->bits[w] = words[w];
    }
}

// SA-mark plane: bit set iff sa[row] % rate == 0
__global__ void __launch_bounds__(256)
    fm_mark_kernel(const uint32_t *__restrict__ sa, uint64_t N, uint32_t rate, uint64_t nblocks,
                   RankBlock *__restrict__ mark) {
    uint64_t blk = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= nblocks) return;
    const unsigned lane = lane_id();
    uint32_t total = 0;
    uint32_t mine = 0;
#pragma unroll
    for (int w = 0; w < 7; w++) {
        uint64_t pos = blk * BLK + w * 32 + lane;
        bool f = pos < N && (sa[pos] % rate == 0);
        uint32_t word = __ballot_sync(TC_FULL, f);
        total += __popc(word);
        if (lane == (unsigned)w + 1) mine = word;
    }
    RankBlock *rb = mark + blk;
    if (lane == 0) rb->cnt = total;
    else if (lane <= 7) rb->bits[lane - 1] = mine;
}

// per plane: exclusive scan of the block counts (one CTA per plane, chunked with a carry)
__global__ void __launch_bounds__(1024) fm_plane_scan_kernel(RankBlock *planes, uint64_t nblocks, uint32_t *totals) {
    __shared__ uint32_t sh[1024 / 32 + 1];
    RankBlock *pl = planes + (uint64_t)blockIdx.x * nblocks;
    uint32_t carry = 0;
    for (uint64_t b = 0; b < nblocks; b += 1024) {
        uint64_t i = b + threadIdx.x;
        uint32_t v = i < nblocks ? pl[i].cnt : 0;
        uint32_t tot;
        uint32_t ex = block_excl_sum<uint32_t, 1024>(v, sh, &tot);
        if (i < nblocks) pl[i].cnt = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && totals) totals[blockIdx.x] = carry;
}

__global__ void fm_samples_kernel(const uint32_t *__restrict__ sa, uint64_t N, uint32_t rate,
                                  const RankBlock *__restrict__ mark, uint32_t *__restrict__ samples) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    uint32_t s = sa[r];
    if (s % rate == 0) samples[rank_in(mark, (uint32_t)r)] = s;
}

// ---- backward search ------------------------------------------------------------
// SURVEY.md A4.  [s,e] are 1-based inclusive SA ranks, Occ(a,k) = rank of a in BWT[1..k].
__device__ __forceinline__ bool backward_search(const FmDev &fm, const uint32_t *sC, const uint16_t *scode,
                                                const uint8_t *pat, uint64_t m, uint32_t &s_out, uint32_t &e_out) {
    if (m == 0) return false; // countFMIndex DS.Empty _ = Nothing (:348)
    int64_t s = -1, e = -1;
    bool started = false, flag = false;
    for (uint64_t k = m; k-- > 0;) {
        if (s > e) { // :385-387
            flag = true;
            break;
        }
        uint32_t j = scode[pat[k]];
        if (j == 0) break; // symbol absent from the text: recursion stops silently (:391,:421-423)
        if (!started) {    // :389-418
            s = (int64_t)sC[j] + 1;
            e = (int64_t)sC[j + 1];
            started = true;
        } else {           // :419-438
            const RankBlock *pl = fm.planes + (uint64_t)(j - 1) * fm.nblocks;
            int64_t ns = (int64_t)sC[j] + rank_in(pl, (uint32_t)(s - 1)) + 1;
            int64_t ne = (int64_t)sC[j] + rank_in(pl, (uint32_t)e);
            s = ns;
            e = ne;
        }
    }
    if (!started || flag || (e - s + 1) == 0) return false; // :366-369
    s_out = (uint32_t)s;
    e_out = (uint32_t)e;
    return true;
}

__global__ void __launch_bounds__(128)
    fm_count_kernel(FmDev fm, FmTables tb, const uint8_t *__restrict__ pats, const uint64_t *__restrict__ off,
                    uint64_t q, int64_t *__restrict__ count, uint32_t *__restrict__ s_arr,
                    uint32_t *__restrict__ n_arr) {
    __shared__ uint32_t sC[258];
    __shared__ uint16_t scode[256];
    for (int j = threadIdx.x; j < 258; j += blockDim.x) sC[j] = tb.C[j];
    for (int j = threadIdx.x; j < 256; j += blockDim.x) scode[j] = tb.code[j];
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint64_t o = off[i], m = off[i + 1] - o;
    uint32_t s = 0, e = 0;
    bool hit = backward_search(fm, sC, scode, pats + o, m, s, e);
    if (count) count[i] = hit ? (int64_t)(e - s + 1) : -1;
    if (s_arr) {
        s_arr[i] = s;
        n_arr[i] = hit ? e - s + 1 : 0;
    }
}

// one thread per hit: map the SA rank to a text position through the sampled SA
__global__ void __launch_bounds__(128)
    fm_locate_kernel(FmDev fm, FmTables tb, const uint32_t *__restrict__ s_arr, const uint64_t *__restrict__ hit_off,
                     uint64_t q, uint64_t total, uint64_t cap, uint64_t *__restrict__ pos_out) {
    __shared__ uint32_t sC[258];
    __shared__ uint16_t scode[256];
    for (int j = threadIdx.x; j < 258; j += blockDim.x) sC[j] = tb.C[j];
    for (int j = threadIdx.x; j < 256; j += blockDim.x) scode[j] = tb.code[j];
    __syncthreads();
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= total || h >= cap) return;
    // query owning hit h: last i with hit_off[i] <= h
    uint64_t lo = 0, hi = q;
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (hit_off[mid] <= h) lo = mid; else hi = mid;
    }
    uint32_t rank1 = s_arr[lo] + (uint32_t)(h - hit_off[lo]); // 1-based SA rank
    uint32_t r = rank1 - 1;
    uint32_t steps = 0;
    for (;; ) {
        if (steps > fm.N) { // cannot happen with a well-formed index; never spin on the device
            pos_out[h] = 0;
            return;
        }
        uint32_t b = r / BLK, rr = r % BLK;
        const RankBlock *mb = fm.mark + b;
        bool marked = (mb->bits[rr >> 5] >> (rr & 31)) & 1;
        if (marked) {
            uint32_t si = rank_in(fm.mark, r);
            pos_out[h] = (uint64_t)fm.samples[si] + steps + 1; // suffixstartpos is 1-based
            return;
        }
        uint32_t j = scode[fm.bwt[r]];
        const RankBlock *pl = fm.planes + (uint64_t)(j - 1) * fm.nblocks;
        r = sC[j] + rank_in(pl, r); // LF step
        steps++;
    }
}

__global__ void fm_export_kernel(FmDev fm, uint64_t primary, int16_t *__restrict__ bwt, uint32_t *__restrict__ sa1) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= fm.N) return;
    bwt[r] = r == primary ? (int16_t)-1 : (int16_t)fm.bwt[r];
    if (sa1) sa1[r] = fm.samples[r] + 1;
}
} // namespace

struct tc_fm {
    int device;
    void *blob;
    bool owns;
    FmHeader hdr;
    FmDev dev() const {
        FmDev d;
        const char *b = (const char *)blob;
        d.planes = (const RankBlock *)(b + hdr.off_planes);
        d.bwt = (const uint8_t *)(b + hdr.off_bwt);
        d.mark = (const RankBlock *)(b + hdr.off_mark);
        d.samples = (const uint32_t *)(b + hdr.off_samples);
        d.nblocks = hdr.nblocks;
        d.N = hdr.N;
        d.sigma = hdr.sigma;
        return d;
    }
    FmTables tables() const {
        FmTables t;
        for (int j = 0; j < 258; j++) t.C[j] = (uint32_t)hdr.C[j];
        memcpy(t.code, hdr.code, sizeof t.code);
        return t;
    }
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int fm_fill_image(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, const CodeLut &lut,
                         uint32_t nplanes, uint64_t nblocks, RankBlock *d_planes, const uint32_t *d_sa, uint32_t rate,
                         RankBlock *d_mark, uint32_t *d_samples) {
    const unsigned gridB = (unsigned)ceil_div_u64(nblocks, 8);
    TC_LAUNCH(ctx, fm_planes_kernel, gridB, 256, 0, d_bwt, N, primary, lut, nplanes, nblocks, d_planes);
    if (nplanes) TC_LAUNCH(ctx, fm_plane_scan_kernel, nplanes, 1024, 0, d_planes, nblocks, (uint32_t *)nullptr);
    TC_LAUNCH(ctx, fm_mark_kernel, gridB, 256, 0, d_sa, N, rate, nblocks, d_mark);
    TC_LAUNCH(ctx, fm_plane_scan_kernel, 1, 1024, 0, d_mark, nblocks, (uint32_t *)nullptr);
    TC_LAUNCH(ctx, fm_samples_kernel, (unsigned)ceil_div_u64(N, 256), 256, 0, d_sa, N, rate, d_mark, d_samples);
    return TC_OK;
}

static int fm_build_dev_impl(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t rate, tc_fm **out) {
    *out = nullptr;
    if (n == 0) return TC_E_ARG; // undefined in the reference; callers guard empty input
    if (rate == 0) rate = 32;
    const uint64_t N = n + 1;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    uint32_t hist[256];
    TC_TRY(tc_byte_hist_dev(ctx, d_text, n, hist));
    FmHeader h;
    memset(&h, 0, sizeof h);
    h.magic = FM_MAGIC;
    h.n = n;
    h.N = N;
    h.rate = rate;
    h.alphabet[0] = -1;
    h.C[0] = 0;
    uint32_t sigma = 1;
    int64_t acc = 1; // the single Nothing sorts first
    for (int c = 0; c < 256; c++) {
        if (hist[c]) {
            h.code[c] = (uint16_t)sigma;
            h.alphabet[sigma] = (int16_t)c;
            h.C[sigma] = acc;
            acc += hist[c];
            sigma++;
        }
    }
    h.C[sigma] = (int64_t)N;
    for (uint32_t j = sigma + 1; j < 258; j++) h.C[j] = (int64_t)N;
    h.sigma = sigma;
    const uint32_t nplanes = sigma - 1;
    h.nblocks = N / BLK + 1;
    h.n_samples = n / rate + 1; // text positions 0, rate, 2*rate, ... <= n
    size_t off = HDR_BYTES;
    h.off_planes = off;
    off = align_up(off + (size_t)nplanes * h.nblocks * sizeof(RankBlock), 512);
    h.off_bwt = off;
    off = align_up(off + N, 512);
    h.off_mark = off;
    off = align_up(off + h.nblocks * sizeof(RankBlock), 512);
    h.off_samples = off;
    off = align_up(off + h.n_samples * sizeof(uint32_t), 512);
    h.blob_bytes = off;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    void *blob = nullptr;
    if (cudaMalloc(&blob, h.blob_bytes) != cudaSuccess) {
        cudaGetLastError();
        snprintf(ctx->err, sizeof ctx->err, "FM-index image: cudaMalloc(%llu) failed",
                 (unsigned long long)h.blob_bytes);
        return TC_E_NOMEM;
    }
    tc_fm *fm = new tc_fm();
    fm->device = ctx->device;
    fm->blob = blob;
    fm->owns = true;
    char *b = (char *)blob;
    uint8_t *d_bwt = (uint8_t *)(b + h.off_bwt);
    RankBlock *d_planes = (RankBlock *)(b + h.off_planes);
    RankBlock *d_mark = (RankBlock *)(b + h.off_mark);
    uint32_t *d_samples = (uint32_t *)(b + h.off_samples);
    int rc = TC_OK;
    do {
        uint32_t *d_sa;
        uint64_t *d_primary;
        if ((rc = ws_alloc(ctx, N, &d_sa)) != TC_OK) break;
        if ((rc = ws_alloc(ctx, 1, &d_primary)) != TC_OK) break;
        if ((rc = tc_suffix_sort_dev(ctx, d_text, n, d_sa)) != TC_OK) break;
        // BWT bytes straight into the image
        uint64_t primary = 0;
        if ((rc = tc_bwt_emit_dev(ctx, d_text, d_sa, N, d_bwt, &primary)) != TC_OK) break;
        h.primary = primary;
        CodeLut lut;
        memcpy(lut.code, h.code, sizeof lut.code);
        if ((rc = fm_fill_image(ctx, d_bwt, N, primary, lut, nplanes, h.nblocks, d_planes, d_sa, rate, d_mark,
                                d_samples)) != TC_OK)
            break;
        fm->hdr = h;
        cudaError_t e = cudaMemcpyAsync(blob, &fm->hdr, sizeof(FmHeader), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            rc = ctx->fail(e, "fm build", __LINE__);
            break;
        }
    } while (0);
    if (rc != TC_OK) {
        cudaFree(blob);
        delete fm;
        return rc;
    }
    *out = fm;
    return TC_OK;
}

extern "C" int tc_fm_build_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t sa_sample_rate, tc_fm **out) {
    if (!ctx || !out) return TC_E_ARG;
    TC_TRY(tc_ws_reset(ctx));
    return fm_build_dev_impl(ctx, d_text, n, sa_sample_rate, out);
}

extern "C" int tc_fm_build(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint32_t sa_sample_rate, tc_fm **out) {
    if (!ctx || !out) return TC_E_ARG;
    TC_TRY(tc_ws_reset(ctx));
    *out = nullptr;
    if (n == 0) return TC_E_ARG;
    uint8_t *d_text;
    TC_TRY(ws_alloc(ctx, n, &d_text));
    TC_CUDA(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, ctx->stream));
    return fm_build_dev_impl(ctx, d_text, n, sa_sample_rate, out);
}

extern "C" void tc_fm_free(tc_fm *fm) {
    if (!fm) return;
    if (fm->owns && fm->blob) {
        cudaSetDevice(fm->device);
        cudaFree(fm->blob);
    }
    delete fm;
}

extern "C" int tc_fm_get_info(const tc_fm *fm, tc_fm_info *info) {
    if (!fm || !info) return TC_E_ARG;
    memset(info, 0, sizeof *info);
    info->n = fm->hdr.n;
    info->N = fm->hdr.N;
    info->primary = fm->hdr.primary;
    info->sigma = fm->hdr.sigma;
    info->sa_sample_rate = fm->hdr.rate;
    memcpy(info->alphabet, fm->hdr.alphabet, sizeof info->alphabet);
    for (int j = 0; j < 257; j++) info->C[j] = fm->hdr.C[j];
    info->blob_bytes = fm->hdr.blob_bytes;
    info->n_samples = fm->hdr.n_samples;
    return TC_OK;
}

extern "C" const void *tc_fm_blob(const tc_fm *fm) { return fm ? fm->blob : nullptr; }

extern "C" int tc_fm_from_blob_dev(tc_ctx *ctx, void *d_blob, uint64_t bytes, int take_ownership, tc_fm **out) {
    if (!ctx || !out || !d_blob || bytes < HDR_BYTES) return TC_E_ARG;
    *out = nullptr;
    tc_fm *fm = new tc_fm();
    fm->device = ctx->device;
    fm->blob = d_blob;
    fm->owns = take_ownership != 0;
    cudaError_t e = cudaMemcpyAsync(&fm->hdr, d_blob, sizeof(FmHeader), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        delete fm;
        return ctx->fail(e, "tc_fm_from_blob_dev", __LINE__);
    }
    if (fm->hdr.magic != FM_MAGIC || fm->hdr.blob_bytes > bytes) {
        delete fm;
        return TC_E_ARG;
    }
    *out = fm;
    return TC_OK;
}

static int fm_count_dev_impl(tc_ctx *ctx, const tc_fm *fm, const uint8_t *d_pats, const uint64_t *d_off, uint64_t q,
                             int64_t *d_count, uint32_t *d_s, uint32_t *d_n) {
    if (q == 0) return TC_OK;
    TC_LAUNCH(ctx, fm_count_kernel, (unsigned)ceil_div_u64(q, 128), 128, 0, fm->dev(), fm->tables(), d_pats, d_off, q,
              d_count, d_s, d_n);
    return TC_OK;
}

extern "C" int tc_fm_count_dev(tc_ctx *ctx, const tc_fm *fm, const uint8_t *d_pats, const uint64_t *d_off, uint64_t q,
                               int64_t *d_count) {
    if (!ctx || !fm) return TC_E_ARG;
    TC_TRY(tc_ws_reset(ctx));
    return fm_count_dev_impl(ctx, fm, d_pats, d_off, q, d_count, nullptr, nullptr);
}

extern "C" int tc_fm_count(tc_ctx *ctx, const tc_fm *fm, const uint8_t *pats, const uint64_t *off, uint64_t q,
                           int64_t *count) {
    if (!ctx || !fm) return TC_E_ARG;
    TC_TRY(tc_ws_reset(ctx));
    if (q == 0) return TC_OK;
    uint64_t total = off[q];
    uint8_t *d_pats;
    uint64_t *d_off;
    int64_t *d_count;
    TC_TRY(ws_alloc(ctx, total ? total : 1, &d_pats));
    TC_TRY(ws_alloc(ctx, q + 1, &d_off));
    TC_TRY(ws_alloc(ctx, q, &d_count));
    if (total) TC_CUDA(cudaMemcpyAsync(d_pats, pats, total, cudaMemcpyHostToDevice, ctx->stream));
    TC_CUDA(cudaMemcpyAsync(d_off, off, (q + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    TC_TRY(fm_count_dev_impl(ctx, fm, d_pats, d_off, q, d_count, nullptr, nullptr));
    TC_CUDA(cudaMemcpyAsync(count, d_count, q * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return TC_OK;
}

static int fm_locate_dev_impl(tc_ctx *ctx, const tc_fm *fm, const uint8_t *d_pats, const uint64_t *d_off, uint64_t q,
                              uint64_t *d_hit_off, uint64_t *d_pos, uint64_t cap, uint64_t *total) {
    *total = 0;
    if (q == 0) {
        TC_CUDA(cudaMemsetAsync(d_hit_off, 0, sizeof(uint64_t), ctx->stream));
        return TC_OK;
    }
    uint32_t *d_s, *d_n;
    TC_TRY(ws_alloc(ctx, q, &d_s));
    TC_TRY(ws_alloc(ctx, q, &d_n));
    TC_TRY(fm_count_dev_impl(ctx, fm, d_pats, d_off, q, nullptr, d_s, d_n));
    TC_TRY(tc_scan_exclusive_u32_to_u64(ctx, d_n, d_hit_off, q, d_hit_off + q));
    TC_TRY(tc_d2h_small(ctx, ctx->h_scal, d_hit_off + q, sizeof(uint64_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t H = ctx->h_scal[0];
    *total = H;
    uint64_t lim = H < cap ? H : cap;
    if (lim)
        TC_LAUNCH(ctx, fm_locate_kernel, (unsigned)ceil_div_u64(lim, 128), 128, 0, fm->dev(), fm->tables(), d_s,
                  d_hit_off, q, H, cap, d_pos);
    return H > cap ? TC_E_CAP : TC_OK;
}

extern "C" int tc_fm_locate_dev(tc_ctx *ctx, const tc_fm *fm, const uint8_t *d_pats, const uint64_t *d_off,
                                uint64_t q, uint64_t *d_hit_off, uint64_t *d_pos_1based, uint64_t cap,
                                uint64_t *total) {
    if (!ctx || !fm || !total) return TC_E_ARG;
    TC_TRY(tc_ws_reset(ctx));
    return fm_locate_dev_impl(ctx, fm, d_pats, d_off, q, d_hit_off, d_pos_1based, cap, total);
}

extern "C" int tc_fm_locate(tc_ctx *ctx, const tc_fm *fm, const uint8_t *pats, const uint64_t *off, uint64_t q,
                            uint64_t *hit_off, uint64_t *pos_1based, uint64_t cap, uint64_t *total) {
    if (!ctx || !fm || !total) return TC_E_ARG;
    TC_TRY(tc_ws_reset(ctx));
    *total = 0;
    if (q == 0) {
        if (hit_off) hit_off[0] = 0;
        return TC_OK;
    }
    uint64_t nbytes = off[q];
    uint8_t *d_pats;
    uint64_t *d_off, *d_hit_off, *d_pos;
    TC_TRY(ws_alloc(ctx, nbytes ? nbytes : 1, &d_pats));
    TC_TRY(ws_alloc(ctx, q + 1, &d_off));
    TC_TRY(ws_alloc(ctx, q + 1, &d_hit_off));
    TC_TRY(ws_alloc(ctx, cap ? cap : 1, &d_pos));
    if (nbytes) TC_CUDA(cudaMemcpyAsync(d_pats, pats, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    TC_CUDA(cudaMemcpyAsync(d_off, off, (q + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    int rc = fm_locate_dev_impl(ctx, fm, d_pats, d_off, q, d_hit_off, d_pos, cap, total);
    if (rc != TC_OK && rc != TC_E_CAP) return rc;
    TC_CUDA(cudaMemcpyAsync(hit_off, d_hit_off, (q + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    uint64_t lim = *total < cap ? *total : cap;
    if (lim) TC_CUDA(cudaMemcpyAsync(pos_1based, d_pos, lim * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return rc;
}

extern "C" int tc_fm_export(tc_ctx *ctx, const tc_fm *fm, int16_t *bwt, uint32_t *sa_1based) {
    if (!ctx || !fm || !bwt) return TC_E_ARG;
    if (sa_1based && fm->hdr.rate != 1) return TC_E_ARG; // full SA only exists with sa_sample_rate == 1
    TC_TRY(tc_ws_reset(ctx));
    const uint64_t N = fm->hdr.N;
    int16_t *d_bwt;
    uint32_t *d_sa = nullptr;
    TC_TRY(ws_alloc(ctx, N, &d_bwt));
    if (sa_1based) TC_TRY(ws_alloc(ctx, N, &d_sa));
    TC_LAUNCH(ctx, fm_export_kernel, (unsigned)ceil_div_u64(N, 256), 256, 0, fm->dev(), fm->hdr.primary, d_bwt, d_sa);
    TC_CUDA(cudaMemcpyAsync(bwt, d_bwt, N * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (sa_1based) TC_CUDA(cudaMemcpyAsync(sa_1based, d_sa, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return TC_OK;
}
