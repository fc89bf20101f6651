// api.cu -- the extern "C" entry points of include/tc_b200.h for BWT / MTF / RLE and the
// composed helpers.  Host entry points stage through the scratch arena (H2D, kernels, D2H on
// the context's stream); `_dev` entry points work on device pointers.  Every entry point
// resets the arena exactly once; the *_impl functions never do.
#include <atomic>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "impl.cuh"

#define TC_ENTER(ctx)                  \
    do {                               \
        if (!(ctx)) return TC_E_ARG;   \
        TC_TRY(tc_ws_reset(ctx));      \
    } while (0)

namespace {
// MTF recency keys are 32-bit distances (mtf.cu): N = n + 1 must stay below 2^31 - 1 when MTF takes part
inline bool too_big(uint64_t n, bool with_mtf) { return n + 1 >= (with_mtf ? 0x7fffffffull : 0xfffffffeull); }

template <typename T>
int h2d(tc_ctx *ctx, T *dst, const T *src, size_t count) {
    if (count) TC_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return TC_OK;
}
template <typename T>
int d2h(tc_ctx *ctx, T *dst, const T *src, size_t count) {
    if (count) TC_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return TC_OK;
}
int sync(tc_ctx *ctx) {
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return TC_OK;
}
} // namespace

// ---- Data.BWT ------------------------------------------------------------------
extern "C" int tc_bwt_encode_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint8_t *d_bwt, uint64_t *primary,
                                 uint32_t *d_sa_1based) {
    TC_ENTER(ctx);
    if (!primary) return TC_E_ARG;
    return bwt_encode_dev_impl(ctx, d_text, n, d_bwt, primary, d_sa_1based);
}

extern "C" int tc_bwt_encode(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint8_t *bwt, uint64_t *primary,
                             uint32_t *sa_1based) {
    TC_ENTER(ctx);
    if (!primary) return TC_E_ARG;
    *primary = 0;
    if (n == 0) return TC_OK;
    if (n + 1 >= 0xfffffffeull) return TC_E_TOOBIG;
    uint8_t *d_text, *d_bwt;
    uint32_t *d_sa1 = nullptr;
    TC_TRY(ws_alloc(ctx, n, &d_text));
    TC_TRY(ws_alloc(ctx, n + 1, &d_bwt));
    if (sa_1based) TC_TRY(ws_alloc(ctx, n + 1, &d_sa1));
    TC_TRY(h2d(ctx, d_text, text, n));
    TC_TRY(bwt_encode_dev_impl(ctx, d_text, n, d_bwt, primary, d_sa1));
    TC_TRY(d2h(ctx, bwt, d_bwt, n + 1));
    if (sa_1based) TC_TRY(d2h(ctx, sa_1based, d_sa1, n + 1));
    return sync(ctx);
}

extern "C" int tc_bwt_decode(tc_ctx *ctx, const int16_t *bwt, uint64_t N, uint8_t *text, uint64_t cap,
                             uint64_t *n_out) {
    TC_ENTER(ctx);
    if (!n_out) return TC_E_ARG;
    *n_out = 0;
    if (N == 0) return TC_OK;
    int16_t *d_bwt;
    uint8_t *d_text;
    TC_TRY(ws_alloc(ctx, N, &d_bwt));
    TC_TRY(ws_alloc(ctx, N, &d_text));
    TC_TRY(h2d(ctx, d_bwt, bwt, N));
    int rc = bwt_decode_i16_dev_impl(ctx, d_bwt, N, d_text, N, n_out);
    if (rc != TC_OK) return rc;
    uint64_t m = *n_out < cap ? *n_out : cap;
    TC_TRY(d2h(ctx, text, d_text, m));
    TC_TRY(sync(ctx));
    return *n_out > cap ? TC_E_CAP : TC_OK;
}

extern "C" int tc_bwt_decode_u8(tc_ctx *ctx, const uint8_t *bwt, uint64_t N, uint64_t primary, uint8_t *text,
                                uint64_t cap, uint64_t *n_out) {
    TC_ENTER(ctx);
    if (!n_out) return TC_E_ARG;
    *n_out = 0;
    if (N == 0) return TC_OK;
    uint8_t *d_bwt, *d_text;
    TC_TRY(ws_alloc(ctx, N, &d_bwt));
    TC_TRY(ws_alloc(ctx, N, &d_text));
    TC_TRY(h2d(ctx, d_bwt, bwt, N));
    int rc = bwt_decode_u8_dev_impl(ctx, d_bwt, N, primary, d_text, N, n_out);
    if (rc != TC_OK) return rc;
    uint64_t m = *n_out < cap ? *n_out : cap;
    TC_TRY(d2h(ctx, text, d_text, m));
    TC_TRY(sync(ctx));
    return *n_out > cap ? TC_E_CAP : TC_OK;
}

// ---- Data.MTF ------------------------------------------------------------------
extern "C" int tc_mtf_encode_u8_dev(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint16_t *d_idx,
                                    int16_t *final_list, uint32_t *sigma) {
    TC_ENTER(ctx);
    if (!sigma || !final_list) return TC_E_ARG;
    return mtf_encode_u8_dev_impl(ctx, d_bwt, N, primary, d_idx, final_list, sigma);
}

extern "C" int tc_mtf_encode(tc_ctx *ctx, const int16_t *sym, uint64_t N, uint16_t *idx, int16_t *final_list,
                             uint32_t *sigma) {
    TC_ENTER(ctx);
    if (!sigma || !final_list) return TC_E_ARG;
    *sigma = 0;
    if (N == 0) return TC_OK;
    int16_t *d_sym;
    uint16_t *d_idx;
    TC_TRY(ws_alloc(ctx, N, &d_sym));
    TC_TRY(ws_alloc(ctx, N, &d_idx));
    TC_TRY(h2d(ctx, d_sym, sym, N));
    TC_TRY(mtf_encode_i16_dev_impl(ctx, d_sym, N, d_idx, final_list, sigma));
    TC_TRY(d2h(ctx, idx, d_idx, N));
    return sync(ctx);
}

extern "C" int tc_mtf_encode_u8(tc_ctx *ctx, const uint8_t *bwt, uint64_t N, uint64_t primary, uint16_t *idx,
                                int16_t *final_list, uint32_t *sigma) {
    TC_ENTER(ctx);
    if (!sigma || !final_list) return TC_E_ARG;
    *sigma = 0;
    if (N == 0) return TC_OK;
    uint8_t *d_bwt;
    uint16_t *d_idx;
    TC_TRY(ws_alloc(ctx, N, &d_bwt));
    TC_TRY(ws_alloc(ctx, N, &d_idx));
    TC_TRY(h2d(ctx, d_bwt, bwt, N));
    TC_TRY(mtf_encode_u8_dev_impl(ctx, d_bwt, N, primary, d_idx, final_list, sigma));
    TC_TRY(d2h(ctx, idx, d_idx, N));
    return sync(ctx);
}

extern "C" int tc_mtf_decode(tc_ctx *ctx, const uint16_t *idx, uint64_t N, const int16_t *final_list, uint32_t sigma,
                             int16_t *sym) {
    TC_ENTER(ctx);
    if (N == 0 || sigma == 0) return TC_OK;
    if (!final_list) return TC_E_ARG;
    uint16_t *d_idx;
    int16_t *d_sym;
    TC_TRY(ws_alloc(ctx, N, &d_idx));
    TC_TRY(ws_alloc(ctx, N, &d_sym));
    TC_TRY(h2d(ctx, d_idx, idx, N));
    TC_TRY(mtf_decode_dev_impl(ctx, d_idx, N, final_list, sigma, d_sym));
    TC_TRY(d2h(ctx, sym, d_sym, N));
    return sync(ctx);
}

// ---- Data.RLE ------------------------------------------------------------------
extern "C" int tc_rle_encode_u8_dev(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary,
                                    uint32_t *d_count, int16_t *d_rsym, uint64_t cap, uint64_t *R) {
    TC_ENTER(ctx);
    if (!R) return TC_E_ARG;
    return rle_encode_u8_dev_impl(ctx, d_bwt, N, primary, d_count, d_rsym, cap, R);
}
extern "C" int tc_rle_encode_u16_dev(tc_ctx *ctx, const uint16_t *d_idx, uint64_t N, uint32_t *d_count,
                                     int16_t *d_rsym, uint64_t cap, uint64_t *R) {
    TC_ENTER(ctx);
    if (!R) return TC_E_ARG;
    return rle_encode_u16_dev_impl(ctx, d_idx, N, d_count, d_rsym, cap, R);
}

namespace {
// shared tail of the host RLE encoders: runs are produced into arena buffers sized for the
// worst case, then the first min(R, cap) are copied out.
template <class F>
int rle_encode_host(tc_ctx *ctx, uint64_t N, uint64_t worst, uint32_t *count, int16_t *rsym, uint64_t cap, uint64_t *R,
                    F run) {
    *R = 0;
    if (N == 0) return TC_OK;
    uint32_t *d_count;
    int16_t *d_rsym;
    TC_TRY(ws_alloc(ctx, worst, &d_count));
    TC_TRY(ws_alloc(ctx, worst, &d_rsym));
    int rc = run(d_count, d_rsym, worst);
    if (rc != TC_OK) return rc;
    uint64_t m = *R < cap ? *R : cap;
    TC_TRY(d2h(ctx, count, d_count, m));
    TC_TRY(d2h(ctx, rsym, d_rsym, m));
    TC_TRY(sync(ctx));
    return *R > cap ? TC_E_CAP : TC_OK;
}
} // namespace

extern "C" int tc_rle_encode(tc_ctx *ctx, const int16_t *sym, uint64_t N, uint32_t *count, int16_t *rsym,
                             uint64_t cap, uint64_t *R) {
    TC_ENTER(ctx);
    if (!R) return TC_E_ARG;
    int16_t *d_sym = nullptr;
    if (N) {
        TC_TRY(ws_alloc(ctx, N, &d_sym));
        TC_TRY(h2d(ctx, d_sym, sym, N));
    }
    return rle_encode_host(ctx, N, 2 * N + 1, count, rsym, cap, R, [&](uint32_t *dc, int16_t *ds, uint64_t w) {
        return rle_encode_i16_dev_impl(ctx, d_sym, N, dc, ds, w, R);
    });
}

extern "C" int tc_rle_encode_u8(tc_ctx *ctx, const uint8_t *bwt, uint64_t N, uint64_t primary, uint32_t *count,
                                int16_t *rsym, uint64_t cap, uint64_t *R) {
    TC_ENTER(ctx);
    if (!R) return TC_E_ARG;
    uint8_t *d_bwt = nullptr;
    if (N) {
        TC_TRY(ws_alloc(ctx, N, &d_bwt));
        TC_TRY(h2d(ctx, d_bwt, bwt, N));
    }
    return rle_encode_host(ctx, N, N + 2, count, rsym, cap, R, [&](uint32_t *dc, int16_t *ds, uint64_t w) {
        return rle_encode_u8_dev_impl(ctx, d_bwt, N, primary, dc, ds, w, R);
    });
}

extern "C" int tc_rle_encode_u16(tc_ctx *ctx, const uint16_t *idx, uint64_t N, uint32_t *count, int16_t *rsym,
                                 uint64_t cap, uint64_t *R) {
    TC_ENTER(ctx);
    if (!R) return TC_E_ARG;
    uint16_t *d_idx = nullptr;
    if (N) {
        TC_TRY(ws_alloc(ctx, N, &d_idx));
        TC_TRY(h2d(ctx, d_idx, idx, N));
    }
    return rle_encode_host(ctx, N, N + 1, count, rsym, cap, R, [&](uint32_t *dc, int16_t *ds, uint64_t w) {
        return rle_encode_u16_dev_impl(ctx, d_idx, N, dc, ds, w, R);
    });
}

extern "C" int tc_rle_decode(tc_ctx *ctx, const uint32_t *count, const int16_t *rsym, uint64_t R, int16_t *sym,
                             uint64_t cap, uint64_t *N) {
    TC_ENTER(ctx);
    if (!N) return TC_E_ARG;
    *N = 0;
    if (R == 0) return TC_OK;
    uint32_t *d_count;
    int16_t *d_rsym, *d_sym;
    TC_TRY(ws_alloc(ctx, R, &d_count));
    TC_TRY(ws_alloc(ctx, R, &d_rsym));
    TC_TRY(ws_alloc(ctx, cap ? cap : 1, &d_sym));
    TC_TRY(h2d(ctx, d_count, count, R));
    TC_TRY(h2d(ctx, d_rsym, rsym, R));
    int rc = rle_decode_dev_impl(ctx, d_count, d_rsym, R, d_sym, cap, N);
    if (rc != TC_OK && rc != TC_E_CAP) return rc;
    uint64_t m = *N < cap ? *N : cap;
    TC_TRY(d2h(ctx, sym, d_sym, m));
    TC_TRY(sync(ctx));
    return rc;
}

// ---- composed helpers --------------------------------------------------------------------
namespace {
void info_clear(tc_block_info *info, uint64_t n) {
    memset(info, 0, sizeof *info);
    info->n = n;
}

// text (device) -> BWT -> [MTF ->] RLE, everything stays in HBM.
int compress_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, bool with_mtf, uint32_t *d_count, int16_t *d_rsym,
                 uint64_t cap, tc_block_info *info, RlePack *pk = nullptr) {
    info_clear(info, n);
    if (n == 0) return TC_OK;
    const uint64_t N = n + 1;
    uint8_t *d_bwt;
    TC_TRY(ws_alloc(ctx, N, &d_bwt));
    TC_TRY(bwt_encode_dev_impl(ctx, d_text, n, d_bwt, &info->primary, nullptr));
    info->N = N;
    if (!with_mtf) return rle_encode_u8_dev_impl(ctx, d_bwt, N, info->primary, d_count, d_rsym, cap, &info->R, pk);
    uint16_t *d_idx;
    TC_TRY(ws_alloc(ctx, N, &d_idx));
    // alphabet of the BWT = the text's bytes (histogram kept by the suffix sort) + the sentinel
    uint8_t present[257];
    present[0] = 1;
    for (int c = 0; c < 256; c++) present[c + 1] = ctx->text_hist[c] != 0;
    MtfRleLink link; // the MTF replay counts the runs of its tiles while the indices are in registers
    TC_TRY(ws_alloc(ctx, ceil_div_u64(N, 4096) + 1, &link.d_tstat));
    TC_TRY(ws_alloc(ctx, ceil_div_u64(N, 4096) + 1, &link.d_toff));
    TC_TRY(ws_alloc(ctx, ceil_div_u64(N, 4096) + 1, &link.d_theadx));
    TC_TRY(ws_alloc(ctx, 2, &link.d_ticket));
    TC_TRY(ws_alloc(ctx, MtfRleLink::SMALL_WORDS, &link.d_R)); // run count, exception count and the final list: one copy back
    link.d_final = reinterpret_cast<uint16_t *>(link.d_R + MtfRleLink::H_FINAL);
    TC_TRY(mtf_encode_u8_dev_impl(ctx, d_bwt, N, info->primary, d_idx, info->final_list, &info->sigma, present, &link));
    int rc = rle_encode_u16_dev_impl(ctx, d_idx, N, d_count, d_rsym, cap, &info->R, pk, &link); // syncs the stream
    int rc2 = mtf_finish_pending(ctx);
    return rc != TC_OK ? rc : rc2;
}

int compress_host(tc_ctx *ctx, const uint8_t *text, uint64_t n, bool with_mtf, uint32_t *count, int16_t *rsym,
                  uint64_t cap, tc_block_info *info) {
    if (!info) return TC_E_ARG;
    info_clear(info, n);
    if (n == 0) return TC_OK;
    if (too_big(n, with_mtf)) return TC_E_TOOBIG;
    uint8_t *d_text;
    uint32_t *d_count;
    int16_t *d_rsym;
    const uint64_t worst = n + 3;
    TC_TRY(ws_alloc(ctx, n, &d_text));
    TC_TRY(ws_alloc(ctx, worst, &d_count));
    TC_TRY(ws_alloc(ctx, worst, &d_rsym));
    TC_TRY(h2d(ctx, d_text, text, n));
    int rc = compress_dev(ctx, d_text, n, with_mtf, d_count, d_rsym, worst, info);
    if (rc != TC_OK) return rc;
    uint64_t m = info->R < cap ? info->R : cap;
    TC_TRY(d2h(ctx, count, d_count, m));
    TC_TRY(d2h(ctx, rsym, d_rsym, m));
    TC_TRY(sync(ctx));
    return info->R > cap ? TC_E_CAP : TC_OK;
}
} // namespace

extern "C" int tc_bwt_rle_encode(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint32_t *count, int16_t *rsym,
                                 uint64_t cap, tc_block_info *info) {
    TC_ENTER(ctx);
    return compress_host(ctx, text, n, false, count, rsym, cap, info);
}
extern "C" int tc_bwt_mtf_rle_encode(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint32_t *count, int16_t *rsym,
                                     uint64_t cap, tc_block_info *info) {
    TC_ENTER(ctx);
    return compress_host(ctx, text, n, true, count, rsym, cap, info);
}
extern "C" int tc_bwt_mtf_rle_encode_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_count,
                                         int16_t *d_rsym, uint64_t cap, tc_block_info *info) {
    TC_ENTER(ctx);
    if (!info) return TC_E_ARG;
    return compress_dev(ctx, d_text, n, true, d_count, d_rsym, cap, info);
}

namespace {
// Every exit of a batch entry point -- errors included -- waits for the copies it has queued: the caller
// frees or reuses text[] / out[] as soon as the call returns, and the next entry point hands out the arena
// memory those copies touch.
struct DrainGuard {
    tc_ctx *c;
    ~DrainGuard() {
        if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
        if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
        cudaStreamSynchronize(c->stream);
    }
};
// copy streams + events of the batch entry points (created on first use)
int blocks_streams(tc_ctx *ctx) {
    if (ctx->s_h2d) return TC_OK;
    TC_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    TC_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        TC_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        TC_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
    }
    return TC_OK;
}
} // namespace

// Pipelined batch: slot = b & 1.  H2D of block b+1 is issued before block b is compressed, the
// D2H of block b right after it (compress_dev returns with the stream drained, so the host
// knows R); compressing block b+2 into the same slot first waits for that D2H.
extern "C" int tc_blocks_encode(tc_ctx *ctx, uint64_t nblocks, const uint8_t *const *text, const uint64_t *n,
                                int with_mtf, uint32_t *const *count, int16_t *const *rsym, const uint64_t *cap,
                                tc_block_info *info) {
    TC_ENTER(ctx);
    if (nblocks == 0) return TC_OK;
    if (!text || !n || !count || !rsym || !cap || !info) return TC_E_ARG;
    uint64_t nmax = 0;
    for (uint64_t b = 0; b < nblocks; b++) {
        if (too_big(n[b], with_mtf != 0)) return TC_E_TOOBIG;
        nmax = n[b] > nmax ? n[b] : nmax;
    }
    TC_TRY(blocks_streams(ctx));
    DrainGuard drain{ctx};
    const uint64_t worst = nmax + 3;
    uint8_t *d_text[2];
    uint32_t *d_count[2];
    int16_t *d_rsym[2];
    for (int s = 0; s < 2; s++) {
        TC_TRY(ws_alloc(ctx, nmax ? nmax : 1, &d_text[s]));
        TC_TRY(ws_alloc(ctx, worst, &d_count[s]));
        TC_TRY(ws_alloc(ctx, worst, &d_rsym[s]));
    }
    // everything queued earlier on the context's stream must be done before the copy streams touch the arena
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    bool d2h_pending[2] = {false, false};
    int rc_all = TC_OK;
    auto issue_h2d = [&](uint64_t b) -> int {
        const int s = (int)(b & 1);
        if (n[b]) TC_CUDA(cudaMemcpyAsync(d_text[s], text[b], n[b], cudaMemcpyHostToDevice, ctx->s_h2d));
        TC_CUDA(cudaEventRecord(ctx->ev_h2d[s], ctx->s_h2d));
        return TC_OK;
    };
    TC_TRY(issue_h2d(0));
    for (uint64_t b = 0; b < nblocks; b++) {
        const int s = (int)(b & 1);
        if (b + 1 < nblocks) TC_TRY(issue_h2d(b + 1)); // its slot's text was consumed by block b-1 (host-synced)
        TC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[s], 0));
        if (d2h_pending[s]) TC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[s], 0));
        WsMark mk = tc_ws_mark(ctx);
        int rc = compress_dev(ctx, d_text[s], n[b], with_mtf != 0, d_count[s], d_rsym[s], worst, &info[b]);
        tc_ws_release(ctx, mk);
        if (rc != TC_OK) return rc;
        TC_CUDA(cudaStreamSynchronize(ctx->stream)); // a no-op when compress_dev already drained it
        uint64_t m = info[b].R < cap[b] ? info[b].R : cap[b];
        if (info[b].R > cap[b]) rc_all = TC_E_CAP;
        if (m) {
            TC_CUDA(cudaMemcpyAsync(count[b], d_count[s], m * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
            TC_CUDA(cudaMemcpyAsync(rsym[b], d_rsym[s], m * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        TC_CUDA(cudaEventRecord(ctx->ev_d2h[s], ctx->s_d2h));
        d2h_pending[s] = true;
    }
    TC_CUDA(cudaStreamSynchronize(ctx->s_d2h));
    return rc_all;
}

namespace {
// runs (device) -> [MTF indices ->] BWT -> text (host).  N == 0: sized by a first pass over the runs.
int decode_runs_dev(tc_ctx *ctx, const uint32_t *d_count, const int16_t *d_rsym, uint64_t R, bool with_mtf,
                    const tc_block_info *info, uint8_t *text, uint64_t cap, uint64_t *n_out) {
    uint64_t N = with_mtf ? info->N : 0; // a block header knows the BWT length: no sizing pass over the runs
    int rc = TC_OK;
    if (N == 0) { // TC_E_CAP with cap 0 returns the length
        rc = rle_decode_dev_impl(ctx, d_count, d_rsym, R, nullptr, 0, &N);
        if (rc != TC_OK && rc != TC_E_CAP) return rc;
    }
    if (N == 0) return TC_OK;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    int16_t *d_a, *d_sym;
    uint8_t *d_text;
    TC_TRY(ws_alloc(ctx, N, &d_a));
    TC_TRY(ws_alloc(ctx, N, &d_text));
    {
        uint64_t Ngot = 0;
        rc = rle_decode_dev_impl(ctx, d_count, d_rsym, R, d_a, N, &Ngot);
        if (rc != TC_OK && rc != TC_E_CAP) return rc;
        if (Ngot != N) return TC_E_ARG; // the runs do not add up to the length the header states
    }
    d_sym = d_a;
    if (with_mtf) { // index stream: int16 == uint16 here
        TC_TRY(ws_alloc(ctx, N, &d_sym));
        TC_TRY(mtf_decode_dev_impl(ctx, (const uint16_t *)d_a, N, info->final_list, info->sigma, d_sym));
    }
    rc = bwt_decode_i16_dev_impl(ctx, d_sym, N, d_text, N, n_out);
    if (rc != TC_OK) return rc;
    uint64_t m = *n_out < cap ? *n_out : cap;
    TC_TRY(d2h(ctx, text, d_text, m));
    TC_TRY(sync(ctx));
    return *n_out > cap ? TC_E_CAP : TC_OK;
}
} // namespace

extern "C" int tc_bwt_rle_decode(tc_ctx *ctx, const uint32_t *count, const int16_t *rsym, uint64_t R, uint8_t *text,
                                 uint64_t cap, uint64_t *n_out) {
    TC_ENTER(ctx);
    if (!n_out) return TC_E_ARG;
    *n_out = 0;
    if (R == 0) return TC_OK;
    uint32_t *d_count;
    int16_t *d_rsym;
    TC_TRY(ws_alloc(ctx, R, &d_count));
    TC_TRY(ws_alloc(ctx, R, &d_rsym));
    TC_TRY(h2d(ctx, d_count, count, R));
    TC_TRY(h2d(ctx, d_rsym, rsym, R));
    return decode_runs_dev(ctx, d_count, d_rsym, R, false, nullptr, text, cap, n_out);
}

extern "C" int tc_bwt_mtf_rle_decode(tc_ctx *ctx, const uint32_t *count, const int16_t *rsym,
                                     const tc_block_info *info, uint8_t *text, uint64_t cap, uint64_t *n_out) {
    TC_ENTER(ctx);
    if (!n_out || !info) return TC_E_ARG;
    *n_out = 0;
    const uint64_t R = info->R;
    if (R == 0) return TC_OK;
    uint32_t *d_count;
    int16_t *d_rsym;
    TC_TRY(ws_alloc(ctx, R, &d_count));
    TC_TRY(ws_alloc(ctx, R, &d_rsym));
    TC_TRY(h2d(ctx, d_count, count, R));
    TC_TRY(h2d(ctx, d_rsym, rsym, R));
    return decode_runs_dev(ctx, d_count, d_rsym, R, true, info, text, cap, n_out);
}

// ---- packed block container (include/tc_b200.h; payload layout in rle.cu) -----------------------
static_assert(sizeof(tc_packed_header) == 640, "tc_packed_header layout");
namespace {
inline uint64_t al16(uint64_t x) { return (x + 15) & ~15ull; }
inline uint64_t big_cap_for(uint64_t n) { return (n + 1) / 16 + 4; }
// section offsets for R runs and n_big exceptions; returns the container size
uint64_t packed_layout(uint64_t R, uint64_t n_big, tc_packed_header *h) {
    uint64_t o = sizeof(tc_packed_header);
    h->off_cnt4 = o, o = al16(o + (R + 1) / 2);
    h->off_sym8 = o, o = al16(o + R);
    h->off_hi = o, o = al16(o + (R + 31) / 32 * 4);
    h->off_big_idx = o, o = al16(o + n_big * 8);
    h->off_big_cnt = o, o = al16(o + n_big * 4);
    return h->total_bytes = o;
}
int packed_check(const void *blob, uint64_t bytes, tc_packed_header *h) {
    if (!blob || bytes < sizeof *h) return TC_E_ARG;
    memcpy(h, blob, sizeof *h);
    if (h->magic != TC_PACKED_MAGIC || h->version != 2 || h->total_bytes > bytes || h->sigma > 257) return TC_E_ARG;
    if (h->R > 0xffffffffull || h->n_big > h->R) return TC_E_ARG;
    tc_packed_header want;
    if (packed_layout(h->R, h->n_big, &want) != h->total_bytes || want.off_cnt4 != h->off_cnt4 ||
        want.off_sym8 != h->off_sym8 || want.off_hi != h->off_hi || want.off_big_idx != h->off_big_idx ||
        want.off_big_cnt != h->off_big_cnt)
        return TC_E_ARG;
    return TC_OK;
}
void packed_to_info(const tc_packed_header &h, tc_block_info *info) {
    memset(info, 0, sizeof *info);
    info->n = h.n, info->N = h.N, info->primary = h.primary, info->sigma = h.sigma, info->R = h.R;
    memcpy(info->final_list, h.final_list, sizeof info->final_list);
}
} // namespace

extern "C" uint64_t tc_packed_bound(uint64_t n) {
    // R runs of total length n + 1 (+ the reference's extra pairs): every exception (count >= 16) costs 12 bytes but
    // removes 15 runs of 1.625 bytes each, so the largest container is the one without exceptions
    tc_packed_header h;
    return packed_layout(n + 3, 4, &h);
}

extern "C" int tc_packed_info(const void *blob, uint64_t bytes, tc_block_info *info, uint32_t *flags) {
    tc_packed_header h;
    TC_TRY(packed_check(blob, bytes, &h));
    if (info) packed_to_info(h, info);
    if (flags) *flags = h.flags;
    return TC_OK;
}

extern "C" int tc_packed_unpack(const void *blob, uint64_t bytes, uint32_t *count, int16_t *rsym, uint64_t cap,
                                tc_block_info *info) {
    tc_packed_header h;
    TC_TRY(packed_check(blob, bytes, &h));
    if (info) packed_to_info(h, info);
    if (h.R > cap) return TC_E_CAP;
    if (h.R && (!count || !rsym)) return TC_E_ARG;
    const uint8_t *base = (const uint8_t *)blob;
    const uint8_t *c4 = base + h.off_cnt4, *s8 = base + h.off_sym8;
    for (uint64_t k = 0; k < h.R; k++) {
        uint32_t w;
        memcpy(&w, base + h.off_hi + (k >> 5) * 4, 4);
        const uint32_t code = s8[k] | (((w >> (k & 31)) & 1u) << 8);
        count[k] = ((c4[k >> 1] >> (4 * (k & 1))) & 15u) + 1u;
        rsym[k] = code == 0x1ffu ? (int16_t)-1 : (int16_t)code;
    }
    uint64_t prev = 0;
    for (uint64_t j = 0; j < h.n_big; j++) {
        uint64_t idx;
        uint32_t c;
        memcpy(&idx, base + h.off_big_idx + j * 8, 8);
        memcpy(&c, base + h.off_big_cnt + j * 4, 4);
        if (idx >= h.R || (j && idx <= prev) || c < 16 || ((c4[idx >> 1] >> (4 * (idx & 1))) & 15u) != 15u) return TC_E_ARG;
        count[idx] = c;
        prev = idx;
    }
    return TC_OK;
}

namespace {
// One lane of the packed batch through one context (its stream, its two copy streams, double buffers:
// slot = k & 1 for the lane's k-th block).  Lanes claim blocks from a shared counter, one block
// ahead of the one they compress (its H2D is in flight meanwhile), so ragged block sizes and
// block counts that are not a multiple of the lane count still keep every lane busy to the end.
// H2D copies of all lanes run one after the other, in issue order: each waits for the previous one's
// event.  Concurrent copies would share the link, and at the start of a batch every lane's first
// block would arrive late (six copies in flight: 1.8 ms instead of 0.3 ms for the first).
struct H2dChain {
    std::mutex m;
    cudaEvent_t last = nullptr;
};

int blocks_packed_lane(tc_ctx *ctx, std::atomic<uint64_t> &next, H2dChain &chain, uint64_t nblocks, uint64_t nmax,
                       const uint8_t *const *text, const uint64_t *n, int with_mtf, uint8_t *const *out,
                       const uint64_t *cap, uint64_t *out_bytes, tc_block_info *info) {
    TC_CUDA(cudaSetDevice(ctx->device)); // helper threads start on device 0
    TC_TRY(tc_ws_reset(ctx));
    TC_TRY(blocks_streams(ctx));
    DrainGuard drain{ctx};
    // an error in any lane stops the others at their next block: claiming past the end ends their loops
    struct Abort {
        std::atomic<uint64_t> &next;
        uint64_t nblocks;
        bool armed = true;
        ~Abort() {
            if (armed) next.store(nblocks);
        }
    } abort_others{next, nblocks};
    const uint64_t worst = nmax + 3;
    uint8_t *d_text[2];
    RlePack pk[2];
    uint32_t *d_count = nullptr; // the 6-byte records are never written: the RLE kernel emits the packed form
    int16_t *d_rsym = nullptr;
    for (int s = 0; s < 2; s++) {
        TC_TRY(ws_alloc(ctx, nmax ? nmax : 1, &d_text[s]));
        TC_TRY(ws_alloc(ctx, al16(worst / 2 + 16), &pk[s].cnt4));
        TC_TRY(ws_alloc(ctx, al16(worst), &pk[s].sym8));
        TC_TRY(ws_alloc(ctx, (worst + 31) / 32, &pk[s].hi));
        pk[s].big_cap = big_cap_for(nmax);
        TC_TRY(ws_alloc(ctx, pk[s].big_cap, &pk[s].big_idx));
        TC_TRY(ws_alloc(ctx, pk[s].big_cap, &pk[s].big_cnt));
    }
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    bool d2h_pending[2] = {false, false}, slot_filled[2] = {false, false};
    int rc_all = TC_OK;
    auto issue_h2d = [&](uint64_t b, int s) -> int {
        std::lock_guard<std::mutex> g(chain.m);
        if (chain.last) TC_CUDA(cudaStreamWaitEvent(ctx->s_h2d, chain.last, 0));
        if (n[b] && !((ctx->diag & 1) && slot_filled[s]))
            TC_CUDA(cudaMemcpyAsync(d_text[s], text[b], n[b], cudaMemcpyHostToDevice, ctx->s_h2d));
        slot_filled[s] = true;
        TC_CUDA(cudaEventRecord(ctx->ev_h2d[s], ctx->s_h2d));
        chain.last = ctx->ev_h2d[s];
        return TC_OK;
    };
    uint64_t b = next.fetch_add(1), b_next;
    if (b < nblocks) {
        // the lane's second block is claimed only once its first has arrived, so that the other
        // lanes' first blocks are next on the link
        TC_TRY(issue_h2d(b, 0));
        TC_CUDA(cudaEventSynchronize(ctx->ev_h2d[0]));
    }
    for (uint64_t k = 0; b < nblocks; b = b_next, k++) {
        const int s = (int)(k & 1);
        b_next = next.fetch_add(1);
        if (b_next < nblocks) TC_TRY(issue_h2d(b_next, s ^ 1)); // that slot's text was consumed by the block before (host-synced)
        TC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[s], 0));
        if (d2h_pending[s]) TC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[s], 0));
        WsMark mk = tc_ws_mark(ctx);
        int rc = compress_dev(ctx, d_text[s], n[b], with_mtf != 0, d_count, d_rsym, worst, &info[b], &pk[s]);
        tc_ws_release(ctx, mk);
        if (rc != TC_OK) return rc;
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        tc_packed_header h;
        memset(&h, 0, sizeof h);
        h.magic = TC_PACKED_MAGIC, h.version = 2, h.flags = with_mtf ? TC_PACKED_MTF : 0;
        h.n = info[b].n, h.N = info[b].N, h.primary = info[b].primary, h.R = info[b].R, h.sigma = info[b].sigma;
        h.n_big = n[b] ? pk[s].n_big : 0;
        memcpy(h.final_list, info[b].final_list, sizeof h.final_list);
        out_bytes[b] = packed_layout(h.R, h.n_big, &h);
        if (out_bytes[b] > cap[b]) {
            rc_all = TC_E_CAP;
            continue;
        }
        uint8_t *o = out[b];
        memcpy(o, &h, sizeof h);
        // every section is copied at its exact length and its tail up to the next section is
        // zeroed on the host, so every byte of the container is defined
        auto section = [&](uint64_t off, const void *d_src, uint64_t len, uint64_t next) -> int {
            if (len && !(ctx->diag & 2)) TC_CUDA(cudaMemcpyAsync(o + off, d_src, len, cudaMemcpyDeviceToHost, ctx->s_d2h));
            memset(o + off + len, 0, next - off - len);
            return TC_OK;
        };
        TC_TRY(section(h.off_cnt4, pk[s].cnt4, (h.R + 1) / 2, h.off_sym8));
        TC_TRY(section(h.off_sym8, pk[s].sym8, h.R, h.off_hi));
        TC_TRY(section(h.off_hi, pk[s].hi, (h.R + 31) / 32 * 4, h.off_big_idx));
        TC_TRY(section(h.off_big_idx, pk[s].big_idx, h.n_big * 8, h.off_big_cnt));
        TC_TRY(section(h.off_big_cnt, pk[s].big_cnt, h.n_big * 4, h.total_bytes));
        TC_CUDA(cudaEventRecord(ctx->ev_d2h[s], ctx->s_d2h));
        d2h_pending[s] = true;
    }
    TC_CUDA(cudaStreamSynchronize(ctx->s_d2h));
    abort_others.armed = false;
    return rc_all;
}
} // namespace

namespace {
// Runs fn(lane context) on `lanes` lanes: lane 0 on the caller's context and thread, the others on
// child contexts driven by helper threads.  TC_E_CAP from any lane is
// reported after all lanes have finished; any other error wins.
template <class F>
int run_lanes(tc_ctx *ctx, uint64_t nblocks, F fn) {
    int lanes = ctx->prof_on ? 1 : ctx->lanes;
    if ((uint64_t)lanes > nblocks) lanes = (int)nblocks;
    for (int l = 1; l < lanes; l++)
        if (!ctx->child[l - 1]) TC_TRY(tc_ctx_create(ctx->device, &ctx->child[l - 1]));
    if (lanes <= 1) return fn(ctx);
    // work queued earlier on the caller's stream (e.g. the producer of device-resident texts) must be
    // visible to the other lanes' streams
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    int rc[tc_ctx::MAX_LANES] = {TC_OK, TC_OK, TC_OK, TC_OK};
    std::thread th[tc_ctx::MAX_LANES - 1];
    for (int l = 1; l < lanes; l++)
        th[l - 1] = std::thread([&, l] { rc[l] = fn(ctx->child[l - 1]); });
    rc[0] = fn(ctx);
    for (int l = 1; l < lanes; l++) th[l - 1].join();
    int out = TC_OK;
    for (int l = 0; l < lanes; l++) {
        if (rc[l] == TC_OK) continue;
        if (rc[l] != TC_E_CAP) {
            if (l) memcpy(ctx->err, ctx->child[l - 1]->err, sizeof ctx->err);
            return rc[l];
        }
        out = TC_E_CAP;
    }
    return out;
}
} // namespace

// Three lanes by default (TC_B200_LANES = 1..4), each taking the next unclaimed block; lane 0 is the caller's
// context and thread, the others are child contexts driven by helper threads.  Each block's kernel
// chain has short serial phases (single-CTA scans, two host syncs); with several blocks in flight
// the other lanes' kernels fill them.
extern "C" int tc_blocks_encode_packed(tc_ctx *ctx, uint64_t nblocks, const uint8_t *const *text, const uint64_t *n,
                                       int with_mtf, uint8_t *const *out, const uint64_t *cap, uint64_t *out_bytes,
                                       tc_block_info *info) {
    if (!ctx) return TC_E_ARG;
    if (nblocks == 0) return TC_OK;
    if (!text || !n || !out || !cap || !out_bytes || !info) return TC_E_ARG;
    uint64_t nmax = 0;
    for (uint64_t b = 0; b < nblocks; b++) {
        if (too_big(n[b], with_mtf != 0)) return TC_E_TOOBIG;
        nmax = n[b] > nmax ? n[b] : nmax;
        out_bytes[b] = 0;
    }
    std::atomic<uint64_t> next{0};
    H2dChain chain;
    int rc = run_lanes(ctx, nblocks, [&](tc_ctx *c) {
        return blocks_packed_lane(c, next, chain, nblocks, nmax, text, n, with_mtf, out, cap, out_bytes, info);
    });
    return rc;
}

// Device-resident batch (bench `value`, HBM-resident callers): texts and run records stay in HBM, no
// copies; the same lanes as tc_blocks_encode_packed, so the serial phases of one block's
// kernel chain are filled by the other block's kernels.
namespace {
int blocks_dev_lane(tc_ctx *ctx, std::atomic<uint64_t> &next, uint64_t nblocks, const uint8_t *const *d_text,
                    const uint64_t *n, int with_mtf, uint32_t *const *d_count, int16_t *const *d_rsym,
                    const uint64_t *cap, tc_block_info *info) {
    int rc_all = TC_OK;
    TC_CUDA(cudaSetDevice(ctx->device)); // helper threads start on device 0; a lane may find no block left to claim
    for (uint64_t b = next.fetch_add(1); b < nblocks; b = next.fetch_add(1)) {
        TC_TRY(tc_ws_reset(ctx));
        int rc = compress_dev(ctx, d_text[b], n[b], with_mtf != 0, d_count[b], d_rsym[b], cap[b], &info[b]);
        if (rc == TC_E_CAP) rc_all = rc;
        else if (rc != TC_OK) return rc;
    }
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return rc_all;
}
} // namespace

extern "C" int tc_blocks_encode_dev(tc_ctx *ctx, uint64_t nblocks, const uint8_t *const *d_text, const uint64_t *n,
                                    int with_mtf, uint32_t *const *d_count, int16_t *const *d_rsym, const uint64_t *cap,
                                    tc_block_info *info) {
    if (!ctx) return TC_E_ARG;
    if (nblocks == 0) return TC_OK;
    if (!d_text || !n || !d_count || !d_rsym || !cap || !info) return TC_E_ARG;
    for (uint64_t b = 0; b < nblocks; b++)
        if (too_big(n[b], with_mtf != 0)) return TC_E_TOOBIG;
    std::atomic<uint64_t> next{0};
    return run_lanes(ctx, nblocks, [&](tc_ctx *c) {
        return blocks_dev_lane(c, next, nblocks, d_text, n, with_mtf, d_count, d_rsym, cap, info);
    });
}

namespace {
// container (host) -> text (host) on one context; the scratch it takes is released again
int packed_decode_one(tc_ctx *ctx, const void *blob, uint64_t bytes, uint8_t *text, uint64_t cap, uint64_t *n_out) {
    *n_out = 0;
    tc_packed_header h;
    TC_TRY(packed_check(blob, bytes, &h));
    if (h.R == 0) return TC_OK;
    tc_block_info info;
    packed_to_info(h, &info);
    const uint8_t *base = (const uint8_t *)blob;
    uint8_t *d_blob; // payload sections keep their 16-byte alignment on the device
    uint32_t *d_count;
    int16_t *d_rsym;
    const uint64_t pay = h.total_bytes - h.off_cnt4;
    WsMark mk = tc_ws_mark(ctx);
    TC_TRY(ws_alloc(ctx, pay, &d_blob));
    TC_TRY(ws_alloc(ctx, h.R, &d_count));
    TC_TRY(ws_alloc(ctx, h.R, &d_rsym));
    TC_TRY(h2d(ctx, d_blob, base + h.off_cnt4, pay));
    const uint64_t o0 = h.off_cnt4;
    TC_TRY(rle_unpack_dev_impl(ctx, d_blob, d_blob + (h.off_sym8 - o0), (const uint32_t *)(d_blob + (h.off_hi - o0)),
                               (const uint64_t *)(d_blob + (h.off_big_idx - o0)),
                               (const uint32_t *)(d_blob + (h.off_big_cnt - o0)), h.n_big, h.R, d_count, d_rsym));
    int rc = decode_runs_dev(ctx, d_count, d_rsym, h.R, (h.flags & TC_PACKED_MTF) != 0, &info, text, cap, n_out);
    tc_ws_release(ctx, mk);
    return rc;
}
} // namespace

extern "C" int tc_packed_decode(tc_ctx *ctx, const void *blob, uint64_t bytes, uint8_t *text, uint64_t cap,
                                uint64_t *n_out) {
    TC_ENTER(ctx);
    if (!n_out) return TC_E_ARG;
    return packed_decode_one(ctx, blob, bytes, text, cap, n_out);
}

// Multi-block decompression: the lanes of tc_blocks_encode_packed, each taking the next unclaimed container and
// running copy in -> unpack -> RLE -> MTF -> BWT inverse -> copy out on its own context and stream, so one lane's copies
// and host syncs are covered by the other lanes' kernels.
extern "C" int tc_blocks_decode_packed(tc_ctx *ctx, uint64_t nblocks, const void *const *blob, const uint64_t *bytes,
                                       uint8_t *const *text, const uint64_t *cap, uint64_t *n_out) {
    if (!ctx) return TC_E_ARG;
    if (nblocks == 0) return TC_OK;
    if (!blob || !bytes || !text || !cap || !n_out) return TC_E_ARG;
    for (uint64_t b = 0; b < nblocks; b++) n_out[b] = 0;
    std::atomic<uint64_t> next{0};
    return run_lanes(ctx, nblocks, [&](tc_ctx *c) -> int {
        if (cudaSetDevice(c->device) != cudaSuccess) return TC_E_CUDA; // helper threads start on device 0
        TC_TRY(tc_ws_reset(c));
        DrainGuard drain{c};
        int rc_all = TC_OK;
        for (uint64_t b = next.fetch_add(1); b < nblocks; b = next.fetch_add(1)) {
            int rc = packed_decode_one(c, blob[b], bytes[b], text[b], cap[b], &n_out[b]);
            if (rc == TC_E_CAP) {
                rc_all = TC_E_CAP;
                continue;
            }
            if (rc != TC_OK) {
                next.store(nblocks); // the other lanes stop at their next block
                return rc;
            }
        }
        return rc_all;
    });
}
