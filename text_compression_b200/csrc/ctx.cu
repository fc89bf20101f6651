// ctx.cu -- context lifecycle, scratch arena, pinned host memory, error strings.
#include <stdlib.h>

#include "common.cuh"
#include "impl.cuh"

static const size_t kAlign = 512;

extern "C" const char *tc_version(void) { return "text_compression_b200 0.1 (sm_100a)"; }

extern "C" const char *tc_strerror(int rc) {
    switch (rc) {
        case TC_OK: return "ok";
        case TC_E_CUDA: return "CUDA runtime error (see tc_last_error)";
        case TC_E_CAP: return "output capacity too small";
        case TC_E_FROMJUST: return "reference semantics: fromJust Nothing";
        case TC_E_INDEX: return "reference semantics: index out of bounds";
        case TC_E_NOMEM: return "out of memory";
        case TC_E_ARG: return "invalid argument";
        case TC_E_TOOBIG: return "input too large (n must be < 2^32-2)";
        case TC_E_NODEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    }
    return "unknown error";
}

extern "C" const char *tc_last_error(const tc_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t tc_ctx_launches(const tc_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t t = ctx->launches;
    for (tc_ctx *c : ctx->child) t += c ? c->launches : 0;
    return t;
}

static int ctx_create_impl(int device, cudaStream_t stream, bool have_stream, tc_ctx **out) {
    if (!out) return TC_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return TC_E_NODEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) return TC_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TC_E_NODEVICE;
    if (prop.major != 10 || prop.minor != 0) return TC_E_NODEVICE; // the cubin is sm_100a only: no image for any other part
    tc_ctx *ctx = new tc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->coop_ok = prop.cooperativeLaunch != 0 && !(getenv("TC_B200_NO_COOP") && getenv("TC_B200_NO_COOP")[0] == '1');
    const char *nm = getenv("TC_B200_NO_MSD");
    ctx->no_msd = nm && nm[0] == '1';
    const char *nr = getenv("TC_B200_NO_RAWKEY");
    ctx->no_rawkey = nr && nr[0] == '1';
    const char *m2 = getenv("TC_B200_MTF_V2");
    ctx->mtf_v2 = m2 && m2[0] == '1';
    const char *d1 = getenv("TC_B200_MTFD_V1");
    ctx->mtfd_v1 = d1 && d1[0] == '1';
    const char *ml = getenv("TC_B200_MTF_L");
    if (ml) ctx->mtf_L = (uint32_t)atoi(ml);
    const char *dg = getenv("TC_B200_DIAG"); // measurement only: 1 = packed batch without its H2D copies (after each slot's first), 2 = without its D2H copies
    ctx->diag = dg ? (uint32_t)atoi(dg) : 0;
    const char *ln = getenv("TC_B200_LANES");
    if (ln && ln[0] >= '1' && ln[0] <= '0' + tc_ctx::MAX_LANES) ctx->lanes = ln[0] - '0';
    if (have_stream) {
        ctx->stream = stream;
        ctx->own_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return TC_E_CUDA;
        }
        ctx->own_stream = true;
    }
    if (cudaMallocHost((void **)&ctx->h_scal, 1024 * sizeof(uint64_t)) != cudaSuccess) {
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return TC_E_NOMEM;
    }
    *out = ctx;
    return TC_OK;
}

extern "C" int tc_ctx_create(int device, tc_ctx **out) { return ctx_create_impl(device, nullptr, false, out); }
extern "C" int tc_ctx_create_on_stream(int device, void *cuda_stream, tc_ctx **out) {
    return ctx_create_impl(device, (cudaStream_t)cuda_stream, true, out);
}

extern "C" void tc_ctx_destroy(tc_ctx *ctx) {
    if (!ctx) return;
    for (tc_ctx *c : ctx->child) tc_ctx_destroy(c);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    mtf_free_tables(ctx);
    for (auto &c : ctx->chunks) cudaFree(c.p);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Per-kernel device timing: on != 0 starts recording one CUDA-event pair per launch on the
// context's stream; tc_ctx_profile_report syncs and writes "name\tlaunches\ttotal_ms\n" lines.
extern "C" int tc_ctx_profile(tc_ctx *ctx, int on) {
    if (!ctx) return TC_E_ARG;
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    ctx->prof_on = on != 0;
    return TC_OK;
}
extern "C" int tc_ctx_profile_report(tc_ctx *ctx, char *buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return TC_E_ARG;
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    struct Agg {
        const char *name;
        uint64_t n;
        double ms;
        uint64_t bytes;
    };
    std::vector<Agg> agg;
    for (auto &r : ctx->prof) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        bool found = false;
        for (auto &a : agg)
            if (a.name == r.name || strcmp(a.name, r.name) == 0) {
                a.n++;
                a.ms += ms;
                a.bytes += r.bytes;
                found = true;
                break;
            }
        if (!found) agg.push_back({r.name, 1, ms, r.bytes});
    }
    size_t o = 0;
    buf[0] = 0;
    for (auto &a : agg) {
        int w = snprintf(buf + o, cap - o, "%s\t%llu\t%.6f\t%llu\n", a.name, (unsigned long long)a.n, a.ms,
                         (unsigned long long)a.bytes);
        if (w < 0 || (size_t)w >= cap - o) break;
        o += (size_t)w;
    }
    return TC_OK;
}

extern "C" int tc_ctx_sync(tc_ctx *ctx) {
    if (!ctx) return TC_E_ARG;
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    return TC_OK;
}

extern "C" void *tc_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void tc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// The arena is a list of device chunks walked by a cursor (cur_chunk, cur_off).  A request that
// does not fit the current chunk moves on to the next chunk that can hold it, and only when none
// can is a new chunk allocated; reset and release just move the cursor back.  A repeated
// allocation sequence (block after block of the same size) therefore reaches a steady state
// after its first pass and never calls cudaMalloc again.
int tc_ws_reset(tc_ctx *ctx) {
    TC_CUDA(cudaSetDevice(ctx->device));
    ctx->cur_chunk = 0;
    ctx->cur_off = 0;
    ctx->used_total = 0;
    return TC_OK;
}

int tc_ws_alloc(tc_ctx *ctx, size_t bytes, void **out) {
    bytes = (bytes + kAlign - 1) / kAlign * kAlign;
    if (bytes == 0) bytes = kAlign;
    while (ctx->cur_chunk < ctx->chunks.size() && ctx->cur_off + bytes > ctx->chunks[ctx->cur_chunk].cap) {
        ctx->cur_chunk++;
        ctx->cur_off = 0;
    }
    if (ctx->cur_chunk >= ctx->chunks.size()) {
        size_t cap = bytes > (size_t(64) << 20) ? bytes : (size_t(64) << 20);
        char *p = nullptr;
        if (cudaMalloc((void **)&p, cap) != cudaSuccess) {
            cudaGetLastError();
            snprintf(ctx->err, sizeof ctx->err, "scratch arena: cudaMalloc(%zu) failed", cap);
            return TC_E_NOMEM;
        }
        ctx->chunks.push_back({p, cap});
        ctx->cur_chunk = ctx->chunks.size() - 1;
        ctx->cur_off = 0;
    }
    *out = ctx->chunks[ctx->cur_chunk].p + ctx->cur_off;
    ctx->cur_off += bytes;
    ctx->used_total += bytes;
    return TC_OK;
}

WsMark tc_ws_mark(tc_ctx *ctx) { return WsMark{ctx->cur_chunk, ctx->cur_off, ctx->used_total}; }
void tc_ws_release(tc_ctx *ctx, WsMark m) {
    ctx->cur_chunk = m.nchunks;
    ctx->cur_off = m.off;
    ctx->used_total = m.used;
}

static __global__ void d2h_small_kernel(unsigned char *dst, const unsigned char *src, size_t bytes) {
    for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}
int tc_d2h_small(tc_ctx *ctx, void *h_pinned_dst, const void *d_src, size_t bytes) {
    if (bytes == 0) return TC_OK;
    d2h_small_kernel<<<1, 256, 0, ctx->stream>>>((unsigned char *)h_pinned_dst, (const unsigned char *)d_src, bytes);
    ctx->launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return ctx->fail(e, "d2h_small_kernel", __LINE__);
    return TC_OK;
}
