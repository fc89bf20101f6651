// bwt.cu -- saToBWT (src/Data/BWT/Internal.hs:98-106) and fromBWT
// (src/Data/BWT.hs:93-104 + sortTB / magicInverseBWT, src/Data/BWT/Internal.hs:144-200).
//
// Inverse: the reference sorts (symbol, index) pairs -- a stable counting
// sort, done here as two 8-bit radix passes over a 9-bit code -- which yields psi, then walks
//   f = psi[0]; while f != 0: emit F[f]; f = psi[f]
// The walk is a linked list over rows; it is parallelised by list ranking with splitters:
// every K-th row is a splitter, each splitter walks to the next splitter (sub-list length),
// the reduced list is ranked by pointer jumping, and a second walk writes the text.
#include <algorithm>
#include <utility>

#include "common.cuh"
#include "impl.cuh"

int tc_byte_hist_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *h_hist);

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

template <class Src>
__global__ void __launch_bounds__(256)
    inv_keys_kernel(Src src, uint64_t N, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                    uint32_t *__restrict__ hist /*257*/) {
    __shared__ uint32_t h[257];
    for (int j = threadIdx.x; j < 257; j += 256) h[j] = 0;
    __syncthreads();
    uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += stride) {
        uint32_t c = src.at(i);
        keys[i] = c;
        vals[i] = (uint32_t)i;
        atomicAdd(&h[c], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 257; j += 256)
        if (h[j]) atomicAdd(&hist[j], h[j]);
}

struct CStart {
    uint32_t c[258]; // c[code] = first row of that code in F; c[257] = N
};

constexpr uint32_t NIL = 0xffffffffu;

// sub-list of splitter s: rows s*K, psi(s*K), ... up to (excluding) the next splitter row
__global__ void inv_walk1_kernel(const uint32_t *__restrict__ psi, uint64_t S, uint32_t K, uint64_t N,
                                 uint32_t *__restrict__ nxt, uint32_t *__restrict__ dist) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    uint32_t cur = (uint32_t)(s * K);
    uint32_t cnt = 0;
    do {
        cur = psi[cur];
        cnt++;
    } while (cur % K != 0 && cnt <= N); // psi is a permutation: the walk returns to a splitter within N steps
    uint32_t succ = cur / K;
    nxt[s] = succ == 0 ? NIL : succ; // the cycle through row 0 is cut just before row 0
    dist[s] = cnt;
}

// one pointer-jumping round (double buffered)
__global__ void inv_jump_kernel(const uint32_t *__restrict__ nxt, const uint32_t *__restrict__ dist, uint64_t S,
                                uint32_t *__restrict__ nxt2, uint32_t *__restrict__ dist2) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    uint32_t nx = nxt[s];
    uint32_t d = dist[s];
    if (nx != NIL) {
        d += dist[nx];
        nx = nxt[nx];
    }
    nxt2[s] = nx;
    dist2[s] = d;
}

// second walk: splitter s starts `total - dist[s]` steps after row 0
__global__ void inv_walk2_kernel(const uint32_t *__restrict__ psi, uint64_t S, uint32_t K,
                                 const uint32_t *__restrict__ nxt_final, const uint32_t *__restrict__ dist_final,
                                 CStart cs, uint8_t *__restrict__ text, uint64_t cap, uint64_t N,
                                 uint32_t *__restrict__ err) {
    __shared__ uint32_t sc[258];
    for (int j = threadIdx.x; j < 258; j += blockDim.x) sc[j] = cs.c[j];
    __syncthreads();
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (nxt_final[s] != NIL) return; // not on the cycle through row 0
    uint64_t total = dist_final[0];
    uint64_t g = total - dist_final[s];
    uint32_t cur = (uint32_t)(s * K);
    bool first = true;
    uint64_t guard = 0;
    while ((first || cur % K != 0) && guard++ <= N) {
        first = false;
        if (g >= 1) {
            // F[cur]: the code whose row range contains cur
            int lo = 0, hi = 257; // sc[lo] <= cur < sc[hi]
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (sc[mid] <= cur) lo = mid; else hi = mid;
            }
            if (lo == 0) {
                atomicMax(err, 1u); // fromJust Nothing (src/Data/BWT/Internal.hs:195)
            } else if (g - 1 < cap) {
                text[g - 1] = (uint8_t)(lo - 1);
            }
        }
        cur = psi[cur];
        g++;
    }
}

template <class Src>
int bwt_decode_impl(tc_ctx *ctx, Src src, uint64_t N, uint8_t *d_text, uint64_t cap, uint64_t *n_out) {
    *n_out = 0;
    if (N == 0) return TC_OK;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    WsMark mk = tc_ws_mark(ctx);
    uint64_t *k0, *k1;
    uint32_t *v0, *v1, *d_hist;
    TC_TRY(ws_alloc(ctx, N, &k0));
    TC_TRY(ws_alloc(ctx, N, &k1));
    TC_TRY(ws_alloc(ctx, N, &v0));
    TC_TRY(ws_alloc(ctx, N, &v1));
    TC_TRY(ws_alloc(ctx, 260, &d_hist));
    TC_CUDA(cudaMemsetAsync(d_hist, 0, 260 * sizeof(uint32_t), ctx->stream));
    unsigned grid = (unsigned)std::min<uint64_t>(ceil_div_u64(N, 256), (uint64_t)ctx->sm_count * 16);
    TC_LAUNCH(ctx, (inv_keys_kernel<Src>), grid, 256, 0, src, N, k0, v0, d_hist);
    uint32_t *h = (uint32_t *)ctx->h_scal;
    TC_TRY(tc_d2h_small(ctx, h, d_hist, 257 * sizeof(uint32_t)));
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
    // stable sort by code == psi
    int shifts[2] = {0, 8};
    uint64_t *ks;
    uint32_t *psi;
    TC_TRY(tc_radix_sort_pairs(ctx, k0, v0, k1, v1, N, shifts, 2, &ks, &psi));
    // list ranking
    const uint32_t K = 32;
    const uint64_t S = ceil_div_u64(N, K);
    uint32_t *nxtA = (uint32_t *)k0, *nxtB = nxtA + S, *dA = nxtB + S, *dB = dA + S; // keys are dead: 4S*4 <= 8N bytes
    uint32_t *d_err;
    TC_TRY(ws_alloc(ctx, 1, &d_err));
    TC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream));
    if (ks == k0) { // two passes: the sorted pairs are back in buffer 0, so buffer 1 is the free one
        nxtA = (uint32_t *)k1;
        nxtB = nxtA + S;
        dA = nxtB + S;
        dB = dA + S;
    }
    const unsigned gridS = (unsigned)ceil_div_u64(S, 128);
    TC_LAUNCH(ctx, inv_walk1_kernel, gridS, 128, 0, psi, S, K, N, nxtA, dA);
    int rounds = 1;
    while ((1ull << rounds) < S) rounds++;
    for (int r = 0; r < rounds; r++) {
        TC_LAUNCH(ctx, inv_jump_kernel, gridS, 128, 0, nxtA, dA, S, nxtB, dB);
        std::swap(nxtA, nxtB);
        std::swap(dA, dB);
    }
    TC_LAUNCH(ctx, inv_walk2_kernel, gridS, 128, 0, psi, S, K, nxtA, dA, cs, d_text, cap, N, d_err);
    TC_TRY(tc_d2h_small(ctx, h, dA, sizeof(uint32_t)));
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
