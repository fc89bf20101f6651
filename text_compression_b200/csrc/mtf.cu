// mtf.cu -- seqToMTF / seqFromMTF (src/Data/MTF/Internal.hs:128-232) as a chunked
// parallel scan over composable move-to-front summaries.
//
// Symbols are first mapped to alphabet ranks 0..sigma-1 (nubSeq', :79-99: sorted,
// Nothing first), so the initial list L0 is the identity permutation.
//
// Encode, sigma <= 8 (ACGT(N) + sentinel): the list is one register, see "small alphabets".
// Encode, generic alphabet (sigma <= 257): the list state at a position is fully described by
// the last occurrence of every rank before it (more recent = closer to the front), so
//   K1  warp per chunk: last occurrence of every rank inside the chunk
//   K2  exclusive max-scan of those rows over chunks (tiles of 64 chunks, then over tiles);
//       the row after the last chunk gives the FINAL list, seqToMTF's second component
//   K3  warp per chunk: replay, 32 positions per step.  The list is never materialised: every
//       rank owns a time slot (virtual slots for the incoming order, then one slot per
//       position) and a bitmap marks the slots that are "latest occurrence of their rank".
//       The MTF index of c is popcount(bitmap above last[c]); ranks touched earlier in the
//       same step are corrected with ballots.
// Decode: same chunk/tile structure; a chunk's summary is the permutation it applies to
//   list positions, composition is a gather.
// Algorithmic bytes: N * (1 + w_idx) (u8 symbols in, u16 indices out here: 3N).
#include <algorithm>

#include "common.cuh"
#include "impl.cuh"
#include "mtfd_select.cuh"

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
    struct Raw {
        uint4 v;
    };
    __device__ __forceinline__ Raw load_raw(uint64_t base) const { return Raw{__ldg(reinterpret_cast<const uint4 *>(p + base))}; }
    __device__ __forceinline__ void decode_li(const Raw &r, uint64_t base, uint32_t *li) const { // 256 = Nothing, else the byte
        const uint4 v = r.v;
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) li[k] = (w[k >> 2] >> ((k & 3) * 8)) & 0xff;
        if (primary - base < 16) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (base + k == primary) li[k] = 256u;
        }
    }
    __device__ __forceinline__ int at(uint64_t i) const { return i == primary ? 0 : (int)p[i] + 1; }
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
    struct Raw {
        uint4 v[2];
    };
    __device__ __forceinline__ Raw load_raw(uint64_t base) const {
        Raw r;
        r.v[0] = __ldg(reinterpret_cast<const uint4 *>(p + base));
        r.v[1] = __ldg(reinterpret_cast<const uint4 *>(p + base + 8));
        return r;
    }
    __device__ __forceinline__ void decode_li(const Raw &r, uint64_t, uint32_t *li) const { // 256 = Nothing, else the byte
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const uint4 v = r.v[q];
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int s = (int)(int16_t)((w[k >> 1] >> ((k & 1) * 16)) & 0xffff);
                li[q * 8 + k] = s < 0 ? 256u : (uint32_t)(s & 0xff);
            }
        }
    }
    __device__ __forceinline__ int at(uint64_t i) const {
        int v = p[i];
        return v < 0 ? 0 : (v & 0xff) + 1;
    }
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
    __device__ __forceinline__ uint16_t get16(int i) const {
        return reinterpret_cast<const uint16_t *>(&sm[(i >> 1) * blockDim.x + threadIdx.x])[i & 1];
    }
    __device__ __forceinline__ void set16(int i, uint16_t v) const {
        reinterpret_cast<uint16_t *>(&sm[(i >> 1) * blockDim.x + threadIdx.x])[i & 1] = v;
    }
};

// Move list entry r (16-bit entries, two per word) to the front and return it.  Entries 0..r-1
// shift up by one: word w becomes (w << 16) | (w-1 >> 16), one load, one funnel shift and one
// store per TWO entries, the previous word carried in a register.
__device__ __forceinline__ uint32_t list_take16(const TState &lst, uint32_t r) {
    const int wr = (int)(r >> 1);
    uint32_t cur = lst.sm[wr * blockDim.x + threadIdx.x];
    const uint32_t sym = (r & 1u) ? (cur >> 16) : (cur & 0xffffu);
    if (r == 0) return sym;
    const unsigned T = blockDim.x, t = threadIdx.x;
    int w = wr;
    if (!(r & 1u) && w > 0) { // even r: entry r+1 (upper half of word wr) stays
        const uint32_t prev = lst.sm[(w - 1) * T + t];
        lst.sm[w * T + t] = (cur & 0xffff0000u) | (prev >> 16);
        cur = prev;
        w--;
    }
    // eight words per trip: the loads are issued before the (possibly aliasing) stores
    for (; w >= 8; w -= 8) {
        uint32_t p[9];
        p[0] = cur;
#pragma unroll
        for (int q = 1; q <= 8; q++) p[q] = lst.sm[(w - q) * T + t];
#pragma unroll
        for (int q = 0; q < 8; q++) lst.sm[(w - q) * T + t] = __funnelshift_l(p[q + 1], p[q], 16); // (p[q] << 16) | (p[q+1] >> 16)
        cur = p[8];
    }
    for (; w > 0; w--) {
        const uint32_t prev = lst.sm[(w - 1) * T + t];
        lst.sm[w * T + t] = __funnelshift_l(prev, cur, 16);
        cur = prev;
    }
    lst.sm[threadIdx.x] = (cur << 16) | sym; // word 0: entry 0 = the moved entry, entry 1 = old entry 0
    return sym;
}

// ---- encode v2: warp per chunk ------------------------------------------------------------
constexpr int VSMAX = 288;           // virtual slots: sigma rounded up to a multiple of 32
constexpr int ENC_WARPS = 8;         // warps (chunks) per CTA
constexpr int WIN_WORDS = 64;        // recency window of the prologue: 2048 positions

// K1: last occurrence (global position + 1, 0 = absent) of every rank inside each chunk
template <class Src>
__global__ void __launch_bounds__(ENC_WARPS * 32)
    mtf2_lastocc_kernel(Src src, Lut lut, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t VS,
                        uint32_t *__restrict__ lastocc) {
    __shared__ uint16_t s_rank[SIGMAX];
    __shared__ uint32_t lo[ENC_WARPS][VSMAX];
    for (int j = threadIdx.x; j < SIGMAX; j += blockDim.x) s_rank[j] = lut.rank[j];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    for (int j = lane; j < VSMAX; j += 32) lo[w][j] = 0;
    __syncthreads();
    uint64_t k = (uint64_t)blockIdx.x * ENC_WARPS + w;
    if (k >= nchunks) return;
    uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    for (uint64_t b0 = beg; b0 < end; b0 += 8 * 32) {
        int c[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint64_t pos = b0 + q * 32 + lane;
            c[q] = pos < end ? (int)s_rank[src.at(pos)] : -1;
        }
        // rows ascend, so a later row simply overwrites; inside a row the largest position must
        // win: plain stores, re-tried by the lanes that read back something smaller
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t v = (uint32_t)(b0 + q * 32 + lane) + 1;
            bool again = c[q] >= 0;
            do {
                if (again) lo[w][c[q]] = v;
                __syncwarp();
                again = again && lo[w][c[q]] < v;
            } while (__any_sync(TC_FULL, again));
        }
    }
    for (int j = lane; j < (int)VS; j += 32) lastocc[k * VS + j] = lo[w][j];
}

// K2a: exclusive running max over the chunks of each tile (in place) + tile totals
__global__ void mtf2_scan_tiles_kernel(uint32_t *__restrict__ lastocc, uint64_t nchunks, uint32_t G, uint32_t VS,
                                       uint32_t *__restrict__ tiletot) {
    uint32_t c = threadIdx.x;
    if (c >= VS) return;
    uint64_t k0 = (uint64_t)blockIdx.x * G, k1 = k0 + G < nchunks ? k0 + G : nchunks;
    uint32_t run = 0;
    for (uint64_t kb = k0; kb < k1; kb += 16) {
        uint32_t v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = (kb + q < k1) ? lastocc[(kb + q) * VS + c] : 0;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            if (kb + q < k1) lastocc[(kb + q) * VS + c] = run;
            run = max(run, v[q]);
        }
    }
    tiletot[(uint64_t)blockIdx.x * VS + c] = run;
}
// K2b: exclusive running max over tiles (in place); finalocc = last occurrence in the whole input
__global__ void mtf2_scan_top_kernel(uint32_t *__restrict__ tiletot, uint64_t ntiles, uint32_t VS,
                                     uint32_t *__restrict__ finalocc) {
    uint32_t c = threadIdx.x;
    if (c >= VS) return;
    uint32_t run = 0;
    for (uint64_t tb = 0; tb < ntiles; tb += 32) {
        uint32_t v[32];
#pragma unroll
        for (int q = 0; q < 32; q++) v[q] = (tb + q < ntiles) ? tiletot[(tb + q) * VS + c] : 0;
#pragma unroll
        for (int q = 0; q < 32; q++) {
            if (tb + q < ntiles) tiletot[(tb + q) * VS + c] = run;
            run = max(run, v[q]);
        }
    }
    finalocc[c] = run;
}

// List position of every rank at text position `start`, from its last occurrence before `start`
// (T[c] = position + 1, 0 = never seen: those keep the alphabet order behind all seen ones).
// pos(c) = number of ranks with a more recent occurrence.  Recent occurrences (within 2048
// positions) are ranked with a bitmap + prefix popcounts; the few older ones pairwise.
// tvals: 9 values per lane, rank j = lane + 32*i.  Writes pos into posv[9].  Warp-collective.
struct WarpPro {
    uint32_t win[WIN_WORDS];
    uint32_t pre[WIN_WORDS];
    uint32_t okey[VSMAX];
    uint32_t ocount;
};
// General form: entry i of a lane is live iff valid[i]; never-seen entries keep the order of ord[i].
__device__ __forceinline__ void list_positions_g(const uint32_t *tvals, const bool *valid, const uint32_t *ord,
                                                 uint64_t start, WarpPro &P, uint32_t *posv);
__device__ __forceinline__ void list_positions(const uint32_t *tvals, uint64_t start, uint32_t sigma, WarpPro &P,
                                               uint32_t *posv) {
    bool valid[9];
    uint32_t ord[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        ord[i] = lane_id() + 32 * i;
        valid[i] = ord[i] < sigma;
    }
    list_positions_g(tvals, valid, ord, start, P, posv);
}
__device__ __forceinline__ void list_positions_g(const uint32_t *tvals, const bool *valid, const uint32_t *ord,
                                                 uint64_t start, WarpPro &P, uint32_t *posv) {
    const unsigned lane = lane_id();
    P.win[lane] = 0;
    P.win[lane + 32] = 0;
    if (lane == 0) P.ocount = 0;
    __syncwarp();
    uint32_t key[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint32_t T = tvals[i];
        // distance back to the last occurrence (>= 1); never-seen ranks sort after every seen one
        key[i] = T ? (uint32_t)(start - (T - 1)) : (0x80000000u + ord[i]);
        if (valid[i] && key[i] <= WIN_WORDS * 32) atomicOr(&P.win[(key[i] - 1) >> 5], 1u << ((key[i] - 1) & 31));
    }
    __syncwarp();
    // prefix popcounts of the window, two words per lane
    uint32_t p0 = __popc(P.win[2 * lane]), p1 = __popc(P.win[2 * lane + 1]);
    uint32_t inc = warp_incl_sum(p0 + p1);
    P.pre[2 * lane] = inc - p0 - p1;
    P.pre[2 * lane + 1] = inc - p1;
    uint32_t m_in = __shfl_sync(TC_FULL, inc, 31);
    // collect the keys outside the window
#pragma unroll
    for (int i = 0; i < 9; i++) {
        bool out = valid[i] && key[i] > WIN_WORDS * 32;
        unsigned m = __ballot_sync(TC_FULL, out);
        uint32_t base = P.ocount;
        __syncwarp();
        if (out) P.okey[base + __popc(m & lanemask_lt())] = key[i];
        if (lane == 0) P.ocount = base + __popc(m);
        __syncwarp();
    }
    uint32_t m_out = P.ocount;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint32_t p = 0;
        if (valid[i]) {
            if (key[i] <= WIN_WORDS * 32) {
                uint32_t b = key[i] - 1;
                p = P.pre[b >> 5] + __popc(P.win[b >> 5] & ((1u << (b & 31)) - 1));
            } else {
                p = m_in;
                for (uint32_t e = 0; e < m_out; e++) p += P.okey[e] < key[i];
            }
        }
        posv[i] = p;
    }
    __syncwarp();
}

// K3: replay, one warp per chunk, 32 positions per step (lane t <-> position base + t).
// Slots: list position j at chunk start owns virtual slot VS-1-j; chunk offset o owns slot VS+o.
// bits[] marks slots that are the latest occurrence of their rank (always exactly sigma bits),
// pre[] holds the exclusive prefix popcount per word, last[r] the slot of rank r.
struct WarpRep {
    uint32_t bits[(VSMAX + 1024) / 32];
    uint32_t pre[(VSMAX + 1024) / 32];
    uint16_t last[VSMAX];
};
template <class Src>
__global__ void __launch_bounds__(ENC_WARPS * 32)
    mtf2_replay_kernel(Src src, Lut lut, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t G, uint32_t sigma,
                       uint32_t VS, const uint32_t *__restrict__ lastocc, const uint32_t *__restrict__ tiletot,
                       uint16_t *__restrict__ idx_out) {
    __shared__ uint16_t s_rank[SIGMAX];
    __shared__ WarpPro pro[ENC_WARPS];
    __shared__ WarpRep rep[ENC_WARPS];
    for (int j = threadIdx.x; j < SIGMAX; j += blockDim.x) s_rank[j] = lut.rank[j];
    __syncthreads();
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    uint64_t k = (uint64_t)blockIdx.x * ENC_WARPS + w;
    if (k >= nchunks) return;
    WarpRep &R = rep[w];
    const uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    const int NW = (int)((VS + L) / 32);
    // ---- prologue: incoming list order -> slots
    {
        uint32_t tv[9], pv[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            uint32_t c = lane + 32 * i;
            tv[i] = c < VS ? max(lastocc[k * VS + c], tiletot[(k / G) * VS + c]) : 0;
        }
        list_positions(tv, beg, sigma, pro[w], pv);
        for (int j = lane; j < NW; j += 32) R.bits[j] = 0;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 9; i++) {
            uint32_t c = lane + 32 * i;
            if (c < sigma) {
                uint32_t slot = VS - 1 - pv[i];
                R.last[c] = (uint16_t)slot;
                atomicOr(&R.bits[slot >> 5], 1u << (slot & 31));
            }
        }
        __syncwarp();
    }
    const unsigned lt = lanemask_lt();
    for (uint32_t o = 0; beg + o < end; o += 32) {
        // prefix popcounts of the bitmap as of the start of this step
        {
            uint32_t a0 = 2 * lane < (unsigned)NW ? R.bits[2 * lane] : 0;
            uint32_t a1 = 2 * lane + 1 < (unsigned)NW ? R.bits[2 * lane + 1] : 0;
            uint32_t q0 = __popc(a0), q1 = __popc(a1);
            uint32_t inc = warp_incl_sum(q0 + q1);
            if (2 * lane < (unsigned)NW) R.pre[2 * lane] = inc - q0 - q1;
            if (2 * lane + 1 < (unsigned)NW) R.pre[2 * lane + 1] = inc - q1;
        }
        __syncwarp();
        const uint64_t pos = beg + o + lane;
        const bool valid = pos < end;
        const uint32_t c = valid ? (uint32_t)s_rank[src.at(pos)] : 0;
        const uint32_t base = VS + o;
        const unsigned peers = match_bits<9>(c, valid);
        const unsigned lower = peers & lt;
        const bool has_prev = valid && lower != 0;
        const bool first = valid && lower == 0;
        const bool is_last = valid && (peers & ~((2u << lane) - 1)) == 0;
        const int u = has_prev ? 31 - __clz(lower) : -1; // previous lane with the same rank
        uint32_t rank = 0;
        uint32_t a = 0;
        uint32_t r0 = 0;
        if (first) {
            a = R.last[c];
            uint32_t wa = a >> 5;
            uint32_t below = R.pre[wa] + __popc(R.bits[wa] & ((2u << (a & 31)) - 1)); // set bits at slots <= a
            r0 = sigma - below;                                                       // list position at step start
        }
        // ranks first touched in this step: r0 minus the earlier-touched ranks that were ahead, plus
        // the number of distinct ranks touched earlier in the step
        const unsigned firstmask = __ballot_sync(TC_FULL, first);
        {
            unsigned ltm = 0, eq = firstmask;
#pragma unroll
            for (int bit = 8; bit >= 0; bit--) {
                bool mb = (r0 >> bit) & 1;
                unsigned Bm = __ballot_sync(TC_FULL, first && mb);
                if (mb) {
                    ltm |= eq & ~Bm;
                    eq &= Bm;
                } else {
                    eq &= ~Bm;
                }
            }
            if (first) rank = r0 - __popc(ltm & lt & firstmask) + __popc(firstmask & lt);
        }
        // ranks seen earlier in this step: distinct ranks strictly between the two occurrences
        unsigned pend = __ballot_sync(TC_FULL, has_prev);
        while (pend) {
            int l = __ffs(pend) - 1;
            int us = __shfl_sync(TC_FULL, u, l);
            unsigned Bq = __ballot_sync(TC_FULL, valid && u < us); // lanes whose previous occurrence is before us
            if ((int)lane == l) rank = __popc(Bq & lt & ~((2u << us) - 1));
            pend &= pend - 1;
        }
        if (valid) idx_out[pos] = (uint16_t)rank;
        // state update
        if (first) atomicAnd(&R.bits[a >> 5], ~(1u << (a & 31)));
        unsigned lastmask = __ballot_sync(TC_FULL, is_last);
        if (lane == 0) R.bits[base >> 5] = lastmask;
        if (is_last) R.last[c] = (uint16_t)(base + lane);
        __syncwarp();
    }
}

// final list = order at the end of the input (seqToMTF's second component)
__global__ void __launch_bounds__(32)
    mtf2_final_kernel(const uint32_t *__restrict__ finalocc, uint64_t N, uint32_t sigma, uint32_t VS,
                      uint16_t *__restrict__ final_list) {
    __shared__ WarpPro pro;
    const unsigned lane = lane_id();
    uint32_t tv[9], pv[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint32_t c = lane + 32 * i;
        tv[i] = c < VS ? finalocc[c] : 0;
    }
    list_positions(tv, N, sigma, pro, pv);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint32_t c = lane + 32 * i;
        if (c < sigma) final_list[pv[i]] = (uint16_t)c;
    }
}

// ---- encode v3: thread per chunk -------------------------------------------------------------
// The warp-per-chunk replay above spends 342 warp instructions per 32 symbols resolving the
// dependencies between the 32 positions of a step.  Here every THREAD replays its own chunk,
// one symbol after the other, so a warp advances 32 chunks per step and nothing has to be
// resolved across lanes.  Per thread, in shared memory (word-interleaved across the CTA, so
// any per-thread index is conflict-free):
//   last[c]   u16: time slot of the latest occurrence of code c (indexed by lidx(c))
//   bits[32]  one bit per time slot (288 virtual slots for the incoming order + one per
//             position of the chunk, at most 1024): set = latest occurrence of its code
//   sfx[g]    8 bytes per group of 8 words: byte k = live bits in words k+1..7 of the group
// and in registers T (4 x u16: live bits in the groups above g) and the bitmap word that is
// still being appended to.  MTF index of c = live bits above last[c] = one popcount + one
// byte of sfx + one field of T: ~50 instructions per symbol per thread, i.e. ~50 warp
// instructions per 32 symbols.
constexpr int R3_VS = 288;                              // virtual slots (a multiple of 32)
constexpr int R3_LMAX = 736;                            // R3_VS + L <= 1024 slots
constexpr int R3_LASTW = 129;                           // u16 pairs for 257 codes (+1 pad)
constexpr int R3_BITW = 32;
constexpr int R3_SFXW = 8;                              // 4 groups x 8 bytes
constexpr int R3_WORDS = R3_LASTW + R3_BITW + R3_SFXW;  // 169 words = 676 B per thread
constexpr int R3_CT = 160;                              // threads per CTA: two CTAs per SM
constexpr int R3_ROW = 288;                             // u16 entries per start row (576 B, 16-byte multiple)

__device__ __forceinline__ uint32_t ldg_u16(const uint16_t *p) { // zero-extended straight into a 32-bit register
    uint32_t r;
    asm volatile("ld.global.u16 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t shr_clamp(uint32_t x, uint32_t s) { // x >> s, 0 for s >= 32
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
}
// code (0 = Nothing, 1 + byte) -> index into last[]: bytes at 0..255, Nothing at 256
__device__ __forceinline__ uint32_t r3_lidx(uint32_t code) { return code == 0 ? 256u : code - 1u; }

// A tile = the 32 consecutive chunks one warp replays (thread t <-> chunk 32 * tile + t).
//   T1  per tile: last occurrence (global position + 1) of every code inside the tile
//   T2  exclusive max-scan of those rows over the tiles; the total gives the FINAL list
//   T3  per tile: incoming list order from the scanned row, then the order at the start of each of
//       its 32 chunks: chunk t moves its codes to the front in the order of their last occurrence q
//       inside the chunk, i.e. every code gets the key  L - q  (seen)  or  L + old position
//       (not seen) and its new position is the number of smaller keys -- a 1024-bit map of the keys,
//       one word per lane, and its prefix popcounts.  Written as "start rows" (slot of every code)
//       that the replay kernel loads.
// State is indexed by li = r3_lidx(code) throughout; lane l owns li = l, l + 32, ... (no alphabet ranks:
// the replay never needs them).  present.w[i] bit l <=> li = l + 32 i occurs in the input.
struct Present {
    uint32_t w[9];
};
constexpr int TP_RLW = 131; // words per row of q values (257 u16 entries; odd, so rows start in different banks)

// A T-tile (the unit of T1 / T2 / T3) is TT_CH = 16 consecutive chunks: half the chain of dependent steps of a
// 32-chunk tile in T3 and twice as many warps to hide its latencies.  In T1 two lanes scan one chunk, half each.
constexpr int TT_CH = 16;
// rows[2 c + h][li] = 1 + offset (inside chunk c of the tile) of the last occurrence of li in half h of the chunk
// (0 = none); the chunk's value is the maximum of its two rows.  Lane l scans half l & 1 of chunk l >> 1.
template <class Src>
__device__ __forceinline__ void tile_positions(const Src &src, uint64_t N, uint32_t L, uint64_t nchunks, uint64_t tile,
                                               uint32_t *rows) {
    const unsigned lane = lane_id();
    for (int j = lane; j < 32 * TP_RLW; j += 32) rows[j] = 0;
    __syncwarp();
    const uint64_t k = tile * TT_CH + (lane >> 1);
    if (k < nchunks) { // forward pass over the own half chunk: later positions overwrite earlier ones
        uint16_t *row = reinterpret_cast<uint16_t *>(rows) + lane * (2 * TP_RLW);
        const uint64_t cbeg = k * L, cend = cbeg + L < N ? cbeg + L : N;
        const uint64_t beg = cbeg + (lane & 1u) * (L / 2);              // L is a multiple of 32
        const uint64_t end = (lane & 1u) ? cend : (cbeg + L / 2 < cend ? cbeg + L / 2 : cend);
        uint64_t pos = beg;
        if (beg < end && src.can_vec(beg)) {
            // each lane streams its own range: keep 64 symbols' worth of loads in flight per lane
            const uint64_t vend = beg + ((end - beg) & ~63ull);
            for (; pos < vend; pos += 64) {
                typename Src::Raw raw[4];
#pragma unroll
                for (int b = 0; b < 4; b++) raw[b] = src.load_raw(pos + 16 * b);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t li[16];
                    src.decode_li(raw[b], pos + 16 * b, li);
                    const uint32_t o = (uint32_t)(pos - cbeg) + 16 * b + 1;
#pragma unroll
                    for (int j = 0; j < 16; j++) row[li[j]] = (uint16_t)(o + j);
                }
            }
        }
        for (; pos < end; pos++) row[r3_lidx((uint32_t)src.at(pos))] = (uint16_t)(pos - cbeg + 1);
    }
    __syncwarp();
}

constexpr int T1_WARPS = 1; // one tile per CTA: 1,366 tiles of a 16 MiB block spread evenly over the SMs
// T1 also leaves q[t][li] in the chunk's start row (global memory): T3 reads it from there and replaces it
// by the slot, so the symbols are scanned once for both.
template <class Src>
__global__ void __launch_bounds__(T1_WARPS * 32)
    mtf3_tile_last_kernel(Src src, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t *__restrict__ trow,
                          uint16_t *__restrict__ start, uint32_t *__restrict__ scan_ticket, uint32_t *__restrict__ rle_ticket) {
    extern __shared__ __align__(16) uint32_t smt1[];
    if (blockIdx.x == 0 && threadIdx.x == 0) scan_ticket[0] = 0, rle_ticket[0] = 0; // the "last CTA" elections of T2 and of the replay
    const unsigned lane = lane_id();
    const uint64_t tile = (uint64_t)blockIdx.x * T1_WARPS + (threadIdx.x >> 5);
    if (tile * TT_CH >= nchunks) return;
    uint32_t *rows = smt1 + (threadIdx.x >> 5) * (32 * TP_RLW);
    tile_positions(src, N, L, nchunks, tile, rows);
    const uint16_t *q16 = reinterpret_cast<const uint16_t *>(rows);
    const uint64_t tbase = tile * TT_CH * L;
    const uint32_t nt = (uint32_t)(nchunks - tile * TT_CH < TT_CH ? nchunks - tile * TT_CH : TT_CH);
    uint32_t last[9];
#pragma unroll
    for (int i = 0; i < 9; i++) last[i] = 0;
#pragma unroll 2
    for (uint32_t t = 0; t < nt; t++) {
        uint16_t *out = start + (tile * TT_CH + t) * R3_ROW;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const uint32_t li = lane + 32 * i;
            uint32_t q = 0;
            if (li < SIGMAX) q = max((uint32_t)q16[(2 * t) * (2 * TP_RLW) + li], (uint32_t)q16[(2 * t + 1) * (2 * TP_RLW) + li]);
            out[li] = (uint16_t)q;
            if (q) last[i] = (t << 10) | q; // chunks ascend: the last one that saw the code wins
        }
    }
#pragma unroll
    for (int i = 0; i < 9; i++)
        trow[tile * R3_ROW + lane + 32 * i] =
            last[i] ? (uint32_t)(tbase + (uint64_t)(last[i] >> 10) * L + (last[i] & 1023u)) : 0u;
}

// final list = codes by descending last occurrence (li space; the host maps li -> symbol)
__device__ __forceinline__ void final_list_from(const uint32_t *finalocc, const Present &present, uint64_t N, WarpPro &pro,
                                                uint16_t *__restrict__ final_list) {
    const unsigned lane = lane_id();
    uint32_t tv[9], pv[9], ord[9];
    bool valid[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t li = lane + 32 * i;
        valid[i] = (present.w[i] >> lane) & 1u;
        ord[i] = li == 256 ? 0u : li + 1u; // alphabet order: Nothing first
        tv[i] = valid[i] ? finalocc[li] : 0;
    }
    list_positions_g(tv, valid, ord, N, pro, pv);
#pragma unroll
    for (int i = 0; i < 9; i++)
        if (valid[i]) final_list[pv[i]] = (uint16_t)(lane + 32 * i);
}

// T2: exclusive running max over the tiles for every column.  The tiles are cut into T2_SEGS segments; CTA
// (column group, segment) scans its segment in place (exclusive, from 0) and publishes the segment total; the CTA
// that finishes last turns the totals into exclusive segment prefixes (segpre) + finalocc (the final list is
// ranked from it by one extra warp of the T3 launch, off the critical path).  A consumer takes max(trow[tile][li], segpre[segment of tile][li]).
constexpr int T2_SEGS = 16;
constexpr int T2_WARPS = 8;
__global__ void __launch_bounds__(T2_WARPS * 32)
    mtf3_tile_scan_kernel(uint32_t *__restrict__ trow, uint64_t ntiles, uint32_t seg_tiles, uint32_t *__restrict__ segpre,
                          uint32_t *__restrict__ finalocc, uint32_t *__restrict__ scan_ticket) {
    __shared__ uint32_t part[T2_WARPS][33];
    __shared__ uint32_t s_last;
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const uint32_t col = blockIdx.x * 32 + lane;
    const uint64_t s0 = (uint64_t)blockIdx.y * seg_tiles, s1 = s0 + seg_tiles < ntiles ? s0 + seg_tiles : ntiles;
    const uint64_t per = (seg_tiles + T2_WARPS - 1) / T2_WARPS; // <= 16 up to 64 MiB: one batch of loads per warp
    const uint64_t t0 = s0 + (uint64_t)w * per < s1 ? s0 + (uint64_t)w * per : s1, t1 = t0 + per < s1 ? t0 + per : s1;
    uint32_t mx = 0;
    for (uint64_t tb = t0; tb < t1; tb += 16) {
        uint32_t v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = tb + q < t1 ? trow[(tb + q) * R3_ROW + col] : 0u;
#pragma unroll
        for (int q = 0; q < 16; q++) mx = max(mx, v[q]);
    }
    part[w][lane] = mx;
    __syncthreads();
    uint32_t run = 0;
    for (int ww = 0; ww < w; ww++) run = max(run, part[ww][lane]);
    for (uint64_t tb = t0; tb < t1; tb += 16) {
        uint32_t v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = tb + q < t1 ? trow[(tb + q) * R3_ROW + col] : 0u;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            if (tb + q < t1) trow[(tb + q) * R3_ROW + col] = run;
            run = max(run, v[q]);
        }
    }
    if (w == T2_WARPS - 1) segpre[(uint64_t)blockIdx.y * R3_ROW + col] = run; // the segment's total for now
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(scan_ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence(); // every CTA's totals are visible: this CTA finishes the job, one column per thread
    for (uint32_t c = threadIdx.x; c < R3_ROW; c += T2_WARPS * 32) {
        uint32_t tot[T2_SEGS];
#pragma unroll
        for (int sg = 0; sg < T2_SEGS; sg++) tot[sg] = __ldcg(&segpre[sg * R3_ROW + c]);
        uint32_t r = 0;
#pragma unroll
        for (int sg = 0; sg < T2_SEGS; sg++) {
            segpre[sg * R3_ROW + c] = r;
            r = max(r, tot[sg]);
        }
        finalocc[c] = r;
    }
}

// T3: start rows.  Warp per tile; lane l owns entries li = l + 32 i, i < 9.  All 288 entries take part
// in every step (entries that are not codes of the input sit behind the sigma real ones and stay
// there), so the step has no predicates.
constexpr int T3_WARPS = 1; // as T1
struct StartsWarp {
    uint32_t map8[2][256]; // the key map of a step, one BYTE per key (plain stores; shared-memory atomics cost 2 cycles
                           // per lane and were the whole kernel), double buffered
    uint2 wp[32];          // the same map as bits: .x = one word per lane, .y = its exclusive prefix popcount
    WarpPro pro;
};
__global__ void __launch_bounds__(T3_WARPS * 32)
    mtf3_starts_kernel(Present present, uint32_t sigma, uint32_t L, uint64_t nchunks, const uint32_t *__restrict__ trow,
                       const uint32_t *__restrict__ segpre, uint32_t seg_tiles, uint16_t *__restrict__ start,
                       const uint32_t *__restrict__ finalocc, uint64_t N, uint16_t *__restrict__ final_list) {
    __shared__ StartsWarp sw[T3_WARPS];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    if (blockIdx.x == gridDim.x - 1) { // the extra CTA: seqToMTF's second component, the list after the last symbol
        if (w == 0) final_list_from(finalocc, present, N, sw[0].pro, final_list);
        return;
    }
    const uint64_t tile = (uint64_t)blockIdx.x * T3_WARPS + w;
    if (tile * TT_CH >= nchunks) return;
    StartsWarp &W = sw[w];
    const unsigned lt = lanemask_lt();
    // list order at the start of the tile
    uint32_t lpos[9];
    {
        uint32_t tv[9], ord[9];
        bool valid[9];
        uint32_t inv_base = sigma;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const uint32_t li = lane + 32 * i;
            valid[i] = (present.w[i] >> lane) & 1u;
            ord[i] = li == 256 ? 0u : li + 1u;
            tv[i] = valid[i] ? max(trow[tile * R3_ROW + li], segpre[(tile / seg_tiles) * R3_ROW + li]) : 0;
        }
        list_positions_g(tv, valid, ord, tile * TT_CH * L, W.pro, lpos);
#pragma unroll
        for (int i = 0; i < 9; i++) { // the other entries: positions sigma .. 287, in any fixed order
            const unsigned b = __ballot_sync(TC_FULL, !valid[i]);
            if (!valid[i]) lpos[i] = inv_base + __popc(b & lt);
            inv_base += __popc(b);
        }
    }
#pragma unroll
    for (int j = 0; j < 16; j++) (&W.map8[0][0])[lane + 32 * j] = 0;
    __syncwarp();
    const uint32_t nt = (uint32_t)(nchunks - tile * TT_CH < TT_CH ? nchunks - tile * TT_CH : TT_CH);
    uint16_t *row = start + tile * TT_CH * R3_ROW + lane;
    // lane-rotated order in which a lane reads the 8 words (32 bytes) of its part of the byte map: no bank conflicts
    uint32_t woff[8], wsh[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        woff[j] = 8 * lane + ((j + lane) & 7u);
        wsh[j] = 4u * ((j + lane) & 7u);
    }
    uint32_t q[9], qn[9];
#pragma unroll
    for (int i = 0; i < 9; i++) qn[i] = ldg_u16(row + 32 * i);
    for (uint32_t t = 0; t < nt; t++, row += R3_ROW) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            q[i] = qn[i];
            row[32 * i] = (uint16_t)(R3_VS - 1 - lpos[i]); // slots of entries behind the sigma codes are never read
        }
        if (t + 1 == nt) break;
#pragma unroll
        for (int i = 0; i < 9; i++) qn[i] = ldg_u16(row + R3_ROW + 32 * i); // next chunk's q while this one is ranked
        uint8_t *map = reinterpret_cast<uint8_t *>(W.map8[t & 1]);
        uint32_t key[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            key[i] = L + lpos[i];          // not seen in the chunk: behind the seen ones, old order kept
            if (q[i]) key[i] = L - q[i];   // seen: by descending q
            map[key[i]] = 1;
        }
        __syncwarp();
        {
            // bytes 32 * lane .. + 31 of the map -> one word of bits
            const uint32_t *mw = W.map8[t & 1];
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) bits |= ((mw[woff[j]] * 0x01020408u) >> 24) << wsh[j];
            const uint32_t pc = __popc(bits);
            W.wp[lane] = make_uint2(bits, warp_incl_sum(pc) - pc);
            uint32_t *other = W.map8[(t & 1) ^ 1]; // last read before the barrier above
#pragma unroll
            for (int j = 0; j < 8; j++) other[lane + 32 * j] = 0;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const uint2 e = W.wp[key[i] >> 5];
            lpos[i] = e.y + __popc(e.x & ((1u << (key[i] & 31u)) - 1u));
        }
    }
}

// Run statistics of the index stream, collected by the replay kernels while the indices are in registers, so
// that the RLE stage that follows in the composed helpers needs no pass of its own to count its runs
// (rle.cu: rle_encode_tiled).  One record per warp tile (32 consecutive chunks):
//   x = run boundaries inside the tile (positions whose index differs from the one before, the tile's first
//       position not counted), y = 1 + position of the last of them (0 = none), z / w = first / last index.
struct RunStat {
    uint32_t npairs = 0, lhead = 0, first = 0, prev = 0xffffffffu;
    __device__ __forceinline__ void see(uint32_t r, uint32_t pos) { // the chunk's first symbol counts as a boundary for now
        const bool ne = r != prev;
        npairs += ne;
        if (ne) lhead = pos + 1;
        if (prev == 0xffffffffu) first = r;
        prev = r;
    }
    // all lanes of the warp; lane t holds chunk t of the tile (inactive lanes hold no chunk)
    __device__ __forceinline__ void finish(bool active, uint32_t beg, uint64_t tile, uint4 *__restrict__ tstat) {
        const unsigned lane = lane_id();
        const uint32_t before = __shfl_up_sync(TC_FULL, prev, 1);
        // a chunk start is a boundary only if the index differs from the last one of the chunk before; for the
        // tile's first chunk that is decided by the scan over the tiles
        if (active && npairs && (lane == 0 || first == before)) {
            npairs--;
            if (lhead == beg + 1) lhead = 0;
        }
        const uint32_t tp = __reduce_add_sync(TC_FULL, active ? npairs : 0u), th = __reduce_max_sync(TC_FULL, active ? lhead : 0u);
        const unsigned act = __ballot_sync(TC_FULL, active);
        const uint32_t tf = __shfl_sync(TC_FULL, first, 0), tl = __shfl_sync(TC_FULL, prev, 31 - __clz(act | 1u));
        if (lane == 0 && act) tstat[tile] = make_uint4(tp, th, tf, tl);
    }
};

struct R3 {
    uint16_t *last;  // per-thread base; entry li is li * CT halfwords further on
    uint32_t *bits;  // per-thread base; consecutive words are CT apart
    uint64_t *sfx;
    uint32_t T;      // 3 x 10 bits: live slots in the groups above group g
    uint32_t ow, e0; // the word being appended to (slots e0 .. e0 + 31), kept in a register
};
// Slots e0 .. e0 + 31 are credited to sfx / T as live when their word is opened ("pre-credit"), so a
// query below the open word reads 32 - (symbols appended so far) too many: the caller subtracts it.
template <int CT>
__device__ __forceinline__ void r3_open_word(R3 &S) {
    const uint32_t we = S.e0 >> 5, ge = we >> 3, ke = we & 7u;
    S.sfx[ge * CT] += 0x0020202020202020ull >> (56u - 8u * ke);
    S.T += 0x02008020u >> (30u - 10u * ge);
    S.ow = 0;
}
// One symbol: last-index li, appended as slot e0 + k (bit = 1 << k, corr = 32 - k).  Returns its MTF index
// = live slots above the slot a of its previous occurrence, and retires a.
template <int CT>
__device__ __forceinline__ uint32_t r3_step(R3 &S, const uint64_t *m8tab, uint32_t a, uint32_t bitk, uint32_t corr) {
    const uint32_t wa = a >> 5, ba = a & 31u, g = a >> 8, kk8 = (a >> 2) & 0x38u, x = 10u * g;
    const uint32_t w = S.bits[wa * CT];
    const uint64_t sv = S.sfx[g * CT];
    const uint64_t m8 = *reinterpret_cast<const uint64_t *>(reinterpret_cast<const char *>(m8tab) + kk8);
    const uint32_t bit = 1u << ba;
    const bool open = a >= S.e0; // in the open word: its bits live in S.ow, the word in memory is still 0
    const uint32_t closed = __popc(shr_clamp(w, ba + 1u)) + ((uint32_t)(sv >> kk8) & 0xffu) + ((S.T >> x) & 0x3ffu) - corr;
    const uint32_t inopen = __popc(shr_clamp(S.ow, ba + 1u));
    S.bits[wa * CT] = w & ~bit;
    S.sfx[g * CT] = sv - m8;
    S.T -= 0x00100401u >> (30u - x);
    S.ow = (S.ow & ~(open ? bit : 0u)) | bitk;
    return open ? inopen : closed;
}

template <class Src, int CT, bool RS>
__global__ void __launch_bounds__(CT)
    mtf3_replay_kernel(Src src, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t sigma,
                       const uint16_t *__restrict__ start, uint16_t *__restrict__ idx_out, uint4 *__restrict__ tstat,
                       uint64_t *__restrict__ toff, uint32_t *__restrict__ theadx, uint32_t *__restrict__ ticket) {
    extern __shared__ __align__(16) uint32_t sm3[];
    __shared__ uint64_t m8tab[8];
    if (threadIdx.x < 8) m8tab[threadIdx.x] = 0x0001010101010101ull >> (56u - 8u * threadIdx.x); // bytes below kk
    __syncthreads();
    const uint64_t k = (uint64_t)blockIdx.x * CT + threadIdx.x;
    RunStat rs;
    const bool active = k < nchunks;
    if (!RS && !active) return;
    if (active) { // (lanes without a chunk still take part in the warp reductions at the end)
    R3 S;
    S.last = reinterpret_cast<uint16_t *>(sm3) + threadIdx.x;
    S.bits = sm3 + R3_LASTW * CT + threadIdx.x;
    S.sfx = reinterpret_cast<uint64_t *>(sm3 + (R3_LASTW + R3_BITW) * CT) + threadIdx.x;
    // incoming order: the chunk's start row
    {
        // 576-byte rows: 8 x 16 bytes in flight per thread
        const uint4 *row = reinterpret_cast<const uint4 *>(start + k * R3_ROW);
#pragma unroll 1
        for (int b = 0; b < 32; b += 8) {
            uint4 v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = __ldg(row + b + q);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t ws[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int li = 8 * (b + q) + 2 * e;
                    S.last[li * CT] = (uint16_t)ws[e];
                    S.last[(li + 1) * CT] = (uint16_t)(ws[e] >> 16);
                }
            }
        }
        S.last[256 * CT] = start[k * R3_ROW + 256];
    }
    // the sigma live slots are R3_VS - sigma .. R3_VS - 1
    {
        const int lo = R3_VS - (int)sigma, hi = R3_VS;
        auto live = [&](int x0, int x1) { // live slots in [x0, x1)
            int a = x0 > lo ? x0 : lo, b = x1 < hi ? x1 : hi;
            return b > a ? b - a : 0;
        };
#pragma unroll 4
        for (int w = 0; w < R3_BITW; w++) {
            int a = 32 * w > lo ? 32 * w : lo, b = 32 * w + 32 < hi ? 32 * w + 32 : hi;
            uint32_t m = 0;
            if (b > a) m = (b - a == 32 ? 0xffffffffu : ((1u << (b - a)) - 1u)) << (a - 32 * w);
            S.bits[w * CT] = m;
        }
        S.T = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint64_t sv = 0;
#pragma unroll
            for (int q = 0; q < 7; q++) sv |= (uint64_t)live(32 * (8 * g + q + 1), 256 * (g + 1)) << (8 * q);
            S.sfx[g * CT] = sv;
            if (g < 3) S.T |= (uint32_t)live(256 * (g + 1), 1024) << (10 * g);
        }
        S.e0 = R3_VS;
        r3_open_word<CT>(S);
    }
    const uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    uint64_t pos = beg;
    const bool vec = src.can_vec(beg) && (reinterpret_cast<uintptr_t>(idx_out + beg) & 15) == 0;
    // the first word was opened above; every later block files the full word and opens the next one first
    // (never past the last block: slot 1024 does not exist)
    bool fresh = true;
    auto next_word = [&]() {
        if (!fresh) {
            S.bits[(S.e0 >> 5) * CT] = S.ow;
            S.e0 += 32;
            r3_open_word<CT>(S);
        }
        fresh = false;
    };
    if (vec) {
        const uint64_t vend = beg + ((end - beg) & ~31ull); // whole words of 32 symbols
        typename Src::Raw nxt;
        if (pos < vend) nxt = src.load_raw(pos);
        for (; pos < vend; pos += 32) {
            next_word();
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                uint32_t li[16];
                const typename Src::Raw now = nxt; // the next 16 symbols are on their way while these are replayed
                if (pos + 16 * half + 16 < vend) nxt = src.load_raw(pos + 16 * half + 16);
                src.decode_li(now, pos + 16 * half, li);
                const uint32_t kb = 16u * half, eb = S.e0 + kb;
                uint32_t o[8];
                uint32_t a_next = S.last[li[0] * CT];
                S.last[li[0] * CT] = (uint16_t)eb;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t a = a_next;
                    if (j + 1 < 16) { // the next symbol's slot is fetched before this symbol's bitmap traffic
                        a_next = S.last[li[j + 1] * CT];
                        S.last[li[j + 1] * CT] = (uint16_t)(eb + j + 1);
                    }
                    const uint32_t r = r3_step<CT>(S, m8tab, a, (1u << j) << kb, 32u - j - kb);
                    o[j >> 1] = (j & 1) ? (o[j >> 1] | (r << 16)) : r;
                    if (RS) rs.see(r, (uint32_t)pos + kb + j);
                }
                uint4 *dst = reinterpret_cast<uint4 *>(idx_out + pos + 16 * half);
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    }
    // tail of the last chunk, and chunks whose buffers are not 16-byte aligned
    for (uint32_t j = 0; pos < end; pos++) {
        if (j == 0) next_word();
        const uint32_t li = r3_lidx((uint32_t)src.at(pos));
        const uint32_t a = S.last[li * CT];
        S.last[li * CT] = (uint16_t)(S.e0 + j);
        const uint32_t r = r3_step<CT>(S, m8tab, a, 1u << j, 32u - j);
        idx_out[pos] = (uint16_t)r;
        if (RS) rs.see(r, (uint32_t)pos);
        j = (j + 1) & 31u;
    }
    } // active
    if (RS) {
        rs.finish(active, (uint32_t)(k * L), k >> 5, tstat);
        // the CTA that finishes last scans the tile records for the RLE stage (no kernel of its own for that)
        if (last_cta_done(ticket, gridDim.x)) runstat_scan<CT>(tstat, (nchunks + 31) / 32, 32 * L, toff, theadx);
    }
}

// ---- encode, small alphabets (sigma <= 8: ACGT(N) + sentinel) -------------------------------------
// The whole list fits one register: nibble j = rank at list position j.  A thread replays a chunk
// of 64 symbols with ~14 integer instructions per symbol (find the nibble, rotate the ones below
// it), so the pass is bound by HBM, not by issue.  Chunk summaries are recency lists (the distinct
// ranks of the chunk, most recent first) + the set of ranks seen;  later . earlier  =  later's list
// followed by earlier's entries that later has not seen -- associative, so incoming lists come from
// an exclusive scan: inside the CTA by shuffles and shared memory (K1), over CTAs by one CTA (K2).
constexpr int SM_L = 64;            // symbols per thread
constexpr int SM_T = 256;           // threads per CTA
struct Summ {
    uint32_t list, mask;
};
__device__ __forceinline__ uint32_t nib_mtf(uint32_t &L, uint32_t s) {
    uint32_t x = L ^ (s * 0x11111111u);
    uint32_t t = (x - 0x11111111u) & ~x & 0x88888888u; // top bit of every zero nibble (lowest one exact)
    uint32_t p4 = (uint32_t)__ffs((int)t) - 4u;        // bit offset of the nibble holding s
    uint32_t low = (1u << p4) - 1u;
    L = (L & ~((low << 4) | 0xFu)) | ((L & low) << 4) | s;
    return p4 >> 2;
}
__device__ __forceinline__ Summ summ_compose(Summ earlier, Summ later) {
    uint32_t pos = 4u * (uint32_t)__popc(later.mask);
    uint32_t out = pos >= 32 ? later.list : (later.list & ((1u << pos) - 1u));
    const uint32_t ke = (uint32_t)__popc(earlier.mask);
    for (uint32_t j = 0; j < ke; j++) {
        uint32_t s = (earlier.list >> (4 * j)) & 15u;
        if (!((later.mask >> s) & 1u)) {
            out |= s << pos;
            pos += 4;
        }
    }
    return Summ{out, earlier.mask | later.mask};
}
__device__ __forceinline__ Summ summ_shfl_up(Summ v, int d) {
    return Summ{__shfl_up_sync(TC_FULL, v.list, d), __shfl_up_sync(TC_FULL, v.mask, d)};
}

template <class Src>
__device__ __forceinline__ void sm_load_ranks(const Src &src, const uint8_t *s_rank, uint64_t base, uint64_t N,
                                              int q, uint32_t *r /*16*/) {
    int c[16];
    uint64_t b0 = base + (uint64_t)q * 16;
    if (b0 + 16 <= N && src.can_vec(b0)) {
        src.load_vec(b0, c);
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = s_rank[c[k]];
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = b0 + k < N ? s_rank[src.at(b0 + k)] : 0xffu;
    }
}

// K1: chunk summaries, exclusive prefix inside the CTA -> part[chunk], CTA total -> tot[cta]
template <class Src>
__global__ void __launch_bounds__(SM_T)
    mtfs_summary_kernel(Src src, Lut lut, uint64_t N, Summ *__restrict__ part, Summ *__restrict__ tot) {
    __shared__ uint8_t s_rank[SIGMAX + 7];
    __shared__ Summ wsum[SM_T / 32];
    for (int j = threadIdx.x; j < SIGMAX; j += SM_T) s_rank[j] = (uint8_t)lut.rank[j];
    __syncthreads();
    const uint64_t chunk = (uint64_t)blockIdx.x * SM_T + threadIdx.x;
    const uint64_t base = chunk * SM_L;
    // recency list of the chunk = ranks in the order they are first met walking backwards
    uint32_t L = 0, mask = 0, pos = 0;
    if (base < N) {
#pragma unroll
        for (int q = SM_L / 16 - 1; q >= 0; q--) {
            uint32_t r[16];
            sm_load_ranks(src, s_rank, base, N, q, r);
#pragma unroll
            for (int k = 15; k >= 0; k--) {
                if (r[k] < 8u && !((mask >> r[k]) & 1u)) {
                    mask |= 1u << r[k];
                    L |= r[k] << pos;
                    pos += 4;
                }
            }
        }
    }
    // inclusive scan over the warp, then over the CTA's warps
    Summ v{L, mask};
    const unsigned lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Summ o = summ_shfl_up(v, d);
        if (lane >= (unsigned)d) v = summ_compose(o, v);
    }
    const int w = threadIdx.x >> 5;
    if (lane == 31) wsum[w] = v;
    Summ ex = summ_shfl_up(v, 1);
    if (lane == 0) ex = Summ{0, 0};
    __syncthreads();
    Summ pre{0, 0};
    for (int ww = 0; ww < w; ww++) pre = summ_compose(pre, wsum[ww]);
    if (base < N) part[chunk] = summ_compose(pre, ex);
    if (threadIdx.x == SM_T - 1) tot[blockIdx.x] = summ_compose(pre, v);
}

// K2: one CTA; full list at the start of every CTA tile (seed = identity = sorted alphabet) and
// the final list of the whole input
__global__ void __launch_bounds__(1024)
    mtfs_top_kernel(const Summ *__restrict__ tot, uint64_t ntiles, uint32_t sigma, uint32_t *__restrict__ start_list,
                    uint16_t *__restrict__ final_list) {
    __shared__ Summ wsum[32];
    __shared__ Summ carry_s;
    const unsigned lane = lane_id();
    const int w = threadIdx.x >> 5;
    const uint32_t full = sigma >= 32 ? 0xffffffffu : ((1u << sigma) - 1u);
    Summ carry{0x76543210u, full};
    for (uint64_t b0 = 0; b0 < ntiles; b0 += 1024) {
        uint64_t t = b0 + threadIdx.x;
        Summ v = t < ntiles ? tot[t] : Summ{0, 0};
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Summ o = summ_shfl_up(v, d);
            if (lane >= (unsigned)d) v = summ_compose(o, v);
        }
        if (lane == 31) wsum[w] = v;
        Summ ex = summ_shfl_up(v, 1);
        if (lane == 0) ex = Summ{0, 0};
        __syncthreads();
        if (w == 0) { // exclusive scan of the 32 warp totals, seeded with the carry
            Summ x = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                Summ o = summ_shfl_up(x, d);
                if (lane >= (unsigned)d) x = summ_compose(o, x);
            }
            Summ xe = summ_shfl_up(x, 1);
            if (lane == 0) xe = Summ{0, 0};
            wsum[lane] = summ_compose(carry, xe);
            if (lane == 31) carry_s = summ_compose(carry, x);
        }
        __syncthreads();
        if (t < ntiles) start_list[t] = summ_compose(wsum[w], ex).list;
        carry = carry_s;
        __syncthreads();
    }
    if (threadIdx.x < sigma) final_list[threadIdx.x] = (uint16_t)((carry.list >> (4 * threadIdx.x)) & 15u);
}

// K3: replay every chunk from its incoming list
template <class Src>
__global__ void __launch_bounds__(SM_T)
    mtfs_replay_kernel(Src src, Lut lut, uint64_t N, const Summ *__restrict__ part,
                       const uint32_t *__restrict__ start_list, uint32_t sigma, uint16_t *__restrict__ idx_out) {
    __shared__ uint8_t s_rank[SIGMAX + 7];
    for (int j = threadIdx.x; j < SIGMAX; j += SM_T) s_rank[j] = (uint8_t)lut.rank[j];
    __syncthreads();
    const uint64_t chunk = (uint64_t)blockIdx.x * SM_T + threadIdx.x;
    const uint64_t base = chunk * SM_L;
    if (base >= N) return;
    const uint32_t full = (1u << sigma) - 1u;
    uint32_t L = summ_compose(Summ{start_list[blockIdx.x], full}, part[chunk]).list;
#pragma unroll
    for (int q = 0; q < SM_L / 16; q++) {
        uint32_t r[16];
        sm_load_ranks(src, s_rank, base, N, q, r);
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t x = r[k] < 8u ? nib_mtf(L, r[k]) : 0u;
            o[k >> 1] = (k & 1) ? (o[k >> 1] | (x << 16)) : x;
        }
        const uint64_t b0 = base + (uint64_t)q * 16;
        if (b0 + 16 <= N && (reinterpret_cast<uintptr_t>(idx_out + b0) & 15) == 0) {
            uint4 *dst = reinterpret_cast<uint4 *>(idx_out + b0);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (b0 + k < N) idx_out[b0 + k] = (uint16_t)((o[k >> 1] >> (16 * (k & 1))) & 0xffffu);
        }
    }
}

// ---- encode, sigma <= 6 (ACGT / ACGTN + sentinel): a permutation automaton in shared memory --------------
// With at most 6 symbols the whole MTF list is one of 720 permutations, so a symbol costs one table lookup
// instead of the ~30 register instructions of the nibble list above:
//   entry[state][column] = (5 * next state) << 3 | MTF index,  state = Lehmer rank of the list
// Rows are 5 words apart (4 words of entries + 1 of padding) so that the 32 lanes' states fall into all 32 banks
// (16-byte rows use 8 classes of 4 banks: measured ~7 wavefronts per lookup), and the column of a symbol is a
// multiplicative hash of its byte found on the host to be injective on the alphabet -- no rank table lookup:
// the kernels are bound by shared-memory instruction issue (~8 cycles per LDS measured), not by arithmetic.
// The summary pass runs the same automaton from the identity and collects the set of columns seen: the
// chunk's recency list is the front of the list it ends with.  The table is rebuilt on the host when the
// alphabet changes (720 x 8 entries).
// Threads own chunks of AU_L consecutive symbols, so a warp's loads and stores would touch 32 different
// 128-byte lines per instruction (2 cycles each in L1TEX: measured 22 us of a 35 us kernel).  Instead the warp
// copies its 32 chunks (4 KB contiguous) through padded shared-memory rows: whole lines in, whole lines out.
constexpr int AU_L = 128; // symbols per thread
constexpr int AU_T = 256; // threads per CTA
constexpr int AU_MAXSIG = 6;
constexpr int AU_ROW = AU_L + 16; // bytes per staged row: consecutive lanes' 16-byte pieces fall into different banks

__host__ __device__ inline uint32_t perm_rank(uint32_t list, uint32_t sigma) { // Lehmer rank of the nibble list
    uint32_t r = 0;
    for (uint32_t i = 0; i < sigma; i++) {
        const uint32_t a = (list >> (4 * i)) & 15u;
        uint32_t c = 0;
        for (uint32_t j = i + 1; j < sigma; j++) c += ((list >> (4 * j)) & 15u) < a;
        r = r * (sigma - i) + c;
    }
    return r;
}

struct AutoHash { // column of a symbol: ((li * mul) >> shift) & 14 is twice its column
    uint32_t mul, shift;
    __host__ __device__ uint32_t col2(uint32_t li) const { return ((li * mul) >> shift) & 14u; }
};
constexpr int AU_RW = 10; // u16 entries per table row (8 + padding): odd number of words
struct AutoSmem { // the dynamic shared memory of both kernels
    uint16_t perm[720 * AU_RW];
    uint32_t list[720];               // state -> nibble list
    uint8_t rows[AU_T / 32][32 * AU_ROW];
};
__device__ __forceinline__ void auto_load_tables(AutoSmem &A, const uint16_t *__restrict__ g_perm,
                                                 const uint32_t *__restrict__ g_list, uint32_t n_perm) {
    // rows of 5 words, word-wise: n_perm * 5 words of table + n_perm lists, all loads of a thread in flight
    const uint32_t *g = reinterpret_cast<const uint32_t *>(g_perm);
    uint32_t *d = reinterpret_cast<uint32_t *>(A.perm);
    const uint32_t nw = n_perm * (AU_RW / 2);
    uint32_t tr[15], tl[3];
#pragma unroll
    for (int q = 0; q < 15; q++)
        if (threadIdx.x + AU_T * q < nw) tr[q] = __ldg(g + threadIdx.x + AU_T * q);
#pragma unroll
    for (int q = 0; q < 3; q++)
        if (threadIdx.x + AU_T * q < n_perm) tl[q] = __ldg(g_list + threadIdx.x + AU_T * q);
#pragma unroll
    for (int q = 0; q < 15; q++)
        if (threadIdx.x + AU_T * q < nw) d[threadIdx.x + AU_T * q] = tr[q];
#pragma unroll
    for (int q = 0; q < 3; q++)
        if (threadIdx.x + AU_T * q < n_perm) A.list[threadIdx.x + AU_T * q] = tl[q];
}
// the warp's 32 chunks (32 * AU_L contiguous bytes at g) -> rows; lane l then owns row l
__device__ __forceinline__ void auto_stage_in(const uint8_t *__restrict__ g, uint8_t *rows) {
    const unsigned lane = lane_id();
#pragma unroll
    for (int i = 0; i < AU_L / 16; i++) {
        const uint32_t off = (i * 32 + lane) * 16;
        *reinterpret_cast<uint4 *>(rows + (off / AU_L) * AU_ROW + (off % AU_L)) = ld_stream_u4(g + off);
    }
    __syncwarp();
}

template <class Src>
__global__ void __launch_bounds__(AU_T, 4)
    mtfa_summary_kernel(Src src, AutoHash hs, uint64_t N, const uint16_t *__restrict__ g_perm,
                        const uint32_t *__restrict__ g_list, uint32_t n_perm, Summ *__restrict__ part,
                        Summ *__restrict__ tot, uint32_t *__restrict__ rle_ticket) {
    extern __shared__ __align__(16) uint32_t sma[];
    AutoSmem &A = *reinterpret_cast<AutoSmem *>(sma);
    __shared__ Summ wsum[AU_T / 32];
    if (rle_ticket && blockIdx.x == 0 && threadIdx.x == 0) *rle_ticket = 0; // the replay kernel's last-CTA election
    auto_load_tables(A, g_perm, g_list, n_perm);
    __syncthreads();
    const uint64_t chunk = (uint64_t)blockIdx.x * AU_T + threadIdx.x;
    const uint64_t base = chunk * AU_L;
    const uint64_t wbase = (chunk & ~31ull) * AU_L; // first symbol of the warp's 32 chunks
    const char *tab = reinterpret_cast<const char *>(A.perm);
    uint32_t st = 0, mask = 0; // the identity is permutation 0
    bool done = false;
    if constexpr (sizeof(*src.p) == 1) {
        if (wbase + 32 * AU_L <= N && src.can_vec(wbase)) { // warp-uniform
            uint8_t *rows = A.rows[threadIdx.x >> 5];
            auto_stage_in(reinterpret_cast<const uint8_t *>(src.p) + wbase, rows);
            const uint4 *mine = reinterpret_cast<const uint4 *>(rows + lane_id() * AU_ROW);
#pragma unroll
            for (int q = 0; q < AU_L / 16; q++) {
                uint32_t li[16], r[16];
                src.decode_li(typename Src::Raw{mine[q]}, base + 16 * q, li);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    r[k] = hs.col2(li[k]);
                    mask |= 1u << r[k];
                }
#pragma unroll
                for (int k = 0; k < 16; k++) st = (*reinterpret_cast<const uint16_t *>(tab + st + r[k]) & 0xfff8u) >> 1;
            }
            done = true;
        }
    }
    if (!done && base < N) {
        const uint64_t end = base + AU_L < N ? base + AU_L : N;
        for (uint64_t pos = base; pos < end; pos++) {
            const uint32_t r = hs.col2(r3_lidx((uint32_t)src.at(pos)));
            mask |= 1u << r;
            st = (*reinterpret_cast<const uint16_t *>(tab + st + r) & 0xfff8u) >> 1;
        }
    }
    // the chunk's recency list = the first popc(mask) entries of the list it leaves behind (mask: columns seen)
    Summ v{0, 0};
    {
        const uint32_t k = (uint32_t)__popc(mask);
        const uint32_t l = A.list[st / (2 * AU_RW)];
        v.list = k >= 8 ? l : (l & ((1u << (4 * k)) - 1u));
        for (uint32_t j = 0; j < k; j++) v.mask |= 1u << ((l >> (4 * j)) & 15u);
    }
    // inclusive scan over the warp, then over the CTA's warps
    const unsigned lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Summ o = summ_shfl_up(v, d);
        if (lane >= (unsigned)d) v = summ_compose(o, v);
    }
    const int w = threadIdx.x >> 5;
    if (lane == 31) wsum[w] = v;
    Summ ex = summ_shfl_up(v, 1);
    if (lane == 0) ex = Summ{0, 0};
    __syncthreads();
    Summ pre{0, 0};
    for (int ww = 0; ww < w; ww++) pre = summ_compose(pre, wsum[ww]);
    if (base < N) part[chunk] = summ_compose(pre, ex);
    if (threadIdx.x == AU_T - 1) tot[blockIdx.x] = summ_compose(pre, v);
}

template <class Src, bool RS>
__global__ void __launch_bounds__(AU_T, 4) // 4 CTAs per SM: the 512 tiles of a 16 MiB block are one wave
    mtfa_replay_kernel(Src src, AutoHash hs, uint64_t N, const uint16_t *__restrict__ g_perm,
                       const uint32_t *__restrict__ g_list, uint32_t n_perm, const Summ *__restrict__ part,
                       const uint32_t *__restrict__ start_list, uint32_t sigma, uint16_t *__restrict__ idx_out,
                       uint4 *__restrict__ tstat, uint64_t *__restrict__ toff, uint32_t *__restrict__ theadx,
                       uint32_t *__restrict__ ticket) {
    extern __shared__ __align__(16) uint32_t sma[];
    AutoSmem &A = *reinterpret_cast<AutoSmem *>(sma);
    auto_load_tables(A, g_perm, g_list, n_perm);
    __syncthreads();
    const uint64_t chunk = (uint64_t)blockIdx.x * AU_T + threadIdx.x;
    const uint64_t base = chunk * AU_L;
    const uint64_t wbase = (chunk & ~31ull) * AU_L;
    if (!RS && wbase >= N) return;
    RunStat rs;
    if (wbase < N) { // (with RS every thread reaches the election at the end)
    const char *tab = reinterpret_cast<const char *>(A.perm);
    const uint32_t full = (1u << sigma) - 1u;
    uint32_t st = 0;
    if (base < N) st = perm_rank(summ_compose(Summ{start_list[blockIdx.x], full}, part[chunk]).list, sigma) * (2 * AU_RW);
    bool done = false;
    if constexpr (sizeof(*src.p) == 1) {
        if (wbase + 32 * AU_L <= N && src.can_vec(wbase) && (reinterpret_cast<uintptr_t>(idx_out + wbase) & 15) == 0) {
            const unsigned lane = lane_id();
            uint8_t *rows = A.rows[threadIdx.x >> 5];
            auto_stage_in(reinterpret_cast<const uint8_t *>(src.p) + wbase, rows);
            uint4 raw[AU_L / 16];
#pragma unroll
            for (int q = 0; q < AU_L / 16; q++) raw[q] = reinterpret_cast<const uint4 *>(rows + lane * AU_ROW)[q];
            __syncwarp(); // the rows are free: they take the indices, 64 symbols (128 bytes) per lane at a time
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int q = 0; q < AU_L / 32; q++) {
                    uint32_t li[16], r[16], o[8];
                    src.decode_li(typename Src::Raw{raw[h * (AU_L / 32) + q]}, base + 16 * (h * (AU_L / 32) + q), li);
#pragma unroll
                    for (int k = 0; k < 16; k++) r[k] = hs.col2(li[k]);
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const uint32_t e = *reinterpret_cast<const uint16_t *>(tab + st + r[k]);
                        const uint32_t x = e & 7u;
                        st = (e & 0xfff8u) >> 1;
                        o[k >> 1] = (k & 1) ? (o[k >> 1] | (x << 16)) : x;
                        if (RS) rs.see(x, (uint32_t)base + 16 * (h * (AU_L / 32) + q) + k);
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(rows + lane * AU_ROW + 32 * q);
                    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
                __syncwarp();
                // 32 rows x 128 bytes of indices -> the 64 symbols at offset 64 h of every chunk (128 B at stride 256 B)
                uint8_t *gout = reinterpret_cast<uint8_t *>(idx_out + wbase) + 2 * (AU_L / 2) * h;
#pragma unroll
                for (int i = 0; i < AU_L / 16; i++) {
                    const uint32_t off = (i * 32 + lane) * 16; // byte offset in the 32 x 128-byte block
                    const uint32_t row = off / AU_L, within = off % AU_L;
                    st_stream_u4(gout + (size_t)row * (2 * AU_L) + within, *reinterpret_cast<const uint4 *>(rows + row * AU_ROW + within));
                }
                __syncwarp();
            }
            done = true;
        }
    }
    if (!done && base < N) {
        const uint64_t end = base + AU_L < N ? base + AU_L : N;
        for (uint64_t pos = base; pos < end; pos++) {
            const uint32_t e = *reinterpret_cast<const uint16_t *>(tab + st + hs.col2(r3_lidx((uint32_t)src.at(pos))));
            idx_out[pos] = (uint16_t)(e & 7u);
            st = (e & 0xfff8u) >> 1;
            if (RS) rs.see(e & 7u, (uint32_t)pos);
        }
    }
    } // wbase < N
    if (RS) {
        rs.finish(base < N, (uint32_t)base, chunk >> 5, tstat);
        if (last_cta_done(ticket, gridDim.x)) runstat_scan<AU_T>(tstat, (N + 32 * AU_L - 1) / (32 * AU_L), 32 * AU_L, toff, theadx);
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
        list_take16(lst, r);
    }
    for (int j = 0; j < (int)sigma; j++) perm[k * sigma + j] = lst.get16(j);
}

// ---- decode, small alphabets (sigma <= 8) -------------------------------------------------------
// Mirror of the small-alphabet encoder: the list of *initial-list positions* is one register
// (nibble j = which entry of the sorted alphabet sits at list position j).  A chunk's effect is
// the permutation it leaves behind when run on the identity; chunk A then chunk B composes as
// (A . B)[j] = A[B[j]] (a nibble gather), so the incoming arrangement of every chunk is an exclusive
// scan: inside the CTA by shuffles + shared memory, over CTAs by one CTA.
__device__ __forceinline__ uint32_t nib_take(uint32_t &P, uint32_t r) { // entry at position r moves to the front
    const uint32_t sh = 4u * r;
    const uint32_t v = (P >> sh) & 15u;
    const uint32_t low = (1u << sh) - 1u;
    P = (P & ~((low << 4) | 0xFu)) | ((P & low) << 4) | v;
    return v;
}
__device__ __forceinline__ uint32_t nib_compose(uint32_t A, uint32_t B) { // arrangement A, then permutation B
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) out |= ((A >> (4u * ((B >> (4 * j)) & 15u))) & 15u) << (4 * j);
    return out;
}
__device__ __forceinline__ void smd_load(const uint16_t *__restrict__ idx, uint64_t base, uint64_t N, int q,
                                         uint32_t *r /*16*/) {
    const uint64_t b0 = base + (uint64_t)q * 16;
    if (b0 + 16 <= N && (reinterpret_cast<uintptr_t>(idx + b0) & 15) == 0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint4 v = *reinterpret_cast<const uint4 *>(idx + b0 + 8 * h);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 8; k++) r[8 * h + k] = (w[k >> 1] >> (16 * (k & 1))) & 0xffffu;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) r[k] = b0 + k < N ? idx[b0 + k] : 0xffffffffu;
    }
}

__global__ void __launch_bounds__(SM_T)
    mtfds_perm_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t sigma, uint32_t *__restrict__ part,
                      uint32_t *__restrict__ tot, uint32_t *__restrict__ err) {
    __shared__ uint32_t wsum[SM_T / 32];
    const uint64_t chunk = (uint64_t)blockIdx.x * SM_T + threadIdx.x;
    const uint64_t base = chunk * SM_L;
    uint32_t P = 0x76543210u;
    bool bad = false;
    if (base < N) {
#pragma unroll
        for (int q = 0; q < SM_L / 16; q++) {
            uint32_t r[16];
            smd_load(idx, base, N, q, r);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (r[k] != 0xffffffffu) {
                    uint32_t x = r[k];
                    if (x >= sigma) {
                        bad = true;
                        x = 0;
                    }
                    nib_take(P, x);
                }
            }
        }
    }
    if (bad) atomicMax(err, 1u);
    const unsigned lane = lane_id();
    uint32_t v = P;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(TC_FULL, v, d);
        if (lane >= (unsigned)d) v = nib_compose(o, v);
    }
    const int w = threadIdx.x >> 5;
    if (lane == 31) wsum[w] = v;
    uint32_t ex = __shfl_up_sync(TC_FULL, v, 1);
    if (lane == 0) ex = 0x76543210u;
    __syncthreads();
    uint32_t pre = 0x76543210u;
    for (int ww = 0; ww < w; ww++) pre = nib_compose(pre, wsum[ww]);
    if (base < N) part[chunk] = nib_compose(pre, ex);
    if (threadIdx.x == SM_T - 1) tot[blockIdx.x] = nib_compose(pre, v);
}

__global__ void __launch_bounds__(1024)
    mtfds_top_kernel(const uint32_t *__restrict__ tot, uint64_t ntiles, uint32_t *__restrict__ start) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const unsigned lane = lane_id();
    const int w = threadIdx.x >> 5;
    uint32_t carry = 0x76543210u;
    for (uint64_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint64_t t = b0 + threadIdx.x;
        uint32_t v = t < ntiles ? tot[t] : 0x76543210u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(TC_FULL, v, d);
            if (lane >= (unsigned)d) v = nib_compose(o, v);
        }
        if (lane == 31) wsum[w] = v;
        uint32_t ex = __shfl_up_sync(TC_FULL, v, 1);
        if (lane == 0) ex = 0x76543210u;
        __syncthreads();
        if (w == 0) {
            uint32_t x = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(TC_FULL, x, d);
                if (lane >= (unsigned)d) x = nib_compose(o, x);
            }
            uint32_t xe = __shfl_up_sync(TC_FULL, x, 1);
            if (lane == 0) xe = 0x76543210u;
            wsum[lane] = nib_compose(carry, xe);
            if (lane == 31) carry_s = nib_compose(carry, x);
        }
        __syncthreads();
        if (t < ntiles) start[t] = nib_compose(wsum[w], ex);
        carry = carry_s;
        __syncthreads();
    }
}

struct List8 {
    int16_t sym[8];
};
__global__ void __launch_bounds__(SM_T)
    mtfds_replay_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t sigma, const uint32_t *__restrict__ part,
                        const uint32_t *__restrict__ start, List8 l0, int16_t *__restrict__ out) {
    const uint64_t chunk = (uint64_t)blockIdx.x * SM_T + threadIdx.x;
    const uint64_t base = chunk * SM_L;
    if (base >= N) return;
    uint32_t P = nib_compose(start[blockIdx.x], part[chunk]);
    // the eight possible symbols, two per register: sym(v) = (pk[v >> 1] >> (16 * (v & 1))) & 0xffff
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; j++) pk[j] = (uint32_t)(uint16_t)l0.sym[2 * j] | ((uint32_t)(uint16_t)l0.sym[2 * j + 1] << 16);
#pragma unroll
    for (int q = 0; q < SM_L / 16; q++) {
        uint32_t r[16];
        smd_load(idx, base, N, q, r);
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t x = r[k] < sigma ? r[k] : 0u; // out-of-range indices were reported by the first pass
            uint32_t v = r[k] != 0xffffffffu ? nib_take(P, x) : 0u;
            uint32_t w = (v & 2u) ? ((v & 4u) ? pk[3] : pk[1]) : ((v & 4u) ? pk[2] : pk[0]);
            uint32_t s = (v & 1u) ? (w >> 16) : (w & 0xffffu);
            o[k >> 1] = (k & 1) ? (o[k >> 1] | (s << 16)) : s;
        }
        const uint64_t b0 = base + (uint64_t)q * 16;
        if (b0 + 16 <= N && (reinterpret_cast<uintptr_t>(out + b0) & 15) == 0) {
            uint4 *dst = reinterpret_cast<uint4 *>(out + b0);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (b0 + k < N) out[b0 + k] = (int16_t)((o[k >> 1] >> (16 * (k & 1))) & 0xffffu);
        }
    }
}

// D2a: exclusive chain of permutations inside a tile: acc'[j] = acc[perm[j]].  Rows of `perm` / `part` are
// `in_stride` / `out_stride` entries apart.  Rows are loaded four steps ahead (a ring of registers): the chain itself
// runs through shared memory only and does not wait for global loads.
constexpr int CH_PER = (SIGMAX + 31) / 32;
constexpr int CH_RING = 4;
__device__ __forceinline__ void chain_load(uint16_t (&dst)[CH_PER], const uint16_t *__restrict__ row, bool ok, uint32_t sigma,
                                           unsigned lane) {
#pragma unroll
    for (int q = 0; q < CH_PER; q++) {
        const uint32_t j = lane + 32u * q;
        dst[q] = (ok && j < sigma) ? row[j] : (uint16_t)0;
    }
}
__global__ void __launch_bounds__(128)
    mtfd_tile_chain_kernel(const uint16_t *__restrict__ perm, uint64_t nchunks, uint32_t G, uint32_t sigma,
                           uint16_t *__restrict__ part, uint16_t *__restrict__ tilesum, uint64_t ntiles,
                           uint32_t in_stride, uint32_t out_stride) {
    __shared__ uint16_t bufs[4][2][LISTPAD];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    uint64_t t = (uint64_t)blockIdx.x * 4 + w;
    if (t >= ntiles) return;
    uint16_t *A = bufs[w][0], *B = bufs[w][1];
    for (int j = lane; j < (int)sigma; j += 32) A[j] = (uint16_t)j;
    __syncwarp();
    const uint64_t k0 = t * G, k1 = k0 + G < nchunks ? k0 + G : nchunks;
    uint16_t ring[CH_RING][CH_PER];
#pragma unroll
    for (int u = 0; u < CH_RING; u++) chain_load(ring[u], perm + (k0 + u) * in_stride, k0 + u < k1, sigma, lane);
    for (uint64_t kb = k0; kb < k1; kb += CH_RING) {
#pragma unroll
        for (int u = 0; u < CH_RING; u++) {
            const uint64_t k = kb + u;
            if (k < k1) { // uniform over the warp
#pragma unroll
                for (int q = 0; q < CH_PER; q++) {
                    const uint32_t j = lane + 32u * q;
                    if (j < sigma) {
                        part[k * out_stride + j] = A[j];
                        B[j] = A[ring[u][q]];
                    }
                }
                __syncwarp();
                uint16_t *tmp = A;
                A = B;
                B = tmp;
                chain_load(ring[u], perm + (k + CH_RING) * in_stride, k + CH_RING < k1, sigma, lane);
            }
        }
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
    uint16_t ring[CH_RING][CH_PER];
#pragma unroll
    for (int u = 0; u < CH_RING; u++) chain_load(ring[u], tilesum + (uint64_t)u * sigma, (uint64_t)u < ntiles, sigma, lane);
    for (uint64_t tb = 0; tb < ntiles; tb += CH_RING) {
#pragma unroll
        for (int u = 0; u < CH_RING; u++) {
            const uint64_t t = tb + u;
            if (t < ntiles) {
#pragma unroll
                for (int q = 0; q < CH_PER; q++) {
                    const uint32_t j = lane + 32u * q;
                    if (j < sigma) {
                        tileprefix[t * sigma + j] = A[j];
                        B[j] = A[ring[u][q]];
                    }
                }
                __syncwarp();
                int16_t *tmp = A;
                A = B;
                B = tmp;
                chain_load(ring[u], tilesum + (t + CH_RING) * sigma, t + CH_RING < ntiles, sigma, lane);
            }
        }
    }
}

// D3: replay with the real symbols.
__global__ void mtfd_replay_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t L, uint64_t nchunks,
                                   uint32_t G, uint32_t sigma, const uint16_t *__restrict__ part,
                                   const uint16_t *__restrict__ part2, const int16_t *__restrict__ superprefix,
                                   int16_t *__restrict__ out) {
    extern __shared__ uint32_t smem[];
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    TState lst{smem};
    // incoming list = (list at the super-tile start) . (tiles before this one) . (chunks before this one)
    const int16_t *sp = superprefix + (k / G / G) * sigma;
    const uint16_t *p2 = part2 + (k / G) * sigma;
    const uint16_t *pp = part + k * sigma;
    for (int j = 0; j < (int)sigma; j++) lst.set16(j, (uint16_t)sp[p2[pp[j]]]);
    uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    for (uint64_t i = beg; i < end; i++) {
        uint32_t r = idx[i];
        if (r >= sigma) r = 0;
        out[i] = (int16_t)list_take16(lst, r);
    }
}

// ---- decode, sigma <= 257, index-independent cost: the select step of mtfd_select.cuh, thread per chunk ---------
// The old kernels above shift a per-thread list, which costs the index value (mean 128 on high-entropy input: 425 M
// warp instructions per pass over 16 Mi symbols); these cost ~90 instructions per symbol whatever the index.
// State per thread, word-interleaved across the CTA (conflict-free for any per-thread index): 128 words of entry bytes
// (one per slot) + 16 bitmap words.  A CTA of D3_CT threads is one tile of the composition chains (G = D3_CT).
constexpr int D3_CT = 160;
constexpr int D3_SYMW = d3::SLOTS / 4;
constexpr int D3_WORDS = D3_SYMW + d3::WORDS; // 144 words = 576 B per thread
constexpr int D3_STRIDE = 264;                // u16 entries per permutation row in global memory (16-byte multiple)
constexpr int D3_G = 32;                      // chunks per tile of the composition chains (divides D3_CT)
struct D3Smem {
    uint32_t st[D3_WORDS * D3_CT];
    int16_t q[D3_CT / D3_G][D3_STRIDE]; // replay: list at the start of each of the CTA's tiles
    uint8_t sel8[2048];
};
struct D3State {
    uint32_t *st; // &smem.st[threadIdx.x]
    const uint8_t *lut;
    __device__ __forceinline__ uint32_t bm_load(uint32_t w) const { return st[(D3_SYMW + w) * D3_CT]; }
    __device__ __forceinline__ void bm_store(uint32_t w, uint32_t x) { st[(D3_SYMW + w) * D3_CT] = x; }
    __device__ __forceinline__ uint32_t sym_load(uint32_t s) const {
        return reinterpret_cast<const uint8_t *>(&st[(s >> 2) * D3_CT])[s & 3u];
    }
    __device__ __forceinline__ void sym_store(uint32_t s, uint32_t v) {
        reinterpret_cast<uint8_t *>(&st[(s >> 2) * D3_CT])[s & 3u] = (uint8_t)v;
    }
    __device__ __forceinline__ uint32_t sel8(uint32_t i) const { return lut[i]; }
};
__device__ __forceinline__ void d3_fill_lut(uint8_t *lut) {
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) lut[i] = d3::sel8_entry(i >> 3, i & 7u);
}
// Walks the chunk's indices eight at a time (128-bit loads when the stream is 16-byte aligned), one load ahead:
// f(i, index, q) with q = 0..7 the position inside a full group of eight (a compile-time constant after unrolling)
// or q = 8 for positions outside full groups (chunk tail, unaligned streams).
template <class F>
__device__ __forceinline__ void d3_walk(const uint16_t *__restrict__ idx, uint64_t beg, uint64_t end, bool wide, F f) {
    uint64_t i = beg;
    if (wide) {
        uint4 nx = make_uint4(0, 0, 0, 0);
        if (i + 8 <= end) nx = *reinterpret_cast<const uint4 *>(idx + i);
        while (i + 8 <= end) {
            const uint4 v = nx;
            if (i + 16 <= end) nx = *reinterpret_cast<const uint4 *>(idx + i + 8);
            f(i + 0, v.x & 0xffffu, 0);
            f(i + 1, v.x >> 16, 1);
            f(i + 2, v.y & 0xffffu, 2);
            f(i + 3, v.y >> 16, 3);
            f(i + 4, v.z & 0xffffu, 4);
            f(i + 5, v.z >> 16, 5);
            f(i + 6, v.w & 0xffffu, 6);
            f(i + 7, v.w >> 16, 7);
            i += 8;
        }
    }
#pragma unroll 1
    for (; i < end; i++) f(i, (uint32_t)idx[i], 8);
}

// D1': the permutation each chunk applies to list positions (select replay on the identity list).
__global__ void __launch_bounds__(D3_CT, 2)
    mtfd3_perm_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t sigma,
                      uint16_t *__restrict__ perm, uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) unsigned char d3_raw[];
    D3Smem &M = *reinterpret_cast<D3Smem *>(d3_raw);
    d3_fill_lut(M.sel8);
    __syncthreads();
    const uint64_t k = (uint64_t)blockIdx.x * D3_CT + threadIdx.x;
    if (k >= nchunks) return;
    D3State S{&M.st[threadIdx.x], M.sel8};
    d3::Regs R;
    d3::init(S, R, sigma);
    // entry of slot s = its initial list position sigma - 1 - s; position 256 (sigma == 257, slot 0) is the special one
    for (uint32_t w4 = 0; 4 * w4 < sigma; w4++) {
        uint32_t word = 0;
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) word |= ((sigma - 1u - (4u * w4 + q)) & 0xffu) << (8 * q);
        S.st[w4 * D3_CT] = word;
    }
    if (sigma == 257) R.special = 0;
    const uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    bool bad = false;
    d3_walk(idx, beg, end, (reinterpret_cast<uintptr_t>(idx) & 15) == 0, [&](uint64_t, uint32_t r, int) {
        if (r >= sigma) bad = true, r = 0;
        d3::take(S, R, r);
    });
    if (bad) atomicMax(err, 1u);
    // the list front to back -> one row of 16-bit entries, eight per store (what lies behind entry sigma - 1 is row
    // padding, never read)
    uint16_t *row = perm + k * D3_STRIDE;
    int w = d3::WORDS;
    uint32_t x = 0;
    auto next = [&]() -> uint32_t {
        while (!x && w > 0) x = S.bm_load((uint32_t)--w);
        if (!x) return 0u;
        const uint32_t bit = 31u - (uint32_t)__clz((int)x);
        x ^= 1u << bit;
        const uint32_t slot = 32u * (uint32_t)w + bit;
        return slot == R.special ? 256u : S.sym_load(slot);
    };
    for (uint32_t j0 = 0; j0 < sigma; j0 += 8) {
        uint32_t e[8];
#pragma unroll
        for (int q = 0; q < 8; q++) e[q] = next();
        *reinterpret_cast<uint4 *>(row + j0) =
            make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
    }
}

// D3': replay with the real symbols.  The list at the start of the CTA's tile (super-tile prefix . tiles before) is
// composed once per CTA into shared memory; a thread's incoming list is that list permuted by its row of `part`.
__global__ void __launch_bounds__(D3_CT, 2)
    mtfd3_replay_kernel(const uint16_t *__restrict__ idx, uint64_t N, uint32_t L, uint64_t nchunks, uint32_t G,
                        uint32_t sigma, const uint16_t *__restrict__ part, const uint16_t *__restrict__ part2,
                        const int16_t *__restrict__ superprefix, int16_t *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char d3_raw[];
    D3Smem &M = *reinterpret_cast<D3Smem *>(d3_raw);
    d3_fill_lut(M.sel8);
    const uint64_t ntiles = (nchunks + G - 1) / G;
    for (uint32_t tl = 0; tl < (uint32_t)(D3_CT / D3_G); tl++) { // G == D3_G: the CTA spans D3_CT / D3_G tiles
        const uint64_t tile = (uint64_t)blockIdx.x * (D3_CT / D3_G) + tl;
        if (tile >= ntiles) break;
        const int16_t *sp = superprefix + (tile / G) * sigma;
        const uint16_t *p2 = part2 + tile * sigma;
        for (uint32_t j = threadIdx.x; j < sigma; j += D3_CT) M.q[tl][j] = sp[p2[j]];
    }
    __syncthreads();
    const uint64_t k = (uint64_t)blockIdx.x * D3_CT + threadIdx.x;
    if (k >= nchunks) return;
    D3State S{&M.st[threadIdx.x], M.sel8};
    d3::Regs R;
    d3::init(S, R, sigma);
    const uint16_t *pp = part + k * D3_STRIDE;
    for (uint32_t j0 = 0; j0 < sigma; j0 += 8) {
        const uint4 v = *reinterpret_cast<const uint4 *>(pp + j0);
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (uint32_t q = 0; q < 8; q++) {
            const uint32_t j = j0 + q;
            if (j < sigma) {
                const int e = M.q[threadIdx.x / D3_G][(wv[q >> 1] >> (16 * (q & 1))) & 0xffffu];
                const uint32_t slot = sigma - 1u - j;
                if (e < 0) R.special = slot;
                S.sym_store(slot, (uint32_t)e & 0xffu);
            }
        }
    }
    const uint64_t beg = k * L, end = beg + L < N ? beg + L : N;
    const bool wide = ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    d3_walk(idx, beg, end, wide, [&](uint64_t i, uint32_t r, int q) {
        if (r >= sigma) r = 0;
        const uint32_t id = d3::take(S, R, r);
        const uint32_t sym16 = id == 256u ? 0xffffu : id; // Nothing = -1
        switch (q) { // q is a constant at every call site
        case 0: o0 = sym16; break;
        case 1: o0 |= sym16 << 16; break;
        case 2: o1 = sym16; break;
        case 3: o1 |= sym16 << 16; break;
        case 4: o2 = sym16; break;
        case 5: o2 |= sym16 << 16; break;
        case 6: o3 = sym16; break;
        case 7:
            o3 |= sym16 << 16;
            *reinterpret_cast<uint4 *>(out + (i - 7u)) = make_uint4(o0, o1, o2, o3);
            break;
        default: out[i] = (int16_t)sym16; break;
        }
    });
}

// The final list (ranks) comes back through the pinned scalars.  In the composed helpers
// (`defer`) the host does not wait here: the RLE stage that follows syncs the stream anyway and
// mtf_finish_pending() then translates the ranks.  Slot 512.. of h_scal is used by nothing else.
// When the RLE stage shares its result words with the list (MtfRleLink::d_final), its one small copy brings the
// list back too and none is launched here.
static int final_buf(tc_ctx *ctx, const MtfRleLink *link, uint16_t **d_final) {
    if (link && link->d_final) {
        *d_final = link->d_final;
        return TC_OK;
    }
    return ws_alloc(ctx, SIGMAX, d_final);
}
int mtf_read_final(tc_ctx *ctx, const uint16_t *d_final, uint32_t sigma, const int16_t *alpha, int16_t *final_list,
                   bool defer, uint32_t nalpha = 0, const MtfRleLink *link = nullptr) {
    // alpha: device list entry -> symbol (nalpha entries, default sigma)
    if (!nalpha) nalpha = sigma;
    const bool shared = defer && link && link->d_final == d_final;
    ctx->mtf_pending.h_off = shared ? MtfRleLink::H_FINAL : 512;
    uint16_t *h_final = (uint16_t *)(ctx->h_scal + ctx->mtf_pending.h_off);
    if (!shared) TC_TRY(tc_d2h_small(ctx, h_final, d_final, sigma * sizeof(uint16_t)));
    if (defer) {
        ctx->mtf_pending.active = true;
        ctx->mtf_pending.sigma = sigma;
        memcpy(ctx->mtf_pending.alpha, alpha, nalpha * sizeof(int16_t));
        ctx->mtf_pending.final_list = final_list;
        return TC_OK;
    }
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t j = 0; j < sigma; j++) final_list[j] = alpha[h_final[j]];
    return TC_OK;
}

// Host side of the automaton: rebuilt when the alphabet changes, kept in device memory.
struct AutoTables {
    uint32_t n_perm = 0;
    uint16_t *d_perm = nullptr;
    uint32_t *d_list = nullptr;
    AutoHash hash{0, 0};
    uint32_t sigma = 0;
    uint16_t li_of_rank[8] = {0};
};
// li_of_rank[r], r < sigma: the symbols of the alphabet in order (Nothing = 256 first if present)
int auto_tables(tc_ctx *ctx, uint32_t sigma, const uint16_t *li_of_rank, AutoTables *out) {
    AutoTables *t = static_cast<AutoTables *>(ctx->mtf_auto[0]);
    if (t && t->sigma == sigma && memcmp(t->li_of_rank, li_of_rank, sigma * sizeof(uint16_t)) == 0) {
        *out = *t;
        return TC_OK;
    }
    // column hash: an odd multiplier and a shift that send the alphabet to distinct columns 0..7
    AutoHash hs{0, 0};
    uint32_t colof[8];
    {
        uint64_t x = 0x9E3779B97F4A7C15ull;
        bool ok = false;
        for (int tries = 0; tries < 200000 && !ok; tries++) {
            x ^= x << 13, x ^= x >> 7, x ^= x << 17; // xorshift64
            hs.mul = (uint32_t)(x >> 16) | 1u;
            for (uint32_t sh = 0; sh < 29 && !ok; sh++) {
                hs.shift = sh;
                uint32_t used = 0;
                ok = true;
                for (uint32_t r = 0; r < sigma && ok; r++) {
                    colof[r] = hs.col2(li_of_rank[r]) >> 1;
                    ok = !((used >> colof[r]) & 1u);
                    used |= 1u << colof[r];
                }
            }
        }
        if (!ok) return TC_E_ARG; // cannot happen for <= 6 symbols and 8 columns
    }
    uint32_t nperm = 1;
    for (uint32_t i = 2; i <= sigma; i++) nperm *= i;
    std::vector<uint16_t> perm((size_t)nperm * AU_RW, 0);
    std::vector<uint32_t> lists(nperm);
    {
        uint32_t p[8];
        for (uint32_t i = 0; i < sigma; i++) p[i] = i;
        do {
            uint32_t l = 0;
            for (uint32_t i = 0; i < sigma; i++) l |= p[i] << (4 * i);
            lists[perm_rank(l, sigma)] = l;
        } while (std::next_permutation(p, p + sigma));
    }
    for (uint32_t id = 0; id < nperm; id++) {
        const uint32_t l = lists[id];
        for (uint32_t c = 0; c < 8; c++) perm[(size_t)id * AU_RW + c] = (uint16_t)((5 * id) << 3); // no-op
        for (uint32_t sy = 0; sy < sigma; sy++) {
            uint32_t pos = 0;
            while (((l >> (4 * pos)) & 15u) != sy) pos++;
            const uint32_t low = (1u << (4 * pos)) - 1u;
            const uint32_t nl = (l & ~((low << 4) | 0xFu)) | ((l & low) << 4) | sy;
            perm[(size_t)id * AU_RW + colof[sy]] = (uint16_t)(((5 * perm_rank(nl, sigma)) << 3) | pos);
        }
    }
    if (!t) {
        t = new AutoTables();
        TC_CUDA(cudaMalloc((void **)&t->d_perm, 720 * AU_RW * sizeof(uint16_t)));
        TC_CUDA(cudaMalloc((void **)&t->d_list, 720 * sizeof(uint32_t)));
        ctx->mtf_auto[0] = t;
    }
    t->sigma = 0; // invalid until the upload below has succeeded
    TC_CUDA(cudaStreamSynchronize(ctx->stream)); // an earlier block may still be reading the old table
    TC_CUDA(cudaMemcpyAsync(t->d_perm, perm.data(), perm.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    TC_CUDA(cudaMemcpyAsync(t->d_list, lists.data(), lists.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    TC_CUDA(cudaStreamSynchronize(ctx->stream)); // the vectors go away
    t->n_perm = nperm;
    t->hash = hs;
    t->sigma = sigma;
    memcpy(t->li_of_rank, li_of_rank, sigma * sizeof(uint16_t));
    *out = *t;
    return TC_OK;
}

// present_hint (257 flags, code = symbol + 1, 0 = Nothing): the alphabet when the caller already
// knows it (the composed helpers do: a BWT has the symbols of its text plus the sentinel)
template <class Src>
int mtf_encode_impl(tc_ctx *ctx, Src src, uint64_t N, uint16_t *d_idx, int16_t *final_list, uint32_t *sigma_out,
                    const uint8_t *present_hint = nullptr, MtfRleLink *link = nullptr) {
    *sigma_out = 0;
    if (link) link->valid = link->scanned = false;
    if (N == 0) return TC_OK;
    if (N >= 0x7fffffffull) return TC_E_TOOBIG; // recency keys are 32-bit distances (see list_positions)
    WsMark mk = tc_ws_mark(ctx);
    // alphabet = nubSeq' (sorted, Nothing first)
    uint32_t h_present_buf[SIGMAX];
    if (present_hint) {
        for (int c = 0; c < SIGMAX; c++) h_present_buf[c] = present_hint[c];
    } else {
        uint32_t *d_present;
        TC_TRY(ws_alloc(ctx, SIGMAX, &d_present));
        TC_CUDA(cudaMemsetAsync(d_present, 0, SIGMAX * sizeof(uint32_t), ctx->stream));
        unsigned pgrid = (unsigned)std::min<uint64_t>(ceil_div_u64(N, 256 * 16), (uint64_t)ctx->sm_count * 8);
        TC_LAUNCH(ctx, (mtf_presence_kernel<Src>), pgrid, 256, 0, src, N, d_present);
        TC_TRY(tc_d2h_small(ctx, ctx->h_scal, d_present, SIGMAX * sizeof(uint32_t)));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(h_present_buf, ctx->h_scal, SIGMAX * sizeof(uint32_t));
    }
    const uint32_t *h_present = h_present_buf;
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
    if (sigma <= AU_MAXSIG && !ctx->mtf_v2) { // automata in shared memory
        AutoTables at;
        uint16_t li_of_rank[8];
        for (uint32_t j = 0; j < sigma; j++) li_of_rank[j] = (uint16_t)(alpha[j] < 0 ? 256 : alpha[j]);
        TC_TRY(auto_tables(ctx, sigma, li_of_rank, &at));
        const uint64_t nchunks = ceil_div_u64(N, AU_L), ntiles = ceil_div_u64(nchunks, AU_T);
        Summ *part, *tot;
        uint32_t *start_list;
        uint16_t *d_final;
        TC_TRY(ws_alloc(ctx, ntiles * AU_T, &part));
        TC_TRY(ws_alloc(ctx, ntiles, &tot));
        TC_TRY(ws_alloc(ctx, ntiles, &start_list));
        TC_TRY(final_buf(ctx, link, &d_final));
        const size_t sma = sizeof(AutoSmem);
        const uint32_t abit = sizeof(*src.p) == 1 ? 4u : 8u;
        if (!(ctx->attr_done & abit)) {
            TC_CUDA(cudaFuncSetAttribute(mtfa_summary_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sma));
            TC_CUDA(cudaFuncSetAttribute(mtfa_replay_kernel<Src, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sma));
            TC_CUDA(cudaFuncSetAttribute(mtfa_replay_kernel<Src, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sma));
            ctx->attr_done |= abit;
        }
        ctx->prof_bytes_next = N * sizeof(*src.p);
        TC_LAUNCH(ctx, (mtfa_summary_kernel<Src>), (unsigned)ntiles, AU_T, sma, src, at.hash, N, at.d_perm, at.d_list, at.n_perm,
                  part, tot, link ? link->d_ticket : (uint32_t *)nullptr);
        TC_LAUNCH(ctx, mtfs_top_kernel, 1, 1024, 0, tot, ntiles, sigma, start_list, d_final);
        ctx->prof_bytes_next = N * (sizeof(*src.p) + 2);
        if (link) {
            TC_LAUNCH(ctx, (mtfa_replay_kernel<Src, true>), (unsigned)ntiles, AU_T, sma, src, at.hash, N, at.d_perm, at.d_list,
                      at.n_perm, part, start_list, sigma, d_idx, link->d_tstat, link->d_toff, link->d_theadx, link->d_ticket);
            link->ntiles = ceil_div_u64(N, 32 * AU_L), link->tile_syms = 32 * AU_L, link->valid = true, link->scanned = true;
        } else {
            TC_LAUNCH(ctx, (mtfa_replay_kernel<Src, false>), (unsigned)ntiles, AU_T, sma, src, at.hash, N, at.d_perm, at.d_list,
                      at.n_perm, part, start_list, sigma, d_idx, (uint4 *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                      (uint32_t *)nullptr);
        }
        *sigma_out = sigma;
        int rc = mtf_read_final(ctx, d_final, sigma, alpha, final_list, present_hint != nullptr, 0, link);
        tc_ws_release(ctx, mk);
        return rc;
    }
    if (sigma <= 8) { // the list fits one register
        const uint64_t nchunks = ceil_div_u64(N, SM_L), ntiles = ceil_div_u64(nchunks, SM_T);
        Summ *part, *tot;
        uint32_t *start_list;
        uint16_t *d_final;
        TC_TRY(ws_alloc(ctx, ntiles * SM_T, &part));
        TC_TRY(ws_alloc(ctx, ntiles, &tot));
        TC_TRY(ws_alloc(ctx, ntiles, &start_list));
        TC_TRY(final_buf(ctx, link, &d_final));
        ctx->prof_bytes_next = N * sizeof(*src.p);
        TC_LAUNCH(ctx, (mtfs_summary_kernel<Src>), (unsigned)ntiles, SM_T, 0, src, lut, N, part, tot);
        TC_LAUNCH(ctx, mtfs_top_kernel, 1, 1024, 0, tot, ntiles, sigma, start_list, d_final);
        ctx->prof_bytes_next = N * (sizeof(*src.p) + 2);
        TC_LAUNCH(ctx, (mtfs_replay_kernel<Src>), (unsigned)ntiles, SM_T, 0, src, lut, N, part, start_list, sigma, d_idx);
        *sigma_out = sigma;
        int rc = mtf_read_final(ctx, d_final, sigma, alpha, final_list, present_hint != nullptr, 0, link);
        tc_ws_release(ctx, mk);
        return rc;
    }
    const uint32_t VS = (sigma + 31) / 32 * 32;
    if (!ctx->mtf_v2) { // thread-per-chunk replay (TC_B200_MTF_V2=1 keeps the warp-per-chunk kernel for comparison)
        // one chunk per resident thread when the input allows it (two CTAs of R3_CT threads per SM)
        uint64_t Lt = ceil_div_u64(N, (uint64_t)ctx->sm_count * 2 * R3_CT);
        if (ctx->mtf_L) Lt = ctx->mtf_L;
        Lt = (Lt + 31) / 32 * 32;
        const uint32_t L = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(Lt, 128), R3_LMAX);
        const uint64_t nchunks = ceil_div_u64(N, L), ntiles = ceil_div_u64(nchunks, TT_CH);
        uint32_t *trow, *finalocc;
        uint16_t *d_final, *start;
        TC_TRY(ws_alloc(ctx, ntiles * R3_ROW, &trow));
        TC_TRY(ws_alloc(ctx, R3_ROW + 32, &finalocc)); // + the tickets of the last-CTA elections
        uint32_t *segpre;
        TC_TRY(ws_alloc(ctx, (size_t)T2_SEGS * R3_ROW, &segpre));
        const uint32_t seg_tiles = (uint32_t)std::max<uint64_t>(1, ceil_div_u64(ntiles, T2_SEGS));
        TC_TRY(final_buf(ctx, link, &d_final));
        TC_TRY(ws_alloc(ctx, nchunks * R3_ROW, &start));
        uint32_t *ticket = finalocc + R3_ROW;
        Present pr;
        memset(&pr, 0, sizeof pr);
        int16_t alpha_li[SIGMAX]; // li -> symbol
        for (int c = 0; c < SIGMAX; c++) {
            const int li = c == 0 ? 256 : c - 1;
            alpha_li[li] = (int16_t)(c - 1);
            if (h_present[c]) pr.w[li >> 5] |= 1u << (li & 31);
        }
        const size_t smem1 = (size_t)T1_WARPS * 32 * TP_RLW * sizeof(uint32_t);
        const size_t smem3 = (size_t)R3_WORDS * R3_CT * sizeof(uint32_t);
        const uint32_t abit = sizeof(*src.p) == 1 ? 1u : 2u;
        if (!(ctx->attr_done & abit)) {
            TC_CUDA(cudaFuncSetAttribute(mtf3_tile_last_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            TC_CUDA(cudaFuncSetAttribute(mtf3_replay_kernel<Src, R3_CT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem3));
            TC_CUDA(cudaFuncSetAttribute(mtf3_replay_kernel<Src, R3_CT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem3));
            ctx->attr_done |= abit;
        }
        TC_LAUNCH(ctx, (mtf3_tile_last_kernel<Src>), (unsigned)ceil_div_u64(ntiles, T1_WARPS), T1_WARPS * 32, smem1, src, N,
                  L, nchunks, trow, start, ticket, link ? link->d_ticket : ticket + 1);
        TC_LAUNCH(ctx, mtf3_tile_scan_kernel, dim3(R3_ROW / 32, T2_SEGS), T2_WARPS * 32, 0, trow, ntiles, seg_tiles, segpre,
                  finalocc, ticket);
        TC_LAUNCH(ctx, mtf3_starts_kernel, (unsigned)ceil_div_u64(ntiles, T3_WARPS) + 1, T3_WARPS * 32, 0, pr, sigma, L,
                  nchunks, trow, segpre, seg_tiles, start, finalocc, N, d_final);
        ctx->prof_bytes_next = N * (sizeof(*src.p) + 2);
        if (link) {
            TC_LAUNCH(ctx, (mtf3_replay_kernel<Src, R3_CT, true>), (unsigned)ceil_div_u64(nchunks, R3_CT), R3_CT, smem3, src, N, L,
                      nchunks, sigma, start, d_idx, link->d_tstat, link->d_toff, link->d_theadx, link->d_ticket);
            link->ntiles = ceil_div_u64(nchunks, 32), link->tile_syms = 32 * L, link->valid = true, link->scanned = true;
        } else {
            TC_LAUNCH(ctx, (mtf3_replay_kernel<Src, R3_CT, false>), (unsigned)ceil_div_u64(nchunks, R3_CT), R3_CT, smem3, src, N,
                      L, nchunks, sigma, start, d_idx, (uint4 *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                      (uint32_t *)nullptr);
        }
        *sigma_out = sigma;
        int rc = mtf_read_final(ctx, d_final, sigma, alpha_li, final_list, present_hint != nullptr, SIGMAX, link);
        tc_ws_release(ctx, mk);
        return rc;
    }
    // chunk length: a multiple of 32, enough chunks to fill the machine with warps
    uint64_t Lw = ceil_div_u64(N, (uint64_t)ctx->sm_count * 64);
    Lw = (Lw + 31) / 32 * 32;
    const uint32_t L = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(Lw, 256), 1024);
    const uint64_t nchunks = ceil_div_u64(N, L);
    const uint32_t G = 64;
    const uint64_t ntiles = ceil_div_u64(nchunks, G);
    uint32_t *lastocc, *tiletot, *finalocc;
    uint16_t *d_final;
    TC_TRY(ws_alloc(ctx, nchunks * VS, &lastocc));
    TC_TRY(ws_alloc(ctx, ntiles * VS, &tiletot));
    TC_TRY(ws_alloc(ctx, VSMAX, &finalocc));
    TC_TRY(final_buf(ctx, link, &d_final));
    unsigned cgrid = (unsigned)ceil_div_u64(nchunks, ENC_WARPS);
    TC_LAUNCH(ctx, (mtf2_lastocc_kernel<Src>), cgrid, ENC_WARPS * 32, 0, src, lut, N, L, nchunks, VS, lastocc);
    TC_LAUNCH(ctx, mtf2_scan_tiles_kernel, (unsigned)ntiles, VSMAX, 0, lastocc, nchunks, G, VS, tiletot);
    TC_LAUNCH(ctx, mtf2_scan_top_kernel, 1, VSMAX, 0, tiletot, ntiles, VS, finalocc);
    ctx->prof_bytes_next = N * (sizeof(*src.p) + 2);
    TC_LAUNCH(ctx, (mtf2_replay_kernel<Src>), cgrid, ENC_WARPS * 32, 0, src, lut, N, L, nchunks, G, sigma, VS, lastocc,
              tiletot, d_idx);
    TC_LAUNCH(ctx, mtf2_final_kernel, 1, 32, 0, finalocc, N, sigma, VS, d_final);
    *sigma_out = sigma;
    int rc = mtf_read_final(ctx, d_final, sigma, alpha, final_list, present_hint != nullptr, 0, link);
    tc_ws_release(ctx, mk);
    return rc;
}
} // namespace

void mtf_free_tables(tc_ctx *ctx) {
    for (void *&p : ctx->mtf_auto) {
        if (!p) continue;
        AutoTables *t = static_cast<AutoTables *>(p);
        cudaFree(t->d_perm);
        cudaFree(t->d_list);
        delete t;
        p = nullptr;
    }
}

// after a stream sync: the deferred final list of the last mtf_encode with an alphabet hint
int mtf_finish_pending(tc_ctx *ctx) {
    if (!ctx->mtf_pending.active) return TC_OK;
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint16_t *h_final = (const uint16_t *)(ctx->h_scal + ctx->mtf_pending.h_off);
    for (uint32_t j = 0; j < ctx->mtf_pending.sigma; j++) // (an RLE stage that failed before its copy leaves garbage here)
        ctx->mtf_pending.final_list[j] = h_final[j] < SIGMAX ? ctx->mtf_pending.alpha[h_final[j]] : (int16_t)-1;
    ctx->mtf_pending.active = false;
    return TC_OK;
}

int mtf_encode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint16_t *d_idx,
                           int16_t *final_list, uint32_t *sigma, const uint8_t *present_hint, MtfRleLink *link) {
    if (N && primary >= N) primary = ~0ull;
    return mtf_encode_impl(ctx, SrcU8{d_bwt, primary}, N, d_idx, final_list, sigma, present_hint, link);
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
    if (sigma <= 8) { // the list fits one register
        const uint64_t nch = ceil_div_u64(N, SM_L), nt = ceil_div_u64(nch, SM_T);
        uint32_t *part, *tot, *start, *d_err;
        TC_TRY(ws_alloc(ctx, nt * SM_T, &part));
        TC_TRY(ws_alloc(ctx, nt, &tot));
        TC_TRY(ws_alloc(ctx, nt, &start));
        TC_TRY(ws_alloc(ctx, 1, &d_err));
        TC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream));
        List8 l8;
        for (int j = 0; j < 8; j++) l8.sym[j] = j < (int)sigma ? l0.sym[j] : (int16_t)0;
        TC_LAUNCH(ctx, mtfds_perm_kernel, (unsigned)nt, SM_T, 0, d_idx, N, sigma, part, tot, d_err);
        TC_LAUNCH(ctx, mtfds_top_kernel, 1, 1024, 0, tot, nt, start);
        ctx->prof_bytes_next = 4 * N;
        TC_LAUNCH(ctx, mtfds_replay_kernel, (unsigned)nt, SM_T, 0, d_idx, N, sigma, part, start, l8, d_sym);
        uint32_t *h_err = (uint32_t *)ctx->h_scal;
        TC_TRY(tc_d2h_small(ctx, h_err, d_err, sizeof(uint32_t)));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        tc_ws_release(ctx, mk);
        return h_err[0] ? TC_E_INDEX : TC_OK;
    }
    if (!ctx->mtfd_v1) {
        // select-based replay (mtfd3_*): whole waves of resident CTAs (2 per SM), chunks of at most 224 symbols
        const uint64_t resident = (uint64_t)ctx->sm_count * 2 * D3_CT;
        const uint64_t waves = std::max<uint64_t>(1, ceil_div_u64(N, resident * d3::LMAX));
        uint64_t Lw = ceil_div_u64(N, waves * resident);
        Lw = (Lw + 15) / 16 * 16;
        const uint32_t L = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(Lw, 32), d3::LMAX);
        const uint64_t nchunks = ceil_div_u64(N, L);
        const uint32_t G = D3_G;
        const uint64_t ntiles = ceil_div_u64(nchunks, G), nsuper = ceil_div_u64(ntiles, G);
        uint16_t *perm, *part, *tilesum, *part2, *supersum;
        int16_t *superprefix;
        uint32_t *d_err;
        TC_TRY(ws_alloc(ctx, nchunks * D3_STRIDE + 8, &perm));
        TC_TRY(ws_alloc(ctx, nchunks * D3_STRIDE + 8, &part));
        TC_TRY(ws_alloc(ctx, ntiles * sigma, &tilesum));
        TC_TRY(ws_alloc(ctx, ntiles * sigma, &part2));
        TC_TRY(ws_alloc(ctx, nsuper * sigma, &supersum));
        TC_TRY(ws_alloc(ctx, nsuper * sigma, &superprefix));
        TC_TRY(ws_alloc(ctx, 1, &d_err));
        TC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream));
        if (!(ctx->attr_done & 16u)) {
            TC_CUDA(cudaFuncSetAttribute(mtfd3_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(D3Smem)));
            TC_CUDA(cudaFuncSetAttribute(mtfd3_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(D3Smem)));
            ctx->attr_done |= 16u;
        }
        const unsigned cgrid = (unsigned)ceil_div_u64(nchunks, D3_CT);
        ctx->prof_bytes_next = 2 * N;
        TC_LAUNCH(ctx, mtfd3_perm_kernel, cgrid, D3_CT, sizeof(D3Smem), d_idx, N, L, nchunks, sigma, perm, d_err);
        TC_LAUNCH(ctx, mtfd_tile_chain_kernel, (unsigned)ceil_div_u64(ntiles, 4), 128, 0, perm, nchunks, G, sigma, part,
                  tilesum, ntiles, (uint32_t)D3_STRIDE, (uint32_t)D3_STRIDE);
        TC_LAUNCH(ctx, mtfd_tile_chain_kernel, (unsigned)ceil_div_u64(nsuper, 4), 128, 0, tilesum, ntiles, G, sigma, part2,
                  supersum, nsuper, sigma, sigma);
        TC_LAUNCH(ctx, mtfd_top_chain_kernel, 1, 32, 0, supersum, nsuper, sigma, l0, superprefix);
        ctx->prof_bytes_next = 4 * N;
        TC_LAUNCH(ctx, mtfd3_replay_kernel, cgrid, D3_CT, sizeof(D3Smem), d_idx, N, L, nchunks, G, sigma, part, part2,
                  superprefix, d_sym);
        uint32_t *h_err = (uint32_t *)ctx->h_scal;
        TC_TRY(tc_d2h_small(ctx, h_err, d_err, sizeof(uint32_t)));
        TC_CUDA(cudaStreamSynchronize(ctx->stream));
        tc_ws_release(ctx, mk);
        return h_err[0] ? TC_E_INDEX : TC_OK;
    }
    // ~512 chunks per SM keep the 12 resident warps per SM (516 B of list per thread) busy
    uint64_t Lw = ceil_div_u64(N, (uint64_t)ctx->sm_count * 512);
    Lw = (Lw + 15) / 16 * 16;
    const uint32_t L = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(Lw, 64), 512);
    const uint64_t nchunks = ceil_div_u64(N, L);
    // three levels of serial composition chains (64 chunks per tile, 64 tiles per super-tile, then
    // the super-tiles): no chain is longer than 64 steps for blocks up to 64 MiB
    const uint32_t G = 64;
    const uint64_t ntiles = ceil_div_u64(nchunks, G), nsuper = ceil_div_u64(ntiles, G);
    uint16_t *perm, *part, *tilesum, *part2, *supersum;
    int16_t *superprefix;
    uint32_t *d_err;
    TC_TRY(ws_alloc(ctx, nchunks * sigma, &perm));
    TC_TRY(ws_alloc(ctx, nchunks * sigma, &part));
    TC_TRY(ws_alloc(ctx, ntiles * sigma, &tilesum));
    TC_TRY(ws_alloc(ctx, ntiles * sigma, &part2));
    TC_TRY(ws_alloc(ctx, nsuper * sigma, &supersum));
    TC_TRY(ws_alloc(ctx, nsuper * sigma, &superprefix));
    TC_TRY(ws_alloc(ctx, 1, &d_err));
    TC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream));
    const int T = 64;
    unsigned cgrid = (unsigned)ceil_div_u64(nchunks, T);
    size_t smem = (size_t)LASTW * T * sizeof(uint32_t);
    TC_LAUNCH(ctx, mtfd_perm_kernel, cgrid, T, smem, d_idx, N, L, nchunks, sigma, perm, d_err);
    TC_LAUNCH(ctx, mtfd_tile_chain_kernel, (unsigned)ceil_div_u64(ntiles, 4), 128, 0, perm, nchunks, G, sigma, part,
              tilesum, ntiles, sigma, sigma);
    TC_LAUNCH(ctx, mtfd_tile_chain_kernel, (unsigned)ceil_div_u64(nsuper, 4), 128, 0, tilesum, ntiles, G, sigma, part2,
              supersum, nsuper, sigma, sigma);
    TC_LAUNCH(ctx, mtfd_top_chain_kernel, 1, 32, 0, supersum, nsuper, sigma, l0, superprefix);
    TC_LAUNCH(ctx, mtfd_replay_kernel, cgrid, T, smem, d_idx, N, L, nchunks, G, sigma, part, part2, superprefix, d_sym);
    uint32_t *h_err = (uint32_t *)ctx->h_scal;
    TC_TRY(tc_d2h_small(ctx, h_err, d_err, sizeof(uint32_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    tc_ws_release(ctx, mk);
    return h_err[0] ? TC_E_INDEX : TC_OK;
}
