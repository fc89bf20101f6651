// rle.cu -- seqToRLE / seqFromRLE (src/Data/RLE/Internal.hs:104-189) as
// flag -> scan -> compaction kernels with 128-bit coalesced HBM access.
//
// Position-local form of the reference's sequential state machine:
//   head(i)  : x[i] is Just and (i == 0 or x[i-1] is Nothing or x[i-1] != x[i])
//   H(i)     : latest head <= i            (max-scan)
//   J(i)     : latest Just position <= i   (max-scan)
//   at i >= 1:  x[i] Nothing            -> 2 pairs: state(i), (1, Nothing)
//               x[i], x[i-1] Just, !=   -> 1 pair : (i - H(i-1), x[i-1])
//   state(i) = x[i-1] Just ? (i - H(i-1), x[i-1]) : (stale, Nothing),
//              stale = J(i-1) - H(i-1) + 1, or 1 when no Just precedes (Q1-Q3)
//   after the last position: one pair state(N).
// Algorithmic bytes: N * w_in read + 6 * R written (u32 count + i16 symbol per run).
#include "common.cuh"
#include "impl.cuh"

namespace {
constexpr int RT = 256;
constexpr int NOPREV = -2;

struct InU8 {
    const uint8_t *p;
    uint64_t primary;
    static constexpr int ITEMS = 16;
    static constexpr bool HAS_NOTHING = true;
    static constexpr int MAX_PER_ITEM = 1; // a single Nothing: at most TILE + 1 pairs + the flush
    __device__ __forceinline__ int at(uint64_t i) const { return i == primary ? -1 : (int)p[i]; }
    __device__ __forceinline__ void load(uint64_t base, uint64_t N, int *c) const {
        if (base + ITEMS <= N && ((reinterpret_cast<uintptr_t>(p + base) & 15) == 0)) {
            uint4 v = ld_stream_u4(p + base);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; k++) c[k] = (w[k >> 2] >> ((k & 3) * 8)) & 0xff;
            if (primary >= base && primary < base + ITEMS) {
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (base + k == primary) c[k] = -1;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; k++) c[k] = (base + k < N) ? at(base + k) : 0;
        }
    }
};
template <bool SIGNED>
struct In16 {
    const uint16_t *p;
    static constexpr int ITEMS = SIGNED ? 8 : 16;
    static constexpr bool HAS_NOTHING = SIGNED; // the unsigned index stream has no Nothing: plain run boundaries
    static constexpr int MAX_PER_ITEM = SIGNED ? 2 : 1; // every position may be a Nothing (2 pairs)
    __device__ __forceinline__ int cvt(uint32_t h) const {
        if (SIGNED) {
            int v = (int)(int16_t)h;
            return v < 0 ? -1 : v;
        }
        return (int)h;
    }
    __device__ __forceinline__ int at(uint64_t i) const { return cvt(p[i]); }
    __device__ __forceinline__ void load(uint64_t base, uint64_t N, int *c) const {
        if (base + ITEMS <= N && ((reinterpret_cast<uintptr_t>(p + base) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < ITEMS / 8; q++) {
                uint4 v = ld_stream_u4(p + base + q * 8);
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 8; k++) c[q * 8 + k] = cvt((w[k >> 1] >> ((k & 1) * 16)) & 0xffff);
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; k++) c[k] = (base + k < N) ? at(base + k) : 0;
        }
    }
};

// emissions of position i given its symbol c and predecessor p (NOPREV at i == 0)
__device__ __forceinline__ int n_emit(int c, int p) {
    if (p == NOPREV) return 0;
    if (c < 0) return 2;
    return (p >= 0 && p != c) ? 1 : 0;
}

template <class In>
__global__ void __launch_bounds__(RT)
    rle_reduce_kernel(In in, uint64_t N, uint32_t *__restrict__ tile_pairs, uint32_t *__restrict__ tile_head,
                      uint32_t *__restrict__ tile_just) {
    constexpr int ITEMS = In::ITEMS;
    __shared__ uint32_t sh[3 * (RT / 32)];
    uint64_t base = ((uint64_t)blockIdx.x * RT + threadIdx.x) * ITEMS;
    int c[ITEMS];
    uint32_t pairs = 0, lh = 0, lj = 0;
    if (base < N) {
        in.load(base, N, c);
        int p = base == 0 ? NOPREV : in.at(base - 1);
        const uint32_t b32 = (uint32_t)base;
        const int lim = base + ITEMS <= N ? ITEMS : (int)(N - base); // items of this thread inside the input
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            if (k < lim) {
                if (In::HAS_NOTHING) {
                    pairs += n_emit(c[k], p);
                    if (c[k] >= 0) {
                        lj = b32 + k + 1;
                        if (p < 0 || p != c[k]) lh = b32 + k + 1;
                    }
                } else { // no Nothing in the stream: a pair per change of symbol
                    if (p != c[k]) {
                        lh = b32 + k + 1;
                        pairs += p != NOPREV;
                    }
                }
                p = c[k];
            }
        }
    }
    // tile totals only: warp reductions (redux.sync) and one round through shared memory
    const uint32_t wp = __reduce_add_sync(TC_FULL, pairs), wh = __reduce_max_sync(TC_FULL, lh);
    const uint32_t wj = In::HAS_NOTHING ? __reduce_max_sync(TC_FULL, lj) : 0u;
    if (lane_id() == 0) {
        sh[threadIdx.x >> 5] = wp;
        sh[RT / 32 + (threadIdx.x >> 5)] = wh;
        sh[2 * (RT / 32) + (threadIdx.x >> 5)] = wj;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tp = 0, th = 0, tj = 0;
#pragma unroll
        for (int w = 0; w < RT / 32; w++) {
            tp += sh[w];
            th = max(th, sh[RT / 32 + w]);
            tj = max(tj, sh[2 * (RT / 32) + w]);
        }
        tile_pairs[blockIdx.x] = tp;
        tile_head[blockIdx.x] = th;
        tile_just[blockIdx.x] = tj;
    }
}

// single block: exclusive sum of pairs (u64) and exclusive max of head / just over tiles.  Every
// thread owns RS_PER consecutive tiles (local scan in registers), so 8 Ki tiles -- a 32 MiB index
// stream -- need one round of three block scans.
constexpr int RS_PER = 8;
__global__ void __launch_bounds__(1024)
    rle_tile_scan_kernel(const uint32_t *__restrict__ tile_pairs, uint32_t *tile_head, uint32_t *tile_just,
                         uint64_t *__restrict__ tile_off, uint64_t tiles) {
    __shared__ uint64_t sh64[1024 / 32 + 1];
    __shared__ uint32_t sh32[1024 / 32 + 1];
    uint64_t carry = 0;
    uint32_t ch = 0, cj = 0;
    for (uint64_t b = 0; b < tiles; b += 1024 * RS_PER) {
        const uint64_t t0 = b + (uint64_t)threadIdx.x * RS_PER;
        uint32_t p[RS_PER], h[RS_PER], j[RS_PER];
        uint64_t v = 0;
        uint32_t hm = 0, jm = 0;
#pragma unroll
        for (int q = 0; q < RS_PER; q++) {
            const bool in = t0 + q < tiles;
            p[q] = in ? tile_pairs[t0 + q] : 0;
            h[q] = in ? tile_head[t0 + q] : 0;
            j[q] = in ? tile_just[t0 + q] : 0;
            v += p[q];
            hm = max(hm, h[q]);
            jm = max(jm, j[q]);
        }
        uint64_t tot;
        uint32_t th, tj;
        uint64_t ex = carry + block_excl_sum<uint64_t, 1024>(v, sh64, &tot);
        uint32_t hx = max(ch, block_excl_max<uint32_t, 1024>(hm, 0u, sh32, &th));
        uint32_t jx = max(cj, block_excl_max<uint32_t, 1024>(jm, 0u, sh32, &tj));
#pragma unroll
        for (int q = 0; q < RS_PER; q++) {
            if (t0 + q < tiles) {
                tile_off[t0 + q] = ex;
                tile_head[t0 + q] = hx;
                tile_just[t0 + q] = jx;
            }
            ex += p[q];
            hx = max(hx, h[q]);
            jx = max(jx, j[q]);
        }
        carry += tot;
        ch = max(ch, th);
        cj = max(cj, tj);
    }
}

struct PackArgs { // device pointers of RlePack, passed by value to the emit kernel
    uint8_t *cnt4, *sym8;
    uint32_t *hi;
    uint64_t *big_idx;
    uint32_t *big_cnt;
    uint64_t big_cap;
    unsigned long long *n_big;
};

// One tile of RT * ITEMS positions starting at tile_base (those below lim_end <= N): goff = runs emitted before it,
// Hx / Jx = latest head / latest Just before it.  Returns the tile's run count and its latest head (for callers that walk several
// tiles in one CTA).  All threads of the CTA must call.
template <class In, bool PACKED>
__device__ __forceinline__ void rle_emit_tile(In in, uint64_t N, uint64_t lim_end, uint64_t tile_base, uint64_t goff, uint32_t Hx,
                                              uint32_t Jx,
                                              uint32_t *__restrict__ count, int16_t *__restrict__ rsym, uint64_t cap,
                                              uint64_t *__restrict__ d_R, PackArgs pk, bool last_tile, uint32_t *out_total,
                                              uint32_t *out_head) {
    constexpr int ITEMS = In::ITEMS;
    // + one Nothing's second pair + final flush + alignment pad, rounded up to the swizzle period
    constexpr int CAP = (In::MAX_PER_ITEM * RT * ITEMS + 3 + 8 + 63) / 64 * 64;
    __shared__ uint32_t sh[RT / 32 + 1];
    __shared__ __align__(16) uint32_t s_cnt_raw[CAP];
    __shared__ __align__(16) int16_t s_sym_raw[CAP];
    // runs are staged so that staged index and global index agree modulo the vector width:
    // the copy-out below then moves 16 bytes per store
    const uint32_t padc = (uint32_t)(goff & 3), pads = (uint32_t)(goff & 7);
    // A thread writes its (up to 16) runs to consecutive slots, so the lanes of a warp hit slots 16
    // apart: 2 banks for 32 lanes.  XOR-ing two slot bits with the thread's position spreads them
    // over 8 banks (16-way -> 4-way conflicts; this staging was the whole kernel time: 60 -> 45 us)
    // and keeps aligned groups of 4 counts / 8 symbols contiguous for the 16-byte copy-out.  A
    // full 5-bit swizzle with a scalar copy-out measured the same (46-50 us).
    auto pc = [&](uint32_t slot) { uint32_t j = padc + slot; return j ^ (((j >> 5) & 3u) << 2); };
    auto ps = [&](uint32_t slot) { uint32_t j = pads + slot; return j ^ (((j >> 6) & 3u) << 3); };
    // PACKED: the same storage holds the packed form instead -- one byte per count, one per symbol
    // (consecutive bytes per thread: lanes 16 bytes apart, 4-way conflicts without any swizzle) and
    // a bitmap of the symbols' bit 8 -- at staged index padp + slot with padp = goff mod 32, so
    // 16-byte groups of the byte streams and 32-bit words of the bitmap line up with global memory
    uint8_t *s_c8 = reinterpret_cast<uint8_t *>(s_cnt_raw);
    uint8_t *s_s8 = reinterpret_cast<uint8_t *>(s_sym_raw);
    __shared__ uint32_t s_hi[PACKED ? (CAP + 32) / 32 + 1 : 1];
    const uint32_t padp = (uint32_t)(goff & 31);
    if (PACKED)
        for (uint32_t j = threadIdx.x; j < (CAP + 32) / 32 + 1; j += RT) s_hi[j] = 0; // block scans below sync
    auto put = [&](uint32_t slot, uint32_t cnt, int sym) {
        if (PACKED) {
            const uint32_t j = padp + slot;
            s_c8[j] = (uint8_t)min(cnt - 1u, 15u); // count - 1 in four bits; 15 = see the exception list
            s_s8[j] = (uint8_t)sym;
            if (sym & 0x100) atomicOr(&s_hi[j >> 5], 1u << (j & 31)); // MTF index 256, Nothing (-1)
            if (cnt >= 16u) {
                const unsigned long long e = atomicAdd(pk.n_big, 1ull);
                if (e < pk.big_cap) {
                    pk.big_idx[e] = goff + slot;
                    pk.big_cnt[e] = cnt;
                }
            }
        } else {
            s_cnt_raw[pc(slot)] = cnt;
            s_sym_raw[ps(slot)] = (int16_t)sym;
        }
    };
    uint64_t base = tile_base + (uint64_t)threadIdx.x * ITEMS;
    int c[ITEMS];
    int p0 = NOPREV;
    uint32_t pairs = 0, lh = 0, lj = 0;
    const uint32_t b32 = (uint32_t)base;
    const int lim = base >= lim_end ? 0 : (base + ITEMS <= lim_end ? ITEMS : (int)(lim_end - base));
    const bool owns_last = lim > 0 && base + lim == N; // this thread holds position N-1
    if (lim > 0) {
        in.load(base, N, c);
        p0 = base == 0 ? NOPREV : in.at(base - 1);
        int p = p0;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            if (k < lim) {
                if (In::HAS_NOTHING) {
                    pairs += n_emit(c[k], p);
                    if (c[k] >= 0) {
                        lj = b32 + k + 1;
                        if (p < 0 || p != c[k]) lh = b32 + k + 1;
                    }
                } else {
                    if (p != c[k]) {
                        lh = b32 + k + 1;
                        pairs += p != NOPREV;
                    }
                }
                p = c[k];
            }
        }
        if (owns_last) pairs += 1; // final flush
    }
    uint32_t tile_total, head_total;
    uint32_t o = block_excl_sum<uint32_t, RT>(pairs, sh, &tile_total);
    uint32_t H = block_excl_max<uint32_t, RT>(lh, 0u, sh, &head_total);
    uint32_t J = 0;
    if (In::HAS_NOTHING) J = block_excl_max<uint32_t, RT>(lj, 0u, sh, (uint32_t *)nullptr);
    H = max(H, Hx);
    if (In::HAS_NOTHING) J = max(J, Jx);
    *out_total = tile_total;
    *out_head = max(head_total, Hx);
    if (lim > 0) {
        int p = p0;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            if (k < lim) {
                const uint32_t i = b32 + k;
                const int ck = c[k];
                if (In::HAS_NOTHING) {
                    if (p != NOPREV) {
                        if (ck < 0) {
                            if (p >= 0) {
                                put(o, i - (H - 1), p);
                            } else {
                                put(o, J == 0 ? 1u : J - H + 1, -1);
                            }
                            put(o + 1, 1u, -1);
                            o += 2;
                        } else if (p >= 0 && p != ck) {
                            put(o, i - (H - 1), p);
                            o += 1;
                        }
                    }
                    if (ck >= 0) {
                        J = i + 1;
                        if (p < 0 || p != ck) H = i + 1;
                    }
                } else if (p != ck) {
                    if (p != NOPREV) {
                        put(o, i - (H - 1), p);
                        o += 1;
                    }
                    H = i + 1;
                }
                p = ck;
            }
        }
        if (owns_last && lim > 0) { // end-of-input flush (src/Data/RLE/Internal.hs:125-130); p = x[N-1]
            if (p >= 0) {
                put(o, (uint32_t)N - (H - 1), p);
            } else {
                put(o, J == 0 ? 1u : J - H + 1, -1);
            }
            o += 1;
        }
    }
    __syncthreads();
    if (PACKED) {
        if (goff + tile_total <= cap) {
            const uint64_t gp = goff - padp; // multiple of 32
            const uint32_t nv = (padp + tile_total + 15) / 16;
            for (uint32_t v = threadIdx.x; v < nv; v += RT) {
                const uint32_t j0 = 16 * v;
                if (j0 >= padp && j0 + 16 <= padp + tile_total) {
                    *reinterpret_cast<uint4 *>(pk.sym8 + gp + j0) = *reinterpret_cast<const uint4 *>(s_s8 + j0);
                } else {
                    for (uint32_t j = j0; j < j0 + 16; j++)
                        if (j >= padp && j < padp + tile_total) pk.sym8[gp + j] = s_s8[j];
                }
            }
            // counts: two per byte (run k in the low nibble of byte k / 2 when k is even), 32 runs = 16 bytes per
            // store; a 16-byte group shared with a neighbouring tile is OR-ed into the (zeroed) plane byte by byte
            const uint32_t nq = (padp + tile_total + 31) / 32;
            for (uint32_t v = threadIdx.x; v < nq; v += RT) {
                const uint32_t j0 = 32 * v;
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t x = 0;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const uint32_t j = j0 + 8 * q + e;
                        const uint32_t c = (j >= padp && j < padp + tile_total) ? s_c8[j] : 0u;
                        x |= c << (4 * e);
                    }
                    w[q] = x;
                }
                uint8_t *dst = pk.cnt4 + ((gp + j0) >> 1);
                if (j0 >= padp && j0 + 32 <= padp + tile_total) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (w[q]) atomicOr(reinterpret_cast<uint32_t *>(dst) + q, w[q]);
                }
            }
            // hi plane in whole words; a word shared with a neighbouring tile is OR-ed into the
            // plane, which is zeroed before the launch (bits outside the tile are 0 in s_hi)
            const uint32_t nh = (padp + tile_total + 31) / 32;
            for (uint32_t v = threadIdx.x; v < nh; v += RT) {
                const uint32_t word = s_hi[v];
                if (32 * v >= padp && 32 * v + 32 <= padp + tile_total) pk.hi[(gp >> 5) + v] = word;
                else if (word) atomicOr(&pk.hi[(gp >> 5) + v], word);
            }
        }
        if (last_tile && threadIdx.x == 0) *d_R = goff + tile_total;
        return;
    }
    const bool vec_ok = goff + tile_total <= cap && (reinterpret_cast<uintptr_t>(count) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(rsym) & 15) == 0;
    if (vec_ok) {
        const uint32_t nvc = (padc + tile_total + 3) / 4;
        uint32_t *gc = count + (goff - padc);
        for (uint32_t v = threadIdx.x; v < nvc; v += RT) {
            const uint32_t j0 = 4 * v;
            if (j0 >= padc && j0 + 4 <= padc + tile_total) {
                *reinterpret_cast<uint4 *>(gc + j0) = *reinterpret_cast<const uint4 *>(s_cnt_raw + (j0 ^ (((j0 >> 5) & 3u) << 2)));
            } else {
                for (uint32_t j = j0; j < j0 + 4; j++)
                    if (j >= padc && j < padc + tile_total) gc[j] = s_cnt_raw[j ^ (((j >> 5) & 3u) << 2)];
            }
        }
        const uint32_t nvs = (pads + tile_total + 7) / 8;
        int16_t *gs = rsym + (goff - pads);
        for (uint32_t v = threadIdx.x; v < nvs; v += RT) {
            const uint32_t j0 = 8 * v;
            if (j0 >= pads && j0 + 8 <= pads + tile_total) {
                *reinterpret_cast<uint4 *>(gs + j0) = *reinterpret_cast<const uint4 *>(s_sym_raw + (j0 ^ (((j0 >> 6) & 3u) << 3)));
            } else {
                for (uint32_t j = j0; j < j0 + 8; j++)
                    if (j >= pads && j < pads + tile_total) gs[j] = s_sym_raw[j ^ (((j >> 6) & 3u) << 3)];
            }
        }
    } else {
        for (uint32_t j = threadIdx.x; j < tile_total; j += RT) {
            uint64_t g = goff + j;
            if (g < cap) {
                count[g] = s_cnt_raw[pc(j)];
                rsym[g] = s_sym_raw[ps(j)];
            }
        }
    }
    if (last_tile && threadIdx.x == 0) *d_R = goff + tile_total;
}

template <class In, bool PACKED>
__global__ void __launch_bounds__(RT)
    rle_emit_kernel(In in, uint64_t N, const uint64_t *__restrict__ tile_off, const uint32_t *__restrict__ tile_headx,
                    const uint32_t *__restrict__ tile_justx, uint32_t *__restrict__ count, int16_t *__restrict__ rsym,
                    uint64_t cap, uint64_t *__restrict__ d_R, PackArgs pk) {
    uint32_t tt, th;
    rle_emit_tile<In, PACKED>(in, N, N, (uint64_t)blockIdx.x * RT * In::ITEMS, tile_off[blockIdx.x], tile_headx[blockIdx.x],
                              In::HAS_NOTHING ? tile_justx[blockIdx.x] : 0u, count, rsym, cap, d_R, pk,
                              blockIdx.x == gridDim.x - 1, &tt, &th);
}

// ---- runs of an MTF index stream whose statistics the MTF replay kernel collected (mtf.cu RunStat) --------------
// The scan over the tile records (common.cuh runstat_scan) normally runs in the last CTA of the replay kernel; this
// kernel is the stand-alone form.
__global__ void __launch_bounds__(1024)
    rle_tstat_scan_kernel(const uint4 *__restrict__ tstat, uint64_t ntiles, uint32_t tile_syms, uint64_t *__restrict__ toff,
                          uint32_t *__restrict__ theadx) {
    runstat_scan<1024>(tstat, ntiles, tile_syms, toff, theadx);
}

// CTA per MTF tile: its tile_syms positions in sub-tiles of RT * ITEMS, run offset and latest head carried along
template <class In, bool PACKED>
__global__ void __launch_bounds__(RT)
    rle_emit_tiled_kernel(In in, uint64_t N, uint32_t tile_syms, const uint64_t *__restrict__ toff,
                          const uint32_t *__restrict__ theadx, uint32_t *__restrict__ count, int16_t *__restrict__ rsym,
                          uint64_t cap, uint64_t *__restrict__ d_R, PackArgs pk) {
    constexpr uint32_t SUB = RT * In::ITEMS;
    const uint64_t t_beg = (uint64_t)blockIdx.x * tile_syms, t_end = t_beg + tile_syms < N ? t_beg + tile_syms : N;
    uint64_t goff = toff[blockIdx.x];
    uint32_t H = theadx[blockIdx.x];
    for (uint64_t sb = t_beg; sb < t_end; sb += SUB) {
        uint32_t tt, th;
        // (only the sub-tile that holds position N - 1 flushes the last run and writes R; the packed path returns
        // from the tile function early, so the barrier below is reached by all threads either way)
        rle_emit_tile<In, PACKED>(in, N, t_end, sb, goff, H, 0u, count, rsym, cap, d_R, pk, t_end == N && sb + SUB >= N, &tt,
                                  &th);
        goff += tt;
        H = th;
        __syncthreads();
    }
}

// ---- packed run payload (block container) ------------------------------------
// Runs leave the device as 2 bytes + 1 bit each instead of the 6-byte record:
//   cnt4: min(count - 1, 15) in four bits per run; sym8[k] = low byte of the symbol; hi bit k = bit 8 of the
//   symbol's 9-bit code (set for MTF index 256 and for Nothing, whose code is 0x1ff);
//   counts >= 16 are listed as (run index, count) exceptions, appended in arbitrary order by
//   the emit kernel and sorted by run index before they leave the device.
// rle_emit_kernel<In, true> stages this form in shared memory and writes it straight out (the
// 6-byte records are never written): byte streams in aligned groups of 16 runs, the hi plane in
// whole 32-run words, with atomicOr on the (zeroed) words a tile shares with its neighbours.

// inverse of the packing for the device-side decoder: one thread per run; exceptions patched after
__global__ void rle_unpack_kernel(const uint8_t *__restrict__ cnt4, const uint8_t *__restrict__ sym8,
                                  const uint32_t *__restrict__ hi, uint64_t R, uint32_t *__restrict__ count,
                                  int16_t *__restrict__ rsym) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= R) return;
    const uint32_t code = sym8[k] | (((hi[k >> 5] >> (k & 31)) & 1u) << 8);
    count[k] = ((cnt4[k >> 1] >> (4 * (k & 1))) & 15u) + 1u;
    rsym[k] = code == 0x1ffu ? (int16_t)-1 : (int16_t)code;
}
__global__ void rle_unpack_big_kernel(const uint64_t *__restrict__ big_idx, const uint32_t *__restrict__ big_cnt,
                                      uint64_t n_big, uint64_t R, uint32_t *__restrict__ count) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_big && big_idx[j] < R) count[big_idx[j]] = big_cnt[j];
}

// exceptions in run order: LSD radix sort of (run index, count) on the 32 index bits
int rle_sort_big(tc_ctx *ctx, RlePack *pk) {
    WsMark mk = tc_ws_mark(ctx);
    uint64_t *k1, *ko;
    uint32_t *v1, *vo;
    TC_TRY(ws_alloc(ctx, pk->n_big, &k1));
    TC_TRY(ws_alloc(ctx, pk->n_big, &v1));
    static const int shifts[4] = {0, 8, 16, 24};
    TC_TRY(tc_radix_sort_pairs(ctx, pk->big_idx, pk->big_cnt, k1, v1, pk->n_big, shifts, 4, &ko, &vo));
    if (ko != pk->big_idx) {
        TC_CUDA(cudaMemcpyAsync(pk->big_idx, ko, pk->n_big * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        TC_CUDA(cudaMemcpyAsync(pk->big_cnt, vo, pk->n_big * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    tc_ws_release(ctx, mk);
    return TC_OK;
}

template <class In>
int rle_encode_impl(tc_ctx *ctx, In in, uint64_t N, uint32_t *d_count, int16_t *d_rsym, uint64_t cap, uint64_t *R,
                    RlePack *pk = nullptr, const MtfRleLink *link = nullptr) {
    *R = 0;
    if (N == 0) return TC_OK;
    if (N >= 0xfffffffeull) return TC_E_TOOBIG;
    constexpr uint64_t TILE = (uint64_t)RT * In::ITEMS;
    const bool tiled = link && link->valid && !In::HAS_NOTHING; // the MTF stage counted the runs of its tiles already
    uint64_t tiles = tiled ? link->ntiles : ceil_div_u64(N, TILE);
    WsMark mk = tc_ws_mark(ctx);
    uint32_t *tp, *th, *tj;
    uint64_t *toff, *d_R;
    TC_TRY(ws_alloc(ctx, tiles, &tp));
    TC_TRY(ws_alloc(ctx, tiles, &th));
    TC_TRY(ws_alloc(ctx, tiles, &tj));
    TC_TRY(ws_alloc(ctx, tiles, &toff));
    const bool shared = link && link->d_R; // result words shared with the MTF stage's final list (impl.cuh)
    if (shared)
        d_R = link->d_R;
    else
        TC_TRY(ws_alloc(ctx, 2, &d_R));
    if (tiled && link->scanned) { // the replay kernel's last CTA has scanned the tile records already
        toff = link->d_toff;
        th = link->d_theadx;
    } else if (tiled) {
        TC_LAUNCH(ctx, rle_tstat_scan_kernel, 1, 1024, 0, link->d_tstat, tiles, link->tile_syms, toff, th);
    } else {
        TC_LAUNCH(ctx, (rle_reduce_kernel<In>), (unsigned)tiles, RT, 0, in, N, tp, th, tj);
        // NB: the reduce kernel's tile pair counts exclude the final flush; the emit kernel adds it
        // for the owner of position N-1, which is always in the last tile, so offsets stay exact.
        TC_LAUNCH(ctx, rle_tile_scan_kernel, 1, 1024, 0, tp, th, tj, toff, tiles);
    }
    PackArgs pa{};
    if (pk) {
        // d_R and the exception counter sit next to each other so that one small copy brings both back
        pk->n_big = 0;
        TC_CUDA(cudaMemsetAsync(d_R + 1, 0, sizeof(uint64_t), ctx->stream));
        TC_CUDA(cudaMemsetAsync(pk->hi, 0, ceil_div_u64(cap, 32) * sizeof(uint32_t), ctx->stream));
        TC_CUDA(cudaMemsetAsync(pk->cnt4, 0, ceil_div_u64(cap, 32) * 16, ctx->stream));
        pa = PackArgs{pk->cnt4, pk->sym8, pk->hi, pk->big_idx, pk->big_cnt, pk->big_cap, (unsigned long long *)(d_R + 1)};
    }
    ctx->prof_bytes_next = N * sizeof(*in.p);
    if (tiled) {
        if (pk)
            TC_LAUNCH(ctx, (rle_emit_tiled_kernel<In, true>), (unsigned)tiles, RT, 0, in, N, link->tile_syms, toff, th, d_count,
                      d_rsym, cap, d_R, pa);
        else
            TC_LAUNCH(ctx, (rle_emit_tiled_kernel<In, false>), (unsigned)tiles, RT, 0, in, N, link->tile_syms, toff, th, d_count,
                      d_rsym, cap, d_R, pa);
    } else if (pk) {
        TC_LAUNCH(ctx, (rle_emit_kernel<In, true>), (unsigned)tiles, RT, 0, in, N, toff, th, tj, d_count, d_rsym, cap, d_R,
                  pa);
    } else {
        TC_LAUNCH(ctx, (rle_emit_kernel<In, false>), (unsigned)tiles, RT, 0, in, N, toff, th, tj, d_count, d_rsym, cap,
                  d_R, pa);
    }
    TC_TRY(tc_d2h_small(ctx, ctx->h_scal, d_R, (shared ? MtfRleLink::SMALL_WORDS : pk ? 2 : 1) * sizeof(uint64_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    *R = ctx->h_scal[0];
    tc_ws_release(ctx, mk);
    if (pk) {
        pk->n_big = ctx->h_scal[1];
        if (pk->n_big > pk->big_cap) return TC_E_CAP;
        if (pk->n_big > 1) TC_TRY(rle_sort_big(ctx, pk));
    }
    return *R > cap ? TC_E_CAP : TC_OK;
}

// ---- decode ------------------------------------------------------------------
__global__ void rle_len_kernel(const uint32_t *__restrict__ count, const int16_t *__restrict__ rsym, uint64_t R,
                               uint32_t *__restrict__ len) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < R) len[k] = rsym[k] < 0 ? 1u : count[k]; // (Just _, Nothing) -> one Nothing, count ignored
}

constexpr int DT = 256;
constexpr int DITEMS = 8;
__global__ void __launch_bounds__(DT)
    rle_expand_kernel(const uint64_t *__restrict__ off, const int16_t *__restrict__ rsym, uint64_t R,
                      const uint64_t *__restrict__ d_total, int16_t *__restrict__ out, uint64_t cap) {
    uint64_t total = *d_total;
    uint64_t lim = total < cap ? total : cap;
    uint64_t o0 = ((uint64_t)blockIdx.x * DT + threadIdx.x) * DITEMS;
    if (o0 >= lim) return;
    // last k with off[k] <= o0
    uint64_t lo = 0, hi = R; // invariant: off[lo] <= o0, (hi == R or off[hi] > o0)
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= o0) lo = mid; else hi = mid;
    }
    uint64_t k = lo;
    uint64_t end = (k + 1 < R) ? off[k + 1] : total;
    int16_t s = rsym[k];
    int16_t v[DITEMS];
#pragma unroll
    for (int j = 0; j < DITEMS; j++) {
        uint64_t o = o0 + j;
        if (o < lim) {
            while (o >= end) {
                k++;
                end = (k + 1 < R) ? off[k + 1] : total;
                s = rsym[k];
            }
        }
        v[j] = s;
    }
    if (o0 + DITEMS <= lim && ((reinterpret_cast<uintptr_t>(out + o0) & 15) == 0)) {
        uint4 w;
        w.x = (uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
        w.y = (uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
        w.z = (uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
        w.w = (uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
        st_stream_u4(out + o0, w);
    } else {
#pragma unroll
        for (int j = 0; j < DITEMS; j++)
            if (o0 + j < lim) out[o0 + j] = v[j];
    }
}
} // namespace

// internal: device pointers
int rle_decode_dev_impl(tc_ctx *ctx, const uint32_t *d_count, const int16_t *d_rsym, uint64_t R, int16_t *d_sym,
                           uint64_t cap, uint64_t *N_out) {
    *N_out = 0;
    if (R == 0) return TC_OK;
    WsMark mk = tc_ws_mark(ctx);
    uint32_t *len;
    uint64_t *off, *d_total;
    TC_TRY(ws_alloc(ctx, R, &len));
    TC_TRY(ws_alloc(ctx, R, &off));
    TC_TRY(ws_alloc(ctx, 1, &d_total));
    TC_LAUNCH(ctx, rle_len_kernel, (unsigned)ceil_div_u64(R, 256), 256, 0, d_count, d_rsym, R, len);
    TC_TRY(tc_scan_exclusive_u32_to_u64(ctx, len, off, R, d_total));
    TC_TRY(tc_d2h_small(ctx, ctx->h_scal, d_total, sizeof(uint64_t)));
    TC_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t total = ctx->h_scal[0];
    *N_out = total;
    uint64_t lim = total < cap ? total : cap;
    if (lim > 0) {
        uint64_t blocks = ceil_div_u64(lim, (uint64_t)DT * DITEMS);
        TC_LAUNCH(ctx, rle_expand_kernel, (unsigned)blocks, DT, 0, off, d_rsym, R, d_total, d_sym, cap);
    }
    tc_ws_release(ctx, mk);
    return total > cap ? TC_E_CAP : TC_OK;
}

int rle_encode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint32_t *d_count,
                           int16_t *d_rsym, uint64_t cap, uint64_t *R, RlePack *pk) {
    if (N && primary >= N) primary = ~0ull; // no Nothing in range: plain byte stream
    return rle_encode_impl(ctx, InU8{d_bwt, primary}, N, d_count, d_rsym, cap, R, pk);
}
int rle_encode_u16_dev_impl(tc_ctx *ctx, const uint16_t *d_idx, uint64_t N, uint32_t *d_count, int16_t *d_rsym,
                            uint64_t cap, uint64_t *R, RlePack *pk, const MtfRleLink *link) {
    return rle_encode_impl(ctx, In16<false>{d_idx}, N, d_count, d_rsym, cap, R, pk, link);
}
int rle_unpack_dev_impl(tc_ctx *ctx, const uint8_t *d_cnt4, const uint8_t *d_sym8, const uint32_t *d_hi,
                        const uint64_t *d_big_idx, const uint32_t *d_big_cnt, uint64_t n_big, uint64_t R,
                        uint32_t *d_count, int16_t *d_rsym) {
    if (R == 0) return TC_OK;
    TC_LAUNCH(ctx, rle_unpack_kernel, (unsigned)ceil_div_u64(R, 256), 256, 0, d_cnt4, d_sym8, d_hi, R, d_count, d_rsym);
    if (n_big)
        TC_LAUNCH(ctx, rle_unpack_big_kernel, (unsigned)ceil_div_u64(n_big, 256), 256, 0, d_big_idx, d_big_cnt, n_big, R,
                  d_count);
    return TC_OK;
}
int rle_encode_i16_dev_impl(tc_ctx *ctx, const int16_t *d_sym, uint64_t N, uint32_t *d_count, int16_t *d_rsym,
                               uint64_t cap, uint64_t *R) {
    return rle_encode_impl(ctx, In16<true>{(const uint16_t *)d_sym}, N, d_count, d_rsym, cap, R);
}
