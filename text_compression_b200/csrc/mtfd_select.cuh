// mtfd_select.cuh -- MTF decode step for alphabets up to 257 symbols whose cost does not depend on the index:
// the encoder's "latest occurrence" bitmap (mtf.cu, thread-per-chunk replay) run backwards.
//
// A chunk's list is kept as TIME SLOTS: slot s holds the entry that was moved to the front s-th; a bit per slot
// says whether that is still the entry's latest position.  The list, front to back, is the live slots from the
// highest down.  Decoding index r = find the (r+1)-th live slot from the top (select), read its entry, clear the
// bit, set the bit of the next free slot and store the entry there.  512 slots (16 words) cover an incoming list of
// up to 288 entries plus 224 moves.  Per-word live counts sit in four registers, one byte per word, most recent word
// in the low byte, so "which word" is a byte-wise prefix sum (one multiply) and a byte-parallel compare; "which bit"
// is the same trick on the byte popcounts of the word and a 2 KB table for the last 8 bits.
//
// The step is written once, over an abstract state (word / byte loads and stores), so that the CPU test suite runs the
// very same code on plain arrays (tests/c/mtfd_select_test.cpp) and the kernels run it on shared memory.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define D3_HD __host__ __device__ __forceinline__
#else
#define D3_HD inline
#endif

namespace d3 {
constexpr int WORDS = 16;             // 512 slots
constexpr int SLOTS = 32 * WORDS;
constexpr int LMAX = 224;             // moves per chunk: 288 + 224 = 512
constexpr uint32_t NO_SLOT = 0xffffffffu;

D3_HD uint32_t popc(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
D3_HD uint32_t ffs_index(uint32_t x) { // index of the lowest set bit; x != 0
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
D3_HD uint32_t bswap(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// sel8[y * 8 + j] = bit position (0..7) of the (j+1)-th set bit of the byte y counted from bit 7 down
D3_HD uint8_t sel8_entry(uint32_t y, uint32_t j) {
    for (int b = 7; b >= 0; b--)
        if ((y >> b) & 1u) {
            if (j == 0) return (uint8_t)b;
            j--;
        }
    return 0;
}

// Registers of one chunk.
struct Regs {
    uint32_t c[4];    // live bits per word: word w is byte (15 - w) & 3 of c[(15 - w) >> 2]
    uint32_t ow;      // copy of the bitmap word that holds slot `top`
    uint32_t top;     // next free slot
    uint32_t special; // slot of the one entry that does not fit a byte (id 256 / Nothing), or NO_SLOT
};

// first byte b (0..3) whose inclusive prefix P_b exceeds r; bytes of P <= 128, r <= 127
D3_HD uint32_t first_byte_above(uint32_t P, uint32_t r) {
    const uint32_t m = (P + (0x7f7f7f7fu - r * 0x01010101u)) & 0x80808080u;
    return ffs_index(m) >> 3;
}

// State: bm_load(w), bm_store(w, x), sym_load(slot), sym_store(slot, v), sel8(idx)
// Returns the entry (0..255, or 256 for the special one) at list position r and moves it to the front.
// r must be below the number of live slots.
template <class State>
D3_HD uint32_t take(State &S, Regs &R, uint32_t r) {
    // ---- which word (branch-free: the compiler turns if / else chains here into divergent code)
    const uint32_t P0 = R.c[0] * 0x01010101u, P1 = R.c[1] * 0x01010101u, P2 = R.c[2] * 0x01010101u,
                   P3 = R.c[3] * 0x01010101u;
    const uint32_t s0 = P0 >> 24, s1 = s0 + (P1 >> 24), s2 = s1 + (P2 >> 24);
    const bool a0 = r >= s0, a1 = r >= s1, a2 = r >= s2;
    const uint32_t g = (uint32_t)a0 + (uint32_t)a1 + (uint32_t)a2;
    const uint32_t base = a2 ? s2 : a1 ? s1 : a0 ? s0 : 0u;
    const uint32_t P = a2 ? P3 : a1 ? P2 : a0 ? P1 : P0;
    uint32_t rr = r - base;
    const uint32_t b = first_byte_above(P, rr);
    rr -= ((P << 8) >> (8 * b)) & 0xffu; // live bits in the more recent words of this register
    const uint32_t w = 15u - (4u * g + b);
    const uint32_t wt = R.top >> 5;
    const uint32_t x = S.bm_load(w); // memory is current for every word, the open one included
    // ---- which bit: byte popcounts, most significant byte first
    const uint32_t xr = bswap(x);
    uint32_t v = xr - ((xr >> 1) & 0x55555555u);
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
    v = (v + (v >> 4)) & 0x0f0f0f0fu;
    const uint32_t Q = v * 0x01010101u;
    const uint32_t bb = first_byte_above(Q, rr); // 0 = most significant byte of x
    rr -= ((Q << 8) >> (8 * bb)) & 0xffu;
    const uint32_t y = (xr >> (8 * bb)) & 0xffu;
    const uint32_t bit = 8u * (3u - bb) + S.sel8(y * 8u + rr);
    const uint32_t slot = 32u * w + bit;
    const bool is_special = slot == R.special;
    const uint32_t id = is_special ? 256u : S.sym_load(slot);
    // ---- move to the front: clear the slot, occupy slot `top`.  Straight-line code (an index of 0 takes the same
    // path: the front entry moves from slot top - 1 to slot top), so that consecutive steps can overlap.
    const uint32_t cleared = x & ~(1u << bit);
    S.bm_store(w, cleared);
    const uint32_t ow = (w == wt ? cleared : R.ow) | (1u << (R.top & 31u));
    S.bm_store(wt, ow); // after the store above: wins when both name the open word
    const uint32_t d = 1u << (8 * b);
    const uint32_t ut = 15u - wt, gt = ut >> 2, dt = 1u << (8 * (ut & 3u));
    R.c[0] += (gt == 0 ? dt : 0u) - (g == 0 ? d : 0u);
    R.c[1] += (gt == 1 ? dt : 0u) - (g == 1 ? d : 0u);
    R.c[2] += (gt == 2 ? dt : 0u) - (g == 2 ? d : 0u);
    R.c[3] += (gt == 3 ? dt : 0u) - (g == 3 ? d : 0u);
    S.sym_store(R.top, id & 0xffu);
    R.special = is_special ? R.top : R.special;
    R.top++;
    R.ow = (R.top & 31u) ? ow : 0u; // a full word stays behind in memory, the next one opens empty
    return id;
}

// Incoming list of `sigma` entries: list position p (0 = front) sits in slot sigma - 1 - p.  The caller stores the
// entries with sym_store and names the special slot; this sets bitmap, counts and the open word.
template <class State>
D3_HD void init(State &S, Regs &R, uint32_t sigma) {
    R.top = sigma;
    R.special = NO_SLOT;
    R.c[0] = R.c[1] = R.c[2] = R.c[3] = 0;
    const uint32_t wt = sigma >> 5;
    R.ow = (sigma & 31u) ? (0xffffffffu >> (32u - (sigma & 31u))) : 0u;
    for (uint32_t w = 0; w < (uint32_t)WORDS; w++) {
        S.bm_store(w, w < wt ? 0xffffffffu : w == wt ? R.ow : 0u); // R.ow mirrors word wt (saves the load of a read-modify-write)
        const uint32_t cnt = w < wt ? 32u : w == wt ? (sigma & 31u) : 0u;
        const uint32_t u = 15u - w;
        R.c[u >> 2] += cnt << (8 * (u & 3u));
    }
}

// Visits the list front to back: f(position, entry).  Leaves the state unchanged.
template <class State, class F>
D3_HD void for_each_entry(State &S, const Regs &R, F f) {
    uint32_t j = 0;
    for (int w = WORDS - 1; w >= 0; w--) {
        uint32_t x = S.bm_load((uint32_t)w);
        while (x) {
            const uint32_t bit = 31u - (uint32_t)
#ifdef __CUDA_ARCH__
                                           __clz((int)x);
#else
                                           __builtin_clz(x);
#endif
            x ^= 1u << bit;
            const uint32_t slot = 32u * (uint32_t)w + bit;
            f(j++, slot == R.special ? 256u : S.sym_load(slot));
        }
    }
}
} // namespace d3
