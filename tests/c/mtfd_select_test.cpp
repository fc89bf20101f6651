// CPU run of the MTF decode step the kernels use (text_compression_b200/csrc/mtfd_select.cuh) on plain arrays,
// against a list that is shifted by hand.  Built and run by tests/test_host_logic.py (no GPU needed).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "mtfd_select.cuh"

struct HostState {
    uint32_t bm[d3::WORDS];
    uint8_t sym[d3::SLOTS];
    uint8_t lut[256 * 8];
    uint32_t bm_load(uint32_t w) const { return bm[w]; }
    void bm_store(uint32_t w, uint32_t x) { bm[w] = x; }
    uint32_t sym_load(uint32_t s) const { return sym[s]; }
    void sym_store(uint32_t s, uint32_t v) { sym[s] = (uint8_t)v; }
    uint32_t sel8(uint32_t i) const { return lut[i]; }
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13, rng_state ^= rng_state >> 7, rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}

int main() {
    HostState S;
    for (uint32_t y = 0; y < 256; y++)
        for (uint32_t j = 0; j < 8; j++) S.lut[y * 8 + j] = d3::sel8_entry(y, j);
    long checked = 0;
    for (int trial = 0; trial < 4000; trial++) {
        const uint32_t sigma = trial < 300 ? 257u - (trial % 3) : 9u + rnd() % 249u; // 9..257
        const uint32_t L = 1u + rnd() % d3::LMAX;
        const int mode = trial % 5; // 0 uniform, 1 mostly small, 2 mostly 0, 3 always the back, 4 alternating far / 0
        // entries: ids 0..sigma-1 in some order; id 256 (special) stands in for id 0 on odd trials when sigma == 257 - ...
        std::vector<uint32_t> list(sigma);
        for (uint32_t j = 0; j < sigma; j++) list[j] = j;
        for (uint32_t j = sigma - 1; j > 0; j--) {
            uint32_t k = rnd() % (j + 1), t = list[j];
            list[j] = list[k], list[k] = t;
        }
        const bool with_special = (trial & 1) != 0 && sigma < 257; // at sigma = 257 the id 256 is there anyway
        if (with_special) {
            for (uint32_t j = 0; j < sigma; j++)
                if (list[j] == 0) list[j] = 256; // one entry that does not fit a byte
        }
        d3::Regs R;
        d3::init(S, R, sigma);
        for (uint32_t p = 0; p < sigma; p++) {
            const uint32_t slot = sigma - 1 - p;
            if (list[p] == 256) R.special = slot, S.sym_store(slot, 0);
            else S.sym_store(slot, list[p]);
        }
        for (uint32_t i = 0; i < L; i++) {
            uint32_t r;
            switch (mode) {
            case 0: r = rnd() % sigma; break;
            case 1: r = (rnd() % 8 == 0) ? rnd() % sigma : rnd() % 4; break;
            case 2: r = (rnd() % 16 == 0) ? rnd() % sigma : 0; break;
            case 3: r = sigma - 1; break;
            default: r = (i & 1) ? 0 : sigma - 1 - (rnd() % 3 < 1 ? 0 : rnd() % sigma % (sigma - 1)); break;
            }
            if (r >= sigma) r = sigma - 1;
            const uint32_t want = list[r];
            for (uint32_t j = r; j > 0; j--) list[j] = list[j - 1];
            list[0] = want;
            const uint32_t got = d3::take(S, R, r);
            if (got != want) {
                printf("FAIL trial %d step %u: sigma %u r %u got %u want %u\n", trial, i, sigma, r, got, want);
                return 1;
            }
            checked++;
        }
        uint32_t bad = 0, seen = 0;
        d3::for_each_entry(S, R, [&](uint32_t j, uint32_t e) {
            seen++;
            if (j >= sigma || list[j] != e) {
                if (!bad) printf("  pos %u: got %u want %u (sigma %u L %u mode %d special %u top %u)\n", j, e, j < sigma ? list[j] : 999u, sigma, L, mode, R.special, R.top);
                bad++;
            }
        });
        if (bad || seen != sigma) {
            printf("FAIL trial %d: final list differs (%u bad, %u entries of %u)\n", trial, bad, seen, sigma);
            return 1;
        }
    }
    printf("ok %ld steps\n", checked);
    return 0;
}
