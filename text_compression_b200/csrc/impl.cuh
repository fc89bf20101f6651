// impl.cuh -- device-level implementations (all data pointers are DEVICE pointers; scalar
// results come back through host pointers after a stream sync).  None of these resets the
// scratch arena: only the extern "C" entry points in api.cu do, once per top-level call.
#pragma once
#include "common.cuh"

int bwt_encode_dev_impl(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint8_t *d_bwt, uint64_t *primary,
                        uint32_t *d_sa_1based);
int bwt_decode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint8_t *d_text,
                           uint64_t cap, uint64_t *n_out);
int bwt_decode_i16_dev_impl(tc_ctx *ctx, const int16_t *d_bwt, uint64_t N, uint8_t *d_text, uint64_t cap,
                            uint64_t *n_out);
// Run statistics handed from the MTF stage to the RLE stage of the composed helpers: the replay kernels hold the
// indices in registers, so they count the run boundaries of their warp tiles on the way (mtf.cu RunStat) and the RLE
// stage needs no counting pass over the index stream (rle.cu rle_emit_tiled_kernel).
struct MtfRleLink {
    uint4 *d_tstat = nullptr; // caller-allocated: ceil(N / 4096) records (a tile has at least 4096 symbols)
    uint64_t *d_toff = nullptr;    // caller-allocated, same count: runs before each tile ...
    uint32_t *d_theadx = nullptr;  // ... and the latest run head before it, both written by the replay kernel's last CTA
    uint32_t *d_ticket = nullptr;  // caller-allocated, 2 words (zeroed by the first MTF kernel)
    bool scanned = false;     // d_toff / d_theadx are filled
    uint64_t ntiles = 0;      // set by the MTF stage
    uint32_t tile_syms = 0;   // symbols per tile
    bool valid = false;       // false: the MTF path taken does not collect them
    // Small results of both stages in one device buffer so that ONE small copy at the end of the RLE stage brings
    // them back: d_R[0] = runs, d_R[1] = exception count, then the final MTF list (257 x u16) at d_final.
    uint64_t *d_R = nullptr;       // caller-allocated, SMALL_WORDS 64-bit words; d_final = (uint16_t *)(d_R + H_FINAL)
    uint16_t *d_final = nullptr;
    static constexpr uint32_t H_FINAL = 2, SMALL_WORDS = 2 + 66;
};
int mtf_encode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint16_t *d_idx,
                           int16_t *final_list, uint32_t *sigma, const uint8_t *present_hint = nullptr,
                           MtfRleLink *link = nullptr);
int mtf_finish_pending(tc_ctx *ctx); // see mtf.cu: deferred final list of the composed helpers
void mtf_free_tables(tc_ctx *ctx);   // device tables of the small-alphabet automata
int mtf_encode_i16_dev_impl(tc_ctx *ctx, const int16_t *d_sym, uint64_t N, uint16_t *d_idx, int16_t *final_list,
                            uint32_t *sigma);
int mtf_decode_dev_impl(tc_ctx *ctx, const uint16_t *d_idx, uint64_t N, const int16_t *final_list, uint32_t sigma_in,
                        int16_t *d_sym);
// Packed form of the runs (payload of the block container; layout in rle.cu).
// All pointers are device pointers sized for `cap` runs; n_big comes back on the host.
struct RlePack {
    uint8_t *cnt4 = nullptr, *sym8 = nullptr; // [cap / 2 rounded up to 16], [cap rounded up to 16]
    uint32_t *hi = nullptr;                   // [ceil(cap / 32)]
    uint64_t *big_idx = nullptr;              // [big_cap] runs with count >= 16, sorted by run index
    uint32_t *big_cnt = nullptr;
    uint64_t big_cap = 0;
    uint64_t n_big = 0;
};
int rle_encode_u8_dev_impl(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint32_t *d_count,
                           int16_t *d_rsym, uint64_t cap, uint64_t *R, RlePack *pk = nullptr);
int rle_encode_u16_dev_impl(tc_ctx *ctx, const uint16_t *d_idx, uint64_t N, uint32_t *d_count, int16_t *d_rsym,
                            uint64_t cap, uint64_t *R, RlePack *pk = nullptr, const MtfRleLink *link = nullptr);
int rle_unpack_dev_impl(tc_ctx *ctx, const uint8_t *d_cnt4, const uint8_t *d_sym8, const uint32_t *d_hi,
                        const uint64_t *d_big_idx, const uint32_t *d_big_cnt, uint64_t n_big, uint64_t R,
                        uint32_t *d_count, int16_t *d_rsym);
int rle_encode_i16_dev_impl(tc_ctx *ctx, const int16_t *d_sym, uint64_t N, uint32_t *d_count, int16_t *d_rsym,
                            uint64_t cap, uint64_t *R);
int rle_decode_dev_impl(tc_ctx *ctx, const uint32_t *d_count, const int16_t *d_rsym, uint64_t R, int16_t *d_sym,
                        uint64_t cap, uint64_t *N_out);
int tc_byte_hist_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *h_hist);
int tc_bwt_emit_dev(tc_ctx *ctx, const uint8_t *d_text, const uint32_t *d_sa, uint64_t N, uint8_t *d_bwt,
                    uint64_t *primary);
