/* abi_layout.c -- what the Haskell shim (haskell/src/Data/TextCompression/B200.hs) assumes about
 * include/tc_b200.h, restated in C99 so that it is CHECKED by a compiler even though no GHC exists in the
 * image: the struct offsets the shim reads with peekByteOff (544 / 552 / 16 ...), and every foreign import's
 * argument list as a function-pointer assignment (arity, order, width and pointer-ness must match the header;
 * a mismatch is a compile error under -Werror).  Compiled and run by tests/test_host_logic.py. */
#include <stddef.h>
#include <stdio.h>

#include "tc_b200.h"

#define SA(cond, msg) _Static_assert(cond, msg)

/* tc_block_info: B200.hs reads R at byte 544 and strides arrays of it by 552 */
SA(offsetof(tc_block_info, n) == 0, "tc_block_info.n");
SA(offsetof(tc_block_info, N) == 8, "tc_block_info.N");
SA(offsetof(tc_block_info, primary) == 16, "tc_block_info.primary");
SA(offsetof(tc_block_info, sigma) == 24, "tc_block_info.sigma");
SA(offsetof(tc_block_info, final_list) == 28, "tc_block_info.final_list");
SA(offsetof(tc_block_info, R) == 544, "tc_block_info.R (B200.hs: peekByteOff pinfo 544)");
SA(sizeof(tc_block_info) == 552, "sizeof tc_block_info (B200.hs: pinned (552 * nb))");
/* tc_packed_header: decodePackedW8 reads n at byte 16 */
SA(offsetof(tc_packed_header, magic) == 0, "tc_packed_header.magic");
SA(offsetof(tc_packed_header, n) == 16, "tc_packed_header.n (B200.hs: peekByteOff p 16)");
SA(offsetof(tc_packed_header, R) == 40, "tc_packed_header.R");
SA(offsetof(tc_packed_header, off_cnt4) == 64, "tc_packed_header.off_cnt4");
SA(offsetof(tc_packed_header, final_list) == 112, "tc_packed_header.final_list");
SA(sizeof(tc_packed_header) == 640, "sizeof tc_packed_header");
/* tc_fm_info (text_compression_b200/_lib.py FmInfo mirrors it) */
SA(offsetof(tc_fm_info, alphabet) == 32, "tc_fm_info.alphabet");
SA(offsetof(tc_fm_info, C) == 552, "tc_fm_info.C");
SA(offsetof(tc_fm_info, blob_bytes) == 2608, "tc_fm_info.blob_bytes");
SA(sizeof(tc_fm_info) == 2624, "sizeof tc_fm_info");
/* Haskell Int is 64-bit on the platforms the shim targets */
SA(sizeof(uint64_t) == 8 && sizeof(size_t) == 8 && sizeof(int) == 4, "LP64");

/* every `foreign import ccall` of B200.hs: Ptr x -> pointer, Word64 / Int64 / CSize -> 64-bit, CInt / Word32 -> 32-bit */
#define IMPORT(name, ret, args) \
    do {                        \
        ret(*f) args = name;    \
        (void)f;                \
    } while (0)

int main(void) {
    IMPORT(tc_ctx_pool_acquire, int, (int, tc_ctx **));
    IMPORT(tc_ctx_pool_release, void, (tc_ctx *));
    IMPORT(tc_device_count, int, (void));
    IMPORT(tc_strerror, const char *, (int));
    IMPORT(tc_last_error, const char *, (const tc_ctx *));
    IMPORT(tc_host_alloc, void *, (size_t));
    IMPORT(tc_host_free, void, (void *));
    IMPORT(tc_bwt_encode, int, (tc_ctx *, const uint8_t *, uint64_t, uint8_t *, uint64_t *, uint32_t *));
    IMPORT(tc_bwt_decode, int, (tc_ctx *, const int16_t *, uint64_t, uint8_t *, uint64_t, uint64_t *));
    IMPORT(tc_mtf_encode, int, (tc_ctx *, const int16_t *, uint64_t, uint16_t *, int16_t *, uint32_t *));
    IMPORT(tc_mtf_decode, int, (tc_ctx *, const uint16_t *, uint64_t, const int16_t *, uint32_t, int16_t *));
    IMPORT(tc_rle_encode, int, (tc_ctx *, const int16_t *, uint64_t, uint32_t *, int16_t *, uint64_t, uint64_t *));
    IMPORT(tc_rle_decode, int, (tc_ctx *, const uint32_t *, const int16_t *, uint64_t, int16_t *, uint64_t, uint64_t *));
    IMPORT(tc_bwt_mtf_rle_encode, int,
           (tc_ctx *, const uint8_t *, uint64_t, uint32_t *, int16_t *, uint64_t, tc_block_info *));
    IMPORT(tc_blocks_encode, int,
           (tc_ctx *, uint64_t, const uint8_t *const *, const uint64_t *, int, uint32_t *const *, int16_t *const *,
            const uint64_t *, tc_block_info *));
    IMPORT(tc_packed_bound, uint64_t, (uint64_t));
    IMPORT(tc_blocks_encode_packed, int,
           (tc_ctx *, uint64_t, const uint8_t *const *, const uint64_t *, int, uint8_t *const *, const uint64_t *,
            uint64_t *, tc_block_info *));
    IMPORT(tc_packed_unpack, int, (const void *, uint64_t, uint32_t *, int16_t *, uint64_t, tc_block_info *));
    IMPORT(tc_bwt_decode_u8, int, (tc_ctx *, const uint8_t *, uint64_t, uint64_t, uint8_t *, uint64_t, uint64_t *));
    IMPORT(tc_mtf_encode_u8, int, (tc_ctx *, const uint8_t *, uint64_t, uint64_t, uint16_t *, int16_t *, uint32_t *));
    IMPORT(tc_packed_decode, int, (tc_ctx *, const void *, uint64_t, uint8_t *, uint64_t, uint64_t *));
    IMPORT(tc_blocks_decode_packed, int, (tc_ctx *, uint64_t, const void *const *, const uint64_t *, uint8_t *const *, const uint64_t *, uint64_t *));
    IMPORT(tc_fm_build, int, (tc_ctx *, const uint8_t *, uint64_t, uint32_t, tc_fm **));
    IMPORT(tc_fm_free, void, (tc_fm *));
    IMPORT(tc_fm_count, int, (tc_ctx *, const tc_fm *, const uint8_t *, const uint64_t *, uint64_t, int64_t *));
    IMPORT(tc_fm_locate, int,
           (tc_ctx *, const tc_fm *, const uint8_t *, const uint64_t *, uint64_t, uint64_t *, uint64_t *, uint64_t,
            uint64_t *));
    IMPORT(tc_mgpu_blocks_encode_packed, int,
           (int, const int *, uint64_t, const uint8_t *const *, const uint64_t *, int, uint8_t *const *,
            const uint64_t *, uint64_t *, tc_block_info *));
    IMPORT(tc_fm_replicate, int, (const tc_fm *, int, const int *, tc_fm **));
    IMPORT(tc_mgpu_fm_count, int,
           (int, const int *, tc_fm *const *, const uint8_t *, const uint64_t *, uint64_t, int64_t *));
    IMPORT(tc_mgpu_fm_locate, int,
           (int, const int *, tc_fm *const *, const uint8_t *, const uint64_t *, uint64_t, uint64_t *, uint64_t *,
            uint64_t, uint64_t *));
    puts("abi layout ok");
    return 0;
}
