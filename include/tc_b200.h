/*
 * tc_b200.h -- C ABI of libtc_b200.so: the B200 (sm_100a) implementation of the
 * text-compression hot path (Data.BWT, Data.MTF, Data.RLE, Data.FMIndex).
 *
 * The reference (Matthew-Mosior/text-compression v0.1.0.25) is pure Haskell and has
 * no FFI today; the boundary a maintainer binds with `foreign import ccall` is the
 * "Internal" function set.  Each entry point below names the reference function it
 * replaces (paths relative to the reference root).  See INTEGRATION.md for the
 * Haskell-side stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns TC_OK (0) or a negative
 *     TC_E_* code and never aborts.  TC_E_FROMJUST / TC_E_INDEX mark the inputs on
 *     which the reference itself throws (`fromJust Nothing`, `DS.index` out of range).
 *   - text symbols are uint8_t.  A `Seq (Maybe b)` is int16_t with -1 == Nothing
 *     (the "$" sentinel, smaller than every Just) and 0..255 == Just byte.
 *   - the `_u8` twins carry a BWT as uint8_t[N] plus `primary`, the 0-based slot
 *     that holds the single Nothing (the byte stored in that slot is ignored).
 *   - Haskell `Int` results are int64_t/uint64_t; ranks and positions are 1-based
 *     wherever the reference's are.
 *   - host entry points take HOST pointers (pageable or tc_host_alloc'ed) and do the
 *     H2D/D2H copies themselves; `_dev` entry points take DEVICE pointers on the
 *     context's device and enqueue on the context's stream.
 *   - one tc_ctx per calling OS thread; distinct contexts are fully concurrent.
 *   - sizes: n < 2^32 - 2 for the BWT / RLE / FM-index entry points, and N = n + 1 < 2^31 - 1 wherever MTF
 *     takes part (tc_mtf_encode*, tc_bwt_mtf_rle_encode*, tc_blocks_encode* with with_mtf != 0): recency
 *     keys are 32-bit distances.  TC_E_TOOBIG otherwise, checked before any work is queued.
 *   - there is NO CPU fallback: without a usable CUDA device tc_ctx_create fails
 *     with TC_E_NODEVICE and nothing else can be called.
 */
#ifndef TC_B200_H
#define TC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TC_OK 0
#define TC_E_CUDA (-1)      /* CUDA runtime error; tc_last_error(ctx) has the text   */
#define TC_E_CAP (-2)       /* output capacity too small; required size is returned  */
#define TC_E_FROMJUST (-3)  /* the reference would throw `fromJust Nothing`           */
#define TC_E_INDEX (-4)     /* the reference would throw `index out of bounds`       */
#define TC_E_NOMEM (-5)
#define TC_E_ARG (-6)
#define TC_E_TOOBIG (-7)
#define TC_E_NODEVICE (-8)

typedef struct tc_ctx tc_ctx; /* device + stream + scratch arena */
typedef struct tc_fm tc_fm;   /* device-resident FM-index */

/* ---- context ------------------------------------------------------------- */
int tc_ctx_create(int device, tc_ctx **out);
/* Same, but enqueue on an existing cudaStream_t (e.g. torch's current stream). */
int tc_ctx_create_on_stream(int device, void *cuda_stream, tc_ctx **out);
void tc_ctx_destroy(tc_ctx *ctx);
int tc_ctx_sync(tc_ctx *ctx);
const char *tc_strerror(int rc);
const char *tc_last_error(const tc_ctx *ctx);
/* CUDA-pinned staging memory for the host entry points. */
void *tc_host_alloc(size_t bytes);
void tc_host_free(void *p);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t tc_ctx_launches(const tc_ctx *ctx);
const char *tc_version(void);
/* Per-kernel device timing for bench.py's roofline: on != 0 records one CUDA-event pair per
 * launch on the context's stream; the report is "name\tlaunches\ttotal_ms\n" lines. */
int tc_ctx_profile(tc_ctx *ctx, int on);
int tc_ctx_profile_report(tc_ctx *ctx, char *buf, size_t cap);

/* ---- Data.BWT -------------------------------------------------------------- */
/* createSuffixArray + saToBWT + toBWT (src/Data/BWT/Internal.hs:98-134,
 * src/Data/BWT.hs:55-64).  bwt[N=n+1]; *primary = slot of the Nothing;
 * sa_1based (nullable) = suffixstartpos per rank.  n == 0: BWT Empty, *primary = 0. */
int tc_bwt_encode(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint8_t *bwt, uint64_t *primary,
                  uint32_t *sa_1based);
/* fromBWT = sortTB + magicInverseBWT (src/Data/BWT.hs:93-104,
 * src/Data/BWT/Internal.hs:144-200) on an arbitrary Seq (Maybe Word8). */
int tc_bwt_decode(tc_ctx *ctx, const int16_t *bwt, uint64_t N, uint8_t *text, uint64_t cap, uint64_t *n_out);
int tc_bwt_decode_u8(tc_ctx *ctx, const uint8_t *bwt, uint64_t N, uint64_t primary, uint8_t *text, uint64_t cap,
                     uint64_t *n_out);

/* ---- Data.MTF -------------------------------------------------------------- */
/* seqToMTF (src/Data/MTF/Internal.hs:128-175): indices + the FINAL list. */
int tc_mtf_encode(tc_ctx *ctx, const int16_t *sym, uint64_t N, uint16_t *idx, int16_t *final_list /*257*/,
                  uint32_t *sigma);
int tc_mtf_encode_u8(tc_ctx *ctx, const uint8_t *bwt, uint64_t N, uint64_t primary, uint16_t *idx,
                     int16_t *final_list /*257*/, uint32_t *sigma);
/* seqFromMTF (src/Data/MTF/Internal.hs:201-232): initial list = sort(final list). */
int tc_mtf_decode(tc_ctx *ctx, const uint16_t *idx, uint64_t N, const int16_t *final_list, uint32_t sigma,
                  int16_t *sym);

/* ---- Data.RLE -------------------------------------------------------------- */
/* seqToRLE (src/Data/RLE/Internal.hs:104-153) including its Nothing quirks.
 * Run k is (count[k], rsym[k]); the reference's flat Seq is [show count, sym]...
 * *R is always the true run count; R > cap gives TC_E_CAP with the first cap runs written. */
int tc_rle_encode(tc_ctx *ctx, const int16_t *sym, uint64_t N, uint32_t *count, int16_t *rsym, uint64_t cap,
                  uint64_t *R);
int tc_rle_encode_u8(tc_ctx *ctx, const uint8_t *bwt, uint64_t N, uint64_t primary, uint32_t *count, int16_t *rsym,
                     uint64_t cap, uint64_t *R);
/* run-length over an MTF index stream (no Nothing can occur). */
int tc_rle_encode_u16(tc_ctx *ctx, const uint16_t *idx, uint64_t N, uint32_t *count, int16_t *rsym, uint64_t cap,
                      uint64_t *R);
/* seqFromRLE (src/Data/RLE/Internal.hs:155-189): (Just _, Nothing) -> one Nothing,
 * else count copies. *N is always the true length; N > cap gives TC_E_CAP. */
int tc_rle_decode(tc_ctx *ctx, const uint32_t *count, const int16_t *rsym, uint64_t R, int16_t *sym, uint64_t cap,
                  uint64_t *N);

/* ---- composed helpers (bytestringToBWTToRLEB-style, device-resident chaining) ---- */
typedef struct {
    uint64_t n;              /* text length                                    */
    uint64_t N;              /* BWT length (n+1, or 0)                         */
    uint64_t primary;        /* slot of the Nothing in the BWT                 */
    uint32_t sigma;          /* MTF alphabet size incl. Nothing (0 if no MTF)  */
    int16_t final_list[257]; /* MTF final list (seqToMTF's second component)   */
    uint64_t R;              /* number of runs                                 */
} tc_block_info;
/* bytestringToBWTToRLEB (src/Data/RLE.hs:83-85): text -> BWT -> seqToRLE(symbols). */
int tc_bwt_rle_encode(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint32_t *count, int16_t *rsym, uint64_t cap,
                      tc_block_info *info);
/* bytestringToBWTToMTFB (src/Data/MTF.hs:82-84) followed by seqToRLE over the index
 * stream (the BWT+MTF+RLE composite of BASELINE.json). */
int tc_bwt_mtf_rle_encode(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint32_t *count, int16_t *rsym, uint64_t cap,
                          tc_block_info *info);
/* The same composite over a batch of independent blocks (BASELINE.json config 5: multi-block
 * compression; the reference maps bytestringToBWTToMTFB over its blocks one by one,
 * src/Data/MTF.hs:82-84).  Block b+1 is copied to the device and block b-1's runs are copied
 * back while block b is being compressed (three streams, double buffers), so a stream of
 * blocks runs at max(copy, compute) instead of their sum.  with_mtf = 0 gives
 * bytestringToBWTToRLEB per block.  count[b] / rsym[b] hold cap[b] entries each; info[b] is
 * filled per block; a block whose runs exceed cap[b] is reported as TC_E_CAP after the whole
 * batch has been processed (info[b].R tells the size needed). */
int tc_blocks_encode(tc_ctx *ctx, uint64_t nblocks, const uint8_t *const *text, const uint64_t *n, int with_mtf,
                     uint32_t *const *count, int16_t *const *rsym, const uint64_t *cap, tc_block_info *info);
/* ---- packed block container (the reference has only Show/Read,
 * src/Data/RLE/Internal.hs:95-96, src/Data/MTF/Internal.hs:67-68) --------------------
 * One compressed block as a single little-endian byte string: the tc_block_info fields, then the
 * runs at 1.625 bytes each instead of the 6-byte (count, symbol) record, so a stream of blocks
 * moves a quarter of the bytes over PCIe (version 2; version 1 spent a whole byte per count):
 *   cnt4[ceil(R/2)]  min(count - 1, 15) in four bits per run, run k in bits 4 (k % 2) .. of byte k / 2
 *   sym8[R]   low byte of the symbol's 9-bit code (code = symbol & 0x1ff: MTF indices 0..256 as
 *             they are, BWT symbols 0..255 as they are, Nothing = 0x1ff)
 *   hi[ceil(R/32)] u32 words, bit k%32 of word k/32 = bit 8 of run k's code
 *   big_idx[n_big] u64, big_cnt[n_big] u32: the runs whose count is >= 16 (nibble 15), ascending run index
 * Sections start at the 16-byte-aligned offsets the header states.  The container is lossless for
 * the run sequence: tc_packed_unpack gives back exactly what tc_blocks_encode returns. */
#define TC_PACKED_MAGIC 0x314b4c4242434254ull /* "TCBBLK1" */
#define TC_PACKED_MTF 1u                      /* flags bit 0: runs are over the MTF index stream */
typedef struct {
    uint64_t magic;
    uint32_t version; /* 2 */
    uint32_t flags;
    uint64_t n, N, primary, R, n_big, total_bytes;
    uint64_t off_cnt4, off_sym8, off_hi, off_big_idx, off_big_cnt; /* from the start of the container */
    uint32_t sigma;
    uint32_t reserved;
    int16_t final_list[257];
    int16_t pad[7];
} tc_packed_header; /* 640 bytes */
/* Upper bound of the container size for a block of n text bytes. */
uint64_t tc_packed_bound(uint64_t n);
/* tc_blocks_encode with container output: out[b] receives cap[b] >= tc_packed_bound(n[b]) bytes at
 * most, out_bytes[b] the size written.  Same pipelining; the packing runs on the device right
 * behind the RLE kernels.  Up to 2 * lanes (at most 8) blocks have device-to-host copies pending at any
 * time, so out[b] .. out[b + 7] must be distinct buffers (as must text[b] .. text[b + 7] if the caller
 * refills them); all copies have completed when the call returns, also when it returns an error. */
int tc_blocks_encode_packed(tc_ctx *ctx, uint64_t nblocks, const uint8_t *const *text, const uint64_t *n, int with_mtf,
                            uint8_t *const *out, const uint64_t *cap, uint64_t *out_bytes, tc_block_info *info);
/* Host-only (no device, no context): header check + tc_block_info of a container. */
int tc_packed_info(const void *blob, uint64_t bytes, tc_block_info *info, uint32_t *flags);
/* Host-only: container -> the run records of tc_blocks_encode.  R > cap gives TC_E_CAP (info->R set). */
int tc_packed_unpack(const void *blob, uint64_t bytes, uint32_t *count, int16_t *rsym, uint64_t cap,
                     tc_block_info *info);
/* Container -> text on the device (unpack kernel, then the inverse chain below). */
int tc_packed_decode(tc_ctx *ctx, const void *blob, uint64_t bytes, uint8_t *text, uint64_t cap, uint64_t *n_out);
/* The same for a list of containers (multi-block decompression, the inverse of tc_blocks_encode_packed): up to
 * TC_B200_LANES containers are in flight, each on its own context and stream, so the copies and host syncs of one are
 * covered by the kernels of the others.  n_out[b] is the length of block b; a block longer than cap[b] gives TC_E_CAP
 * after every block has been decoded (n_out[b] set), any other error stops the batch.  text[b] must be distinct
 * buffers; all copies have completed when the call returns. */
int tc_blocks_decode_packed(tc_ctx *ctx, uint64_t nblocks, const void *const *blob, const uint64_t *bytes,
                            uint8_t *const *text, const uint64_t *cap, uint64_t *n_out);

/* inverses: runs -> (MTF indices ->) BWT -> text. */
int tc_bwt_rle_decode(tc_ctx *ctx, const uint32_t *count, const int16_t *rsym, uint64_t R, uint8_t *text,
                      uint64_t cap, uint64_t *n_out);
int tc_bwt_mtf_rle_decode(tc_ctx *ctx, const uint32_t *count, const int16_t *rsym, const tc_block_info *info,
                          uint8_t *text, uint64_t cap, uint64_t *n_out);

/* device-pointer twins used when the data is already in HBM (bench `value`). */
int tc_bwt_encode_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint8_t *d_bwt, uint64_t *primary,
                      uint32_t *d_sa_1based);
int tc_mtf_encode_u8_dev(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint16_t *d_idx,
                         int16_t *final_list, uint32_t *sigma);
int tc_rle_encode_u8_dev(tc_ctx *ctx, const uint8_t *d_bwt, uint64_t N, uint64_t primary, uint32_t *d_count,
                         int16_t *d_rsym, uint64_t cap, uint64_t *R);
int tc_rle_encode_u16_dev(tc_ctx *ctx, const uint16_t *d_idx, uint64_t N, uint32_t *d_count, int16_t *d_rsym,
                          uint64_t cap, uint64_t *R);
int tc_bwt_mtf_rle_encode_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t *d_count, int16_t *d_rsym,
                              uint64_t cap, tc_block_info *info);

/* tc_blocks_encode on device pointers: d_text[b], d_count[b], d_rsym[b] live in HBM, nothing is copied.
 * Blocks are compressed `lanes` at a time (3 by default, TC_B200_LANES = 1..4: the caller's context plus
 * child contexts on helper threads), which fills the short serial phases of one block's kernel chain
 * with the other blocks' kernels; blocks b .. b + lanes - 1 run together and must not share output
 * buffers. */
int tc_blocks_encode_dev(tc_ctx *ctx, uint64_t nblocks, const uint8_t *const *d_text, const uint64_t *n, int with_mtf,
                         uint32_t *const *d_count, int16_t *const *d_rsym, const uint64_t *cap, tc_block_info *info);

/* ---- Data.FMIndex ------------------------------------------------------------ */
/* *ToBWTToFMIndex* (src/Data/FMIndex.hs:108-183): C[c] (seqToCc, Internal.hs:275-316),
 * Occ (seqToOccCK, :195-259; stored as rank-blocks instead of the dense table) and the
 * suffix array (sampled every sa_sample_rate text positions; 1 keeps the full SA). */
int tc_fm_build(tc_ctx *ctx, const uint8_t *text, uint64_t n, uint32_t sa_sample_rate, tc_fm **out);
int tc_fm_build_dev(tc_ctx *ctx, const uint8_t *d_text, uint64_t n, uint32_t sa_sample_rate, tc_fm **out);
void tc_fm_free(tc_fm *fm);
typedef struct {
    uint64_t n, N, primary;
    uint32_t sigma;       /* alphabet incl. Nothing */
    uint32_t sa_sample_rate;
    int16_t alphabet[257]; /* sorted, alphabet[0] == -1 */
    int64_t C[257];        /* Cc: C[j] for alphabet[j] */
    uint64_t blob_bytes;   /* size of the device image (tc_fm_blob) */
    uint64_t n_samples;
} tc_fm_info;
int tc_fm_get_info(const tc_fm *fm, tc_fm_info *info);
/* countFMIndex (src/Data/FMIndex/Internal.hs:347-438) for q patterns; pattern i is
 * pats[off[i] .. off[i+1]).  count[i] == -1 is Nothing. */
int tc_fm_count(tc_ctx *ctx, const tc_fm *fm, const uint8_t *pats, const uint64_t *off, uint64_t q, int64_t *count);
int tc_fm_count_dev(tc_ctx *ctx, const tc_fm *fm, const uint8_t *d_pats, const uint64_t *d_off, uint64_t q,
                    int64_t *d_count);
/* locateFMIndex (:448-542) + the rank->position map of the wrappers
 * (src/Data/FMIndex.hs:496,526,562,598): pattern i's hits are
 * pos_1based[hit_off[i] .. hit_off[i+1]) in SA-rank order (the reference's order, unsorted).
 * *total is always the true number of hits; total > cap gives TC_E_CAP. */
int tc_fm_locate(tc_ctx *ctx, const tc_fm *fm, const uint8_t *pats, const uint64_t *off, uint64_t q,
                 uint64_t *hit_off /*q+1*/, uint64_t *pos_1based, uint64_t cap, uint64_t *total);
int tc_fm_locate_dev(tc_ctx *ctx, const tc_fm *fm, const uint8_t *d_pats, const uint64_t *d_off, uint64_t q,
                     uint64_t *d_hit_off /*q+1*/, uint64_t *d_pos_1based, uint64_t cap, uint64_t *total);
/* Materialise what the public FMIndex value holds (small N): BWT as maybe-symbols and,
 * when built with sa_sample_rate == 1, the full 1-based suffix array. */
int tc_fm_export(tc_ctx *ctx, const tc_fm *fm, int16_t *bwt /*N*/, uint32_t *sa_1based /*N, nullable*/);
/* Replication across GPUs: the index is one contiguous device image.  Broadcast the
 * image (NCCL) and re-open it on the peer with tc_fm_from_blob_dev. */
const void *tc_fm_blob(const tc_fm *fm);
int tc_fm_from_blob_dev(tc_ctx *ctx, void *d_blob, uint64_t bytes, int take_ownership, tc_fm **out);

/* ---- multi-GPU below the C ABI ------------------------------------------------------------------
 * The reference's only parallelism is parListChunk over the cores inside its ...P functions
 * (src/Data/FMIndex.hs:417-423, 544-553): the LIBRARY CALL fans out.  These entry points do the same over the
 * GPUs of one box from a single process (one host thread and one pooled context per device), so a caller
 * that is not Python + torch.distributed gets all of them through one call.  devices[i] are CUDA ordinals. */
int tc_device_count(void);
/* The process-wide context of a device (created on first use, never destroyed): exclusive until released.
 * For callers that would otherwise create and destroy a context per call (the Haskell shim's pure functions).
 * Do not acquire the same device twice from one thread. */
int tc_ctx_pool_acquire(int device, tc_ctx **out);
void tc_ctx_pool_release(tc_ctx *ctx);
/* tc_blocks_encode_packed over several devices (BASELINE.json config 5): block b runs on devices[b % ndev];
 * independent blocks, no exchange step. */
int tc_mgpu_blocks_encode_packed(int ndev, const int *devices, uint64_t nblocks, const uint8_t *const *text,
                                 const uint64_t *n, int with_mtf, uint8_t *const *out, const uint64_t *cap,
                                 uint64_t *out_bytes, tc_block_info *info);
/* Copies the index image of `root` to devices[0 .. ndev) (peer copies over NVLink where the devices are peers)
 * and opens it there: replicas[i] lives on devices[i] and is freed with tc_fm_free. */
int tc_fm_replicate(const tc_fm *root, int ndev, const int *devices, tc_fm **replicas);
/* tc_fm_count / tc_fm_locate with the q patterns split into ndev contiguous chunks (the reference's
 * parListChunk), chunk i answered by replicas[i] on devices[i]; results in input order, exactly what the
 * single-device calls return. */
int tc_mgpu_fm_count(int ndev, const int *devices, tc_fm *const *replicas, const uint8_t *pats, const uint64_t *off,
                     uint64_t q, int64_t *count);
int tc_mgpu_fm_locate(int ndev, const int *devices, tc_fm *const *replicas, const uint8_t *pats, const uint64_t *off,
                      uint64_t q, uint64_t *hit_off /*q+1*/, uint64_t *pos_1based, uint64_t cap, uint64_t *total);

#ifdef __cplusplus
}
#endif
#endif /* TC_B200_H */
