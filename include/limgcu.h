/* include/limgcu.h -- C ABI of the B200-native (sm_100a) limg encode/decode hot path.
 *
 * This is the drop-in boundary a reference maintainer binds against: plain C,
 * plain pointers and sizes, no torch / C++ types. Two layers:
 *
 *   limgcu_*        device-buffer entry points. All `d_` pointers are CUDA device memory of the context's
 *                   device; work is enqueued on the context's stream and is stream ordered. No hidden
 *                   synchronisation except where a function returns a host value (documented per function).
 *   limgcu_host_*   host-buffer entry points with the reference's argument meaning (H2D, kernels, D2H inside).
 *                   include/limg_dropin.h declares the reference's own C++ signatures on top of these.
 *
 * There is NO CPU fallback: every entry point fails with LIMGCU_ERROR_NO_DEVICE / a CUDA error if the
 * kernels cannot run.
 *
 * Reference interfaces replaced (file:line relative to the reference's src/):
 *   limg_blocked_encode3d_test        limg.h:46,  limg.cpp:2329-2453   -> limgcu_blocked_encode3d / limgcu_host_blocked_encode3d
 *   limg_encode3d_test                limg.h:35,  limg.cpp:2175-2265   -> limgcu_encode3d / limgcu_host_encode3d
 *   limg_encode3d_test_perf           limg.h:37,  limg.cpp:2267-2327   -> limgcu_encode3d with no output planes
 *   limg_decode_block_from_factors_3d limg_decode.h:326-340 (static)   -> limgcu_decode / limgcu_host_decode
 *   limg_compare                      limg.h:48,  limg.cpp:2455-2491   -> limgcu_compare / limgcu_host_compare
 *   limg_encode3d_blocked_test_y_range (pass 1) limg.cpp:1088-1119     -> limgcu_pass1
 *   limg_encode_find_block_3d + drivers         limg.cpp:1390-1496,1814-1878 -> limgcu_merge
 */
#ifndef LIMGCU_H
#define LIMGCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as the reference's enum limg_result (limg.h:9-18), plus device errors. */
enum
{
  LIMGCU_SUCCESS = 0,
  LIMGCU_ERROR_GENERIC = 100,
  LIMGCU_ERROR_INVALID_PARAMETER = 101,
  LIMGCU_ERROR_ARGUMENT_NULL = 102,
  LIMGCU_ERROR_OUT_OF_BOUNDS = 103,
  LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE = 104,
  LIMGCU_ERROR_NO_DEVICE = 200,
  LIMGCU_ERROR_CUDA = 201
};

/* Per-block / per-area decomposition (reference: limg_encode_3d_output<channels>, limg_internal.h:343-353).
 * Channel-agnostic: four slots, unused slots zero. 64 bytes. */
typedef struct limgcu_decomp
{
  float avg[4];
  int16_t dirA_min[4], dirA_max[4];
  int16_t dirB_offset[4], dirB_mag[4];
  int16_t dirC_offset[4], dirC_mag[4];
} limgcu_decomp;

/* One area of the area table, in the reference's emission order: large merges (stage 0), remaining merges
 * (stage 1), leftover 8x8 blocks in raster order (stage 2). The area's pixels are addressed row-major inside
 * its pixel rectangle ("area-contiguous order", limg.cpp:1752-1753). 120 bytes. */
typedef struct limgcu_area
{
  uint32_t ox, oy, rx, ry;          /* rectangle in 8x8-block units */
  uint32_t stage;
  uint32_t px_x, px_y, px_w, px_h;  /* pixel rectangle after the edge fit (limg.cpp:1722-1739) */
  uint8_t shift[3];                 /* bits dropped per factor, 0..8 (8 = factor dropped) */
  uint8_t pad;
  uint64_t ditherBefore, ditherAfter; /* dither chain state around this area (limg_internal.h:711, limg.cpp:1539-1549) */
  limgcu_decomp decomp;
} limgcu_area;

/* The reference's limg_blocked_encode3d_info (limg.h:39-44): sizeX*sizeY elements each. Any pointer may be NULL
 * (plane skipped). pBlockError is accepted and never written, exactly like the reference's 3D paths. */
typedef struct limgcu_planes
{
  uint32_t *pDecoded;
  uint8_t *pFactorsA, *pFactorsB, *pFactorsC, *pBlockError, *pBitsPerPixel;
  uint32_t *pShiftABCX, *pColAMin, *pColAMax, *pColBMin, *pColBMax, *pColCMin, *pColCMax, *pBlockIndex;
} limgcu_planes;

/* Compact encoder output ("the stream"): area table + block->area map + three factor planes in image layout holding the
 * RIGHT-ALIGNED codes (enc = dithered factor >> shift; a dropped factor keeps its raw byte, Q7). All device pointers;
 * areas / block_to_area need blockX*blockY entries. Any of the members may be NULL. */
typedef struct limgcu_stream
{
  limgcu_area *areas;
  uint32_t *area_count;     /* one uint32 in device memory */
  uint32_t *block_to_area;
  uint8_t *codesA, *codesB, *codesC;
} limgcu_stream;

enum
{
  LIMGCU_FLAG_FAST_BIT_CRUSH = 1u << 0,  /* reference `fastBitCrushing` (limg.h:46); clear = --accurate-bit-crushing */
  LIMGCU_FLAG_NO_MERGE = 1u << 1,        /* every 8x8 block is its own area (limg_encode3d_test) */
  LIMGCU_FLAG_DITHER_AES = 1u << 2       /* the dither noise of a reference running on a host with SSE4.1 + AES-NI (limg.cpp:824-879) instead of the
                                            LCG it uses elsewhere (limg.cpp:799-822). The AES-round chain has no skip-ahead: it is walked on the host
                                            (AES-NI when the host has it, software otherwise) once the shifts are known, so a call with this flag
                                            synchronises with the stream. Areas, shifts and decompositions are the same in both modes. */
};

typedef struct limgcu_ctx limgcu_ctx;

/* context ------------------------------------------------------------------------------------------------------- */
int limgcu_create(int device, limgcu_ctx **out);
void limgcu_destroy(limgcu_ctx *ctx);
const char *limgcu_last_error(const limgcu_ctx *ctx);
int limgcu_device_count(void);
/* x86 RSQRTPS table the fit emulates (2048 entries, see limg_b200/csrc/rsqrt_lut.h). NULL restores the built-in table. */
int limgcu_set_rsqrt_lut(limgcu_ctx *ctx, const uint16_t *lut2048);
void *limgcu_stream_handle(limgcu_ctx *ctx); /* cudaStream_t */
int limgcu_sync(limgcu_ctx *ctx);
/* Synchronises with the stream and reports the hard errors of the last merge scan that the stream-ordered device entry points
 * (limgcu_merge, limgcu_blocked_encode3d, limgcu_encode_areas) cannot return themselves: LIMGCU_ERROR_GENERIC (watchdog: a block
 * row waited too long) or LIMGCU_ERROR_OUT_OF_BOUNDS (a row's rectangle list overflowed); the area map is then truncated and nothing
 * computed from it is valid. The host-buffer entry points and limgcu_finalize_rows check the same flags themselves. */
int limgcu_status(limgcu_ctx *ctx);
/* dither generator of the entry points that have no flags argument (limgcu_host_blocked_encode3d, limgcu_host_encode3d): 0 = LCG (default),
 * 1 = AES (as LIMGCU_FLAG_DITHER_AES). limgcu_host_has_aesni: 1 when the reference itself would pick the AES generator on this host. */
int limgcu_set_dither_mode(limgcu_ctx *ctx, int aes);
int limgcu_host_has_aesni(void);
/* The reference's non-merged encoder (limg_encode3d_test, limg.cpp:2108-2137) restarts its dither chain at the top of every y-band of its thread
 * pool, so its output depends on the pool size: `threads` > 0 reproduces the run with a pool of that many threads in the NO_MERGE paths
 * (limgcu_host_encode3d, LIMGCU_FLAG_NO_MERGE), 0 (default) the pool-less run. The blocked (merged) encoder has one chain either way. */
int limgcu_set_pool_threads(limgcu_ctx *ctx, int threads);
/* number of kernels launched through this context since creation (bench.py's gpu_launches) */
uint64_t limgcu_launch_count(const limgcu_ctx *ctx);
/* milliseconds the kernels of one named phase took during the last limgcu_*encode3d call when profiling was enabled
 * (phase: 0 pass1, 1 predicate windows, 2 merge scan, 3 area encode, 4 dither scan, 5 finalize). Synchronises. */
int limgcu_enable_phase_timing(limgcu_ctx *ctx, int enable);
float limgcu_phase_ms(limgcu_ctx *ctx, int phase);
/* 32 internal counters of the last encode / merge (area counts, merge iterations ...), for profiling. Synchronises. */
int limgcu_debug_counters(limgcu_ctx *ctx, uint32_t *out32);
/* detailed counters of the last merge scan (kernels_wave.cuh, WaveArgs::dbg): 256 words */
int limgcu_debug_wave(limgcu_ctx *ctx, uint32_t *out256);
/* self check of the merge predicate's guard-banded shortcut (kernels_merge.cuh, predicate_shortcut) against the reference-order
 * evaluation on a pass-1 table (host memory): out4 = { pairs that needed the 27-sample score, pairs the shortcut decided,
 * disagreements (must be 0), reserved } */
int limgcu_debug_predicate_check(limgcu_ctx *ctx, const limgcu_decomp *table, size_t sizeX, size_t sizeY, int hasAlpha, uint64_t *out4);
/* per block row of the last merge scan, both stages: [2][blockY][4] time stamps in ns (ticket, first decision, last decision, done);
 * followed by 512 words of per-decision events of four rows; recorded only when LIMGCU_MERGE_ROWTIMES is set at limgcu_create
 * time (1: row stamps, n > 1: also the decisions of stage-0 rows n .. n+3). `out` needs 8 * blockY + 512 words. */
int limgcu_debug_wave_rows(limgcu_ctx *ctx, uint32_t *out, size_t blockY);

/* selects the reconstruction kernel of limgcu_decode (tuning / A-B timing, tools/decode_time.py): 0 = generic k_decode,
 * 2 / 4 / 8 = rows of an 8x8 block one thread of k_decode_tile reconstructs (kernels_decode.cuh; needs sizeX % 8 == 0, otherwise the
 * generic kernel runs). Results are identical for every value. */
int limgcu_debug_set_decode_variant(limgcu_ctx *ctx, int variant);

/* device-buffer entry points ------------------------------------------------------------------------------------ */

/* pass 1: three-factor fit of every 8x8 block -> blockX*blockY records (limg.cpp:1088-1119). */
int limgcu_pass1(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_decomp *d_table);

/* area expansion: pass-1 table -> emission-ordered area rectangles (ox,oy,rx,ry,stage and the pixel rectangle filled in)
 * + block->area map (limg.cpp:1390-1496, 1814-1878). sizeX/sizeY are needed for the edge fit of the pixel rectangles. */
int limgcu_merge(limgcu_ctx *ctx, const limgcu_decomp *d_table, size_t sizeX, size_t sizeY, int hasAlpha,
                 limgcu_area *d_areas, uint32_t *d_area_count, uint32_t *d_block_to_area);

/* the whole encode path. `stream` and/or `planes` (device pointers inside) select what is written. */
int limgcu_blocked_encode3d(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags,
                            const limgcu_stream *stream, const limgcu_planes *planes);

/* streaming reconstruction of a stream produced by limgcu_blocked_encode3d (or by the reference). */
int limgcu_decode(limgcu_ctx *ctx, const limgcu_area *d_areas, const uint32_t *d_block_to_area, const uint8_t *d_codesA, const uint8_t *d_codesB, const uint8_t *d_codesC,
                  size_t sizeX, size_t sizeY, int hasAlpha, uint32_t *d_dst);

/* rebuilds block_to_area from an area table that was produced elsewhere (e.g. by the reference). */
int limgcu_build_block_map(limgcu_ctx *ctx, const limgcu_area *d_areas, uint32_t area_count, size_t sizeX, size_t sizeY, uint32_t *d_block_to_area);

/* limg_compare on device images; synchronises and returns host values. */
int limgcu_compare(limgcu_ctx *ctx, const uint32_t *d_a, const uint32_t *d_b, size_t sizeX, size_t sizeY, int hasAlpha, double *psnr, double *mse, double *maxError);

/* host-buffer entry points (reference argument meaning) ----------------------------------------------------------- */

int limgcu_host_blocked_encode3d(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, const limgcu_planes *pInfo, uint32_t errorFactor, int fastBitCrushing);
int limgcu_host_encode3d(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, const limgcu_planes *pInfo, uint32_t errorFactor, int fastBitCrushing);
/* host stream out: areas (capacity blockX*blockY), *area_count, codes in image layout; any may be NULL. */
int limgcu_host_encode_stream(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags,
                              limgcu_area *areas, uint32_t *area_count, uint8_t *codesA, uint8_t *codesB, uint8_t *codesC, uint32_t *pDecoded);
int limgcu_host_decode(limgcu_ctx *ctx, const limgcu_area *areas, uint32_t area_count, const uint8_t *codesA, const uint8_t *codesB, const uint8_t *codesC,
                       size_t sizeX, size_t sizeY, int hasAlpha, uint32_t *pOut);
int limgcu_host_pass1(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_decomp *table);
int limgcu_host_merge(limgcu_ctx *ctx, const limgcu_decomp *table, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_area *areas, uint32_t *area_count);
double limgcu_host_compare(limgcu_ctx *ctx, const uint32_t *pImageA, const uint32_t *pImageB, size_t sizeX, size_t sizeY, int hasAlpha, double *pMeanSquaredError, double *pMaxPossibleSquaredError);

/* .limg container ------------------------------------------------------------------------------------------------ */

/* The reference defines no bitstream (it only accounts for one: limg.cpp:1629-1636, a per-area header plus
 * rangeSize * ((8 - shiftA) + (8 - shiftB) + (8 - shiftC)) payload bits). Container "LIMGB200" version 1, little endian:
 *
 *   header (48 bytes)   char magic[8] = "LIMGB200"; u32 version = 1; u32 flags (bit 0: alpha); u32 sizeX, sizeY; u32 areaCount;
 *                       u32 recordBytes (48 RGB / 60 RGBA); u64 payloadBytes; u64 reserved = 0
 *   area table          areaCount records in emission order: u16 ox, oy, rx, ry (8x8-block units); u8 shift[3]; u8 stage;
 *                       int16 dirA_min[ch], dirA_max[ch], dirB_offset[ch], dirB_mag[ch], dirC_offset[ch], dirC_mag[ch] (un-clamped)
 *   payload             per area, in table order: factor A, factor B, factor C; per factor the area's pixels in area-contiguous
 *                       order (row-major inside the pixel rectangle), (8 - shift) bits per code, LSB first; every 8-pixel run of a
 *                       row starts on a byte boundary (a ragged last run is padded with zero codes). A dropped factor (shift 8)
 *                       takes 0 bits for RGB and keeps its raw byte for RGBA (the reference's RGBA reconstruction reads it).
 *
 * The pixel rectangle of an area follows from its block rectangle (edge fit, limg.cpp:1722-1739). */

/* worst-case size of a container for an image of this size */
size_t limgcu_container_bound(size_t sizeX, size_t sizeY, int hasAlpha);
/* parses and validates the header; any out pointer may be NULL. No device needed. */
int limgcu_container_info(const void *data, size_t bytes, size_t *sizeX, size_t *sizeY, int *hasAlpha, uint32_t *areaCount, uint64_t *payloadBytes);
/* encode (limg_blocked_encode3d_test's path; LIMGCU_FLAG_NO_MERGE for limg_encode3d_test's) straight into a container in host memory */
int limgcu_host_encode_container(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags, void *out, size_t capacity,
                                 size_t *written);
/* container in host memory -> sizeX * sizeY pixels (bit-identical to the encoder's pDecoded) */
int limgcu_host_decode_container(limgcu_ctx *ctx, const void *data, size_t bytes, uint32_t *pOut, size_t outPixels);
/* device level: code planes <-> payload. d_offsets receives area_count + 1 byte offsets (the last one is the payload size);
 * d_payload needs 3 * ceil(sizeX / 8) * 8 * sizeY bytes. d_area_count (device) may be NULL, then area_count (host value) is used. */
int limgcu_pack_payload(limgcu_ctx *ctx, const limgcu_area *d_areas, const uint32_t *d_area_count, uint32_t area_count, const uint32_t *d_block_to_area, const uint8_t *d_codesA,
                        const uint8_t *d_codesB, const uint8_t *d_codesC, size_t sizeX, size_t sizeY, int hasAlpha, uint8_t *d_payload, uint64_t *d_offsets);
int limgcu_unpack_payload(limgcu_ctx *ctx, const limgcu_area *d_areas, uint32_t area_count, const uint32_t *d_block_to_area, const uint8_t *d_payload, uint64_t *d_offsets,
                          size_t sizeX, size_t sizeY, int hasAlpha, uint8_t *d_codesA, uint8_t *d_codesB, uint8_t *d_codesC);

/* whole-image-exact row bands ------------------------------------------------------------------- */

/* One very large image on several GPUs with the SAME result as one limg_blocked_encode3d_test call over the whole image (areas may cross
 * the bands, one dither chain). Every rank holds the whole source (all-gathered over NVLink) and runs, with its own context:
 *   1. limgcu_pass1 on its band of block rows -> all-gather of the table (64 B per block)
 *   2. limgcu_merge on the full table: the scan is deterministic, every rank gets the identical area table and block map
 *   3. limgcu_encode_areas: refit + shift search of the areas whose first block row lies in [rowLo, rowHi) (the rank's band);
 *      d_results receives limgcu_area_result_words() uint32 per area, zero for the areas of other ranks
 *   4. SUM all-reduce of d_results over the ranks (every area has exactly one owner)
 *   5. limgcu_finalize_rows: results -> area table, dither chain states of all areas, codes / planes of the pixel rows [yLo, yHi)
 * The collectives are the caller's (limg_b200/shard.py: torch.distributed over NCCL). LCG dither only. limgcu_finalize_rows synchronises. */
size_t limgcu_area_result_words(void);
int limgcu_encode_areas(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags, const limgcu_decomp *d_table,
                        limgcu_area *d_areas, uint32_t rowLo, uint32_t rowHi, uint32_t *d_results);
int limgcu_finalize_rows(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t flags, limgcu_area *d_areas, const uint32_t *d_results,
                         const uint32_t *d_block_to_area, const limgcu_stream *stream, const limgcu_planes *planes, size_t yLo, size_t yHi);

/* batches of independent frames (batch mode) ---------------------------------------------------------- */

/* One frame does not fill a B200 (the area scan is a latency-bound chain), so a batch runs on `lanes` contexts at once -- of one device,
 * or of several -- each driven by its own host thread; frame i is handled by ctxs[i % lanes], the frames of one lane in order. Every frame
 * is an independent limg_blocked_encode3d_test call (own dither chain): results equal the one-at-a-time results bit for bit.
 * Returns the first failing lane's code (limgcu_last_error of that lane's context has the text). */
int limgcu_batch_host_encode_containers(limgcu_ctx *const *ctxs, int lanes, const uint32_t *const *frames, int count, size_t sizeX, size_t sizeY, int hasAlpha,
                                        uint32_t errorFactor, uint32_t flags, void *const *outs, const size_t *capacities, size_t *written);
int limgcu_batch_host_decode_containers(limgcu_ctx *const *ctxs, int lanes, const void *const *containers, const size_t *bytes, int count, uint32_t *const *outs,
                                        const size_t *outPixels);

#ifdef __cplusplus
}
#endif

#endif /* LIMGCU_H */
