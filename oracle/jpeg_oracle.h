/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * CPU restatement, in plain C, of the baseline-JPEG encode path of
 * rusty-electron/jpeg-encoder-opencl (its CPU implementation, src/utils.cpp,
 * driven in the order of src/OpenCLProject_JpegEncoder.cpp:59-225).  Every
 * function cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product library (libjpegb200.so) never does.
 *
 * Pinning: functions marked [pinned] are checked bit-for-bit against the
 * reference's own code compiled unmodified (oracle/_ref/libjpegref.so, see
 * oracle/Makefile) and against the digests recorded in SURVEY.md section 8c /
 * tests/golden/.  Functions marked [unpinned] implement surface the reference
 * does not have (byte packing, 0xFF stuffing, restart markers, JFIF segments,
 * quality scaling, true 4:2:0 MCUs); they are pinned instead by ITU-T T.81 /
 * JFIF conformance, i.e. by decoding with independent decoders (PIL, OpenCV).
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Reference quirks (SURVEY.md section 8 "Quirks"), OR-able. */
#define ORC_Q1_INPLACE_DCT 1 /* utils.cpp:342-345 writes into the block it reads */
#define ORC_Q2_TYPO_TABLES 2 /* huffman.hpp:92-98 seven 17-bit luma AC codes       */
#define ORC_Q3_ALWAYS_EOB 4  /* utils.cpp:607-608 EOB even when coefficient 63 != 0 */
#define ORC_REFERENCE_AS_WRITTEN (ORC_Q1_INPLACE_DCT | ORC_Q2_TYPO_TABLES | ORC_Q3_ALWAYS_EOB)

/* Chroma handling. */
#define ORC_SUB_444 0     /* no chroma averaging, 3 blocks per 8x8 MCU                       */
#define ORC_SUB_REPL420 1 /* reference: 2x2 mean replicated at full res, coded as 4:4:4 (Q4) */
#define ORC_SUB_420 2     /* true 4:2:0: 16x16 MCU = Y00 Y01 Y10 Y11 Cb Cr                   */

/* ---- per-stage functions, reference layouts [pinned] -------------------- */
void orc_csc(uint8_t *px, size_t npixels);                 /* utils.cpp:92-110   */
void orc_cds(uint8_t *px, size_t W, size_t H);             /* utils.cpp:113-141  */
void orc_padded_size(size_t W, size_t H, size_t mult, size_t *nW, size_t *nH); /* :184-187, cpp:93-98 */
int orc_pad_mirror(const uint8_t *src, size_t W, size_t H, uint8_t *dst, size_t nW, size_t nH); /* :199-233 */
void orc_u8_to_double(const uint8_t *src, double *dst, size_t n);  /* utils.cpp:236-246 */
void orc_subtract(double *img, size_t n, double val);              /* utils.cpp:190-196 */
void orc_dct_image(double *img, size_t W, size_t H, int inplace);  /* utils.cpp:262-270, 314-347 */
void orc_quantize_image(double *img, size_t W, size_t H, const unsigned ql[64], const unsigned qc[64]); /* :454-467 */
void orc_blockify(const double *img, size_t W, size_t H, int32_t *linear); /* utils.cpp:482-498 */
void orc_zigzag(const int32_t *linear, int32_t *zz, size_t rows);          /* utils.cpp:539-558 */
void orc_aos_to_planar_u32(const uint8_t *px, size_t W, size_t H, uint32_t *v);       /* utils.cpp:700-707 */
void orc_planar_u32_interleave(const uint32_t *in, size_t W, size_t H, uint32_t *out); /* utils.cpp:745-754 */
size_t orc_rle_block(const int32_t zz[64], int32_t *pairs, int always_eob); /* utils.cpp:572-609 */
int orc_category(int v);                                   /* utils.cpp:623-627 */
int orc_value_bits(int v, uint32_t *bits);                 /* utils.cpp:630-653 */
/* utils.cpp:656-698: zz is the reference layout int32[3*rpc][64]; output is
 * MSB-first packed bits (no padding, no stuffing); returns the bit count. */
uint64_t orc_huffman_ref(const int32_t *zz, size_t rpc, int quirks, uint8_t *packed, size_t cap_bytes);
/* Code tables as (code,len); table 0 DC luma, 1 DC chroma, 2 AC luma, 3 AC chroma.
 * Returns len (0 = undefined symbol).  huffman.hpp:9,26,43,250. */
int orc_table_code(int table, int run, int cat, int typo, uint32_t *code);
extern const unsigned orc_q50_lum[64];   /* utils.hpp:42-51 */
extern const unsigned orc_q50_chrom[64]; /* utils.hpp:53-62 */
extern uint8_t orc_zigzag_order[64]; /* valid after orc_init() */
void orc_init(void);

/* ---- new surface [unpinned by the reference; T.81/JFIF conformance] ------ */
void orc_quality_tables(int quality, unsigned ql[64], unsigned qc[64]); /* IJG scaling of utils.hpp:42-62 */
/* CSC (+CDS unless 444) + mirror pad to a multiple of the MCU size. dst must hold nW*nH*3. */
int orc_ycc_padded(const uint8_t *rgb, size_t W, size_t H, int sub, uint8_t *dst, size_t *nW, size_t *nH);
/* Quantised zigzag coefficients in scan order: int16[n_mcu][blocks_per_mcu][64].
 * quirks: only ORC_Q1_INPLACE_DCT is looked at. */
int orc_transform(const uint8_t *rgb, size_t W, size_t H, int sub, const unsigned ql[64],
                  const unsigned qc[64], int quirks, int16_t *coef);
int orc_transform_ycc(const uint8_t *ycc, size_t W, size_t H, int sub, const unsigned ql[64], const unsigned qc[64],
                      int quirks, int16_t *coef); /* orc_transform for an image that is already Y,Cb,Cr (after utils.cpp:141) */
size_t orc_num_mcus(size_t W, size_t H, int sub);
int orc_blocks_per_mcu(int sub);
/* Entropy-code scan-order coefficients.  restart_interval in MCUs (0 = none).
 * raw_bits != 0: reference-style bit string (no byte padding, no stuffing, no
 * markers; restart_interval must be 0) packed MSB-first, *nbits = bit count.
 * raw_bits == 0: T.81 entropy segment bytes (1-padding, FF00 stuffing, RSTn).
 * rst_phase = index of the first interval (for strips). final_rst: also emit a
 * RST marker after the last interval.  Returns bytes written or (size_t)-1. */
size_t orc_entropy(const int16_t *coef, size_t n_mcu, int sub, int restart_interval, int quirks,
                   int raw_bits, int rst_phase, int final_rst, uint8_t *out, size_t cap, uint64_t *nbits);
size_t orc_jfif_header(size_t W, size_t H, int sub, const unsigned ql[64], const unsigned qc[64],
                       int restart_interval, uint8_t *out, size_t cap);
/* Whole image -> JFIF file bytes. Returns size or (size_t)-1. */
size_t orc_encode_jfif(const uint8_t *rgb, size_t W, size_t H, int sub, const unsigned ql[64],
                       const unsigned qc[64], int restart_interval, int quirks, uint8_t *out, size_t cap);
/* Deterministic integer-only synthetic image. Fills rows
 * [y0, y0+rows) of a W-wide image into out (rows*W*3 bytes). */
/* optimised Huffman tables (T.81 K.2 as in libjpeg's jpeg_gen_optimal_table); spec index 0 DC lum, 1 DC chr, 2 AC lum, 3 AC chr */
void orc_reset_huffman_specs(void);
void orc_set_huffman_specs(const uint8_t bits[4][16], const uint8_t vals[4][256]);
void orc_symbol_histogram(const int16_t *coef, size_t n_mcu, int sub, int restart_interval, int quirks,
                          uint64_t hist[4][256]);
int orc_optimal_spec(const uint64_t freq[256], uint8_t bits[16], uint8_t vals[256]);
size_t orc_encode_jfif_optimized(const uint8_t *rgb, size_t W, size_t H, int sub, const unsigned ql[64],
                                 const unsigned qc[64], int restart_interval, int quirks, uint8_t *out, size_t cap);
void orc_synth_rgb(uint64_t seed, size_t W, size_t y0, size_t rows, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
