/* jpegb200.h -- C ABI of libjpegb200.so, the B200-native (sm_100a) baseline-JPEG
 * encode path that stands in for the CPU/OpenCL path of
 * rusty-electron/jpeg-encoder-opencl.
 *
 * The reference has no plugin/FFI layer: its boundary is the set of C++ free
 * functions declared in src/utils.hpp:77-137 and called, in order, by
 * JpegEncoderHost (src/OpenCLProject_JpegEncoder.cpp:59-225).  Every "staged"
 * entry point below replaces one of those functions and keeps its data layout
 * (AoS rgb_pixel_t / rgb_pixel_d_t images, int[rows][64] block arrays); the
 * citation after each prototype is the reference declaration it replaces.
 * host/utils_compat.hpp re-declares the reference names on top of this ABI so
 * the reference driver can be re-linked unchanged.
 *
 * Conventions: plain pointers and sizes only; caller owns every in/out buffer;
 * the library owns device memory, pinned staging and streams inside jb_ctx;
 * every call returns a jb_status (0 = ok) and never throws; jb_last_error()
 * gives the message.  One jb_ctx per host thread; any number per GPU.  There is
 * no CPU fallback: without a CUDA device jb_create fails with JB_E_CUDA.
 */
#ifndef JPEGB200_H
#define JPEGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jb_ctx jb_ctx;

typedef enum {
    JB_OK = 0,
    JB_E_INVALID = 1,     /* bad argument                                       */
    JB_E_CUDA = 2,        /* CUDA runtime error / no device                     */
    JB_E_NOSPACE = 3,     /* output capacity too small (see jb_required_bytes)  */
    JB_E_NOMEM = 4,       /* allocation failed                                  */
    JB_E_UNSUPPORTED = 5, /* e.g. dimension > 65535 without JB_FLAG_CLAMP_SOF   */
    JB_E_INTERNAL = 6
} jb_status;

/* Chroma handling. */
#define JB_SUB_444 0     /* no chroma averaging; 8x8 MCU = Y Cb Cr                               */
#define JB_SUB_REPL420 1 /* the reference's "4:2:0": 2x2 mean replicated (utils.cpp:113-141),   *
                          * coded 4:4:4 in the order of HuffmanEncoder (utils.cpp:667-695)      */
#define JB_SUB_420 2     /* true 4:2:0: 16x16 MCU = Y00 Y01 Y10 Y11 Cb Cr                         */

/* Flags. The three REF_* flags reproduce the reference's defects for verbatim comparisons. */
#define JB_FLAG_REF_INPLACE_DCT 0x1u /* Q1 utils.cpp:342-345: staged jb_dct_f64, and the fused path  *
                                      * (tcgen05 transform with the in-place map as its matrix)     */
#define JB_FLAG_REF_TYPO_TABLES 0x2u /* Q2 huffman.hpp:92-98 17-bit luma AC codes 3/4..3/A       */
#define JB_FLAG_REF_ALWAYS_EOB 0x4u  /* Q3 utils.cpp:607-608 EOB also after a full block         */
#define JB_FLAG_CLAMP_SOF 0x8u       /* declare min(dim,65535) in SOF0                */
#define JB_FLAG_NO_TIE_FIXUP 0x10u   /* skip the binary64 replay of near-tie coefficients        */
#define JB_FLAG_TENSOR_DCT 0x20u     /* accepted, no effect: the tcgen05 transform (FDCT + scale  *
                                      * + zigzag as one tensor-core contraction per block, fp16   *
                                      * 2-split operands, fp32 in TMEM) is the default of every   *
                                      * subsampling mode                                          */
#define JB_FLAG_FMA_DCT 0x40u        /* use the CUDA-core kernel (register AAN FDCT, near-tie     *
                                      * band proven analytically) instead of the tcgen05 one      */

#define JB_FLAG_TMA 0x100u            /* 4:2:0 tcgen05 transform: stage the image tiles with TMA (cp.async.bulk.tensor  *
                                      * boxes of a 3-D tensor map, k_transform_tma) instead of per-lane cp.async.     *
                                      * Bit-identical output; needs 16-byte aligned base / pitch / frame stride (else *
                                      * ignored).  Measured 6 % slower than the cp.async kernel on the B200 (DESIGN    *
                                      * 3.1d), hence opt-in                                                           */
#define JB_FLAG_ENTROPY_LDG 0x200u     /* entropy coder: stage the coefficient tiles with per-thread loads instead of the   *
                                      * TMA box (the path taken when the driver offers no tensor-map encoder).           *
                                      * Bit-identical output; exists so that the tests exercise that path                 */
#define JB_FLAG_OPTIMIZE_HUFFMAN 0x80u /* two passes like libjpeg's optimize_coding: the symbols of the call's   *
                                      * coefficients are counted on the GPU, optimal tables (T.81 K.2) are built  *
                                      * per call (shared by the frames of a batch), written into the DHT          *
                                      * segments and used by the coder; files are 5-12 % smaller, decoded pixels  *
                                      * identical.  JFIF entry points only (not strips / jb_entropy / jb_huffman); *
                                      * the header is then shorter than jb_header_bytes() says                    */

typedef struct {
    int32_t subsampling;      /* JB_SUB_*                                                   */
    int32_t restart_interval; /* MCUs per restart interval, 0 = none (<= 65535)            */
    uint32_t flags;           /* JB_FLAG_*                                                  */
    uint32_t qlum[64];        /* row-major [v][u], as quant_mat_lum  (utils.hpp:42-51)      */
    uint32_t qchrom[64];      /* row-major [v][u], as quant_mat_chrom (utils.hpp:53-62)     */
} jb_params;

/* Per-stage device times since jb_reset_counters (recorded with CUDA events when jb_set_profiling is on), in
 * microseconds; the first nine fields mirror CPUTelemetry (utils.hpp:65-75) -- a jb_timings* can be read as a
 * CPUTelemetry*.  Every staged entry point adds its kernel's time to the field of the reference stage it replaces
 * (jb_csc_rgb8_aos -> CSCTime, ... jb_huffman -> HuffmanTime; TotalCopyTime = padding / conversion kernels +
 * host<->device transfers).  In the fused path CSC, CDS, level shift, DCT, quantisation and zigzag are ONE kernel:
 * its time goes to DCTTime, the entropy coder's (RLE + Huffman + packing) to HuffmanTime. */
typedef struct {
    double CSCTime, CDSTime, levelShiftTime, DCTTime, QuantTime, TotalCopyTime, zigZagTime, RLETime, HuffmanTime;
    double transform_us; /* fused transform kernel (== DCTTime)                     */
    double fixup_us;     /* near-tie binary64 replay                                */
    double entropy_us;   /* all entropy-coder kernels (lengths, scans, pack, stuff) */
    double h2d_us, d2h_us;
    double edge_us;      /* generic kernel for MCUs that need mirror padding            */
    uint64_t transform_launches, total_launches; /* kernels launched by the library since jb_reset_counters */
    uint64_t tie_fixups;                         /* coefficients replayed in binary64 in the last call      */
    /* staged entry points (kernel time only, with jb_set_profiling): what the sums above are made of */
    double staged_dct_us;     /* jb_dct_f64                                          (in DCTTime)       */
    double staged_copy_us;    /* jb_pad_mirror_aos, jb_u8_to_f64, layout conversions (in TotalCopyTime) */
    double staged_huffman_us; /* jb_huffman                                          (in HuffmanTime)   */
} jb_timings;

/* ---- context ------------------------------------------------------------- */
int jb_create(int device, jb_ctx **ctx);
void jb_destroy(jb_ctx *ctx);
const char *jb_last_error(const jb_ctx *ctx);
int jb_sync(jb_ctx *ctx);
void *jb_stream(jb_ctx *ctx);                 /* the cudaStream_t kernels are launched on */
int jb_set_profiling(jb_ctx *ctx, int on);    /* record CUDA events around each kernel    */
int jb_get_timings(jb_ctx *ctx, jb_timings *t);
int jb_reset_counters(jb_ctx *ctx);
int jb_version(void);

/* pinned host memory for the e2e path (replaces the reference's malloc'd buffers, utils.cpp:50) */
int jb_host_alloc(void **p, size_t bytes);
int jb_host_free(void *p);
int jb_device_alloc(jb_ctx *ctx, void **p, size_t bytes);
int jb_device_free(jb_ctx *ctx, void *p);
int jb_memcpy_h2d(jb_ctx *ctx, void *dst, const void *src, size_t bytes);
int jb_memcpy_d2h(jb_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- staged entry points: host pointers, synchronous, reference layouts --- */
int jb_csc_rgb8_aos(jb_ctx *ctx, uint8_t *px, size_t W, size_t H);
/*      replaces void performCSC(ppm_t*)                                   utils.hpp:81  */
int jb_cds_aos(jb_ctx *ctx, uint8_t *ycc, size_t W, size_t H);
/*      replaces void performCDS(ppm_t*)                                   utils.hpp:82  */
int jb_padded_size(size_t W, size_t H, size_t mult, size_t *nW, size_t *nH);
/*      replaces void getNearest8x8ImageSize(size_t,size_t,size_t*,size_t*) utils.hpp:98 */
int jb_pad_mirror_aos(jb_ctx *ctx, const uint8_t *src, size_t W, size_t H, uint8_t *dst, size_t nW, size_t nH);
/*      replaces copyToLargerImage + addReversedPadding                    utils.hpp:97,99 */
int jb_u8_to_f64(jb_ctx *ctx, const uint8_t *src, double *dst, size_t n);
/*      replaces void copyUIntToDoubleImage(ppm_t*, ppm_d_t*)              utils.hpp:94  */
int jb_levelshift_f64(jb_ctx *ctx, double *img, size_t n, double val);
/*      replaces void substractfromAll(ppm_d_t*, double)                   utils.hpp:100 */
int jb_dct_f64(jb_ctx *ctx, double *img, size_t W, size_t H, uint32_t flags);
/*      replaces void performDCT(ppm_d_t*)                                 utils.hpp:102 */
int jb_quantize_f64(jb_ctx *ctx, double *img, size_t W, size_t H, const uint32_t qlum[64], const uint32_t qchrom[64]);
/*      replaces void performQuantization(ppm_d_t*, const unsigned[][8], const unsigned[][8])  utils.hpp:106 */
int jb_blockify(jb_ctx *ctx, const double *img, size_t W, size_t H, int32_t *linear);
/*      replaces void everyMCUisnow2DArray(ppm_d_t*, int[][64])            utils.hpp:122 */
int jb_zigzag(jb_ctx *ctx, const int32_t *linear, int32_t *zz, size_t rows);
/*      replaces void performZigZag(int[][64], int[][64], int)             utils.hpp:127 */
/* The image layout of the reference's OpenCL half: planar uint32, R plane, G plane, B plane of W*H words each
 * (its kernels read d_input[c*W*H + y*W + x], src/OpenCLProject_JpegEncoder.cl:17-19).
 * replaces: void copyImageToVector(ppm_t*, std::vector<cl_uint>&)  utils.hpp:116 (utils.cpp:700-707) */
int jb_planar_u32_from_aos(jb_ctx *ctx, const uint8_t *px, size_t W, size_t H, uint32_t *planar);
/* replaces: void switchVectorChannelOrdering(std::vector<cl_uint>&, std::vector<cl_uint>&, unsigned, unsigned)
 *           utils.hpp:119 (utils.cpp:745-754): planes -> interleaved words */
int jb_planar_u32_interleave(jb_ctx *ctx, const uint32_t *planar, size_t W, size_t H, uint32_t *interleaved);
/* Device-side hop from that layout to the pitched RGB8 frame the fused entry points read (d_* are device
 * pointers; asynchronous on jb_stream()): feed the result to jb_encode_batch_device. */
int jb_planar_u32_to_rgb8_device(jb_ctx *ctx, const uint32_t *d_planar, size_t W, size_t H, uint8_t *d_rgb, size_t pitch);
/* jb_encode_jfif for a host image in that layout (what the reference's main() holds after copyImageToVector,
 * src/OpenCLProject_JpegEncoder.cpp:325): upload, convert on the device, fused encode. */
int jb_encode_jfif_planar_u32(jb_ctx *ctx, const uint32_t *planar, size_t W, size_t H, const jb_params *p, uint8_t *out,
                              size_t cap, size_t *out_len);
int jb_pad_mirror_planar_u32(jb_ctx *ctx, const uint32_t *planar, size_t W, size_t H, uint32_t *out, size_t nW, size_t nH);
/*      replaces void copyOntoLargerVectorWithPadding(std::vector<cl_uint>&, std::vector<cl_uint>&, ...)  utils.hpp:118
 *      (mirror padding of the planar uint32 image; the bottom-right corner mirrors both ways like
 *      addReversedPadding, where the reference's loop reads past the row, utils.cpp:733-740) */
int jb_blockify_planar_i32(jb_ctx *ctx, const int32_t *planar, size_t W, size_t H, int32_t *linear);
/*      replaces void everyMCUisnow1DArray(std::vector<int>&, int[], unsigned, unsigned)   utils.hpp:123 */
int jb_f64_to_u8(jb_ctx *ctx, const double *src, uint8_t *dst, size_t n);
/*      replaces void copyDoubleToUIntImage(ppm_d_t*, ppm_t*)              utils.hpp:95  */
int jb_remove_red_aos(jb_ctx *ctx, uint8_t *px, size_t W, size_t H);
/*      replaces void removeRedChannel(ppm_t*)  ("TEST FUNCTION")          utils.hpp:79  */
int jb_value_categories(jb_ctx *ctx, const int16_t *v, size_t n, uint8_t *cat, uint16_t *bits);
/*      replaces const int16_t getValueCategory(const int16_t) and const std::string valueToBitString(const int16_t)
 *      utils.hpp:134-135, for n values at once: cat[i] = category (bit length of |v|), bits[i] = the cat[i] value
 *      bits (MSB first when written out), computed by the device function the entropy coder itself uses */
int jb_rle(jb_ctx *ctx, const int32_t *zz, size_t rows, uint32_t flags, int32_t *pairs, uint32_t *counts);
/*      replaces void performRLE(int[][64], vector<vector<int>>&, int)     utils.hpp:132
 *      pairs: rows x 128 ints (run,value,...); counts[r] = ints used by row r */
int jb_huffman(jb_ctx *ctx, const int32_t *zz, size_t rows_per_channel, uint32_t flags, uint8_t *bits, size_t cap_bytes,
               uint64_t *nbits);
/*      replaces std::string HuffmanEncoder(int[][64], vector<vector<int>>&, int)  utils.hpp:137
 *      zz = int[3*rows_per_channel][64] (planar by channel); output = the same bit
 *      sequence packed MSB-first, *nbits bits, no padding/stuffing/markers */

/* ---- fused path ------------------------------------------------------------ */
int jb_quality_tables(int quality, uint32_t qlum[64], uint32_t qchrom[64]); /* IJG scaling of utils.hpp:42-62 */
/* The table generator of JB_FLAG_OPTIMIZE_HUFFMAN (host only): BITS / HUFFVAL for 256 symbol counts, T.81 K.2 as
 * libjpeg's jpeg_gen_optimal_table applies it (reserved all-ones code, code lengths <= 16). */
int jb_optimal_huffman_spec(const uint64_t counts[256], uint8_t bits[16], uint8_t vals[256], int *n_vals);
size_t jb_num_mcus(size_t W, size_t H, int subsampling);
int jb_blocks_per_mcu(int subsampling);
size_t jb_header_bytes(const jb_params *p);
size_t jb_required_bytes(jb_ctx *ctx); /* after JB_E_NOSPACE: bytes the last call needed */

/* RGB8 AoS (host) -> quantised zigzag int16 coefficients in scan order
 * [n_mcu][blocks_per_mcu][64] (host).  One kernel does performCSC .. performZigZag. */
int jb_transform(jb_ctx *ctx, const uint8_t *rgb, size_t W, size_t H, size_t pitch, const jb_params *p, int16_t *coef);
/* scan-order coefficients (host) -> entropy segment bytes (host): 1-padding, FF00 stuffing, RSTn */
int jb_entropy(jb_ctx *ctx, const int16_t *coef, size_t n_mcu, const jb_params *p, uint8_t *out, size_t cap,
               size_t *out_len);
/* RGB8 (host) -> complete JFIF file (host) */
int jb_encode_jfif(jb_ctx *ctx, const uint8_t *rgb, size_t W, size_t H, size_t pitch, const jb_params *p, uint8_t *out,
                   size_t cap, size_t *out_len);
/* n_frames equally sized RGB8 frames (host, frame f at rgb + f*frame_stride) -> n_frames JFIF
 * files packed back to back in out; offsets[f], sizes[f] (host arrays) locate them.
 * Host<->device copies are pipelined with the kernels over several streams. */
int jb_encode_batch(jb_ctx *ctx, const uint8_t *rgb, size_t n_frames, size_t W, size_t H, size_t pitch,
                    size_t frame_stride, const jb_params *p, uint8_t *out, size_t cap, uint64_t *offsets,
                    uint64_t *sizes);
/* One image of ANY size as a grid of independent JFIF files -- the way past SOF0's 16-bit X / Y (DNL cannot
 * help: its NL field is 16 bits too, T.81 B.2.5).  Tiles are tile_w x tile_h pixels (multiples of the MCU, <= 65535),
 * the last column / row of tiles is narrower / shorter and mirror-padded like any image; tile (tx, ty) is file
 * ty * ceil(W / tile_w) + tx: offsets[] / sizes[] (host arrays of *n_tiles entries) locate it in `out`.  Every tile
 * column is coded as one batch (the tiles of a column are equally sized frames tile_h * pitch bytes apart). */
int jb_encode_tiles(jb_ctx *ctx, const uint8_t *rgb, size_t W, size_t H, size_t pitch, size_t tile_w, size_t tile_h,
                    const jb_params *p, uint8_t *out, size_t cap, uint64_t *offsets, uint64_t *sizes, size_t *n_tiles);
/* Same with every buffer resident in HBM (d_* are device pointers); asynchronous on jb_stream(),
 * complete after jb_sync().  d_total receives the total bytes (1 x uint64). */
int jb_encode_batch_device(jb_ctx *ctx, const uint8_t *d_rgb, size_t n_frames, size_t W, size_t H, size_t pitch,
                           size_t frame_stride, const jb_params *p, uint8_t *d_out, size_t cap, uint64_t *d_offsets,
                           uint64_t *d_sizes, uint64_t *d_total);
/* NV12-style device input: the frames are YCbCr 4:2:0 already -- a full-resolution Y plane (pitch_y)
 * and a plane of interleaved Cb,Cr pairs at half resolution (ceil(W/2) pairs x ceil(H/2) rows, pitch_uv), as video
 * decoders and camera pipelines leave them in HBM.  No colour conversion and no chroma averaging happen; the
 * reference's stages from the mirror padding on apply unchanged (the pairs replicated over their 2x2 cells are the
 * plane performCDS, utils.cpp:113-141, leaves behind).  JB_SUB_420 only; everything else as jb_encode_batch_device. */
int jb_encode_nv12_device(jb_ctx *ctx, const uint8_t *d_y, size_t pitch_y, size_t frame_stride_y, const uint8_t *d_uv,
                          size_t pitch_uv, size_t frame_stride_uv, size_t n_frames, size_t W, size_t H, const jb_params *p,
                          uint8_t *d_out, size_t cap, uint64_t *d_offsets, uint64_t *d_sizes, uint64_t *d_total);
/* The same from HOST planes to host JFIF files (the NV12 counterpart of jb_encode_batch: pinned buffers recommended, groups of
 * frames pipelined over the context's streams).  Half the bytes of RGB8 cross the host link: where jb_encode_batch is bound by
 * the H2D copy (3 B/px), this path moves 1.5 B/px.  Output layout, JB_E_NOSPACE / jb_required_bytes as jb_encode_batch. */
int jb_encode_nv12_batch(jb_ctx *ctx, const uint8_t *y, size_t pitch_y, size_t frame_stride_y, const uint8_t *uv, size_t pitch_uv,
                         size_t frame_stride_uv, size_t n_frames, size_t W, size_t H, const jb_params *p, uint8_t *out, size_t cap,
                         uint64_t *offsets, uint64_t *sizes);
/* RGB8 -> NV12 on the device with the reference's own arithmetic: performCSC (utils.cpp:92-110) and performCDS
 * (truncated mean of every complete 2x2 cell; a cell cut by an odd edge keeps its top-left pixel's chroma). */
int jb_rgb8_to_nv12_device(jb_ctx *ctx, const uint8_t *d_rgb, size_t W, size_t H, size_t pitch, uint8_t *d_y, size_t pitch_y,
                           uint8_t *d_uv, size_t pitch_uv);
/* One horizontal strip of a large image whose restart intervals are whole MCU rows groups
 * (multi-GPU: strip s on GPU s).  Produces only entropy bytes (no header, no EOI); RST
 * numbering starts at first_interval; a RST marker follows the last interval unless
 * last_strip.  rgb is a device pointer when d_out is (device_io != 0). */
int jb_encode_strip(jb_ctx *ctx, const uint8_t *rgb, size_t W, size_t strip_rows, size_t pitch, const jb_params *p,
                    uint64_t first_interval, int last_strip, int device_io, uint8_t *out, size_t cap, size_t *out_len);
/* The same strip in two asynchronous halves, for a stitch that never visits the host (all pointers are device
 * pointers; both calls only enqueue work on jb_stream(); status at jb_sync()):
 *   begin   transform + entropy coder up to the sizes; *d_len = byte count of the strip
 *   finish  the final placement kernel writes the strip to d_out + *d_off (cap = bytes available from d_out).
 * Between the two the caller all-gathers the lengths and prefix-sums them on the device (NCCL, 8 bytes per rank).
 * d_out may be memory of ANOTHER GPU mapped with jb_ipc_open: the placement kernel's coalesced 128-bit stores then
 * travel over NVLink and land at the strip's final position in the stitching rank's file -- compute and exchange in
 * one kernel, no gather, no second copy. */
int jb_encode_strip_begin(jb_ctx *ctx, const uint8_t *d_rgb, size_t W, size_t strip_rows, size_t pitch, const jb_params *p,
                          uint64_t first_interval, int last_strip, uint64_t *d_len);
int jb_encode_strip_finish(jb_ctx *ctx, uint8_t *d_out, size_t cap, const uint64_t *d_off);
/* d_dst[*d_dst_off ...] = d_src[0 .. *d_len) with device-side length and offset (d_dst may be peer memory): pushes the
 * strips a rank coded in several calls, stitched locally, to their place in the stitching rank's file. */
int jb_copy_bytes_device(jb_ctx *ctx, uint8_t *d_dst, size_t cap, const uint64_t *d_dst_off, const uint8_t *d_src,
                         const uint64_t *d_len);
/* The exchange step itself, without a collective library.  d_ctl = a zero-initialised 512-byte control block on the
 * stitching rank (jb_device_alloc + memset; every rank maps it with jb_ipc_open).  epoch = 1, 2, 3, ... per step.
 *   exchange  publishes *d_len as this rank's strip length (stores over NVLink), waits until all `world` ranks have
 *             published theirs, and writes d_off[0] = base + lengths of the ranks before this one (where this
 *             rank's strip goes), d_off[1] = base + all lengths (where the data ends: EOI goes there)
 *   complete  ranks other than dst: signal that their bytes have landed (call after jb_encode_strip_finish /
 *             jb_copy_bytes_device); dst: wait for every other rank's signal.
 * Both only enqueue a one-warp kernel on jb_stream().  A peer that never reports is given up on after ~20 s
 * (JB_E_INTERNAL / JB_E_NOSPACE at jb_sync; nothing is written). */
int jb_stitch_exchange(jb_ctx *ctx, uint64_t *d_ctl, int rank, int world, uint64_t epoch, uint64_t base,
                       const uint64_t *d_len, uint64_t *d_off);
int jb_stitch_complete(jb_ctx *ctx, uint64_t *d_ctl, int rank, int world, int dst, uint64_t epoch);
/* CUDA IPC for one process per GPU: export a jb_device_alloc'd buffer, map it in another process (peer access over
 * NVLink is enabled on open), unmap. */
int jb_ipc_export(jb_ctx *ctx, void *d_ptr, uint8_t handle[64]);
int jb_ipc_open(jb_ctx *ctx, const uint8_t handle[64], void **d_ptr);
int jb_ipc_close(jb_ctx *ctx, void *d_ptr);
/* JFIF header (SOI..SOS) for a W x H image with these parameters (host) */
int jb_write_header(const jb_params *p, size_t W, size_t H, uint8_t *out, size_t cap, size_t *out_len);

/* ---- decode path -------------------------------------------------------------------------------
 * Baseline JFIF (3 components, 4:4:4 or 4:2:0, one interleaved scan: what this library writes, and what libjpeg
 * writes by default) -> quantised coefficients -> RGB8, on the GPU: the PSNR loop of the parity report without a CPU
 * decoder, and a round-trip check of the byte stream (decoded coefficients == the coefficients that were coded).
 * The reference has no decoder; the reconstruction repeats libjpeg's default one operation for operation (integer
 * "islow" IDCT, h2v2 "fancy" upsampling, fixed-point colour conversion), so pixels equal PIL's / OpenCV's exactly. */
typedef struct {
    uint32_t W, H;
    int32_t subsampling;       /* JB_SUB_444 or JB_SUB_420 (the replicated 4:2:0 mode is coded as 4:4:4) */
    uint32_t restart_interval; /* MCUs, 0 = none */
    uint64_t scan_offset;      /* first byte of the entropy-coded data */
} jb_jfif_info;
int jb_jfif_info_host(const uint8_t *jfif, size_t len, jb_jfif_info *info); /* host buffer, host only (the marker parser) */
int jb_jfif_info_device(jb_ctx *ctx, const uint8_t *d_jfif, size_t len, jb_jfif_info *info);
/* d_rgb (W*3 <= pitch) and / or d_coef (int16 [n_mcu][blocks_per_mcu][64], zigzag order: the layout jb_transform
 * produces) may be null; device pointers, synchronous.  One thread decodes one restart interval. */
int jb_decode_jfif_device(jb_ctx *ctx, const uint8_t *d_jfif, size_t len, uint8_t *d_rgb, size_t pitch, int16_t *d_coef);
/* host buffers; *W, *H are set even when cap is too small (JB_E_NOSPACE, jb_required_bytes) */
int jb_decode_jfif(jb_ctx *ctx, const uint8_t *jfif, size_t len, uint8_t *rgb, size_t cap, size_t *W, size_t *H);
/* PSNR (dB) over the three channels and / or the sum of squared differences of two RGB8 images in HBM */
int jb_psnr_device(jb_ctx *ctx, const uint8_t *d_a, size_t pitch_a, const uint8_t *d_b, size_t pitch_b, size_t W, size_t H,
                   double *psnr, uint64_t *sq_err);

/* Deterministic synthetic RGB8 image rows [y0, y0+rows) generated on the device */
int jb_synth_rgb_device(jb_ctx *ctx, uint64_t seed, size_t W, size_t y0, size_t rows, size_t pitch, uint8_t *d_out);

#ifdef __cplusplus
}
#endif
#endif /* JPEGB200_H */
