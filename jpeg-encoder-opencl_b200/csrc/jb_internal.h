// Internal declarations shared by the translation units of libjpegb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/jpegb200.h"
#include "jb_math.h"

namespace jb {

// Geometry of one frame (or one strip) in MCUs / blocks.
struct Geometry {
    int W, H;      // source size in pixels (unpadded)
    int sub;       // JB_SUB_*
    int mcu_px;    // 8 (444, REPL420) or 16 (420)
    int mcux, mcuy;
    int bpm;       // blocks per MCU: 3 or 6
    int n_mcu;     // per frame
    int ri;        // restart interval in MCUs (== n_mcu when none)
    int n_int;     // restart intervals per frame
};

inline Geometry make_geometry(size_t W, size_t H, int sub, int restart_interval) {
    Geometry g;
    g.W = (int)W;
    g.H = (int)H;
    g.sub = sub;
    g.mcu_px = sub == JB_SUB_420 ? 16 : 8;
    g.mcux = (int)((W + g.mcu_px - 1) / g.mcu_px);
    g.mcuy = (int)((H + g.mcu_px - 1) / g.mcu_px);
    g.bpm = sub == JB_SUB_420 ? 6 : 3;
    g.n_mcu = g.mcux * g.mcuy;
    g.ri = restart_interval > 0 ? restart_interval : g.n_mcu;
    g.n_int = (g.n_mcu + g.ri - 1) / g.ri;
    return g;
}

// Quantisation constants in natural [v][u] order; table 0 = luma, 1 = chroma.
struct QuantConst {
    float mul[2][64];    // alpha(u)alpha(v) / (4 q aan(u) aan(v))
    float band[2][64];   // 0.5 - (worst-case binary32 error of coefficient n, in quotient units)
    uint32_t dc_d[2];    // 8 q[0]: the DC coefficient is S / dc_d with S the exact integer sample sum
    uint32_t dc_m[2];    // ceil(2^32 / (2 dc_d))
    uint32_t dc_exact;   // 1 when "ties go towards zero" reproduces the reference's binary64 DC for every S
};

struct QuantTables {
    uint32_t q[2][64];  // the caller's integer tables, natural order
};

struct TransformArgs {
    const uint8_t* rgb;
    size_t pitch, frame_stride;
    int n_frames;
    Geometry g;
    int16_t* coef;          // [n_frames][n_mcu][bpm][64], zigzag order
    const uint32_t* ydown;  // 2048 words, see jb_math.h
    uint32_t* tie_list;     // near-tie coefficient indices (block*64 + zigzag position)
    uint32_t* tie_count;
    uint32_t tie_cap;
    int units_per_row;
    uint32_t total_units;
    int fast_mcux, fast_mcuy;  // MCU columns / rows the hot kernel takes (the rest goes to k_transform_edge)
    uint32_t tc_mcus;          // tcgen05 kernel: MCUs it covers (all frames); its units are runs of 16 of them
    uint32_t tc_per_frame, tc_row_len;         // tcgen05 kernels: MCUs (4:2:0) or MCU pairs (8x8 MCUs) per frame / per MCU row
    uint32_t tc_magic_frame, tc_magic_row;     // floor(2^32 / tc_per_frame), floor(2^32 / tc_row_len)
    uint32_t* unit_counter;    // zeroed per call: dynamic hand-out of strips in the tcgen05 kernel
    const uint8_t* tc_mat;     // tensor-core variant: 6 pre-swizzled bf16 matrices (null = FMA kernel)
    float tband[2][64];        // tensor-core variant: near-tie bands in zigzag order
    int inplace_dct;           // Q1 (utils.cpp:342-345): W holds the in-place map; edge blocks go to the replay whole
    const uint8_t* uv;         // NV12-style input (jb_encode_nv12_device): rgb / pitch / frame_stride describe the Y plane,
    size_t pitch_uv, frame_stride_uv;  // these the plane of interleaved Cb,Cr pairs (null: RGB input)
    int use_tma;               // JB_FLAG_TMA: stage pixels with TMA boxes (k_transform_tma) instead of per-lane cp.async
    QuantConst qc;
};

struct FixupArgs {
    const uint8_t* rgb;
    size_t pitch, frame_stride;
    Geometry g;
    int16_t* coef;
    const uint32_t* ydown;
    const uint32_t* tie_list;
    const uint32_t* tie_count;
    uint32_t tie_cap;
    const double* costab;  // [u][x] = cos((2x+1) u pi / 16), from the host's libm
    const double* scale;   // [u][v] = alpha(u) alpha(v) / 4.0
    int inplace_dct;       // Q1: replay the reference's in-place block transform up to the flagged output
    int rgb_align4;        // base, pitch and frame stride are multiples of 4: rows are read as words
    uint64_t m_bpf, m_mcux;  // ceil(2^52 / blocks per frame), ceil(2^52 / MCUs per row): divisions by multiplication
    const uint8_t* uv;     // NV12-style input: see TransformArgs
    size_t pitch_uv, frame_stride_uv;
    QuantTables qt;
};

// Huffman tables as the kernels consume them.
struct HuffDev {
    // every entry: the code LEFT-ALIGNED in the word (first bit = bit 31), its length in the low five bits, 0 = no
    // such symbol.  Lengths stay <= 27 (16 + 11 value bits; 17 + 10 with the reference's typo codes), so the two
    // fields never meet, a funnel shift can take the length from the entry itself, and (x : e) << len moves the
    // code in below x (jb_entropy.cu: SlotSink).
    uint32_t ac[2][256];  // [luma/chroma][run<<4 | cat]
    uint32_t dc[2][16];   // [luma/chroma][cat]
    uint32_t small[2][512];  // [luma/chroma][run<<5 | (v & 31)], |v| <= 15: code + value bits
};

// How entropy segments are framed in the output.
struct Framing {
    uint32_t hdr_bytes;   // JFIF header written in front of each frame's first interval (0 = none)
    uint32_t emit_eoi;    // append FFD9 after each frame's last interval
    uint32_t final_rst;   // append RSTn after the last interval instead (strips)
    uint32_t rst_phase;   // index of the first interval (strips)
    uint32_t raw_bits;    // reference-style bit string: no padding, no stuffing, no markers
};

// What the final placement kernel needs to know about one tile of 256 chunks (4 KB of unstuffed bytes), computed by
// one thread per tile (k_stuff_plan) so that no thread of k_stuff walks dependent loads while 255 others wait.
struct StuffPlan {
    uint64_t g0;    // first output byte the tile owns
    uint64_t off0;  // offset of the tile's first chunk inside its restart interval
    uint64_t nb;    // data bytes of that interval
    uint32_t i0, i1;  // restart intervals of the tile's first and last chunk
    uint32_t k;       // index of interval i0 inside its frame
    uint32_t hdr_first;  // the tile starts a frame that carries a header
};

// What k_pack needs to know about one tile of 256 blocks (k_pack_plan, one thread per tile).  A tile that lies inside
// one restart interval ("uniform": the usual case, an interval is much longer than a tile) and whose codes fit 4 KB
// leaves k_encode as a TILE STREAM (its blocks' codes concatenated, in the tile's 4 KB of the slots array) and is
// placed by plain stores ("fast"); any other tile keeps one 128-bit slot per block and is placed with atomics.
enum : uint32_t {
    PACK_UNIFORM = 1u,    // all blocks of the tile belong to one restart interval
    PACK_FAST = 2u,       // uniform and in stream form
    PACK_STARTS = 4u,     // the tile's first block starts its interval
    PACK_ENDS = 8u,       // the tile's last block ends its interval
    PACK_PREV_FAST = 16u, // the tile before continues the same interval and is fast: its last bits arrive in prev_tail
    PACK_NEXT_FAST = 32u, // the tile after continues the same interval and is fast: it writes the word the two share
};
constexpr uint32_t STREAM_MAX_BITS = 256u * 128u;  // a tile's stream lives in its 256 slots
struct PackPlan {
    uint64_t c;              // bit position of the tile's first block in the unstuffed buffer
    uint64_t slot_end_word;  // first 32-bit word after the interval's reservation (zero fill up to it when the tile ends the
                             // interval); for a tile placed with atomics: first word after the range it may touch
    uint32_t last_b;         // last block of that interval (it appends the 1-padding)
    uint32_t flags;          // PACK_*
    uint32_t prev_tail;      // the previous tile's bits of the word the two tiles share, in place (fast after fast)
    uint32_t tile_bits;
};

// Device work arrays of the entropy coder (all sized by the context).
struct EntropyWork {
    uint32_t* blk_prefix;   // [n_blocks] exclusive bit prefix inside its 256-block tile
    uint32_t* blk_len;      // [n_blocks] code length of every block in bits (tiles in slot form only)
    uint4* slots;           // [n_tiles * 256] per tile: its stream, or one right-aligned 128-bit slot per block (PackPlan)
    uint32_t* long_list;    // [n_blocks] blocks whose code does not fit a slot
    uint32_t* n_long;       // device scalar
    uint32_t* any_slow;     // device scalar: number of tiles placed with atomics (their part of the buffer is cleared first)
    uint32_t* slow_list;    // [n_tiles] those tiles, in no particular order
    uint32_t* tile_bits;    // [n_tiles]
    uint64_t* tile_base;    // [n_tiles + 1] exclusive scan of tile_bits
    uint32_t* int_slot;     // [n_int_total] bytes reserved in the unstuffed buffer (multiple of 16)
    uint64_t* int_bits;     // [n_int_total]
    uint64_t* int_ubase;    // [n_int_total + 1] byte offset of each interval in the unstuffed buffer
    uint8_t* ubuf;          // unstuffed bytes
    uint64_t ubuf_cap;
    uint32_t* ff_prefix;    // [ubuf_cap / 16] exclusive 0xFF count inside its chunk tile
    uint32_t* ff_tile;      // [ubuf_cap / 16 / 256 + 1]
    uint64_t* ff_tile_base; // same + 1
    uint32_t* n_ff_tiles;   // device scalar
    uint32_t* int_osize;    // [n_int_total]
    uint64_t* int_obase;    // [n_int_total + 1]
    StuffPlan* stuff_plan;  // [ubuf_cap / 16 / 256 + 2]
    PackPlan* pack_plan;    // [n_tiles]
    uint64_t* scan_tmp;     // scratch of the multi-CTA scans
    uint64_t* status;       // [0] error bits, [1] required ubuf bytes, [2] required out bytes, [3] total bits
};

struct EntropyArgs {
    const int16_t* coef;
    Geometry g;
    int n_frames;
    uint32_t n_blocks;     // total blocks in the batch
    uint32_t n_int_total;  // n_frames * g.n_int
    const HuffDev* huff;
    uint64_t m_bpf, m_ri;  // ceil(2^52 / blocks per frame), ceil(2^52 / restart interval): see div_magic
    uint32_t m32_ri;       // floor(2^32 / restart interval), saturated: quotient estimate at most one too small
    uint32_t always_eob;
    uint32_t no_tma;       // JB_FLAG_ENTROPY_LDG: k_encode stages its tile with loads (the no-tensor-map fallback)
    Framing fr;
    EntropyWork w;
    const uint8_t* hdr;    // device copy of the JFIF header
    uint8_t* out;
    uint64_t out_cap;
    uint64_t* frame_off;   // [n_frames] (may be null)
    uint64_t* frame_size;  // [n_frames] (may be null)
    uint64_t* total_out;   // device scalar (may be null)
    const uint64_t* out_off;  // device scalar (may be null): the segment goes to out + *out_off (strip stitch)
};

#define JB_STATUS_UBUF_OVERFLOW 1ull
#define JB_STATUS_OUT_OVERFLOW 2ull
#define JB_STATUS_TIE_OVERFLOW 4ull
#define JB_STATUS_PEER_TIMEOUT 8ull

// ---- decode path (jb_decode.cu) -------------------------------------------------------------------------------------
struct DecTable {          // one DHT table, T.81 F.2.2.3 decoding arrays + a 9-bit look-ahead
    uint16_t look[512];    // (code length << 8) | symbol for codes of up to 9 bits, 0 = longer
    int32_t maxcode[18], mincode[17];
    uint16_t valptr[17];
    uint8_t vals[256];
};
struct DecTables {
    DecTable t[4];  // DC 0, DC 1, AC 0, AC 1
};
struct JfifInfo {
    uint32_t W, H;
    int32_t sub;  // JB_SUB_444 (also what the replicated 4:2:0 mode is coded as) or JB_SUB_420
    uint32_t restart_interval;
    uint64_t scan_offset;       // first byte of the entropy-coded data
    uint32_t dc_tab[3], ac_tab[3];
    uint32_t q[3][64];          // quantisation table of every component, natural order
};
int parse_jfif(const uint8_t* d, size_t n, JfifInfo* o, DecTables* tabs);  // JB_E_NOSPACE: the header continues past n
int launch_rst_index(const uint8_t* d_scan, size_t n, uint32_t* d_cnt, uint32_t* d_groups, uint64_t* d_start, uint32_t max_int, cudaStream_t s);
int launch_huff_decode(const uint8_t* d_scan, size_t n, const uint64_t* d_start, uint32_t n_int, uint32_t ri, uint32_t n_mcu, int bpm,
                       const DecTables* d_tabs, const JfifInfo& info, int16_t* d_coef, cudaStream_t s);
int launch_reconstruct(const int16_t* d_coef, uint32_t n_mcu, int mcux, int bpm, const JfifInfo& info, uint8_t* py, uint8_t* pcb, uint8_t* pcr,
                       size_t pitch_y, size_t pitch_c, uint8_t* d_rgb, size_t pitch, cudaStream_t s);
int launch_sq_err(const uint8_t* a, size_t pitch_a, const uint8_t* b, size_t pitch_b, size_t W, size_t H, unsigned long long* d_sum, cudaStream_t s);

// ---- launchers (each returns the number of kernels it launched) -------------
int launch_transform(const TransformArgs& a, cudaStream_t s);       // MCUs inside the image (hot kernel)
int launch_transform_edge(const TransformArgs& a, cudaStream_t s);  // MCUs that need mirror padding
int launch_transform_nv12(const TransformArgs& a, cudaStream_t s);       // NV12-style input: tcgen05 kernel, or all MCUs on CUDA cores
int launch_transform_nv12_edge(const TransformArgs& a, cudaStream_t s);  // ... the MCUs the tcgen05 kernel leaves out
int launch_rgb_to_nv12(const uint8_t* rgb, size_t W, size_t H, size_t pitch, const uint32_t* ydown, uint8_t* y, size_t pitch_y, uint8_t* uv,
                       size_t pitch_uv, cudaStream_t s);
int launch_fixup(const FixupArgs& a, cudaStream_t s);
void* tensor_map_encode_fn();  // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, or null (jb_transform.cu)
int launch_entropy(const EntropyArgs& a, cudaStream_t s, int phase = 0);  // 1: up to the sizes, 2: final placement only
int launch_stitch_exchange(uint64_t* ctl, int rank, int world, uint64_t epoch, uint64_t base, const uint64_t* d_len, uint64_t* d_off,
                           cudaStream_t s);
int launch_stitch_complete(uint64_t* ctl, int rank, int world, int dst, uint64_t epoch, uint64_t* status, cudaStream_t s);
int launch_copy_bytes(uint8_t* dst_base, const uint64_t* d_dst_off, const uint8_t* src, const uint64_t* d_len, uint64_t cap,
                      uint64_t* status, cudaStream_t s);
int launch_synth(uint64_t seed, size_t W, size_t y0, size_t rows, size_t pitch, uint8_t* d_out, cudaStream_t s);

// staged kernels (device pointers)
int launch_csc(uint8_t* px, size_t n, const uint32_t* ydown, cudaStream_t s);
int launch_cds(uint8_t* px, size_t W, size_t H, cudaStream_t s);
int launch_pad(const uint8_t* src, size_t W, size_t H, uint8_t* dst, size_t nW, size_t nH, cudaStream_t s);
int launch_u8_to_f64(const uint8_t* src, double* dst, size_t n, cudaStream_t s);
int launch_sub_f64(double* img, size_t n, double val, cudaStream_t s);
int launch_dct_f64(double* img, size_t W, size_t H, int inplace, const double* costab, const double* scale,
                   cudaStream_t s);
int launch_quant_f64(double* img, size_t W, size_t H, const QuantTables& qt, cudaStream_t s);
int launch_blockify(const double* img, size_t W, size_t H, int32_t* linear, cudaStream_t s);
int launch_aos_to_planar_u32(const uint8_t* px, size_t n, uint32_t* out, cudaStream_t s);
int launch_planar_u32_interleave(const uint32_t* in, size_t n, uint32_t* out, cudaStream_t s);
int launch_planar_u32_to_rgb8(const uint32_t* in, size_t W, size_t H, uint8_t* out, size_t pitch, cudaStream_t s);
int launch_pad_planar_u32(const uint32_t* in, size_t W, size_t H, uint32_t* out, size_t nW, size_t nH, cudaStream_t s);
int launch_blockify_planar_i32(const int32_t* in, size_t W, size_t H, int32_t* linear, cudaStream_t s);
int launch_f64_to_u8(const double* src, uint8_t* dst, size_t n, cudaStream_t s);
int launch_value_categories(const int16_t* v, size_t n, uint8_t* cat, uint16_t* bits, cudaStream_t s);
int launch_remove_red(uint8_t* px, size_t n, cudaStream_t s);
int launch_zigzag(const int32_t* linear, int32_t* zz, size_t rows, cudaStream_t s);
int launch_rle(const int32_t* zz, size_t rows, int always_eob, int32_t* pairs, uint32_t* counts, cudaStream_t s);
int launch_planar_to_scan(const int32_t* zz, size_t rpc, int16_t* coef, cudaStream_t s);

// host helpers
void build_quant_const(const uint32_t ql[64], const uint32_t qc[64], QuantConst* out);
void aan_error_bound(double err[64], double amax[64]);
void build_tc_matrices(const uint32_t ql[64], const uint32_t qc[64], double step_ulps, int repl_chroma, int inplace_dct,
                       uint8_t* out /* 32768 B */,
                       float tband[2][64]);  // worst-case |binary32 - exact| per AAN output
void build_huff(bool typo, HuffDev* out);
void build_ydown(uint32_t ydown[2048]);
void build_dct_tables(double costab[64], double scale[64]);
// BITS / HUFFVAL lists of the four tables in DHT order: DC luma, AC luma, DC chroma, AC chroma
struct HuffSpecs {
    uint8_t bits[4][16];
    uint8_t vals[4][256];
    int n[4];
};
void annex_k_specs(HuffSpecs* sp);
void optimal_spec(const uint64_t counts[256], uint8_t bits[16], uint8_t vals[256], int* n);
void optimal_huff_specs(const uint64_t counts[4][256], HuffSpecs* sp);  // T.81 K.2 as in libjpeg (counts in DHT order)
void build_huff_from_specs(const HuffSpecs& sp, HuffDev* out);
size_t build_header(const jb_params* p, size_t W, size_t H, uint8_t* out, const HuffSpecs* custom = nullptr);  // out >= 2048 bytes
int launch_symbol_hist(const EntropyArgs& a, uint32_t* d_hist /* [4][256], DHT order */, cudaStream_t s);
extern const uint8_t kZigzag[64];  // zigzag position -> natural index

}  // namespace jb
