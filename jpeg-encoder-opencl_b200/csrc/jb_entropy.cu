// GPU entropy coder: RLE + Huffman (utils.cpp:572-698) as a multi-pass parallel
// coder, plus what the reference lacks for a real JPEG stream: byte packing,
// 1-padding, 0xFF00 stuffing, RSTn/EOI markers.
//
//   k_encode    one thread per 8x8 block; the tile of 256 blocks is staged in shared memory by one TMA box
//               (cp.async.bulk.tensor, SWIZZLE_128B): non-zero mask, blocks sorted by their number of non-zeros,
//               sparse walk of the run/size symbols into a 128-bit register sink, code lengths scanned inside the
//               tile, the tile's codes assembled as ONE stream (a tile inside one restart interval with at most
//               4 KB of codes) or kept as one 128-bit slot per block
//   k_scan      exclusive scan of the tile totals (device-wide bit offsets)
//   k_intervals bits / reserved bytes of every restart interval
//   k_scan      byte offset of every interval in the unstuffed buffer
//   k_pack_plan per tile: its place in the unstuffed buffer, what it shares with its neighbours
//   k_zero      clear what the tiles in slot form are going to OR into (nothing, a few ranges, or everything)
//   k_pack      a warp per tile: the stream as a funnel-shifted copy with plain 128-bit stores (tiles in slot
//               form: a lane per block, atomics; + 1-padding at interval ends)
//   k_pack_long blocks longer than a slot are re-walked
//   k_ff_count  0xFF bytes per 16-byte chunk, scan inside 256-chunk tiles
//   k_scan      device-wide 0xFF prefix
//   k_int_out   output bytes of every interval (data + stuffing + marker + header)
//   k_scan      output offset of every interval / frame
//   k_finalize  capacity check, frame table
//   k_stuff_plan / k_stuff  assemble the final bytes per 4 KB tile in (word-swizzled) shared memory (0x00 after
//               0xFF, markers, JFIF headers) and store them with coalesced 128-bit stores
// Bit order is MSB first; the unstuffed buffer is addressed as big-endian words.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>

#include "jb_pixels.cuh"  // cat_bits (utils.cpp:623-653)

namespace jb {

constexpr int TILE = 256;

// ------------------------------------------------------------ block walker --
struct BitSink {
    uint64_t acc;
    int n;
    uint32_t* wp;
    __device__ __forceinline__ void init(uint8_t* base, uint64_t bitpos) {
        wp = reinterpret_cast<uint32_t*>(base) + (bitpos >> 5);
        n = (int)(bitpos & 31);
        acc = 0;
    }
    __device__ __forceinline__ void put(uint32_t code, int len) {
        acc = (acc << len) | code;
        n += len;
        if (n >= 32) {
            n -= 32;
            atomicOr(wp, __byte_perm((uint32_t)(acc >> n), 0, 0x0123));
            ++wp;
        }
    }
    // table-entry form (HuffDev): the code left-aligned in the word, its length in the low five bits
    __device__ __forceinline__ void put(uint32_t e) {
        const int len = (int)(e & 31u);
        if (len) put(e >> (32 - len), len);
    }
    __device__ __forceinline__ void finish() {
        if (n > 0) atomicOr(wp, __byte_perm((uint32_t)(acc << (32 - n)), 0, 0x0123));
    }
};

struct BlockInfo {
    int comp;          // 0 Y, 1 Cb, 2 Cr
    uint32_t prev;     // block whose DC is the predictor (valid when has_prev)
    bool has_prev;     // false at the start of a restart interval: predictor 0
    uint32_t interval; // global restart-interval index
    bool last_in_interval;
};

// x / d for x, d < 2^26 with the host-computed magic m = ceil(2^52 / d): exact because x * d < 2^52
__device__ __forceinline__ uint32_t div_magic(uint32_t x, uint64_t m) {
    return (uint32_t)__umul64hi((uint64_t)x << 12, m);
}

__device__ __forceinline__ BlockInfo block_info(const EntropyArgs& a, uint32_t b) {
    BlockInfo bi;
    const uint32_t bpm = (uint32_t)a.g.bpm, bpf = (uint32_t)a.g.n_mcu * bpm, ri = (uint32_t)a.g.ri;
    uint32_t f = div_magic(b, a.m_bpf), rb = b - f * bpf;
    uint32_t mcu = __umulhi(rb, 0xAAAAAAABu) >> (bpm == 3 ? 1 : 2), j = rb - mcu * bpm;  // bpm is 3 or 6
    uint32_t k = div_magic(mcu, a.m_ri);
    bool first = mcu - k * ri == 0;  // first MCU of its restart interval: predictors are 0
    bi.interval = f * (uint32_t)a.g.n_int + k;
    uint32_t mcu_end = min((k + 1) * ri, (uint32_t)a.g.n_mcu);
    bi.last_in_interval = (mcu == mcu_end - 1) && (j == bpm - 1);
    if (bpm == 3) {
        bi.comp = (int)j;
        bi.has_prev = !first;
        bi.prev = b - 3;
    } else {
        bi.comp = j < 4 ? 0 : (int)j - 3;
        if (j >= 1 && j <= 3) {
            bi.has_prev = true;
            bi.prev = b - 1;
        } else {
            bi.has_prev = !first;
            bi.prev = b - (j == 0 ? 3u : 6u);
        }
    }
    return bi;
}

__device__ __forceinline__ void load_tables(const EntropyArgs& a, uint32_t (&s_ac)[2][256], uint32_t (&s_dc)[2][16]) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_ac[0][0])[i] = (&a.huff->ac[0][0])[i];
    if (threadIdx.x < 32) (&s_dc[0][0])[threadIdx.x] = (&a.huff->dc[0][0])[threadIdx.x];
    __syncthreads();
}

// exclusive scan of one value per thread over a 256-thread CTA; total in *total.  REUSE: s_warp is written again later
// (a second barrier keeps this call's readers ahead of that)
template <bool REUSE = true>
__device__ __forceinline__ uint32_t cta_scan_256(uint32_t x, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t t = s_warp[i];
        if (i < wid) base += t;
        tot += t;
    }
    if (REUSE) __syncthreads();
    total = tot;
    return base + inc - x;
}

// ---- sparse block walk ---------------------------------------------------------------
// 0xFFFF in each half of the result whose int16 half of w is non-zero
__device__ __forceinline__ uint32_t nz_flags(uint32_t w) {
    const __half2 z = __half2half2(__ushort_as_half((unsigned short)0));
    return __hneu2_mask(*reinterpret_cast<const __half2*>(&w), z);
}
// bits 0..15 of x -> even bits, bits 16..31 -> odd bits of the result
__device__ __forceinline__ uint32_t interleave16(uint32_t x) {
    uint32_t e = x & 0xFFFFu, o = x >> 16;
    e = (e | (e << 8)) & 0x00FF00FFu; o = (o | (o << 8)) & 0x00FF00FFu;
    e = (e | (e << 4)) & 0x0F0F0F0Fu; o = (o | (o << 4)) & 0x0F0F0F0Fu;
    e = (e | (e << 2)) & 0x33333333u; o = (o | (o << 2)) & 0x33333333u;
    e = (e | (e << 1)) & 0x55555555u; o = (o | (o << 1)) & 0x55555555u;
    return e | (o << 1);
}

// eight mask bits (bit i = coefficient 8 pc + i is non-zero) of piece pc of a block, inserted into m[pc >> 2]
__device__ __forceinline__ void nz_mask_piece(const uint4 q, int pc, uint32_t (&m)[2]) {
    const uint32_t fa = __byte_perm(nz_flags(q.x), nz_flags(q.y), 0x6420);
    const uint32_t fb = __byte_perm(nz_flags(q.z), nz_flags(q.w), 0x6420);
    const uint32_t x = (fa & 0x08040201u) | (fb & 0x80402010u);
    const uint32_t sel = (pc & 3) == 0 ? 0x3217u : (pc & 3) == 1 ? 0x3270u : (pc & 3) == 2 ? 0x3710u : 0x7210u;
    m[pc >> 2] = __byte_perm(m[pc >> 2], x * 0x01010101u, sel);
}

// Sink of k_encode: the block's code as ONE right-aligned 128-bit number in four registers (s3 most significant,
// the first code bit the most significant one of the n bits) plus the bit count.  A symbol is four funnel shifts:
// no flush branch, no shared-memory traffic inside the walk.  Codes longer than 128 bits lose their first bits here;
// k_pack_long re-walks those blocks.
struct SlotSink {
    uint32_t s0, s1, s2, s3, n;
    __device__ __forceinline__ void init() { s0 = s1 = s2 = s3 = n = 0u; }
    // e = code left-aligned | length (HuffDev): the funnel shifts take the length from the low five bits of e
    // themselves, and (s0 : e) << len drops the length field while it moves the code in below s0
    __device__ __forceinline__ void put(uint32_t e) {
        s3 = __funnelshift_l(s2, s3, e);
        s2 = __funnelshift_l(s1, s2, e);
        s1 = __funnelshift_l(s0, s1, e);
        s0 = __funnelshift_l(e, s0, e);
        n += e & 31u;
    }
};

// Huffman code t (left-aligned | length hl) followed by the cat value bits of v (utils.cpp:623-653: v for v > 0, the
// low cat bits of v - 1 for v < 0), as one entry of the same form: hl + cat <= 27 for every table the coder builds.
template <bool MAY_BE_ZERO>
__device__ __forceinline__ uint32_t with_value_bits(uint32_t t, int v, int cat) {
    const uint32_t x = (uint32_t)(v + (v >> 31));
    uint32_t vbl = x << ((32 - cat) & 31);  // value bits left-aligned (bits above cat fall off)
    if (MAY_BE_ZERO && cat == 0) vbl = 0u;
    return (t + (uint32_t)cat) | (vbl >> (t & 31u));
}

// The block as HuffmanEncoder codes it (utils.cpp:667-694), visiting only the non-zero AC
// coefficients: bit k of (mlo, mhi) = coefficient k != 0, value(k) fetches coefficient k.  s_small (may
// be null) maps (run <= 15, |v| <= 15) straight to code + value bits: the common case is one
// look-up instead of category, value bits and code assembly.
template <bool SMALL, class Sink, class Fetch>
__device__ __forceinline__ void encode_sparse(uint32_t mlo, uint32_t mhi, Fetch value, int dc_diff, const uint32_t* s_ac,
                                              const uint32_t* s_dc, const uint32_t* s_small, bool always_eob, Sink& s) {
    {
        const int cat = 32 - __clz(abs(dc_diff));
        s.put(with_value_bits<true>(s_dc[cat], dc_diff, cat));
    }
    int cur = 1;  // next AC position to account for
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {  // positions 1..31, then 32..63: 32-bit mask arithmetic
        uint32_t m = half ? mhi : (mlo & ~1u);
        const int base = 32 * half - 1;
        while (m) {
            const int pos = __ffs((int)m) + base;
            m &= m - 1;
            int run = pos - cur;
            cur = pos + 1;
            const int v = value(pos);
            uint32_t e;
            if (SMALL && run < 16 && (uint32_t)(v + 15) <= 30u) {
                // byte offset (run * 32 + (v & 31)) * 4: the two fields do not overlap
                e = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(s_small) + ((run << 7) | ((v << 2) & 0x7C)));
            } else {
#pragma unroll 1
                while (run >= 16) {  // ZRL, utils.cpp:592-597
                    s.put(s_ac[0xF0]);
                    run -= 16;
                }
                const int cat = 32 - __clz(abs(v));
                e = with_value_bits<false>(s_ac[(run << 4) | cat], v, cat);
            }
            s.put(e);
        }
    }
    if (cur < 64 || always_eob) s.put(s_ac[0]);  // EOB, utils.cpp:607-608 (Q3 when always_eob)
}

// ---- k_encode's walk: the same coding as encode_sparse, on shared-memory addresses ----------------------------
__device__ __forceinline__ int lds_s16(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// cblk: shared address of the block's 128 bytes (XOR-swizzled by swz = (block & 7) << 4); dc_tab / small_tab: shared
// addresses of the block's DC and small-value tables (small_tab 128-byte aligned); g_ac: its full AC table (global
// memory / L1: symbols outside the small-value table, ZRL, EOB).  Positions and runs are kept DOUBLED (byte offsets
// of int16 coefficients): the doubled position is the coefficient's address, the doubled run times 64 its table row.
__device__ __forceinline__ void walk_block(uint32_t mlo, uint32_t mhi, uint32_t cblk, uint32_t swz, int dc_diff,
                                           const uint32_t* __restrict__ g_ac, uint32_t dc_tab, uint32_t small_tab, bool always_eob,
                                           SlotSink& s) {
    {
        const int cat = 32 - __clz(abs(dc_diff));
        s.put(with_value_bits<true>(lds_u32(dc_tab + 4u * (uint32_t)cat), dc_diff, cat));
    }
    int cur2 = 2;  // (doubled) next AC position to account for
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {  // positions 1..31, then 32..63: 32-bit mask arithmetic
        uint32_t m = half ? mhi : (mlo & ~1u);
        const int base2 = 64 * half - 2;
        while (m) {
            const int pos2 = 2 * __ffs((int)m) + base2;
            m &= m - 1;
            int run2 = pos2 - cur2;
            cur2 = pos2 + 2;
            const int v = lds_s16(cblk | ((uint32_t)pos2 ^ swz));
            uint32_t e;
            if (((uint32_t)(v + 15) | ((uint32_t)run2 & ~31u)) <= 30u) {  // |v| <= 15 and run < 16: one look-up
                e = lds_u32((((uint32_t)v << 2) & 0x7Cu) | (small_tab + (uint32_t)run2 * 64u));
            } else {
#pragma unroll 1
                while (run2 >= 32) {  // ZRL, utils.cpp:592-597
                    s.put(__ldg(g_ac + 0xF0));
                    run2 -= 32;
                }
                const int cat = 32 - __clz(abs(v));
                e = with_value_bits<false>(__ldg(g_ac + (((uint32_t)run2 << 3) | (uint32_t)cat)), v, cat);
            }
            s.put(e);
        }
    }
    if (cur2 < 128 || always_eob) s.put(__ldg(g_ac));  // EOB, utils.cpp:607-608 (Q3 when always_eob)
}

__device__ __forceinline__ void or_slot(uint32_t* words, long long pos, uint32_t len, const uint4 q, bool shared_space) {
    // the slot is a right-aligned 128-bit number: its (all-zero) first bit sits 128 - len bits before pos --
    // possibly before the buffer / window; zero words are never written
    const uint32_t w[6] = {0u, q.x, q.y, q.z, q.w, 0u};
    const long long p0 = pos + (long long)len - 128;
    const uint32_t sh = (uint32_t)p0 & 31u;
    uint32_t* dst = words + (p0 >> 5);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const uint32_t o = __funnelshift_r(w[j + 1], w[j], sh);  // bits of (w[j]:w[j+1]) >> sh
        if (o) {
            if (shared_space)
                atomicOr(dst + j, o);  // (bytes are swapped on the way out)
            else
                atomicOr(dst + j, __byte_perm(o, 0, 0x0123));
        }
    }
}

#ifndef ENC_CTAS
#define ENC_CTAS 5
#endif
#ifndef ENC_PREFETCH
#define ENC_PREFETCH 740  // tiles ahead: 148 SMs x 5 resident CTAs = one CTA lifetime (0: 733 us, 148 - 740: 717 us, 1480: ~790 us)
#endif
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One thread per block.  The tile's coefficients (256 blocks, 32 KB, contiguous in global memory) are staged in shared
// memory with the 128-byte XOR swizzle (piece p of block t at t*8 + (p ^ (t&7)): a thread that reads its own block and
// its neighbours theirs hit different banks) -- by ONE cp.async.bulk.tensor (TMA = true: a 2-D tensor map over the
// coefficient array, box = 256 rows x 128 bytes, SWIZZLE_128B produces exactly that layout; completion on an
// mbarrier) or, for coefficient arrays the tensor map cannot describe, by coalesced 128-bit loads and stores.
// Each thread builds the non-zero mask of its block; the blocks are counting-sorted by their number of non-zero
// coefficients; thread t walks the non-zero coefficients of block perm[t]; code lengths are scanned per tile.
template <bool TMA>
__global__ void __launch_bounds__(TILE, ENC_CTAS) k_encode(const __grid_constant__ EntropyArgs a, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(1024) uint4 s_coef[TILE * 8];
    __shared__ __align__(16) uint32_t s_slot[TILE * 4];
    __shared__ __align__(8) uint2 s_mask[TILE];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(128) uint32_t s_small[2][512];
    __shared__ __align__(16) uint32_t s_dc[2][16];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_len[TILE];
    __shared__ uint32_t s_hist[64], s_start[64];
    __shared__ uint16_t s_perm[TILE];
    __shared__ uint32_t s_uniform;
    const uint32_t t = threadIdx.x, b0 = blockIdx.x * TILE, b = b0 + t;
    const uint32_t n_blk = min((uint32_t)TILE, a.n_blocks - b0);
    if (t == 32)  // (k_pack_plan's definition; read after the barriers below)
        s_uniform = block_info(a, b0).interval == block_info(a, b0 + n_blk - 1).interval ? 1u : 0u;
    if (TMA) {
        if (t == 0) {
            const uint32_t bar = smem_addr(&s_bar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            // rows past the end of the array are filled with zeros and count as transferred bytes; the two tables
            // the walk reads from shared memory travel as plain bulk copies on the same barrier
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(TILE * 128 + 4096 + 128) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             smem_addr(s_coef)),
                         "l"(&tmap), "r"(0), "r"((int)b0), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s_small)),
                         "l"(&a.huff->small[0][0]), "n"(4096), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s_dc)),
                         "l"(&a.huff->dc[0][0]), "n"(128), "r"(bar)
                         : "memory");
#if ENC_PREFETCH
            // the tile a CTA that starts about one CTA lifetime from now will ask for: into L2 ahead of its TMA
            const uint32_t b_pf = b0 + ENC_PREFETCH * TILE;
            if (b_pf + TILE <= a.n_blocks)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.coef + (size_t)b_pf * 64), "n"(TILE * 128) : "memory");
#endif
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(a.coef) + (size_t)b0 * 8;
        const uint32_t n_here = n_blk * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t g = i * TILE + t, blk = g >> 3, pc = g & 7;
            if (g < n_here) s_coef[blk * 8 + (pc ^ (blk & 7))] = __ldg(src + g);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) (&s_small[0][0])[i * TILE + t] = (&a.huff->small[0][0])[i * TILE + t];
        if (t < 32) (&s_dc[0][0])[t] = (&a.huff->dc[0][0])[t];
    }
    if (t < 64) s_hist[t] = 0;
    __syncthreads();
    if (TMA) {  // (the barrier above made the mbarrier's initialisation visible to every thread)
        const uint32_t bar = smem_addr(&s_bar);
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(bar), "r"(0)
                         : "memory");
    }
    // ---- non-zero mask of this thread's block, then a counting sort of the tile's blocks by
    // their number of non-zero AC coefficients: the walk below costs one loop iteration per
    // non-zero, so warps made of similar blocks do not wait for their busiest lane.
    uint32_t cnt = 0;
    if (t < n_blk) {
        // One half2 "not equal (unordered) to zero" compare flags both int16 halves of a word (every non-zero
        // coefficient |c| <= 2047 is a non-zero, possibly subnormal or NaN, binary16 pattern): 0xFFFF per half.  Two
        // byte permutes line up one flag byte per coefficient of a piece (eight coefficients, zigzag order), two
        // LOP3 keep bit i of coefficient i, and a multiplication by 0x01010101 adds the four bytes into the top one:
        // the piece's eight mask bits, inserted into the mask by a third permute.  (The first version of this loop
        // collected even and odd coefficients in two bit planes and interleaved them afterwards: +50 instructions.)
        uint32_t m[2] = {0u, 0u};
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) nz_mask_piece(s_coef[t * 8 + (pc ^ (t & 7))], pc, m);
        s_mask[t] = make_uint2(m[0], m[1]);
        cnt = (uint32_t)(__popc(m[0] & ~1u) + __popc(m[1]));
        atomicAdd(&s_hist[cnt], 1u);
    }
    __syncthreads();
    if (t < 32) {  // exclusive scan of the 64 bins, busiest blocks first
        uint32_t h0 = s_hist[63 - 2 * t], h1 = s_hist[62 - 2 * t], inc = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (t >= (uint32_t)o) inc += y;
        }
        s_start[63 - 2 * t] = inc - h0 - h1;
        s_start[62 - 2 * t] = inc - h1;
    }
    __syncthreads();
    if (t < n_blk) s_perm[atomicAdd(&s_start[cnt], 1u)] = (uint16_t)t;
    __syncthreads();
    // ---- encode block s_perm[t] ------------------------------------------------------------
    if (t < n_blk) {
        const uint32_t k = s_perm[t], bk = b0 + k;
        // coefficient pos of block k: byte k*128 + (((pos >> 3) ^ (k & 7)) << 4) + (pos & 7) * 2 = k*128 | ((2 pos) ^ swz)
        const uint32_t cblk = smem_addr(s_coef) + k * 128u, swz = (k & 7u) << 4;
        // Component, DC predictor and "first MCU of a restart interval" of block bk.  The tile's place in its frame
        // depends on blockIdx alone (uniform registers); a tile crosses at most one frame boundary unless frames are
        // smaller than a tile.  mcu % ri == 0 comes from a 32-bit reciprocal whose quotient is at most one too
        // small: the remainder estimate is then 0 or ri.
        const uint32_t bpm = (uint32_t)a.g.bpm, bpf = (uint32_t)a.g.n_mcu * bpm, ri = (uint32_t)a.g.ri;
        uint32_t rb;
        if (bpf >= (uint32_t)TILE) {
            rb = (b0 - div_magic(b0, a.m_bpf) * bpf) + k;
            if (rb >= bpf) rb -= bpf;
        } else {
            rb = bk - div_magic(bk, a.m_bpf) * bpf;
        }
        const uint32_t mcu = __umulhi(rb, 0xAAAAAAABu) >> (bpm == 3 ? 1 : 2), j = rb - mcu * bpm;  // bpm is 3 or 6
        const uint32_t r_est = mcu - __umulhi(mcu, a.m32_ri) * ri;
        const bool first = r_est == 0u || r_est == ri;  // first MCU of its restart interval: predictors are 0
        // per position j in the MCU, one nibble / bit each: distance to the block whose DC is the predictor
        // (Y00 <- Y11 of the MCU before, Y01..Y11 <- the block before, Cb / Cr <- one MCU back), and whether that
        // block lies in the MCU before (no predictor at the start of an interval)
        const uint32_t dist = ((bpm == 3 ? 0x333u : 0x661113u) >> (4u * j)) & 15u;
        const bool crosses = ((bpm == 3 ? 0x7u : 0x31u) >> j) & 1u;
        const uint32_t tab = j >= (bpm == 3 ? 1u : 4u) ? 1u : 0u;
        int pred = 0;  // DC of the previous block of the component (utils.cpp:669-670)
        if (!(first && crosses)) {
            const uint32_t prev = bk - dist;
            if (prev >= b0) {
                const uint32_t pt = prev - b0;  // coefficient 0 of block pt: piece 0 ^ (pt & 7)
                pred = lds_s16(smem_addr(s_coef) + pt * 128u + ((pt & 7u) << 4));
            } else {
                pred = (int)a.coef[(size_t)prev * 64];
            }
        }
        const uint2 mk = s_mask[k];
        SlotSink s;
        s.init();
        walk_block(mk.x, mk.y, cblk, swz, lds_s16(cblk | swz) - pred, a.huff->ac[tab], smem_addr(s_dc[tab]), smem_addr(s_small[tab]),
                   a.always_eob != 0, s);
        s_len[k] = s.n;
        *reinterpret_cast<uint4*>(s_slot + k * 4) = make_uint4(s.s3, s.s2, s.s1, s.s0);  // stream order
        if (s.n > 128) {  // too long for a slot: k_pack_long re-walks it
            uint32_t idx = atomicAdd(a.w.n_long, 1u);
            a.w.long_list[idx] = bk;
        }
    }
    __syncthreads();  // every walk has ended: the coefficients are dead, s_coef becomes the stream window
    uint32_t bits = 0;
    uint4 slot = make_uint4(0u, 0u, 0u, 0u);
    if (t < n_blk) {
        bits = s_len[t];
        slot = *reinterpret_cast<const uint4*>(s_slot + t * 4);
    }
    // (the stream window, cleared here: the barriers inside the scan order these stores before the ORs below)
    uint32_t* win = reinterpret_cast<uint32_t*>(s_coef);
    reinterpret_cast<uint4*>(win)[t] = make_uint4(0u, 0u, 0u, 0u);
    uint32_t total;
    const uint32_t ex = cta_scan_256<false>(bits, s_warp, total);  // (the CTA scans once)
    a.w.blk_prefix[b] = ex;  // padded to a whole tile
    if (threadIdx.x == 0) a.w.tile_bits[blockIdx.x] = total;
    if (s_uniform && total <= STREAM_MAX_BITS) {
        // Tile stream (PackPlan): the codes of the tile's blocks concatenated at their in-tile bit offsets (blocks
        // longer than a slot leave a gap of zeros for k_pack_long), assembled with shared-memory atomics and stored
        // as whole 128-bit words into the tile's 4 KB of the slots array: 1/3 of the bytes that one slot, one
        // length and one offset per block take, and k_pack places a tile with plain stores.
        const uint32_t n_vec = (((max(total, 1u) + 31) >> 5) + 3) >> 2;  // (<= 256: one 128-bit word per thread)
        if (t < n_blk && bits <= 128) or_slot(win, (long long)ex, bits, slot, true);
        __syncthreads();
        if (t < n_vec) a.w.slots[(size_t)blockIdx.x * TILE + t] = reinterpret_cast<const uint4*>(win)[t];
    } else if (t < n_blk) {
        a.w.blk_len[b] = bits;
        a.w.slots[b] = slot;
    }
}

// Single-CTA exclusive scan: out[i] = sum in[0..i), out[n] = total.
template <class T>
__global__ void __launch_bounds__(1024) k_scan(const T* __restrict__ in, uint64_t* __restrict__ out, uint32_t n,
                                               const uint32_t* n_dev) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (n_dev) n = *n_dev;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096) {
        uint32_t i0 = base + threadIdx.x * 4;
        uint64_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = i0 + j < n ? (uint64_t)in[i0 + j] : 0ull;
        uint64_t sum = v[0] + v[1] + v[2] + v[3], inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint64_t wbase = 0, tot = 0;
        for (int i = 0; i < 32; ++i) {
            uint64_t t = s_warp[i];
            if (i < wid) wbase += t;
            tot += t;
        }
        uint64_t ex = s_carry + wbase + inc - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) out[i0 + j] = ex;
            ex += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

// Device-wide exclusive scan for long arrays: per-chunk sums (k_scan_partial), single-CTA scan of
// the chunk sums (k_scan<uint64_t>), then every chunk rescans itself from its base (k_scan_apply).
constexpr uint32_t SCAN_CHUNK = 4096;

__device__ __forceinline__ uint64_t cta_scan_256_u64(uint64_t x, uint64_t* s_warp, uint64_t& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = s_warp[i];
        if (i < wid) base += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return base + inc - x;
}

__global__ void __launch_bounds__(256) k_scan_partial(const uint32_t* __restrict__ in, uint64_t* __restrict__ partial,
                                                      uint32_t n, const uint32_t* n_dev) {
    __shared__ uint64_t s_warp[8];
    if (n_dev) n = *n_dev;
    const uint32_t i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 16;
    uint64_t sum = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) sum += i0 + j < n ? (uint64_t)in[i0 + j] : 0ull;
    uint64_t total;
    cta_scan_256_u64(sum, s_warp, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t* __restrict__ in, uint64_t* __restrict__ out,
                                                    const uint64_t* __restrict__ chunk_base, uint32_t n,
                                                    const uint32_t* n_dev) {
    __shared__ uint64_t s_warp[8];
    if (n_dev) n = *n_dev;
    const uint32_t i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 16;
    uint32_t v[16];
    uint64_t sum = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        v[j] = i0 + j < n ? in[i0 + j] : 0u;
        sum += v[j];
    }
    uint64_t total;
    uint64_t ex = chunk_base[blockIdx.x] + cta_scan_256_u64(sum, s_warp, total);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (i0 + j <= n) out[i0 + j] = ex;  // index n receives the grand total
        ex += v[j];
    }
}

// out[0..n] = exclusive scan of in[0..n); n is known on the host or read from *n_dev (then n_cap bounds it)
static int scan_u32(const uint32_t* in, uint64_t* out, uint32_t n_cap, const uint32_t* n_dev, uint64_t* tmp,
                    cudaStream_t s) {
    if (n_cap <= 4 * SCAN_CHUNK) {
        k_scan<uint32_t><<<1, 1024, 0, s>>>(in, out, n_dev ? 0u : n_cap, n_dev);
        return 1;
    }
    const uint32_t chunks = n_cap / SCAN_CHUNK + 1;  // covers index n itself
    k_scan_partial<<<chunks, 256, 0, s>>>(in, tmp, n_dev ? 0u : n_cap, n_dev);
    k_scan<uint64_t><<<1, 1024, 0, s>>>(tmp, tmp + chunks + 1, chunks, nullptr);
    k_scan_apply<<<chunks, 256, 0, s>>>(in, out, tmp + chunks + 1, n_dev ? 0u : n_cap, n_dev);
    return 3;
}

// bit offset (inside the whole batch) of block x; x may equal n_blocks
__device__ __forceinline__ uint64_t bit_prefix(const EntropyArgs& a, uint32_t x) {
    uint64_t g = a.w.tile_base[x >> 8];
    if (x & 255u) g += a.w.blk_prefix[x];
    return g;
}

__device__ __forceinline__ void interval_blocks(const EntropyArgs& a, uint32_t i, uint32_t& s, uint32_t& e) {
    const uint32_t bpm = (uint32_t)a.g.bpm, bpf = (uint32_t)a.g.n_mcu * bpm, ri = (uint32_t)a.g.ri;
    uint32_t f = i / (uint32_t)a.g.n_int, k = i - f * (uint32_t)a.g.n_int;
    s = f * bpf + k * ri * bpm;
    e = f * bpf + min((k + 1) * ri, (uint32_t)a.g.n_mcu) * bpm;
}

__global__ void k_intervals(const __grid_constant__ EntropyArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_int_total) return;
    uint32_t s, e;
    interval_blocks(a, i, s, e);
    uint64_t bits = bit_prefix(a, e) - bit_prefix(a, s);
    a.w.int_bits[i] = bits;
    uint64_t nb = (bits + 7) >> 3;
    uint64_t slot = (nb + 15) & ~15ull;
    if (slot == 0) slot = 16;
    a.w.int_slot[i] = (uint32_t)slot;
}

// Clears what the tiles placed with atomics are going to OR into (fast tiles store whole words and need nothing
// cleared): the listed tiles' ranges when they are few (the usual case: the tiles that straddle two frames), the
// whole used part of the buffer when they are many (short restart intervals: every tile holds several).
__global__ void __launch_bounds__(256) k_zero(const __grid_constant__ EntropyArgs a) {
    uint64_t total = a.w.int_ubase[a.n_int_total];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.w.status[1] = total;
        if (total > a.w.ubuf_cap) atomicOr((unsigned long long*)&a.w.status[0], JB_STATUS_UBUF_OVERFLOW);
    }
    if (total > a.w.ubuf_cap) return;
    const uint32_t n_slow = *a.w.any_slow, n_tiles = (a.n_blocks + TILE - 1) / TILE;
    if (n_slow == 0) return;
    if ((uint64_t)n_slow * 16 > n_tiles) {
        uint4* p = reinterpret_cast<uint4*>(a.w.ubuf);
        uint64_t n = total >> 4;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
            p[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    uint32_t* words = reinterpret_cast<uint32_t*>(a.w.ubuf);
    for (uint32_t i = blockIdx.x; i < n_slow; i += gridDim.x) {
        const PackPlan pl = a.w.pack_plan[a.w.slow_list[i]];
        const uint64_t w0 = pl.c >> 5, w1 = min(pl.slot_end_word, total >> 2);
        for (uint64_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) words[w] = 0u;
    }
}

// One thread per tile of 256 blocks: what k_pack's threads would otherwise each derive again (frame, MCU, restart
// interval of the block; start of the interval in the unstuffed buffer; two 64-bit bit prefixes), and what a fast
// tile has to know about its neighbours (PackPlan).
__global__ void __launch_bounds__(256) k_pack_plan(const __grid_constant__ EntropyArgs a) {
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x, n_tiles = (a.n_blocks + TILE - 1) / TILE;
    if (tile >= n_tiles) return;
    const uint32_t b_first = tile * TILE, b_last = min(b_first + TILE - 1, a.n_blocks - 1);
    const BlockInfo f = block_info(a, b_first), l = block_info(a, b_last);
    uint32_t s0, e0;
    interval_blocks(a, f.interval, s0, e0);
    PackPlan pl;
    pl.c = a.w.int_ubase[f.interval] * 8 + (a.w.tile_base[tile] - bit_prefix(a, s0));
    pl.slot_end_word = a.w.int_ubase[f.interval + 1] >> 2;
    pl.last_b = e0 - 1;
    pl.tile_bits = a.w.tile_bits[tile];
    pl.prev_tail = 0;
    uint32_t fl = 0;
    if (f.interval == l.interval) {
        fl |= PACK_UNIFORM;
        if (pl.tile_bits <= STREAM_MAX_BITS) fl |= PACK_FAST;
        if (b_first == s0) fl |= PACK_STARTS;
        if (b_last == e0 - 1) fl |= PACK_ENDS;
    }
    if (fl & PACK_FAST) {
        // the neighbours inside the same interval are uniform when they lie inside it entirely
        if (!(fl & PACK_STARTS) && b_first - s0 >= (uint32_t)TILE) {
            const uint32_t tb_prev = a.w.tile_bits[tile - 1], sh = (uint32_t)pl.c & 31u;
            if (tb_prev <= STREAM_MAX_BITS) {
                fl |= PACK_PREV_FAST;
                if (sh) {  // the last sh bits of the previous tile's stream, as the top bits of the shared word
                    const uint32_t* sp = reinterpret_cast<const uint32_t*>(a.w.slots + (size_t)(tile - 1) * TILE);
                    const uint32_t start = tb_prev - sh, wi = start >> 5, n_prev = (tb_prev + 31) >> 5;
                    const uint32_t w0 = sp[wi], w1 = wi + 1 < n_prev ? sp[wi + 1] : 0u;
                    pl.prev_tail = __funnelshift_l(w1, w0, start & 31u) & ~(0xFFFFFFFFu >> sh);
                }
            }
        }
        if (!(fl & PACK_ENDS) && min(b_last + TILE, a.n_blocks - 1) <= e0 - 1 && a.w.tile_bits[tile + 1] <= STREAM_MAX_BITS)
            fl |= PACK_NEXT_FAST;
    } else {
        // placed with atomics on cleared memory (k_zero, the next kernel): from the tile's first bit to the end of its
        // last block, or of that block's interval reservation when it ends the interval
        a.w.slow_list[atomicAdd(a.w.any_slow, 1u)] = tile;
        uint32_t s1, e1;
        interval_blocks(a, l.interval, s1, e1);
        const uint64_t end_bit = a.w.int_ubase[l.interval] * 8 + (bit_prefix(a, b_last + 1) - bit_prefix(a, s1));
        pl.slot_end_word = b_last == e1 - 1 ? a.w.int_ubase[l.interval + 1] >> 2 : (end_bit + 31) >> 5;
    }
    pl.flags = fl;
    a.w.pack_plan[tile] = pl;
}

// k_pack and k_stuff are bound by the latency of dependent loads (plan -> stream / prefix -> length -> slot; the
// compiler does not hoist loads over the branches between them, and volatile loads -- LDG.STRONG.SYS -- cost more
// than they save: 195 -> 315 us).  A CTA takes several tiles and prefetches the next one's lines into L2.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#ifndef PACK_UNROLL
#define PACK_UNROLL 2
#endif

// Places the tiles' codes in the unstuffed buffer (big-endian words); 1-padding after the last block of an interval.
// Blocks longer than a slot are left to k_pack_long.  ONE WARP PER TILE, several tiles per warp: the work on a tile is
// a short chain of dependent loads (plan -> stream -> stores), so what counts is how many tiles an SM has in flight
// (64 with a warp each, 8 with a CTA each: 165 us per 1.06 Gpx whatever the bytes moved), and a warp needs no barrier.
//   fast tile: output word W0 + k = (S[k-1] : S[k]) >> (c mod 32) of its stream S -- a funnel-shifted copy, a lane
//     per aligned group of four output words (five stream words, read through L1), stored as whole 128-bit words.
//     The word it shares with the tile before is completed with that tile's last bits (PackPlan::prev_tail), the word
//     it shares with the tile after is left to that tile; the tile that ends an interval adds the padding and clears
//     the rest of the interval's reservation.  No atomics, and nothing has to be cleared beforehand.
//   any other tile (several intervals in the tile, or more than 4 KB of codes): a lane per block shifts the block's
//     128-bit slot into place with global atomics on memory k_zero cleared; a fast tile next to one of these ORs its
//     share of the common word instead of storing it.
__global__ void __launch_bounds__(TILE) k_pack(const __grid_constant__ EntropyArgs a) {
    const uint32_t lane = threadIdx.x & 31u, n_tiles = (a.n_blocks + TILE - 1) / TILE;
    const uint32_t warp = blockIdx.x * (TILE / 32) + (threadIdx.x >> 5), n_warps = gridDim.x * (TILE / 32);
    // (128 bytes per lane, as far as the stream is expected to reach: the size of the tile just placed, plus a margin --
    // prefetching all 4 KB of every tile read 0.19 GB of unused slots per 1.06 Gpx)
    uint32_t expect_bytes = 4096;
    auto prefetch_tile = [&](uint32_t tile) {
        const uint4* sp = a.w.slots + (size_t)tile * TILE;
        if (128u * lane < expect_bytes) prefetch_l2(sp + 8 * lane);
        if (lane == 0) prefetch_l2(a.w.pack_plan + tile);
    };
    if (warp < n_tiles) prefetch_tile(warp);
    if (a.w.int_ubase[a.n_int_total] > a.w.ubuf_cap) return;
    uint32_t* words = reinterpret_cast<uint32_t*>(a.w.ubuf);
    for (uint32_t tile = warp; tile < n_tiles; tile += n_warps) {
        if (tile + n_warps < n_tiles) prefetch_tile(tile + n_warps);
        const PackPlan pl = a.w.pack_plan[tile];
        expect_bytes = (pl.flags & PACK_FAST) ? min(4096u, (pl.tile_bits >> 3) + 384u) : 4096u;
        if (pl.flags & PACK_FAST) {  // (uniform over the warp)
            const uint32_t tb = max(pl.tile_bits, 1u), sh = (uint32_t)pl.c & 31u;
            const uint32_t n_s4 = ((((tb + 31) >> 5) + 3) >> 2) << 2;  // stream words incl. the zero padding of its last 128-bit word
            const uint64_t W0 = pl.c >> 5;
            const uint32_t m = (uint32_t)(((pl.c + tb - 1) >> 5) - W0);  // the stream ends in output word W0 + m
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(a.w.slots + (size_t)tile * TILE);
            auto S = [&](long long k) { return k >= 0 && k < (long long)n_s4 ? __ldg(sp + k) : 0u; };
            const bool ends = (pl.flags & PACK_ENDS) != 0, starts = (pl.flags & PACK_STARTS) != 0;
            const bool shared_prev = !starts && sh != 0, shared_next = !ends && ((pl.c + tb) & 31u) != 0;
            uint32_t pad_bits = 0;
            if (ends && !a.fr.raw_bits) {  // pad the interval to a byte boundary with 1s (T.81 F.1.2.3)
                const uint32_t pp = (uint32_t)(pl.c + tb) & 31u;  // an interval starts on a byte boundary: pp mod 8 = its bits mod 8
                const uint32_t pad = (8u - (pp & 7u)) & 7u;
                if (pad) pad_bits = ((1u << pad) - 1u) << (32 - pp - pad);
            }
            // the tile writes the words W0 .. last (zeros from W0 + m + 1 on), as aligned groups of four
            const uint64_t last = ends ? max(W0 + m, pl.slot_end_word - 1) : W0 + m, vec0 = W0 >> 2;
            const uint32_t n_out = (uint32_t)(last - 4 * vec0) + 1;
            const long long k_first = (long long)(4 * vec0) - (long long)W0;  // stream index of the first word of group 0
            // PACK_UNROLL groups per lane at a time: all their stream words are requested before the first is used
            for (uint32_t vb = lane; 4 * vb < n_out; vb += 32 * PACK_UNROLL) {
                uint32_t sw[PACK_UNROLL][5];
                bool inside[PACK_UNROLL];
#pragma unroll
                for (int i = 0; i < PACK_UNROLL; ++i) {
                    const uint32_t v = vb + 32 * i;
                    const long long k0 = k_first + 4 * (long long)v;
                    // four words inside the stream: no neighbour, no padding
                    inside[i] = 4 * v < n_out && k0 >= 1 && k0 + 3 < (long long)m && k0 + 3 < (long long)n_s4;
                    if (inside[i]) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) sw[i][j] = __ldg(sp + (k0 - 1) + j);
                    }
                }
#pragma unroll
                for (int i = 0; i < PACK_UNROLL; ++i) {
                    const uint32_t v = vb + 32 * i;
                    if (4 * v >= n_out) break;
                    const long long k0 = k_first + 4 * (long long)v;
                    uint32_t* gw = words + 4 * (vec0 + v);
                    if (inside[i]) {
                        *reinterpret_cast<uint4*>(gw) = make_uint4(
                            __byte_perm(__funnelshift_r(sw[i][1], sw[i][0], sh), 0, 0x0123), __byte_perm(__funnelshift_r(sw[i][2], sw[i][1], sh), 0, 0x0123),
                            __byte_perm(__funnelshift_r(sw[i][3], sw[i][2], sh), 0, 0x0123), __byte_perm(__funnelshift_r(sw[i][4], sw[i][3], sh), 0, 0x0123));
                        continue;
                    }
                    uint32_t o[4];
                    uint32_t act = 0;  // per word, two bits: 0 leave, 1 store, 2 atomic OR
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const long long k = k0 + j;
                        o[j] = 0u;
                        if (k < 0 || k > (long long)(last - W0)) continue;
                        uint32_t a_j = 1u;
                        if (k <= (long long)m) {
                            o[j] = __funnelshift_r(S(k), S(k - 1), sh);
                            if (k == 0 && shared_prev) {
                                if (pl.flags & PACK_PREV_FAST)
                                    o[j] |= pl.prev_tail;
                                else
                                    a_j = 2u;
                            }
                            if (k == (long long)m) {
                                o[j] |= pad_bits;
                                if (shared_next) a_j = (pl.flags & PACK_NEXT_FAST) ? 0u : 2u;
                            }
                        }
                        o[j] = __byte_perm(o[j], 0, 0x0123);
                        act |= a_j << (2 * j);
                    }
                    if (act == 0x55u) {
                        *reinterpret_cast<uint4*>(gw) = make_uint4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t a_j = (act >> (2 * j)) & 3u;
                            if (a_j == 1u)
                                gw[j] = o[j];
                            else if (a_j == 2u && o[j])
                                atomicOr(gw + j, o[j]);
                        }
                    }
                }
            }
            continue;
        }
        for (uint32_t b = tile * TILE + lane; b < min((tile + 1) * TILE, a.n_blocks); b += 32) {
            const uint32_t len = a.w.blk_len[b];
            uint64_t pos;
            bool last_blk;
            if (pl.flags & PACK_UNIFORM) {
                pos = pl.c + a.w.blk_prefix[b];
                last_blk = b == pl.last_b;
            } else {
                const BlockInfo bi = block_info(a, b);
                uint32_t s0, e0;
                interval_blocks(a, bi.interval, s0, e0);
                pos = a.w.int_ubase[bi.interval] * 8 + (bit_prefix(a, b) - bit_prefix(a, s0));
                last_blk = bi.last_in_interval;
            }
            if (len <= 128) or_slot(words, (long long)pos, len, a.w.slots[b], false);
            if (last_blk && !a.fr.raw_bits) {
                const uint64_t pp = pos + len;
                const uint32_t pad = (8u - ((uint32_t)pp & 7u)) & 7u;
                if (pad) {
                    uint32_t o = ((1u << pad) - 1u) << (32 - ((uint32_t)pp & 31) - pad);
                    atomicOr(words + (pp >> 5), __byte_perm(o, 0, 0x0123));
                }
            }
        }
    }
}

// Blocks whose code is longer than a slot (rare at usual qualities): re-walk the
// coefficients and OR the codes straight into the unstuffed buffer.
__global__ void __launch_bounds__(TILE) k_pack_long(const __grid_constant__ EntropyArgs a) {
    __shared__ uint32_t s_ac[2][256], s_dc[2][16];
    load_tables(a, s_ac, s_dc);
    if (a.w.int_ubase[a.n_int_total] > a.w.ubuf_cap) return;
    const uint32_t n_long = *a.w.n_long;
    for (uint32_t i = blockIdx.x * TILE + threadIdx.x; i < n_long; i += gridDim.x * TILE) {
        const uint32_t b = a.w.long_list[i];
        BlockInfo bi = block_info(a, b);
        uint32_t s0, e0;
        interval_blocks(a, bi.interval, s0, e0);
        const uint64_t pos = a.w.int_ubase[bi.interval] * 8 + (bit_prefix(a, b) - bit_prefix(a, s0));
        const short* c = reinterpret_cast<const short*>(a.coef) + (size_t)b * 64;
        uint32_t m[2] = {0u, 0u};
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) nz_mask_piece(__ldg(reinterpret_cast<const uint4*>(c) + pc), pc, m);
        auto value = [&](int p) { return (int)__ldg(c + p); };
        const int pred = bi.has_prev ? (int)a.coef[(size_t)bi.prev * 64] : 0;
        const int tab = bi.comp ? 1 : 0;
        BitSink s;
        s.init(a.w.ubuf, pos);
        encode_sparse<false>(m[0], m[1], value, value(0) - pred, s_ac[tab], s_dc[tab], nullptr, a.always_eob != 0, s);
        s.finish();
    }
}

__device__ __forceinline__ uint32_t count_ff(uint4 q) {
    return (__popc(__vcmpeq4(q.x, 0xFFFFFFFFu)) + __popc(__vcmpeq4(q.y, 0xFFFFFFFFu)) +
            __popc(__vcmpeq4(q.z, 0xFFFFFFFFu)) + __popc(__vcmpeq4(q.w, 0xFFFFFFFFu))) >> 3;
}

// 0xFF bytes per 16-byte chunk of the unstuffed buffer, as exclusive prefixes inside tiles of 256 chunks (the unit
// k_stuff works on) plus the tile totals.  A CTA takes four tiles per step: a thread counts four consecutive chunks
// (64 contiguous bytes) and stores their four prefixes as one 128-bit word, a warp covers half a tile, and the halves
// meet through one shared-memory word per warp -- one barrier per 16 KB (the first version scanned one chunk per
// thread with two barriers per 4 KB: 63 us per 0.19 GB).
__global__ void __launch_bounds__(TILE) k_ff_count(const __grid_constant__ EntropyArgs a) {
    __shared__ uint32_t s_warp[2][8];
    uint64_t total = a.w.int_ubase[a.n_int_total];
    if (total > a.w.ubuf_cap) total = 0;
    const uint64_t n_chunks = total >> 4;
    const uint32_t n_tiles = (uint32_t)((n_chunks + TILE - 1) / TILE), n_super = (n_tiles + 3) / 4;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.w.n_ff_tiles = n_tiles;
    const uint4* p = reinterpret_cast<const uint4*>(a.w.ubuf);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t par = 0;
    for (uint32_t st = blockIdx.x; st < n_super; st += gridDim.x, par ^= 1) {
        const uint64_t c0 = (uint64_t)st * (4 * TILE) + 4 * threadIdx.x;
        uint32_t cnt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cnt[j] = c0 + j < n_chunks ? count_ff(p[c0 + j]) : 0u;
        const uint32_t mine = cnt[0] + cnt[1] + cnt[2] + cnt[3];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        if (lane == 31) s_warp[par][wid] = inc;
        __syncthreads();  // (the two parities make a second barrier unnecessary)
        const uint32_t tile = st * 4 + (wid >> 1);
        if (tile < n_tiles) {
            const uint32_t lower = s_warp[par][wid & ~1u];  // the tile's first half
            uint32_t ex = inc - mine + ((wid & 1) ? lower : 0u);
            uint4 o;
            o.x = ex;
            o.y = ex += cnt[0];
            o.z = ex += cnt[1];
            o.w = ex += cnt[2];
            *reinterpret_cast<uint4*>(a.w.ff_prefix + c0) = o;
            if ((wid & 1) && lane == 31) a.w.ff_tile[tile] = lower + inc;
        }
    }
}

__device__ __forceinline__ uint64_t ff_prefix(const EntropyArgs& a, uint64_t c) {
    uint64_t g = a.w.ff_tile_base[c >> 8];
    if (c & 255u) g += a.w.ff_prefix[c];
    return g;
}

// marker that follows interval k of a frame: 0 = none, else the second marker byte
__device__ __forceinline__ uint32_t marker_after(const EntropyArgs& a, uint32_t k) {
    if (k + 1 < (uint32_t)a.g.n_int || a.fr.final_rst) return 0xD0u + ((k + a.fr.rst_phase) & 7u);
    return a.fr.emit_eoi ? 0xD9u : 0u;
}

__global__ void k_int_out(const __grid_constant__ EntropyArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_int_total) return;
    uint32_t k = i % (uint32_t)a.g.n_int;
    uint64_t ff = ff_prefix(a, a.w.int_ubase[i + 1] >> 4) - ff_prefix(a, a.w.int_ubase[i] >> 4);
    uint64_t nb = (a.w.int_bits[i] + 7) >> 3;
    uint64_t sz = nb + ff + (marker_after(a, k) ? 2u : 0u) + (k == 0 ? a.fr.hdr_bytes : 0u);
    a.w.int_osize[i] = (uint32_t)sz;
}

__global__ void k_finalize(const __grid_constant__ EntropyArgs a) {
    uint64_t total = a.w.int_obase[a.n_int_total];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        a.w.status[2] = total;
        if (!a.out_off && total > a.out_cap) atomicOr((unsigned long long*)&a.w.status[0], JB_STATUS_OUT_OVERFLOW);
        if (a.total_out) *a.total_out = total;
    }
    if (i < (uint32_t)a.n_frames) {
        uint64_t o0 = a.w.int_obase[i * (uint32_t)a.g.n_int], o1 = a.w.int_obase[(i + 1) * (uint32_t)a.g.n_int];
        if (a.frame_off) a.frame_off[i] = o0;
        if (a.frame_size) a.frame_size[i] = o1 - o0;
    }
}

// Final placement.  One CTA per tile of 256 chunks (4 KB of unstuffed bytes).  The output
// stream is partitioned among chunks -- a chunk owns its data bytes with their stuffing
// zeros, the marker after its interval if it is the interval's last chunk, and the JFIF
// header if it is the first chunk of a frame -- so a tile owns one contiguous byte range
// [g0, g1) of the output.  The tile is assembled in shared memory (placed so that shared and
// global addresses agree modulo 16) and then stored with 128-bit coalesced stores; a tile
// whose range does not fit the window (pathological: thousands of tiny frames) writes bytes
// straight to global memory.
constexpr int STUFF_WIN = 12288;

// k_stuff's window is word-swizzled: 32-bit word i of the window lives at word i ^ ((i >> 5) & 3).  A warp's lanes write
// at a stride of 16 bytes (a chunk each), i.e. into words 4 apart -- four lanes per bank, every byte store four
// wavefronts (17.8 M bank conflicts per 0.19 GB); with the swizzle the lanes 8, 16 and 24 apart land in the other three
// words of the same 16-byte group.  An aligned 16-byte group stays one: the copy-out reads it whole and puts its words
// back in order.  (The window is 512-byte aligned, so the swizzle can be taken from the shared address itself.)
__device__ __forceinline__ uint32_t stuff_swz(uint32_t addr) { return addr ^ ((addr >> 5) & 0xCu); }
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(stuff_swz(addr)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(stuff_swz(addr)) : "memory");
    return v;
}
// the 16-byte group at the (16-byte aligned) shared address addr, words in logical order
__device__ __forceinline__ uint4 lds128_unswz(uint32_t addr) {
    uint4 q;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr) : "memory");
    const uint32_t k = (addr >> 7) & 3u;
    if (k & 1u) q = make_uint4(q.y, q.x, q.w, q.z);
    if (k & 2u) q = make_uint4(q.z, q.w, q.x, q.y);
    return q;
}

// interval i with int_ubase[i] <= pos < int_ubase[i+1], searched in [lo, hi)
__device__ __forceinline__ uint32_t find_interval(const EntropyArgs& a, uint64_t pos, uint32_t lo, uint32_t hi) {
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (a.w.int_ubase[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// One thread per 4 KB tile of the unstuffed buffer: everything k_stuff needs that is uniform over the tile.
__global__ void __launch_bounds__(256) k_stuff_plan(const __grid_constant__ EntropyArgs a) {
    const uint64_t total = a.w.int_ubase[a.n_int_total];
    if (total > a.w.ubuf_cap) return;
    const uint64_t n_chunks = total >> 4;
    const uint32_t n_tiles = (uint32_t)((n_chunks + TILE - 1) / TILE);
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    const uint64_t c = (uint64_t)tile * TILE, c_last = min(c + TILE - 1, n_chunks - 1);
    const uint32_t i = find_interval(a, c << 4, 0, a.n_int_total), k = i % (uint32_t)a.g.n_int;
    const uint64_t ub = a.w.int_ubase[i], off0 = (c << 4) - ub;
    StuffPlan pl;
    pl.i0 = i;
    pl.i1 = find_interval(a, c_last << 4, i, a.n_int_total);
    pl.k = k;
    pl.off0 = off0;
    pl.nb = (a.w.int_bits[i] + 7) >> 3;
    pl.hdr_first = k == 0 && off0 == 0 && a.fr.hdr_bytes != 0;
    pl.g0 = a.w.int_obase[i] + (k == 0 && off0 != 0 ? a.fr.hdr_bytes : 0u) + off0 +
            (a.w.ff_tile_base[tile] - ff_prefix(a, ub >> 4));  // chunk tile * 256 has in-tile prefix 0
    a.w.stuff_plan[tile] = pl;
}

__global__ void __launch_bounds__(TILE) k_stuff(const __grid_constant__ EntropyArgs a) {
    __shared__ __align__(512) uint8_t win[STUFF_WIN + 32];
    __shared__ uint64_t s_g0, s_g1;
    __shared__ uint32_t s_n;
    uint64_t total = a.w.int_ubase[a.n_int_total];
    // The segment may be placed at a byte offset that is only known on the device (a.out_off): a strip's place in
    // the stitched file follows from the lengths of the strips before it, and a.out may then be another GPU's
    // memory mapped over NVLink -- the stores below are the stitch.
    const uint64_t out_off = a.out_off ? *a.out_off : 0ull;
    if (total > a.w.ubuf_cap) return;
    if (out_off + a.w.int_obase[a.n_int_total] > a.out_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.w.status[2] = out_off + a.w.int_obase[a.n_int_total];
            atomicOr((unsigned long long*)&a.w.status[0], JB_STATUS_OUT_OVERFLOW);
        }
        return;
    }
    uint8_t* const out = a.out + out_off;
    const uint64_t n_chunks = total >> 4;
    const uint32_t n_tiles = (uint32_t)((n_chunks + TILE - 1) / TILE);
    const uint4* p = reinterpret_cast<const uint4*>(a.w.ubuf);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t c = (uint64_t)tile * TILE + threadIdx.x;
        const bool have = c < n_chunks;
        // what is uniform over the tile comes from k_stuff_plan: the restart intervals of the tile's first and last
        // chunk (they bracket every other chunk's interval search -- mostly to nothing: an interval is usually much
        // longer than a 4 KB tile), and for a tile that lies inside one interval its place in that interval
        const StuffPlan pl = a.w.stuff_plan[tile];
        const uint4 q_early = have ? p[c] : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t ffp_early = have ? a.w.ff_prefix[c] : 0u;
        {  // the next tile of this CTA: on its way while this one is assembled
            const uint64_t cn = c + (uint64_t)gridDim.x * TILE;
            if (cn < n_chunks) {
                if ((threadIdx.x & 1u) == 0) prefetch_l2(p + cn);
                if ((threadIdx.x & 7u) == 0) prefetch_l2(a.w.ff_prefix + cn);
                if (threadIdx.x == 0) prefetch_l2(a.w.stuff_plan + tile + gridDim.x);
            }
        }
        const uint32_t s_i0 = pl.i0, s_i1 = pl.i1, s_k = pl.k;
        const uint64_t s_off0 = pl.off0, s_nb = pl.nb;
        const bool s_hdr_first = pl.hdr_first != 0;
        if (s_i0 == s_i1 && !s_hdr_first) {
            // ---- fast path: the whole tile lies in one interval and starts no frame.  Per thread: its chunk and
            // its in-tile 0xFF prefix; everything else is tile-uniform, offsets are 32 bits relative to the tile.
            // (Measured and dropped: every warp assembling and storing its 32 chunks on its own, without the two CTA
            // barriers -- 170 -> 181 us; the kernel is bound by its byte-granular shared-memory stores.)
            const uint64_t g0 = pl.g0, off = s_off0 + 16u * threadIdx.x, nb = s_nb;
            uint32_t n_here = 0;
            if (s_off0 + 16u * TILE < s_nb && (uint64_t)(tile + 1) * TILE <= n_chunks) {
                // ---- and every chunk of it is 16 data bytes, no marker follows (a tile in the middle of an interval:
                // nearly all): per word one test for a 0xFF byte, four byte stores
                const uint32_t wds[4] = {q_early.x, q_early.y, q_early.z, q_early.w};
                const uint32_t rel = 16u * threadIdx.x + (threadIdx.x ? ffp_early : 0u);
                uint32_t d = (uint32_t)__cvta_generic_to_shared(win) + (uint32_t)((reinterpret_cast<uintptr_t>(out) + g0) & 15) + rel;
                const uint32_t d_begin = d;
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) {
                    const uint32_t w = wds[wq];
                    if ((((~w) - 0x01010101u) & w & 0x80808080u) == 0u) {  // no byte of w is 0xFF
                        sts8(d, w);
                        sts8(d + 1, w >> 8);
                        sts8(d + 2, w >> 16);
                        sts8(d + 3, w >> 24);
                        d += 4;
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const uint32_t byte = (w >> (jj * 8)) & 0xFFu;
                            sts8(d, byte);
                            if (byte == 0xFFu) sts8(d + 1, 0u);  // T.81 F.1.2.3 byte stuffing
                            d += byte == 0xFFu ? 2u : 1u;
                        }
                    }
                }
                if (threadIdx.x == TILE - 1) s_n = rel + (d - d_begin);
            } else if (have) {
                const uint4 q = q_early;
                const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
                const int valid = off >= nb ? 0 : (nb - off < 16 ? (int)(nb - off) : 16);
                const uint32_t marker = off + 16 >= nb ? marker_after(a, s_k) : 0u;  // last chunk of the interval
                const uint32_t rel = 16u * threadIdx.x + (threadIdx.x ? ffp_early : 0u);
                uint32_t d = (uint32_t)__cvta_generic_to_shared(win) + (uint32_t)((reinterpret_cast<uintptr_t>(out) + g0) & 15) + rel;
                const uint32_t d_begin = d;
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) {
                    const uint32_t w = wds[wq];
                    if (valid >= 4 * wq + 4 && __vcmpeq4(w, 0xFFFFFFFFu) == 0u) {  // four plain bytes: no tests
                        sts8(d, w);
                        sts8(d + 1, w >> 8);
                        sts8(d + 2, w >> 16);
                        sts8(d + 3, w >> 24);
                        d += 4;
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = 4 * wq + jj;
                            const uint32_t byte = (w >> (jj * 8)) & 0xFFu;
                            const bool on = j < valid, ff = on && byte == 0xFFu;
                            if (on) sts8(d, byte);
                            if (ff) sts8(d + 1, 0u);  // T.81 F.1.2.3 byte stuffing
                            d += on ? (ff ? 2u : 1u) : 0u;
                        }
                    }
                }
                if (marker) {
                    sts8(d, 0xFFu);
                    sts8(d + 1, marker);
                    d += 2;
                }
                n_here = rel + (d - d_begin);
                if (threadIdx.x == TILE - 1 || c + 1 == n_chunks) s_n = n_here;
            }
            __syncthreads();
            {
                const uint32_t base = (uint32_t)((reinterpret_cast<uintptr_t>(out) + g0) & 15), n = s_n;
                const uint32_t head = min(n, (16u - base) & 15u);          // bytes before the first 16-byte boundary
                const uint32_t n16 = (n - head) >> 4, tail = (n - head) & 15u;
                const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(win) + base;
                if (threadIdx.x < head) out[g0 + threadIdx.x] = (uint8_t)lds8(wbase + threadIdx.x);
                uint4* gdst = reinterpret_cast<uint4*>(out + g0 + head);
                for (uint32_t j = threadIdx.x; j < n16; j += TILE) gdst[j] = lds128_unswz(wbase + head + 16 * j);
                if (threadIdx.x < tail) out[g0 + head + 16 * (uint64_t)n16 + threadIdx.x] = (uint8_t)lds8(wbase + head + 16 * n16 + threadIdx.x);
            }
            __syncthreads();
            continue;
        }
        uint64_t dst = 0, start = 0, end = 0;
        uint32_t wds[4] = {0, 0, 0, 0}, marker = 0;
        int valid = 0;
        bool hdr = false;
        if (have) {
            uint64_t pos = c << 4;
            uint32_t i = find_interval(a, pos, s_i0, s_i1 + 1), k = i % (uint32_t)a.g.n_int;
            uint64_t ub = a.w.int_ubase[i];
            uint64_t off = pos - ub, nb = (a.w.int_bits[i] + 7) >> 3;
            hdr = k == 0 && off == 0 && a.fr.hdr_bytes != 0;
            start = a.w.int_obase[i] + (k == 0 && off != 0 ? a.fr.hdr_bytes : 0u) + off +
                    (ff_prefix(a, c) - ff_prefix(a, ub >> 4));
            dst = start + (hdr ? a.fr.hdr_bytes : 0u);
            const uint4 q = q_early;
            wds[0] = q.x; wds[1] = q.y; wds[2] = q.z; wds[3] = q.w;
            valid = off >= nb ? 0 : (nb - off < 16 ? (int)(nb - off) : 16);
            if (off + 16 >= nb) marker = marker_after(a, k);  // last chunk of the interval
            end = dst + (uint64_t)valid + count_ff(q) + (marker ? 2u : 0u);
        }
        if (threadIdx.x == 0) s_g0 = start;
        if (have && (threadIdx.x == TILE - 1 || c + 1 == n_chunks)) s_g1 = end;
        __syncthreads();
        const uint64_t g0 = s_g0, g1 = s_g1;
        const uintptr_t P0 = reinterpret_cast<uintptr_t>(out) + g0;
        const uint32_t base = (uint32_t)(P0 & 15);
        const bool fits = g1 - g0 <= STUFF_WIN;
        if (fits) {  // the usual case: 32-bit shared-memory addresses, predicated byte stores
            if (have) {
                uint32_t d = (uint32_t)__cvta_generic_to_shared(win) + base + (uint32_t)(start - g0);
                if (hdr) {
                    for (uint32_t j = 0; j < a.fr.hdr_bytes; ++j) sts8(d + j, a.hdr[j]);
                    d += a.fr.hdr_bytes;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t byte = (wds[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                    const bool on = j < valid, ff = on && byte == 0xFFu;
                    if (on) sts8(d, byte);
                    if (ff) sts8(d + 1, 0u);  // T.81 F.1.2.3 byte stuffing
                    d += on ? (ff ? 2u : 1u) : 0u;
                }
                if (marker) {
                    sts8(d, 0xFFu);
                    sts8(d + 1, marker);
                }
            }
        } else if (have) {  // pathological (thousands of tiny frames): bytes go straight to global memory
            uint8_t* o = out;
            if (hdr)
                for (uint32_t j = 0; j < a.fr.hdr_bytes; ++j) o[start + j] = a.hdr[j];
            for (int j = 0; j < valid; ++j) {
                uint32_t byte = (wds[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                o[dst++] = (uint8_t)byte;
                if (byte == 0xFFu) o[dst++] = 0;
            }
            if (marker) {
                o[dst] = 0xFF;
                o[dst + 1] = (uint8_t)marker;
            }
        }
        __syncthreads();
        if (fits) {
            const uint32_t n = (uint32_t)(g1 - g0);
            const uint32_t head = min(n, (16u - base) & 15u);          // bytes before the first 16-byte boundary
            const uint32_t n16 = (n - head) >> 4, tail = (n - head) & 15u;
            const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(win) + base;
            if (threadIdx.x < head) out[g0 + threadIdx.x] = (uint8_t)lds8(wbase + threadIdx.x);
            uint4* gdst = reinterpret_cast<uint4*>(out + g0 + head);
            for (uint32_t j = threadIdx.x; j < n16; j += TILE) gdst[j] = lds128_unswz(wbase + head + 16 * j);
            if (threadIdx.x < tail) out[g0 + head + 16 * (uint64_t)n16 + threadIdx.x] = (uint8_t)lds8(wbase + head + 16 * n16 + threadIdx.x);
        }
        __syncthreads();
    }
}

// Byte copy whose length and destination offset are only known on the device: pushes a rank's locally stitched strips
// to their place in another GPU's buffer (peer memory over NVLink) when the rank coded them in several calls.
// Destination-aligned 128-bit stores; the source is read as aligned words and realigned with funnel shifts.
__global__ void __launch_bounds__(256) k_copy_bytes(uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ d_dst_off,
                                                    const uint8_t* __restrict__ src, const uint64_t* __restrict__ d_len, uint64_t cap,
                                                    uint64_t* status) {
    const uint64_t off = d_dst_off ? *d_dst_off : 0ull, n = *d_len;
    if (off + n > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && status) {
            status[2] = off + n;
            atomicOr((unsigned long long*)&status[0], JB_STATUS_OUT_OVERFLOW);
        }
        return;
    }
    uint8_t* dst = dst_base + off;
    const uint64_t head = min(n, (uint64_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    const uint64_t n16 = (n - head) >> 4, tail = (n - head) & 15u;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
    if (tid < head) dst[tid] = src[tid];
    if (tid < tail) dst[head + 16 * n16 + tid] = src[head + 16 * n16 + tid];
    const uint8_t* s0 = src + head;                                  // source of the first aligned destination vector
    const uint32_t r = (uint32_t)(reinterpret_cast<uintptr_t>(s0) & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s0 - (r >> 3));
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (uint64_t j = tid; j < n16; j += nthreads) {
        const uint32_t* w = sw + 4 * j;
        // (the fifth word may lie past the last source byte when r == 0: it is then not needed, and not read)
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = r ? w[4] : 0u;
        d4[j] = make_uint4(__funnelshift_r(w0, w1, r), __funnelshift_r(w1, w2, r), __funnelshift_r(w2, w3, r),
                           __funnelshift_r(w3, w4, r));
    }
}

// ---- the strip lengths' all-gather and the completion signal as plain stores / polls on peer memory --------------
// One control block (64 x uint64, zero-initialised) lives on the stitching rank; every rank maps it over NVLink.
//   words  0..31  len[parity][rank]   byte count of the rank's strip in step `epoch` (parity = epoch & 1)
//   words 32..47  len_epoch[rank]     epoch of the newest published length
//   words 48..63  done_epoch[rank]    epoch up to which the rank's bytes have landed in the stitched file
// k_stitch_exchange (one warp): publish this rank's length, wait until every rank has published the step's length,
// prefix-sum them: off[0] = base + lengths of the ranks before this one, off[1] = base + all lengths.  A rank can be
// at most one step ahead of the slowest one (it needs everybody's length to pass), hence the two parities.
// Polling gives up after ~20 s (a peer died): the offsets become 2^62, which the placement kernel's capacity check
// rejects -- nothing is written and jb_sync reports the overflow.
constexpr int STITCH_MAX_RANKS = 16;
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool poll_at_least(const uint64_t* p, uint64_t epoch) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < epoch) {
        if (clock64() - t0 > 40000000000ll) return false;  // ~20 s at 1.9 GHz
        __nanosleep(200);
    }
    return true;
}

__global__ void __launch_bounds__(32) k_stitch_exchange(uint64_t* ctl, int rank, int world, uint64_t epoch, uint64_t base,
                                                        const uint64_t* __restrict__ d_len, uint64_t* __restrict__ d_off) {
    const int lane = threadIdx.x;
    uint64_t* len = ctl + (epoch & 1) * STITCH_MAX_RANKS;
    if (lane == 0) {
        len[rank] = *d_len;
        st_release_sys(ctl + 32 + rank, epoch);  // the length is visible before the epoch
    }
    bool ok = true;
    uint64_t mine = 0;
    if (lane < world) {
        ok = poll_at_least(ctl + 32 + lane, epoch);
        mine = ok ? len[lane] : 0ull;
    }
    ok = __all_sync(0xffffffffu, ok);
    uint64_t inc = mine;  // inclusive scan over the ranks
#pragma unroll
    for (int o = 1; o < STITCH_MAX_RANKS; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    const uint64_t before = __shfl_sync(0xffffffffu, inc - mine, rank), all = __shfl_sync(0xffffffffu, inc, world - 1);
    if (lane == 0) {
        d_off[0] = ok ? base + before : (1ull << 62);
        d_off[1] = ok ? base + all : (1ull << 62);
    }
}

// after the placement kernel (stream order: its stores have been performed): tell the stitching rank
__global__ void k_stitch_done(uint64_t* ctl, int rank, uint64_t epoch) {
    __threadfence_system();
    st_release_sys(ctl + 48 + rank, epoch);
}

// stitching rank: every other rank's bytes of this step have landed; *d_ok = 0 if a peer never reported
__global__ void __launch_bounds__(32) k_stitch_wait(const uint64_t* ctl, int rank, int world, uint64_t epoch, uint64_t* status) {
    const int lane = threadIdx.x;
    bool ok = true;
    if (lane < world && lane != rank) ok = poll_at_least(ctl + 48 + lane, epoch);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok && lane == 0 && status) atomicOr((unsigned long long*)&status[0], JB_STATUS_PEER_TIMEOUT);
}

int launch_stitch_exchange(uint64_t* ctl, int rank, int world, uint64_t epoch, uint64_t base, const uint64_t* d_len, uint64_t* d_off,
                           cudaStream_t s) {
    k_stitch_exchange<<<1, 32, 0, s>>>(ctl, rank, world, epoch, base, d_len, d_off);
    return 1;
}
int launch_stitch_complete(uint64_t* ctl, int rank, int world, int dst, uint64_t epoch, uint64_t* status, cudaStream_t s) {
    if (rank == dst)
        k_stitch_wait<<<1, 32, 0, s>>>(ctl, rank, world, epoch, status);
    else
        k_stitch_done<<<1, 1, 0, s>>>(ctl, rank, epoch);
    return 1;
}

int launch_copy_bytes(uint8_t* dst_base, const uint64_t* d_dst_off, const uint8_t* src, const uint64_t* d_len, uint64_t cap,
                      uint64_t* status, cudaStream_t s) {
    k_copy_bytes<<<148 * 8, 256, 0, s>>>(dst_base, d_dst_off, src, d_len, cap, status);
    return 1;
}

static uint64_t magic52(uint64_t d) { return ((1ull << 52) + d - 1) / d; }
static uint32_t magic32_floor(uint64_t d) { return d <= 1 ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / d); }

// k_encode with its tile staged by TMA when a tensor map can describe the coefficient array (16-byte aligned base;
// the library's own arena always is), otherwise with per-thread loads.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void launch_encode(const EntropyArgs& a, uint32_t n_tiles, cudaStream_t s) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    EncodeTiledFn encode = (EncodeTiledFn)tensor_map_encode_fn();
    if (encode && !a.no_tma && (reinterpret_cast<uintptr_t>(a.coef) & 15u) == 0) {
        const cuuint64_t dims[2] = {32, (cuuint64_t)a.n_blocks};  // uint32 elements: one block = one 128-byte row
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {32, TILE}, estr[2] = {1, 1};
        if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<int16_t*>(a.coef), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
            k_encode<true><<<n_tiles, TILE, 0, s>>>(a, tm);
            return;
        }
    }
    k_encode<false><<<n_tiles, TILE, 0, s>>>(a, tm);
}

// ---- symbol histogram for per-call optimal Huffman tables (JB_FLAG_OPTIMIZE_HUFFMAN) ------------------
// One thread per block, the same symbols k_encode would emit (utils.cpp:572-609, 667-694): DC difference
// category, (run, size) of every non-zero AC coefficient, ZRL, EOB.  hist = [4][256] in DHT order
// (DC luma, AC luma, DC chroma, AC chroma); counted in shared memory, flushed with one atomic per used bin.
__global__ void __launch_bounds__(TILE) k_symbol_hist(const __grid_constant__ EntropyArgs a, uint32_t* __restrict__ hist) {
    __shared__ uint4 s_coef[TILE * 8];  // the tile's coefficients, staged with coalesced loads as in k_encode
    __shared__ uint32_t s_h[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += TILE) s_h[i] = 0;
    const uint32_t n_tiles = (a.n_blocks + TILE - 1) / TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t t = threadIdx.x, b0 = tile * TILE, b = b0 + t;
        const uint4* src = reinterpret_cast<const uint4*>(a.coef) + (size_t)b0 * 8;
        const uint32_t n_blk = min((uint32_t)TILE, a.n_blocks - b0), n_here = n_blk * 8;
        __syncthreads();  // the previous tile has been read (and, the first time, the bins are zero)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t g = i * TILE + t, blk = g >> 3, pc = g & 7;
            if (g < n_here) s_coef[blk * 8 + (pc ^ (blk & 7))] = __ldg(src + g);
        }
        __syncthreads();
        if (t >= n_blk) continue;
        const BlockInfo bi = block_info(a, b);
        uint32_t* dc = s_h + (bi.comp ? 2 : 0) * 256;
        uint32_t* ac = dc + 256;
        const int pred = bi.has_prev ? (int)a.coef[(size_t)bi.prev * 64] : 0;
        int run = 0, last_nz = 0;
#pragma unroll 1
        for (int pc = 0; pc < 8; ++pc) {
            const uint4 q = s_coef[t * 8 + (pc ^ (t & 7))];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            if (pc && !(q.x | q.y | q.z | q.w)) {  // eight zeros
                run += 8;
                continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = pc * 8 + i;
                const int v = (int)(short)(w[i >> 1] >> ((i & 1) * 16));
                if (k == 0) {
                    const int d = v - pred;
                    atomicAdd(&dc[32 - __clz(abs(d))], 1u);
                } else if (v == 0) {
                    ++run;
                } else {
                    while (run >= 16) {  // ZRL
                        atomicAdd(&ac[0xF0], 1u);
                        run -= 16;
                    }
                    atomicAdd(&ac[(run << 4) | (32 - __clz(abs(v)))], 1u);
                    run = 0;
                    last_nz = k;
                }
            }
        }
        if (last_nz < 63 || a.always_eob) atomicAdd(&ac[0], 1u);  // EOB
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += TILE)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

int launch_symbol_hist(const EntropyArgs& a_in, uint32_t* d_hist, cudaStream_t s) {
    if (a_in.n_blocks == 0) return 0;
    EntropyArgs a = a_in;
    a.m_bpf = magic52((uint64_t)a.g.n_mcu * (uint64_t)a.g.bpm);
    a.m_ri = magic52((uint64_t)a.g.ri);
    uint32_t n_tiles = (a.n_blocks + TILE - 1) / TILE;
    k_symbol_hist<<<n_tiles < 148u * 8u ? n_tiles : 148u * 8u, TILE, 0, s>>>(a, d_hist);
    return 1;
}

// phase 0: the whole coder; 1: everything up to the sizes (k_finalize: the segment's length is known on the device,
// nothing has been written to a.out); 2: the final placement alone (k_stuff) -- a.out / a.out_cap / a.out_off may
// differ from phase 1: between the two the caller learns where the segment goes (multi-GPU strip stitch).
int launch_entropy(const EntropyArgs& a_in, cudaStream_t s, int phase) {
    if (a_in.n_blocks == 0) return 0;
    EntropyArgs a = a_in;
    a.m_bpf = magic52((uint64_t)a.g.n_mcu * (uint64_t)a.g.bpm);
    a.m_ri = magic52((uint64_t)a.g.ri);
    a.m32_ri = magic32_floor((uint64_t)a.g.ri);
    int launches = 0;
    uint32_t n_tiles = (a.n_blocks + TILE - 1) / TILE;
    uint32_t gi = (a.n_int_total + 255) / 256;
    const uint32_t chunk_tiles = (uint32_t)(a.w.ubuf_cap / 16 / TILE + 1);
    if (phase == 2) {
        k_stuff_plan<<<(chunk_tiles + 255) / 256, 256, 0, s>>>(a);
        k_stuff<<<chunk_tiles < 1184u ? chunk_tiles : 1184u, TILE, 0, s>>>(a);
        return 2;
    }
    launch_encode(a, n_tiles, s);
    launches += scan_u32(a.w.tile_bits, a.w.tile_base, n_tiles, nullptr, a.w.scan_tmp, s);
    k_intervals<<<gi, 256, 0, s>>>(a);
    launches += scan_u32(a.w.int_slot, a.w.int_ubase, a.n_int_total, nullptr, a.w.scan_tmp, s);
    // grids of the grid-stride kernels are capped by the work the plan allows: a small image launches few CTAs
    k_pack_plan<<<(n_tiles + 255) / 256, 256, 0, s>>>(a);
    k_zero<<<chunk_tiles < 592u ? chunk_tiles : 592u, 256, 0, s>>>(a);  // (sparse or full: decided on the device)
    {
        static int pack_ctas = 0;  // resident CTAs of k_pack per SM: one wave, every warp takes its share of the tiles
        if (!pack_ctas && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pack_ctas, k_pack, TILE, 0) != cudaSuccess || pack_ctas < 1))
            pack_ctas = 4;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const uint32_t cap = (uint32_t)sms * (uint32_t)pack_ctas;
        k_pack<<<(n_tiles + 7) / 8 < cap ? (n_tiles + 7) / 8 : cap, TILE, 0, s>>>(a);
    }
    k_pack_long<<<n_tiles < 1184u ? n_tiles : 1184u, TILE, 0, s>>>(a);  // (blocks past the list end leave at once)
    launches += 6;
    if (a.fr.raw_bits) return launches;
    k_ff_count<<<(chunk_tiles + 3) / 4 < 1184u ? (chunk_tiles + 3) / 4 : 1184u, TILE, 0, s>>>(a);
    launches += scan_u32(a.w.ff_tile, a.w.ff_tile_base, (uint32_t)(a.w.ubuf_cap / 16 / TILE + 1), a.w.n_ff_tiles,
                         a.w.scan_tmp, s);
    k_int_out<<<gi, 256, 0, s>>>(a);
    launches += scan_u32(a.w.int_osize, a.w.int_obase, a.n_int_total, nullptr, a.w.scan_tmp, s);
    uint32_t gf = ((uint32_t)a.n_frames + 255) / 256;
    k_finalize<<<gf, 256, 0, s>>>(a);
    launches += 3;
    if (phase == 1) return launches;
    k_stuff_plan<<<(chunk_tiles + 255) / 256, 256, 0, s>>>(a);
    k_stuff<<<chunk_tiles < 1184u ? chunk_tiles : 1184u, TILE, 0, s>>>(a);
    return launches + 2;
}

}  // namespace jb
