// C ABI of libjpegb200.so (include/jpegb200.h): context, device workspaces,
// CUDA streams with pinned staging (replacing the reference's OpenCL
// context/queue/buffer plumbing, src/OpenCLProject_JpegEncoder.cpp:257-315 and
// its per-stage blocking enqueueWrite/enqueueRead pairs, cpp:336-616), the
// staged per-function entry points and the fused/batched encode calls.
// There is no CPU fallback anywhere in this file: every entry point runs CUDA
// kernels or fails with JB_E_CUDA.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "jb_internal.h"

using namespace jb;

namespace {

constexpr int kSlots = 3;               // in-flight groups of the batched host path
constexpr int kDevStages = 4;           // header / status staging blocks of back-to-back asynchronous device calls
constexpr size_t kDevStageBytes = 2048; // 1 KB JFIF header + result words
constexpr size_t kUbufBytesPerBlock = 64;     // first unstuffed-buffer budget (typical use: 5-40 B/block)
constexpr size_t kUbufBytesPerBlockMax = 256;  // no block codes longer: 63 x (16 + 10) + 16 + 11 bits = 209 bytes

struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, used = 0;
};

struct EventPair {
    cudaEvent_t a, b;
    int kind;  // 0 transform, 1 fixup, 2 entropy, 3 h2d, 4 d2h, 5 edge MCUs
};

// One in-flight unit of work: stream + device workspace + pinned result words.
struct Slot {
    cudaStream_t st = nullptr;
    Arena arena;
    // carved from the arena by plan()
    uint8_t* d_rgb = nullptr;
    int16_t* d_coef = nullptr;
    uint32_t* d_tie_list = nullptr;
    uint32_t* d_scalars = nullptr;  // [0] tie_count, [1] n_ff_tiles, [2] n_long, [3] any_slow, status[4] (u64) at +16 bytes, [12] unit counter
    EntropyWork w{};
    uint8_t* d_out = nullptr;
    size_t d_out_cap = 0;
    uint64_t* d_frame_off = nullptr;
    uint64_t* d_frame_size = nullptr;
    uint64_t* d_total = nullptr;
    uint8_t* d_hdr = nullptr;
    uint8_t* d_tc = nullptr;  // tcgen05 transform: the four fp16 W matrices
    uint32_t* d_hist = nullptr;     // JB_FLAG_OPTIMIZE_HUFFMAN: symbol counts [4][256]
    HuffDev* d_huff_opt = nullptr;  // ... and the tables derived from them
    uint8_t* tc_loaded_at = nullptr;  // where, and which generation of, the matrices this slot last uploaded
    uint64_t tc_loaded_gen = 0;
    uint32_t tie_cap = 0;
    // small images: the entropy coder's ~17 launches replayed as one CUDA graph while the arguments repeat
    cudaGraphExec_t ent_graph = nullptr;
    EntropyArgs ent_key;
    int ent_launches = 0;
    // pinned host mirror: [0..3] status, [4] total, [5] tie_count, then frame_off[n], frame_size[n]
    uint64_t* h_res = nullptr;
    size_t h_res_cap = 0;
    cudaEvent_t ev_scalars = nullptr, ev_done = nullptr;
    bool busy = false;
    // Asynchronous device-resident calls (jb_encode_batch_device): the pinned header and status words of a call must
    // stay untouched until its copies have run, so every call takes the next of kDevStages staging blocks
    // (1 KB header + result words) and waits only for the call that used that block kDevStages calls ago.
    EntropyArgs strip_ea;        // jb_encode_strip_begin -> jb_encode_strip_finish
    bool strip_pending = false;
    uint8_t* h_dev = nullptr;
    cudaEvent_t dev_ev[kDevStages] = {};
    bool dev_used[kDevStages] = {};
    unsigned dev_seq = 0;
    // bookkeeping of the group in flight
    size_t first_frame = 0, n_frames = 0;
};

}  // namespace

struct jb_ctx {
    int device = 0;
    char err[512] = {0};
    Slot slot[kSlots];
    uint32_t* d_ydown = nullptr;
    double* d_costab = nullptr;
    double* d_scale = nullptr;
    HuffDev* d_huff[2] = {nullptr, nullptr};
    Arena scratch;  // staged entry points
    bool profiling = false;
    std::vector<EventPair> events;
    std::vector<cudaEvent_t> event_pool;
    jb_timings tm{};
    uint64_t required = 0;
    uint64_t pending_status_slot = 0;
    size_t ubuf_per_block = kUbufBytesPerBlock;  // grows (sticky) when the entropy workspace overflows
    bool ubuf_grew = false;                       // set by the overflow: synchronous entry points run again
    bool out_internal = false;  // the output buffer of the call in flight is the library's own staging (host-output paths)
    bool out_full = false;      // sticky: size that staging for the worst case (every byte stuffed) after it overflowed once
    // host-side constants derived from the quantisation tables, rebuilt only when the tables change
    struct TableCache {
        bool valid = false, tc_valid = false;
        uint64_t gen = 0;  // bumped whenever the W matrices are rebuilt
        uint32_t q[128];
        double tc_scale = 0;
        int tc_repl = -1;  // the chroma slots hold the K = 16 cell matrices of the replicated 4:2:0 mode
        int tc_inplace = -1;  // the matrices hold the reference's in-place transform (Q1)
        QuantConst qc;
        float tband[2][64];
        std::vector<uint8_t> tc;  // 32768 bytes: the four fp16 W matrices of the tcgen05 kernel
    } tables;
};

namespace {

int fail(jb_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, JB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int arena_reserve(jb_ctx* ctx, Arena& a, size_t bytes) {
    if (bytes <= a.cap) {
        a.used = 0;
        return JB_OK;
    }
    if (a.base) CK(cudaFree(a.base));
    a.base = nullptr;
    a.cap = 0;
    size_t want = bytes + bytes / 8 + (1u << 20);
    cudaError_t e = cudaMalloc(&a.base, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, JB_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    a.cap = want;
    a.used = 0;
    return JB_OK;
}

template <class T>
T* carve(Arena& a, size_t count) {
    size_t off = align_up(a.used, 256);
    a.used = off + count * sizeof(T);
    return reinterpret_cast<T*>(a.base + off);
}

// NV12-style input: d_rgb / pitch / frame_stride of enqueue_encode describe the Y plane, this the Cb,Cr plane
struct Nv12Src {
    const uint8_t* uv;
    size_t pitch_uv, frame_stride_uv;
};

struct Plan {
    Geometry g;
    size_t n_frames;
    size_t n_blocks, n_tiles, n_int_total;
    size_t rgb_bytes;      // 0 when the input is already on the device
    size_t d_pitch, d_frame_stride;
    size_t ubuf_cap, chunks_cap, out_cap;  // out_cap 0 when the output is a caller's device buffer
};

size_t plan_bytes(const Plan& p) {
    size_t b = 0;
    auto add = [&](size_t n) { b = align_up(b, 256) + n; };
    add(p.rgb_bytes);
    add(p.n_blocks * 128);
    add(p.n_blocks * 64 * 4);              // tie list: every coefficient may be flagged (never overflows)
    add(64);                               // scalars
    add((p.n_tiles * 256 + 1) * 4);        // blk_prefix
    add(p.n_tiles * 256 * 4);              // blk_len
    add(p.n_tiles * 256 * 16);             // slots
    add(p.n_tiles * 256 * 4);              // long_list
    add(((p.n_tiles + p.n_int_total + p.chunks_cap / 256) / 4096 + 8) * 2 * 8);  // scan_tmp
    add(p.n_tiles * 4);
    add((p.n_tiles + 1) * 8);
    add(p.n_int_total * 4);
    add(p.n_int_total * 8);
    add((p.n_int_total + 1) * 8);
    add(p.ubuf_cap);
    add((p.chunks_cap + 256) * 4);
    add((p.chunks_cap / 256 + 2) * 4);
    add((p.chunks_cap / 256 + 3) * 8);
    add((p.chunks_cap / 256 + 2) * sizeof(StuffPlan));
    add((p.n_tiles + 1) * sizeof(PackPlan));
    add((p.n_tiles + 1) * 4);
    add(p.n_int_total * 4);
    add((p.n_int_total + 1) * 8);
    add(p.out_cap);
    add(p.n_frames * 8);
    add(p.n_frames * 8);
    add(8);
    add(1024);
    add(32768);
    add(4096);
    add(sizeof(HuffDev));
    return align_up(b, 256) + 4096;
}

int slot_prepare(jb_ctx* ctx, Slot& s, const Plan& p) {
    const void* base_before = s.arena.base;
    const size_t cap_before = s.arena.cap;
    int rc = arena_reserve(ctx, s.arena, plan_bytes(p));
    if (rc) return rc;
    Arena& a = s.arena;
    // new allocation: nothing of the old arena survives -- also when cudaMalloc hands the old base address back
    // (a re-allocation always changes the capacity)
    if (a.base != base_before || a.cap != cap_before) s.tc_loaded_at = nullptr;
    s.d_tc = carve<uint8_t>(a, 32768);                    // first, so that it keeps its place while the plan varies
    s.d_rgb = carve<uint8_t>(a, p.rgb_bytes);
    s.d_coef = carve<int16_t>(a, p.n_blocks * 64);
    s.d_tie_list = carve<uint32_t>(a, p.n_blocks * 64);
    s.tie_cap = (uint32_t)(p.n_blocks * 64);
    s.d_scalars = carve<uint32_t>(a, 16);
    s.w.blk_prefix = carve<uint32_t>(a, p.n_tiles * 256 + 1);
    s.w.blk_len = carve<uint32_t>(a, p.n_tiles * 256);
    s.w.slots = carve<uint4>(a, p.n_tiles * 256);
    s.w.long_list = carve<uint32_t>(a, p.n_tiles * 256);
    s.w.scan_tmp = carve<uint64_t>(a, ((p.n_tiles + p.n_int_total + p.chunks_cap / 256) / 4096 + 8) * 2);
    s.w.tile_bits = carve<uint32_t>(a, p.n_tiles);
    s.w.tile_base = carve<uint64_t>(a, p.n_tiles + 1);
    s.w.int_slot = carve<uint32_t>(a, p.n_int_total);
    s.w.int_bits = carve<uint64_t>(a, p.n_int_total);
    s.w.int_ubase = carve<uint64_t>(a, p.n_int_total + 1);
    s.w.ubuf = carve<uint8_t>(a, p.ubuf_cap);
    s.w.ubuf_cap = p.ubuf_cap;
    s.w.ff_prefix = carve<uint32_t>(a, p.chunks_cap + 256);
    s.w.ff_tile = carve<uint32_t>(a, p.chunks_cap / 256 + 2);
    s.w.ff_tile_base = carve<uint64_t>(a, p.chunks_cap / 256 + 3);
    s.w.stuff_plan = carve<StuffPlan>(a, p.chunks_cap / 256 + 2);
    s.w.pack_plan = carve<PackPlan>(a, p.n_tiles + 1);
    s.w.slow_list = carve<uint32_t>(a, p.n_tiles + 1);
    s.w.int_osize = carve<uint32_t>(a, p.n_int_total);
    s.w.int_obase = carve<uint64_t>(a, p.n_int_total + 1);
    s.d_out = carve<uint8_t>(a, p.out_cap);
    s.d_out_cap = p.out_cap;
    s.d_frame_off = carve<uint64_t>(a, p.n_frames);
    s.d_frame_size = carve<uint64_t>(a, p.n_frames);
    s.d_total = carve<uint64_t>(a, 1);
    s.d_hdr = carve<uint8_t>(a, 1024);
    s.d_hist = carve<uint32_t>(a, 1024);
    s.d_huff_opt = carve<HuffDev>(a, 1);
    s.w.n_ff_tiles = s.d_scalars + 1;
    s.w.n_long = s.d_scalars + 2;
    s.w.any_slow = s.d_scalars + 3;
    s.w.status = reinterpret_cast<uint64_t*>(s.d_scalars + 4);
    // pinned result block: result words, then (last 1 KB) the staging area of the JFIF header
    // ... before it the W matrices (32 KB), the symbol counts (4 KB) and the optimised Huffman tables (8 KB)
    size_t need = (8 + 2 * p.n_frames) * sizeof(uint64_t) + 2048 + 32768 + 4096 + 8192;
    if (need > s.h_res_cap) {
        if (s.h_res) cudaFreeHost(s.h_res);
        s.h_res = nullptr;
        s.h_res_cap = 0;
        CK(cudaMallocHost(&s.h_res, need * 2));
        s.h_res_cap = need * 2;
    }
    return JB_OK;
}

int make_plan(jb_ctx* ctx, size_t n_frames, size_t W, size_t H, const jb_params* p, bool host_in, bool host_out,
              Plan* out) {
    if (!p || W == 0 || H == 0 || n_frames == 0) return fail(ctx, JB_E_INVALID, "empty image or null parameters");
    if (p->subsampling < 0 || p->subsampling > 2) return fail(ctx, JB_E_INVALID, "bad subsampling %d", p->subsampling);
    if (p->restart_interval < 0 || p->restart_interval > 65535)
        return fail(ctx, JB_E_INVALID, "restart interval %d out of range", p->restart_interval);
    for (int i = 0; i < 64; ++i)
        if (p->qlum[i] < 1 || p->qlum[i] > 255 || p->qchrom[i] < 1 || p->qchrom[i] > 255)
            return fail(ctx, JB_E_INVALID, "quantisation table entries must be 1..255");
    if ((W > 65535 || H > 65535) && !(p->flags & JB_FLAG_CLAMP_SOF))
        return fail(ctx, JB_E_UNSUPPORTED, "dimension > 65535 needs JB_FLAG_CLAMP_SOF (SOF0 sizes are 16 bit)");
    if (W >= (1u << 24) || H >= (1u << 24)) return fail(ctx, JB_E_UNSUPPORTED, "image too large");
    Plan pl;
    pl.g = make_geometry(W, H, p->subsampling, p->restart_interval);
    size_t m = (size_t)pl.g.mcu_px;
    if ((size_t)pl.g.mcux * m - W > W || (size_t)pl.g.mcuy * m - H > H)
        return fail(ctx, JB_E_UNSUPPORTED, "image smaller than the mirror padding it needs (utils.cpp:211-233)");
    pl.n_frames = n_frames;
    pl.n_blocks = n_frames * (size_t)pl.g.n_mcu * (size_t)pl.g.bpm;
    if (pl.n_blocks >= (1u << 26))
        return fail(ctx, JB_E_UNSUPPORTED, "more than 2^26 blocks in one call; split the batch");
    pl.n_tiles = (pl.n_blocks + 255) / 256;
    pl.n_int_total = n_frames * (size_t)pl.g.n_int;
    pl.d_pitch = align_up(W * 3, 16);
    pl.d_frame_stride = pl.d_pitch * H;
    pl.rgb_bytes = host_in ? pl.d_frame_stride * n_frames + 64 : 0;
    pl.ubuf_cap = align_up(pl.n_blocks * ctx->ubuf_per_block + pl.n_int_total * 16, 4096);
    pl.chunks_cap = pl.ubuf_cap / 16;
    pl.out_cap = host_out ? pl.ubuf_cap + (ctx->out_full ? pl.ubuf_cap : pl.ubuf_cap / 8) + n_frames * 1024 : 0;
    *out = pl;
    return JB_OK;
}

cudaEvent_t get_event(jb_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct Timed {
    jb_ctx* ctx;
    cudaStream_t st;
    EventPair ep;
    bool on;
    Timed(jb_ctx* c, cudaStream_t s, int kind) : ctx(c), st(s), on(c->profiling) {
        if (on) {
            ep.a = get_event(c);
            ep.b = get_event(c);
            ep.kind = kind;
            cudaEventRecord(ep.a, st);
        }
    }
    ~Timed() {
        if (on) {
            cudaEventRecord(ep.b, st);
            ctx->events.push_back(ep);
        }
    }
};

void resolve_events(jb_ctx* ctx) {
    for (auto& ep : ctx->events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) {
            double us = ms * 1000.0;
            switch (ep.kind) {
                case 0: ctx->tm.transform_us += us; break;
                case 1: ctx->tm.fixup_us += us; break;
                case 2: ctx->tm.entropy_us += us; break;
                case 3: ctx->tm.h2d_us += us; break;
                case 5: ctx->tm.edge_us += us; break;
                // staged entry points: one reference stage each (CPUTelemetry, utils.hpp:65-75)
                case 6: ctx->tm.CSCTime += us; break;
                case 7: ctx->tm.CDSTime += us; break;
                case 8: ctx->tm.levelShiftTime += us; break;
                case 9: ctx->tm.staged_dct_us += us; break;
                case 10: ctx->tm.QuantTime += us; break;
                case 11: ctx->tm.zigZagTime += us; break;
                case 12: ctx->tm.RLETime += us; break;
                case 13: ctx->tm.staged_copy_us += us; break;
                case 14: break;  // helpers outside the telemetry set
                case 15: ctx->tm.staged_huffman_us += us; break;
                default: ctx->tm.d2h_us += us; break;
            }
        }
        ctx->event_pool.push_back(ep.a);
        ctx->event_pool.push_back(ep.b);
    }
    ctx->events.clear();
    // CPUTelemetry view: the fused transform kernel is reported as DCT time, the whole entropy coder as Huffman time;
    // TotalCopyTime = the staged copy / padding / conversion kernels (the reference's meaning, cpp:93-139) + transfers
    ctx->tm.DCTTime = ctx->tm.transform_us + ctx->tm.staged_dct_us;
    ctx->tm.HuffmanTime = ctx->tm.entropy_us + ctx->tm.staged_huffman_us;
    ctx->tm.TotalCopyTime = ctx->tm.staged_copy_us + ctx->tm.h2d_us + ctx->tm.d2h_us;
}

// The entropy coder is ~17 small launches; for a small image (the reference's own use case) their launch
// overhead is the whole cost.  While a slot sees the same arguments again (same shape, same buffers: a stream of
// equal frames) the chain is replayed as one CUDA graph captured from the very same launcher.
constexpr uint32_t kGraphMaxBlocks = 1u << 16;  // ~2.8 Mpx of 4:2:0: beyond that the kernels dominate the launches

int launch_entropy_graphed(jb_ctx* ctx, Slot& s, const EntropyArgs& ea) {
    if (ea.n_blocks > kGraphMaxBlocks) return launch_entropy(ea, s.st);
    if (s.ent_graph && memcmp(&s.ent_key, &ea, sizeof(ea)) == 0) {
        if (cudaGraphLaunch(s.ent_graph, s.st) == cudaSuccess) return s.ent_launches;
        cudaGetLastError();
    }
    if (s.ent_graph) {
        cudaStreamSynchronize(s.st);  // an asynchronous caller may still be running the old graph
        cudaGraphExecDestroy(s.ent_graph);
        s.ent_graph = nullptr;
    }
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return launch_entropy(ea, s.st);
    }
    const int n = launch_entropy(ea, s.st);
    if (cudaStreamEndCapture(s.st, &g) != cudaSuccess || !g || cudaGraphInstantiate(&s.ent_graph, g, 0) != cudaSuccess) {
        cudaGetLastError();
        if (g) cudaGraphDestroy(g);
        s.ent_graph = nullptr;
        return launch_entropy(ea, s.st);  // nothing ran during the failed capture
    }
    cudaGraphDestroy(g);
    memcpy(&s.ent_key, &ea, sizeof(ea));
    s.ent_launches = n;
    if (cudaGraphLaunch(s.ent_graph, s.st) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphExecDestroy(s.ent_graph);
        s.ent_graph = nullptr;
        return launch_entropy(ea, s.st);
    }
    return n;
}

// Enqueue transform (+fix-up) + entropy coder for frames already in device memory.
int enqueue_encode(jb_ctx* ctx, Slot& s, const Plan& pl, const jb_params* p, const uint8_t* d_rgb, size_t pitch,
                   size_t frame_stride, const Framing& fr, size_t W, size_t H, uint8_t* d_out, size_t out_cap,
                   uint64_t* d_off, uint64_t* d_size, uint64_t* d_total, uint8_t* h_hdr_stage = nullptr,
                   EntropyArgs* two_phase = nullptr, const Nv12Src* nv = nullptr) {
    CK(cudaMemsetAsync(s.d_scalars, 0, 64, s.st));
    TransformArgs ta{};
    ta.rgb = d_rgb;
    ta.pitch = pitch;
    ta.frame_stride = frame_stride;
    ta.n_frames = (int)pl.n_frames;
    ta.g = pl.g;
    ta.coef = s.d_coef;
    ta.ydown = ctx->d_ydown;
    ta.tie_list = s.d_tie_list;
    ta.tie_count = s.d_scalars;
    ta.unit_counter = s.d_scalars + 12;
    ta.tie_cap = s.tie_cap;
    ta.inplace_dct = (p->flags & JB_FLAG_REF_INPLACE_DCT) ? 1 : 0;
    ta.use_tma = (p->flags & JB_FLAG_TMA) ? 1 : 0;
    if (nv) {
        if (pl.g.sub != JB_SUB_420 || (p->flags & JB_FLAG_REF_INPLACE_DCT))
            return fail(ctx, JB_E_UNSUPPORTED, "NV12-style input is 4:2:0 (JB_SUB_420) and takes the true DCT");
        ta.uv = nv->uv;
        ta.pitch_uv = nv->pitch_uv;
        ta.frame_stride_uv = nv->frame_stride_uv;
    }
    if (ta.inplace_dct) {
        // Q1 in the fused path is a different W matrix of the tcgen05 contraction: the CUDA-core kernels (AAN
        // factorisation) cannot express it, and the near-tie replay must be on
        const uintptr_t bits = (uintptr_t)d_rgb | (uintptr_t)pitch | (uintptr_t)frame_stride;
        if ((p->flags & (JB_FLAG_FMA_DCT | JB_FLAG_NO_TIE_FIXUP)) || (bits & 3))
            return fail(ctx, JB_E_UNSUPPORTED, "JB_FLAG_REF_INPLACE_DCT needs the tcgen05 transform (4-byte aligned input, no JB_FLAG_FMA_DCT) and the tie replay");
    }
    {
        jb_ctx::TableCache& tc = ctx->tables;
        if (!tc.valid || memcmp(tc.q, p->qlum, 256) || memcmp(tc.q + 64, p->qchrom, 256)) {
            memcpy(tc.q, p->qlum, 256);
            memcpy(tc.q + 64, p->qchrom, 256);
            build_quant_const(p->qlum, p->qchrom, &tc.qc);
            tc.valid = true;
            tc.tc_valid = false;
        }
        ta.qc = tc.qc;
        if (!(p->flags & JB_FLAG_FMA_DCT)) {  // tcgen05 transform, every subsampling mode and NV12-style input (launch falls back if unaligned)
#ifdef JB_DEBUG_KNOBS  // tests/tools/tc_band_scan.py only (a separate build): the shipped library reads no environment
            const char* e = getenv("JB_TC_STEP_ULPS");
            const double scale = e ? atof(e) : JB_TC_STEP_ULPS;
#else
            const double scale = JB_TC_STEP_ULPS;
#endif
            const int repl = pl.g.sub == JB_SUB_REPL420 ? 1 : 0, inplace = (p->flags & JB_FLAG_REF_INPLACE_DCT) ? 1 : 0;
            if (!tc.tc_valid || tc.tc_scale != scale || tc.tc_repl != repl || tc.tc_inplace != inplace) {
                for (Slot& o : ctx->slot)  // an earlier asynchronous call may still be copying the old matrices
                    if (o.st) CK(cudaStreamSynchronize(o.st));
                tc.tc.resize(32768);
                build_tc_matrices(p->qlum, p->qchrom, scale, repl, inplace, tc.tc.data(), tc.tband);
                tc.tc_scale = scale;
                tc.tc_repl = repl;
                tc.tc_inplace = inplace;
                tc.tc_valid = true;
                ++tc.gen;
            }
            memcpy(ta.tband, tc.tband, sizeof(ta.tband));
            if (s.tc_loaded_at != s.d_tc || s.tc_loaded_gen != tc.gen) {
                // the W matrices travel through the pinned result block (asynchronous copy on the slot's stream)
                uint8_t* h = reinterpret_cast<uint8_t*>(s.h_res) + s.h_res_cap - 1024 - 32768;
                memcpy(h, tc.tc.data(), 32768);
                CK(cudaMemcpyAsync(s.d_tc, h, 32768, cudaMemcpyHostToDevice, s.st));
                s.tc_loaded_at = s.d_tc;
                s.tc_loaded_gen = tc.gen;
            }
            ta.tc_mat = s.d_tc;
        }
    }
    if (p->flags & JB_FLAG_NO_TIE_FIXUP)
        for (int t = 0; t < 2; ++t)
            for (int i = 0; i < 64; ++i) ta.qc.band[t][i] = ta.tband[t][i] = 1.0f;  // never flag
    if (nv) {
        {
            Timed t(ctx, s.st, 0);
            int n = launch_transform_nv12(ta, s.st);
            ctx->tm.transform_launches += n;
            ctx->tm.total_launches += n;
        }
        {
            Timed t(ctx, s.st, 5);
            ctx->tm.total_launches += launch_transform_nv12_edge(ta, s.st);
        }
    } else {
        {
            Timed t(ctx, s.st, 0);
            int n = launch_transform(ta, s.st);
            ctx->tm.transform_launches += n;
            ctx->tm.total_launches += n;
        }
        {
            Timed t(ctx, s.st, 5);
            ctx->tm.total_launches += launch_transform_edge(ta, s.st);
        }
    }
    if (!(p->flags & JB_FLAG_NO_TIE_FIXUP)) {
        FixupArgs fa{};
        fa.rgb = d_rgb;
        fa.pitch = pitch;
        fa.frame_stride = frame_stride;
        fa.g = pl.g;
        fa.coef = s.d_coef;
        fa.ydown = ctx->d_ydown;
        fa.tie_list = s.d_tie_list;
        fa.tie_count = s.d_scalars;
        fa.tie_cap = s.tie_cap;
        fa.inplace_dct = ta.inplace_dct;
        fa.uv = ta.uv;
        fa.pitch_uv = ta.pitch_uv;
        fa.frame_stride_uv = ta.frame_stride_uv;
        fa.costab = ctx->d_costab;
        fa.scale = ctx->d_scale;
        memcpy(fa.qt.q[0], p->qlum, sizeof(fa.qt.q[0]));
        memcpy(fa.qt.q[1], p->qchrom, sizeof(fa.qt.q[1]));
        Timed t(ctx, s.st, 1);
        ctx->tm.total_launches += launch_fixup(fa, s.st);
    }
    if (!d_out) return JB_OK;  // transform only
    EntropyArgs ea;
    memset(&ea, 0, sizeof(ea));  // padding too: the graph cache compares the bytes
    ea.coef = s.d_coef;
    ea.g = pl.g;
    ea.n_frames = (int)pl.n_frames;
    ea.n_blocks = (uint32_t)pl.n_blocks;
    ea.n_int_total = (uint32_t)pl.n_int_total;
    ea.huff = ctx->d_huff[(p->flags & JB_FLAG_REF_TYPO_TABLES) ? 1 : 0];
    ea.always_eob = (p->flags & JB_FLAG_REF_ALWAYS_EOB) ? 1u : 0u;
    ea.no_tma = (p->flags & JB_FLAG_ENTROPY_LDG) ? 1u : 0u;
    ea.fr = fr;
    ea.w = s.w;
    ea.hdr = s.d_hdr;
    ea.out = d_out;
    ea.out_cap = out_cap;
    ea.frame_off = d_off;
    ea.frame_size = d_size;
    ea.total_out = d_total;
    HuffSpecs specs;
    const HuffSpecs* custom = nullptr;
    if (p->flags & JB_FLAG_OPTIMIZE_HUFFMAN) {
        // Two passes (like libjpeg's optimize_coding): count the symbols of this call's coefficients on the GPU,
        // derive the optimal tables on the host (T.81 K.2), code with them and write them into the DHT segments.
        // The counts make this a synchronisation point of the stream.
        if (!fr.hdr_bytes || (p->flags & JB_FLAG_REF_TYPO_TABLES))
            return fail(ctx, JB_E_UNSUPPORTED, "optimised Huffman tables need a JFIF header to travel in (no strips, no raw bit strings)");
        uint8_t* stage = reinterpret_cast<uint8_t*>(s.h_res) + s.h_res_cap - 1024 - 32768 - 4096;
        uint32_t* h_hist = reinterpret_cast<uint32_t*>(stage);
        HuffDev* h_huff = reinterpret_cast<HuffDev*>(stage - 8192);
        {
            Timed t(ctx, s.st, 2);
            CK(cudaMemsetAsync(s.d_hist, 0, 4096, s.st));
            ctx->tm.total_launches += launch_symbol_hist(ea, s.d_hist, s.st);
            CK(cudaMemcpyAsync(h_hist, s.d_hist, 4096, cudaMemcpyDeviceToHost, s.st));
        }
        CK(cudaStreamSynchronize(s.st));
        uint64_t counts[4][256];
        for (int t = 0; t < 4; ++t)
            for (int i = 0; i < 256; ++i) counts[t][i] = h_hist[t * 256 + i];
        optimal_huff_specs(counts, &specs);
        build_huff_from_specs(specs, h_huff);
        CK(cudaMemcpyAsync(s.d_huff_opt, h_huff, sizeof(HuffDev), cudaMemcpyHostToDevice, s.st));
        ea.huff = s.d_huff_opt;
        custom = &specs;
    }
    if (fr.hdr_bytes) {
        // the header travels through the pinned result block so that the copy is truly asynchronous
        uint8_t* h = h_hdr_stage ? h_hdr_stage : reinterpret_cast<uint8_t*>(s.h_res) + s.h_res_cap - 1024;
        size_t n = build_header(p, W, H, h, custom);
        if (custom)
            ea.fr.hdr_bytes = (uint32_t)n;  // fewer symbols than Annex K lists: a shorter header
        else if (n != fr.hdr_bytes)
            return fail(ctx, JB_E_INTERNAL, "header size mismatch");
        CK(cudaMemcpyAsync(s.d_hdr, h, n, cudaMemcpyHostToDevice, s.st));
    }
    if (two_phase) {  // strips placed later (jb_encode_strip_finish): stop once the segment's length is known
        ea.out_off = reinterpret_cast<const uint64_t*>(ea.total_out);  // any non-null pointer: k_finalize leaves the capacity check to k_stuff
        Timed t(ctx, s.st, 2);
        ctx->tm.total_launches += launch_entropy(ea, s.st, 1);
        *two_phase = ea;
    } else {
        Timed t(ctx, s.st, 2);
        ctx->tm.total_launches += launch_entropy_graphed(ctx, s, ea);
    }
    CK(cudaGetLastError());
    return JB_OK;
}

int status_to_rc(jb_ctx* ctx, const uint64_t* st, uint64_t tie_count, uint32_t tie_cap) {
    ctx->tm.tie_fixups = tie_count;
    if (tie_count > tie_cap) return fail(ctx, JB_E_INTERNAL, "near-tie list overflow (%llu > %u)", (unsigned long long)tie_count, tie_cap);
    if (st[0] & JB_STATUS_PEER_TIMEOUT) return fail(ctx, JB_E_INTERNAL, "strip stitch: a peer rank never reported its strip (timeout)");
    if (st[0] & JB_STATUS_UBUF_OVERFLOW) {
        // the library's own workspace, not the caller's buffer: enlarge the per-block budget (sticky) so that the
        // synchronous entry points can run the call again, and an asynchronous caller's next call succeeds
        ctx->required = st[1];
        if (ctx->ubuf_per_block < kUbufBytesPerBlockMax) {
            ctx->ubuf_per_block = std::min(kUbufBytesPerBlockMax, ctx->ubuf_per_block * 2);
            ctx->ubuf_grew = true;
        }
        return fail(ctx, JB_E_NOSPACE, "entropy workspace too small (need %llu bytes): it has been enlarged, run the call again",
                    (unsigned long long)st[1]);
    }
    if (st[0] & JB_STATUS_OUT_OVERFLOW) {
        ctx->required = st[2];
        if (ctx->out_internal && !ctx->out_full) {
            // the library's own staging buffer (sized for 1/8 of stuffing), not the caller's: size it for the worst case
            // from now on and let the synchronous entry points run the call again
            ctx->out_full = true;
            ctx->ubuf_grew = true;
            return fail(ctx, JB_E_NOSPACE, "output staging too small (need %llu bytes): it has been enlarged, run the call again",
                        (unsigned long long)st[2]);
        }
        return fail(ctx, JB_E_NOSPACE, "output buffer too small: need %llu bytes", (unsigned long long)st[2]);
    }
    return JB_OK;
}

int scratch(jb_ctx* ctx, size_t bytes) { return arena_reserve(ctx, ctx->scratch, bytes + 4096); }

// Run a synchronous entry point again while its only failure is the library's own entropy workspace (very
// high-entropy content at q100 can need more than the first budget per block; status_to_rc enlarges it).
template <class F>
int with_workspace_retry(jb_ctx* ctx, F call) {
    int rc = JB_OK;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (ctx) ctx->ubuf_grew = false;
        rc = call();
        if (rc != JB_E_NOSPACE || !ctx || !ctx->ubuf_grew) break;
    }
    return rc;
}

// ---- tiling: images beyond SOF0's 16-bit dimensions as a grid of independent JFIF files -------
// Tiles of one tile column are equally sized frames `tile_h * pitch` bytes apart, i.e. a batch; the bottom row of
// tiles (when H is not a multiple of tile_h) is a second, shorter batch of one frame per column.
template <class Encode>
static int for_each_tile_batch(jb_ctx* ctx, size_t W, size_t H, size_t tile_w, size_t tile_h, const jb_params* p, Encode encode) {
    if (!p) return fail(ctx, JB_E_INVALID, "null parameters");
    const size_t m = p->subsampling == JB_SUB_420 ? 16 : 8;
    if (W == 0 || H == 0 || tile_w == 0 || tile_h == 0 || tile_w % m || tile_h % m || tile_w > 65535 || tile_h > 65535)
        return fail(ctx, JB_E_INVALID, "tile sizes must be multiples of the MCU (%zu) and at most 65535", m);
    const size_t ntx = (W + tile_w - 1) / tile_w, nty = (H + tile_h - 1) / tile_h, full_y = H / tile_h;
    for (size_t tx = 0; tx < ntx; ++tx) {
        const size_t w = std::min(tile_w, W - tx * tile_w);
        int rc;
        if (full_y && (rc = encode(tx, 0, w, tile_h, full_y, ntx))) return rc;                             // tiles (tx, 0 .. full_y-1)
        if (nty > full_y && (rc = encode(tx, full_y, w, H - full_y * tile_h, (size_t)1, ntx))) return rc;  // bottom tile
    }
    return JB_OK;
}

}  // namespace

// =============================================================== context ====
extern "C" {

int jb_version(void) { return 100; }

int jb_create(int device, jb_ctx** out) {
    if (!out) return JB_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return JB_E_CUDA;
    if (device < 0 || device >= n) return JB_E_INVALID;
    jb_ctx* ctx = new (std::nothrow) jb_ctx();
    if (!ctx) return JB_E_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return JB_E_CUDA;
    }
    for (auto& s : ctx->slot) {
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_scalars, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming) != cudaSuccess) {
            jb_destroy(ctx);
            return JB_E_CUDA;
        }
    }
    {
        Slot& s0 = ctx->slot[0];
        bool ok = cudaMallocHost(&s0.h_dev, kDevStages * kDevStageBytes) == cudaSuccess;
        for (int k = 0; ok && k < kDevStages; ++k) ok = cudaEventCreateWithFlags(&s0.dev_ev[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            jb_destroy(ctx);
            return JB_E_CUDA;
        }
    }
    std::vector<uint32_t> ydown(2048);
    double costab[64], scale[64];
    build_ydown(ydown.data());
    build_dct_tables(costab, scale);
    HuffDev hd[2];
    build_huff(false, &hd[0]);
    build_huff(true, &hd[1]);
    bool ok = cudaMalloc(&ctx->d_ydown, 2048 * 4) == cudaSuccess && cudaMalloc(&ctx->d_costab, 64 * 8) == cudaSuccess &&
              cudaMalloc(&ctx->d_scale, 64 * 8) == cudaSuccess && cudaMalloc(&ctx->d_huff[0], sizeof(HuffDev)) == cudaSuccess &&
              cudaMalloc(&ctx->d_huff[1], sizeof(HuffDev)) == cudaSuccess;
    ok = ok && cudaMemcpy(ctx->d_ydown, ydown.data(), 2048 * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(ctx->d_costab, costab, sizeof(costab), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(ctx->d_scale, scale, sizeof(scale), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(ctx->d_huff[0], &hd[0], sizeof(HuffDev), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(ctx->d_huff[1], &hd[1], sizeof(HuffDev), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        jb_destroy(ctx);
        return JB_E_CUDA;
    }
    *out = ctx;
    return JB_OK;
}

void jb_destroy(jb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& s : ctx->slot) {
        if (s.ent_graph) cudaGraphExecDestroy(s.ent_graph);
        if (s.arena.base) cudaFree(s.arena.base);
        if (s.h_res) cudaFreeHost(s.h_res);
        if (s.h_dev) cudaFreeHost(s.h_dev);
        for (auto e : s.dev_ev)
            if (e) cudaEventDestroy(e);
        if (s.ev_scalars) cudaEventDestroy(s.ev_scalars);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.st) cudaStreamDestroy(s.st);
    }
    if (ctx->scratch.base) cudaFree(ctx->scratch.base);
    for (auto& ep : ctx->events) {
        cudaEventDestroy(ep.a);
        cudaEventDestroy(ep.b);
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    cudaFree(ctx->d_ydown);
    cudaFree(ctx->d_costab);
    cudaFree(ctx->d_scale);
    cudaFree(ctx->d_huff[0]);
    cudaFree(ctx->d_huff[1]);
    delete ctx;
}

const char* jb_last_error(const jb_ctx* ctx) { return ctx ? ctx->err : "no context"; }

void* jb_stream(jb_ctx* ctx) { return ctx ? (void*)ctx->slot[0].st : nullptr; }

int jb_sync(jb_ctx* ctx) {
    if (!ctx) return JB_E_INVALID;
    for (auto& s : ctx->slot) CK(cudaStreamSynchronize(s.st));
    resolve_events(ctx);
    // device-resident calls report their status here
    Slot& s = ctx->slot[0];
    int rc = JB_OK;
    if (s.busy) {
        s.busy = false;
        ctx->out_internal = false;  // device-resident calls write the caller's buffer
        for (int k = 0; k < kDevStages; ++k) {  // every call since the last jb_sync; the first failure is reported
            if (!s.dev_used[k]) continue;
            s.dev_used[k] = false;
            const uint64_t* res = reinterpret_cast<const uint64_t*>(s.h_dev + k * kDevStageBytes + 1024);
            const int r = status_to_rc(ctx, res, res[5], s.tie_cap);
            if (rc == JB_OK) rc = r;
        }
    }
    return rc;
}

int jb_set_profiling(jb_ctx* ctx, int on) {
    if (!ctx) return JB_E_INVALID;
    ctx->profiling = on != 0;
    return JB_OK;
}

int jb_get_timings(jb_ctx* ctx, jb_timings* t) {
    if (!ctx || !t) return JB_E_INVALID;
    *t = ctx->tm;
    return JB_OK;
}

int jb_reset_counters(jb_ctx* ctx) {
    if (!ctx) return JB_E_INVALID;
    ctx->tm = jb_timings{};
    return JB_OK;
}

size_t jb_required_bytes(jb_ctx* ctx) { return ctx ? (size_t)ctx->required : 0; }

int jb_host_alloc(void** p, size_t bytes) {
    if (!p) return JB_E_INVALID;
    return cudaMallocHost(p, bytes) == cudaSuccess ? JB_OK : JB_E_NOMEM;
}
int jb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? JB_OK : JB_E_CUDA; }
int jb_device_alloc(jb_ctx* ctx, void** p, size_t bytes) {
    if (!ctx || !p) return JB_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, JB_E_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return JB_OK;
}
int jb_device_free(jb_ctx* ctx, void* p) {
    CK(cudaFree(p));
    return JB_OK;
}
int jb_memcpy_h2d(jb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->slot[0].st));
    CK(cudaStreamSynchronize(ctx->slot[0].st));
    return JB_OK;
}
int jb_memcpy_d2h(jb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->slot[0].st));
    CK(cudaStreamSynchronize(ctx->slot[0].st));
    return JB_OK;
}

// ================================================================ staged ====
// Each call: host -> device, one kernel, device -> host, synchronous -- the
// reference's per-stage protocol (cpp:336-616) with the blocking OpenCL
// transfers replaced by stream-ordered copies.

#define STAGE_BEGIN(bytes)                           \
    if (!ctx) return JB_E_INVALID;                   \
    CK(cudaSetDevice(ctx->device));                  \
    {                                                \
        int rc_ = scratch(ctx, (bytes));             \
        if (rc_) return rc_;                         \
    }                                                \
    cudaStream_t st = ctx->slot[0].st;               \
    Arena& A = ctx->scratch;                         \
    (void)A;

#define STAGE_END()                   \
    CK(cudaGetLastError());           \
    CK(cudaStreamSynchronize(st));    \
    resolve_events(ctx);              \
    return JB_OK;

int jb_csc_rgb8_aos(jb_ctx* ctx, uint8_t* px, size_t W, size_t H) {
    size_t n = W * H;
    if (!px || n == 0) return fail(ctx, JB_E_INVALID, "empty image");
    STAGE_BEGIN(n * 3)
    uint8_t* d = carve<uint8_t>(A, n * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, px, n * 3, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 6); ctx->tm.total_launches += launch_csc(d, n, ctx->d_ydown, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(px, d, n * 3, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_cds_aos(jb_ctx* ctx, uint8_t* px, size_t W, size_t H) {
    size_t n = W * H;
    if (!px || n == 0) return fail(ctx, JB_E_INVALID, "empty image");
    STAGE_BEGIN(n * 3)
    uint8_t* d = carve<uint8_t>(A, n * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, px, n * 3, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 7); if (W >= 2 && H >= 2) ctx->tm.total_launches += launch_cds(d, W, H, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(px, d, n * 3, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_padded_size(size_t W, size_t H, size_t mult, size_t* nW, size_t* nH) {
    if (!nW || !nH || mult == 0) return JB_E_INVALID;
    *nW = (W + mult - 1) / mult * mult;
    *nH = (H + mult - 1) / mult * mult;
    return JB_OK;
}

int jb_pad_mirror_aos(jb_ctx* ctx, const uint8_t* src, size_t W, size_t H, uint8_t* dst, size_t nW, size_t nH) {
    if (!src || !dst || W == 0 || H == 0 || nW < W || nH < H) return fail(ctx, JB_E_INVALID, "bad sizes");
    if (nW - W > W || nH - H > H) return fail(ctx, JB_E_UNSUPPORTED, "padding larger than the image");
    STAGE_BEGIN(W * H * 3 + nW * nH * 3 + 512)
    uint8_t* ds = carve<uint8_t>(A, W * H * 3);
    uint8_t* dd = carve<uint8_t>(A, nW * nH * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(ds, src, W * H * 3, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_pad(ds, W, H, dd, nW, nH, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(dst, dd, nW * nH * 3, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_u8_to_f64(jb_ctx* ctx, const uint8_t* src, double* dst, size_t n) {
    if (!src || !dst || n == 0) return fail(ctx, JB_E_INVALID, "empty input");
    STAGE_BEGIN(n * 9 + 512)
    double* dd = carve<double>(A, n);
    uint8_t* ds = carve<uint8_t>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(ds, src, n, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_u8_to_f64(ds, dd, n, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(dst, dd, n * 8, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_levelshift_f64(jb_ctx* ctx, double* img, size_t n, double val) {
    if (!img || n == 0) return fail(ctx, JB_E_INVALID, "empty input");
    STAGE_BEGIN(n * 8)
    double* d = carve<double>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, img, n * 8, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 8); ctx->tm.total_launches += launch_sub_f64(d, n, val, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(img, d, n * 8, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_dct_f64(jb_ctx* ctx, double* img, size_t W, size_t H, uint32_t flags) {
    if (!img || W == 0 || H == 0 || (W & 7) || (H & 7)) return fail(ctx, JB_E_INVALID, "size must be a multiple of 8");
    size_t n = W * H * 3;
    STAGE_BEGIN(n * 8)
    double* d = carve<double>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, img, n * 8, cudaMemcpyHostToDevice, st)); }
    {
        Timed t_(ctx, st, 9);
        ctx->tm.total_launches += launch_dct_f64(d, W, H, (flags & JB_FLAG_REF_INPLACE_DCT) ? 1 : 0, ctx->d_costab, ctx->d_scale, st);
    }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(img, d, n * 8, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_quantize_f64(jb_ctx* ctx, double* img, size_t W, size_t H, const uint32_t ql[64], const uint32_t qc[64]) {
    if (!img || !ql || !qc || W == 0 || H == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    size_t n = W * H * 3;
    STAGE_BEGIN(n * 8)
    double* d = carve<double>(A, n);
    QuantTables qt;
    memcpy(qt.q[0], ql, sizeof(qt.q[0]));
    memcpy(qt.q[1], qc, sizeof(qt.q[1]));
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, img, n * 8, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 10); ctx->tm.total_launches += launch_quant_f64(d, W, H, qt, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(img, d, n * 8, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_blockify(jb_ctx* ctx, const double* img, size_t W, size_t H, int32_t* linear) {
    if (!img || !linear || W == 0 || H == 0 || (W & 7) || (H & 7)) return fail(ctx, JB_E_INVALID, "bad arguments");
    size_t n = W * H * 3;
    STAGE_BEGIN(n * 12 + 512)
    double* d = carve<double>(A, n);
    int32_t* o = carve<int32_t>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, img, n * 8, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 11); ctx->tm.total_launches += launch_blockify(d, W, H, o, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(linear, o, n * 4, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_zigzag(jb_ctx* ctx, const int32_t* linear, int32_t* zz, size_t rows) {
    if (!linear || !zz || rows == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    size_t n = rows * 64;
    STAGE_BEGIN(n * 8 + 512)
    int32_t* a = carve<int32_t>(A, n);
    int32_t* b = carve<int32_t>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(a, linear, n * 4, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 11); ctx->tm.total_launches += launch_zigzag(a, b, rows, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(zz, b, n * 4, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_f64_to_u8(jb_ctx* ctx, const double* src, uint8_t* dst, size_t n) {
    if (!src || !dst || n == 0) return fail(ctx, JB_E_INVALID, "empty input");
    STAGE_BEGIN(n * 9 + 512)
    double* ds = carve<double>(A, n);
    uint8_t* dd = carve<uint8_t>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(ds, src, n * 8, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_f64_to_u8(ds, dd, n, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(dst, dd, n, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_remove_red_aos(jb_ctx* ctx, uint8_t* px, size_t W, size_t H) {
    size_t n = W * H;
    if (!px || n == 0) return fail(ctx, JB_E_INVALID, "empty image");
    STAGE_BEGIN(n * 3)
    uint8_t* d = carve<uint8_t>(A, n * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, px, n * 3, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_remove_red(d, n, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(px, d, n * 3, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_value_categories(jb_ctx* ctx, const int16_t* v, size_t n, uint8_t* cat, uint16_t* bits) {
    if (!v || !cat || !bits || n == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    STAGE_BEGIN(n * 5 + 1024)
    int16_t* dv = carve<int16_t>(A, n);
    uint16_t* db = carve<uint16_t>(A, n);
    uint8_t* dc = carve<uint8_t>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(dv, v, n * 2, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 14); ctx->tm.total_launches += launch_value_categories(dv, n, dc, db, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(cat, dc, n, cudaMemcpyDeviceToHost, st)); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(bits, db, n * 2, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_pad_mirror_planar_u32(jb_ctx* ctx, const uint32_t* in, size_t W, size_t H, uint32_t* out, size_t nW, size_t nH) {
    if (!in || !out || W == 0 || H == 0 || nW < W || nH < H) return fail(ctx, JB_E_INVALID, "bad sizes");
    if (nW - W > W || nH - H > H) return fail(ctx, JB_E_UNSUPPORTED, "padding larger than the image");
    STAGE_BEGIN((W * H + nW * nH) * 12 + 512)
    uint32_t* a = carve<uint32_t>(A, W * H * 3);
    uint32_t* b = carve<uint32_t>(A, nW * nH * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(a, in, W * H * 12, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_pad_planar_u32(a, W, H, b, nW, nH, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(out, b, nW * nH * 12, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_blockify_planar_i32(jb_ctx* ctx, const int32_t* planar, size_t W, size_t H, int32_t* linear) {
    if (!planar || !linear || W == 0 || H == 0 || (W & 7) || (H & 7)) return fail(ctx, JB_E_INVALID, "bad arguments");
    size_t n = W * H * 3;
    STAGE_BEGIN(n * 8 + 512)
    int32_t* a = carve<int32_t>(A, n);
    int32_t* b = carve<int32_t>(A, n);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(a, planar, n * 4, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 11); ctx->tm.total_launches += launch_blockify_planar_i32(a, W, H, b, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(linear, b, n * 4, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

// ---- the image layout of the reference's OpenCL half: planar uint32 ----------
int jb_planar_u32_from_aos(jb_ctx* ctx, const uint8_t* px, size_t W, size_t H, uint32_t* planar) {
    size_t n = W * H;
    if (!px || !planar || n == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    STAGE_BEGIN(n * 15 + 512)
    uint8_t* d = carve<uint8_t>(A, n * 3);
    uint32_t* o = carve<uint32_t>(A, n * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d, px, n * 3, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_aos_to_planar_u32(d, n, o, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(planar, o, n * 12, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_planar_u32_interleave(jb_ctx* ctx, const uint32_t* planar, size_t W, size_t H, uint32_t* interleaved) {
    size_t n = W * H;
    if (!planar || !interleaved || n == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    STAGE_BEGIN(n * 24 + 512)
    uint32_t* a = carve<uint32_t>(A, n * 3);
    uint32_t* b = carve<uint32_t>(A, n * 3);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(a, planar, n * 12, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 13); ctx->tm.total_launches += launch_planar_u32_interleave(a, n, b, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(interleaved, b, n * 12, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

int jb_planar_u32_to_rgb8_device(jb_ctx* ctx, const uint32_t* d_planar, size_t W, size_t H, uint8_t* d_rgb, size_t pitch) {
    if (!ctx) return JB_E_INVALID;
    if (!d_planar || !d_rgb || W * H == 0 || pitch < W * 3) return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    ctx->tm.total_launches += launch_planar_u32_to_rgb8(d_planar, W, H, d_rgb, pitch, ctx->slot[0].st);
    CK(cudaGetLastError());
    return JB_OK;
}

int jb_encode_jfif_planar_u32(jb_ctx* ctx, const uint32_t* planar, size_t W, size_t H, const jb_params* p, uint8_t* out,
                              size_t cap, size_t* out_len) {
    size_t n = W * H;
    if (!planar || !p || !out || n == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    const size_t pitch = align_up(W * 3, 16);
    STAGE_BEGIN(n * 12 + pitch * H + cap + 4096)
    uint32_t* d_pl = carve<uint32_t>(A, n * 3);
    uint8_t* d_rgb = carve<uint8_t>(A, pitch * H);
    uint8_t* d_out = carve<uint8_t>(A, cap);
    uint64_t* d_meta = carve<uint64_t>(A, 4);  // offset, size, total
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(d_pl, planar, n * 12, cudaMemcpyHostToDevice, st)); }
    ctx->tm.total_launches += launch_planar_u32_to_rgb8(d_pl, W, H, d_rgb, pitch, st);
    int rc = with_workspace_retry(ctx, [&] {
        int r = jb_encode_batch_device(ctx, d_rgb, 1, W, H, pitch, pitch * H, p, d_out, cap, d_meta, d_meta + 1, d_meta + 2);
        return r == JB_OK ? jb_sync(ctx) : r;
    });
    if (rc != JB_OK) {
        if (out_len) *out_len = rc == JB_E_NOSPACE ? (size_t)jb_required_bytes(ctx) : 0;
        return rc;
    }
    uint64_t size = 0;
    CK(cudaMemcpy(&size, d_meta + 1, 8, cudaMemcpyDeviceToHost));
    if (size > cap) return fail(ctx, JB_E_INTERNAL, "frame larger than its buffer");
    CK(cudaMemcpy(out, d_out, size, cudaMemcpyDeviceToHost));
    if (out_len) *out_len = (size_t)size;
    return JB_OK;
}

int jb_rle(jb_ctx* ctx, const int32_t* zz, size_t rows, uint32_t flags, int32_t* pairs, uint32_t* counts) {
    if (!zz || !pairs || !counts || rows == 0) return fail(ctx, JB_E_INVALID, "bad arguments");
    STAGE_BEGIN(rows * (256 + 512 + 4) + 1024)
    int32_t* a = carve<int32_t>(A, rows * 64);
    int32_t* b = carve<int32_t>(A, rows * 128);
    uint32_t* c = carve<uint32_t>(A, rows);
    { Timed t_(ctx, st, 3); CK(cudaMemcpyAsync(a, zz, rows * 256, cudaMemcpyHostToDevice, st)); }
    { Timed t_(ctx, st, 12); ctx->tm.total_launches += launch_rle(a, rows, (flags & JB_FLAG_REF_ALWAYS_EOB) ? 1 : 0, b, c, st); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(pairs, b, rows * 512, cudaMemcpyDeviceToHost, st)); }
    { Timed t_(ctx, st, 4); CK(cudaMemcpyAsync(counts, c, rows * 4, cudaMemcpyDeviceToHost, st)); }
    STAGE_END()
}

// ================================================================= fused ====

size_t jb_num_mcus(size_t W, size_t H, int sub) { return (size_t)make_geometry(W, H, sub, 0).n_mcu; }
int jb_blocks_per_mcu(int sub) { return sub == JB_SUB_420 ? 6 : 3; }

size_t jb_header_bytes(const jb_params* p) {
    if (!p) return 0;
    uint8_t tmp[2048];
    return build_header(p, 8, 8, tmp);
}

int jb_write_header(const jb_params* p, size_t W, size_t H, uint8_t* out, size_t cap, size_t* out_len) {
    if (!p || !out || !out_len) return JB_E_INVALID;
    uint8_t tmp[2048];
    size_t n = build_header(p, W, H, tmp);
    *out_len = n;
    if (n > cap) return JB_E_NOSPACE;
    memcpy(out, tmp, n);
    return JB_OK;
}

static int upload_frames(jb_ctx* ctx, Slot& s, const Plan& pl, const uint8_t* rgb, size_t W, size_t H, size_t pitch,
                         size_t frame_stride) {
    Timed t(ctx, s.st, 3);
    if (pitch == pl.d_pitch && (frame_stride == pl.d_frame_stride || pl.n_frames == 1)) {
        CK(cudaMemcpyAsync(s.d_rgb, rgb, pl.d_frame_stride * pl.n_frames, cudaMemcpyHostToDevice, s.st));
    } else if (frame_stride == pitch * H) {
        CK(cudaMemcpy2DAsync(s.d_rgb, pl.d_pitch, rgb, pitch, W * 3, H * pl.n_frames, cudaMemcpyHostToDevice, s.st));
    } else {
        for (size_t f = 0; f < pl.n_frames; ++f)
            CK(cudaMemcpy2DAsync(s.d_rgb + f * pl.d_frame_stride, pl.d_pitch, rgb + f * frame_stride, pitch, W * 3, H,
                                 cudaMemcpyHostToDevice, s.st));
    }
    return JB_OK;
}

int jb_transform(jb_ctx* ctx, const uint8_t* rgb, size_t W, size_t H, size_t pitch, const jb_params* p, int16_t* coef) {
    if (!ctx || !rgb || !coef) return fail(ctx, JB_E_INVALID, "null argument");
    if (pitch < W * 3) return fail(ctx, JB_E_INVALID, "pitch smaller than a row");
    CK(cudaSetDevice(ctx->device));
    Plan pl;
    int rc = make_plan(ctx, 1, W, H, p, true, false, &pl);
    if (rc) return rc;
    Slot& s = ctx->slot[0];
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    if ((rc = upload_frames(ctx, s, pl, rgb, W, H, pitch, pitch * H))) return rc;
    Framing fr{};
    if ((rc = enqueue_encode(ctx, s, pl, p, s.d_rgb, pl.d_pitch, pl.d_frame_stride, fr, W, H, nullptr, 0, nullptr,
                             nullptr, nullptr)))
        return rc;
    CK(cudaMemcpyAsync(coef, s.d_coef, pl.n_blocks * 128, cudaMemcpyDeviceToHost, s.st));
    s.h_res[5] = 0;
    CK(cudaMemcpyAsync(s.h_res + 5, s.d_scalars, 4, cudaMemcpyDeviceToHost, s.st));
    CK(cudaStreamSynchronize(s.st));
    resolve_events(ctx);
    uint64_t zero[4] = {0, 0, 0, 0};
    return status_to_rc(ctx, zero, s.h_res[5], s.tie_cap);
}

// Entropy-code coefficients that are already in the slot's d_coef.
static int entropy_from_slot(jb_ctx* ctx, Slot& s, const Plan& pl, const jb_params* p, const Framing& fr, uint8_t* out,
                             size_t cap, size_t* out_len, uint64_t* nbits) {
    EntropyArgs ea{};
    ea.coef = s.d_coef;
    ea.g = pl.g;
    ea.n_frames = (int)pl.n_frames;
    ea.n_blocks = (uint32_t)pl.n_blocks;
    ea.n_int_total = (uint32_t)pl.n_int_total;
    ea.huff = ctx->d_huff[(p->flags & JB_FLAG_REF_TYPO_TABLES) ? 1 : 0];
    ea.always_eob = (p->flags & JB_FLAG_REF_ALWAYS_EOB) ? 1u : 0u;
    ea.no_tma = (p->flags & JB_FLAG_ENTROPY_LDG) ? 1u : 0u;
    ea.fr = fr;
    ea.w = s.w;
    ea.hdr = s.d_hdr;
    ea.out = s.d_out;
    ea.out_cap = s.d_out_cap;
    ea.total_out = s.d_total;
    {
        Timed t(ctx, s.st, 2);
        ctx->tm.total_launches += launch_entropy(ea, s.st);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s.h_res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(s.h_res + 4, s.d_total, 8, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(s.h_res + 6, s.w.int_bits, 8, cudaMemcpyDeviceToHost, s.st));
    CK(cudaStreamSynchronize(s.st));
    int rc = status_to_rc(ctx, s.h_res, 0, s.tie_cap);
    if (rc) return rc;
    size_t n;
    const uint8_t* src;
    if (fr.raw_bits) {
        if (nbits) *nbits = s.h_res[6];
        n = (size_t)((s.h_res[6] + 7) >> 3);
        src = s.w.ubuf;
    } else {
        n = (size_t)s.h_res[4];
        src = s.d_out;
    }
    if (out_len) *out_len = n;
    if (n > cap) {
        ctx->required = n;
        return fail(ctx, JB_E_NOSPACE, "output buffer too small: need %zu bytes", n);
    }
    CK(cudaMemcpyAsync(out, src, n, cudaMemcpyDeviceToHost, s.st));
    CK(cudaStreamSynchronize(s.st));
    resolve_events(ctx);
    return JB_OK;
}

static int entropy_once(jb_ctx* ctx, const int16_t* coef, size_t n_mcu, const jb_params* p, uint8_t* out, size_t cap,
                        size_t* out_len) {
    if (!ctx || !coef || !out || !p || n_mcu == 0) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    // geometry only matters through n_mcu / blocks per MCU / restart interval here
    size_t m = p->subsampling == JB_SUB_420 ? 16 : 8;
    Plan pl;
    int rc = make_plan(ctx, 1, m, m, p, false, true, &pl);
    if (rc) return rc;
    if (n_mcu * (size_t)pl.g.bpm >= (1u << 26)) return fail(ctx, JB_E_UNSUPPORTED, "too many blocks");
    {
        pl.g.n_mcu = (int)n_mcu;
        pl.g.ri = p->restart_interval > 0 ? p->restart_interval : (int)n_mcu;
        pl.g.n_int = (pl.g.n_mcu + pl.g.ri - 1) / pl.g.ri;
        pl.n_blocks = n_mcu * (size_t)pl.g.bpm;
        pl.n_tiles = (pl.n_blocks + 255) / 256;
        pl.n_int_total = (size_t)pl.g.n_int;
        pl.ubuf_cap = align_up(pl.n_blocks * ctx->ubuf_per_block + pl.n_int_total * 16, 4096);
        pl.chunks_cap = pl.ubuf_cap / 16;
        pl.out_cap = pl.ubuf_cap + (ctx->out_full ? pl.ubuf_cap : pl.ubuf_cap / 8) + 1024;
    }
    ctx->out_internal = true;
    Slot& s = ctx->slot[0];
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    CK(cudaMemsetAsync(s.d_scalars, 0, 64, s.st));
    CK(cudaMemcpyAsync(s.d_coef, coef, pl.n_blocks * 128, cudaMemcpyHostToDevice, s.st));
    Framing fr{};
    return entropy_from_slot(ctx, s, pl, p, fr, out, cap, out_len, nullptr);
}

int jb_entropy(jb_ctx* ctx, const int16_t* coef, size_t n_mcu, const jb_params* p, uint8_t* out, size_t cap,
               size_t* out_len) {
    return with_workspace_retry(ctx, [&] { return entropy_once(ctx, coef, n_mcu, p, out, cap, out_len); });
}

int jb_huffman(jb_ctx* ctx, const int32_t* zz, size_t rpc, uint32_t flags, uint8_t* bits, size_t cap, uint64_t* nbits) {
    if (!ctx || !zz || !bits || rpc == 0) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    jb_params p{};
    p.subsampling = JB_SUB_REPL420;
    p.flags = flags;
    for (int i = 0; i < 64; ++i) p.qlum[i] = p.qchrom[i] = 1;
    Plan pl;
    int rc = make_plan(ctx, 1, 8, 8, &p, false, true, &pl);
    if (rc) return rc;
    pl.g.n_mcu = (int)rpc;
    pl.g.ri = (int)rpc;
    pl.g.n_int = 1;
    pl.n_blocks = rpc * 3;
    if (pl.n_blocks >= (1u << 26)) return fail(ctx, JB_E_UNSUPPORTED, "too many blocks");
    pl.n_tiles = (pl.n_blocks + 255) / 256;
    pl.n_int_total = 1;
    pl.ubuf_cap = align_up(pl.n_blocks * 256 + 16, 4096);  // worst case: the bit string is the product here
    pl.chunks_cap = pl.ubuf_cap / 16;
    pl.out_cap = 4096;
    pl.rgb_bytes = rpc * 3 * 64 * 4;  // staging for the int32 planar input
    ctx->out_internal = true;
    Slot& s = ctx->slot[0];
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    CK(cudaMemsetAsync(s.d_scalars, 0, 64, s.st));
    int32_t* d_zz = reinterpret_cast<int32_t*>(s.d_rgb);
    CK(cudaMemcpyAsync(d_zz, zz, rpc * 3 * 64 * 4, cudaMemcpyHostToDevice, s.st));
    ctx->tm.total_launches += launch_planar_to_scan(d_zz, rpc, s.d_coef, s.st);
    Framing fr{};
    fr.raw_bits = 1;
    size_t n = 0;
    return entropy_from_slot(ctx, s, pl, &p, fr, bits, cap, &n, nbits);
}

// Harvest a finished group of the batched host path: read its totals, copy its bytes out.
static int harvest(jb_ctx* ctx, Slot& s, uint8_t* out, size_t cap, uint64_t* offsets, uint64_t* sizes,
                   uint64_t* running, bool* overflow) {
    CK(cudaEventSynchronize(s.ev_scalars));
    int rc = status_to_rc(ctx, s.h_res, s.h_res[5], s.tie_cap);
    if (rc) return rc;
    uint64_t total = s.h_res[4];
    if (*overflow || *running + total > cap) {
        // the caller's buffer is too small: keep going without copying, so that jb_required_bytes() is the size
        // of the whole batch and not of the groups up to here
        *overflow = true;
        *running += total;
        CK(cudaEventRecord(s.ev_done, s.st));
        s.busy = false;
        return JB_OK;
    }
    {
        Timed t(ctx, s.st, 4);
        CK(cudaMemcpyAsync(out + *running, s.d_out, total, cudaMemcpyDeviceToHost, s.st));
    }
    CK(cudaEventRecord(s.ev_done, s.st));
    for (size_t f = 0; f < s.n_frames; ++f) {
        if (offsets) offsets[s.first_frame + f] = *running + s.h_res[6 + f];
        if (sizes) sizes[s.first_frame + f] = s.h_res[6 + s.n_frames + f];
    }
    *running += total;
    s.busy = false;
    return JB_OK;
}

// Host planes of `rows` rows of `width` bytes each, n frames -> device planes: as few copies as the strides allow.
static int upload_planes(jb_ctx* ctx, cudaStream_t st, uint8_t* d, size_t d_pitch, size_t d_stride, const uint8_t* h, size_t h_pitch,
                         size_t h_stride, size_t width, size_t rows, size_t n) {
    if (h_pitch == d_pitch && (h_stride == d_stride || n == 1)) {
        CK(cudaMemcpyAsync(d, h, d_stride * (n - 1) + d_pitch * (rows - 1) + width, cudaMemcpyHostToDevice, st));
    } else if (h_stride == h_pitch * rows && d_stride == d_pitch * rows) {
        CK(cudaMemcpy2DAsync(d, d_pitch, h, h_pitch, width, rows * n, cudaMemcpyHostToDevice, st));
    } else {
        for (size_t f = 0; f < n; ++f)
            CK(cudaMemcpy2DAsync(d + f * d_stride, d_pitch, h + f * h_stride, h_pitch, width, rows, cudaMemcpyHostToDevice, st));
    }
    return JB_OK;
}

// `nvh` set: `rgb` / `pitch` / `frame_stride` describe the Y planes of NV12-style host frames and nvh the CbCr planes.
static int encode_batch_once(jb_ctx* ctx, const uint8_t* rgb, size_t n_frames, size_t W, size_t H, size_t pitch,
                             size_t frame_stride, const jb_params* p, uint8_t* out, size_t cap, uint64_t* offsets,
                             uint64_t* sizes, const Nv12Src* nvh = nullptr) {
    if (!ctx || !rgb || !out || (nvh && !nvh->uv)) return fail(ctx, JB_E_INVALID, "null argument");
    const size_t cw = 2 * ((W + 1) / 2), ch = (H + 1) / 2;  // bytes per row / rows of a CbCr plane
    if (nvh) {
        if (pitch < W || nvh->pitch_uv < cw || (n_frames > 1 && (frame_stride < pitch * H || nvh->frame_stride_uv < nvh->pitch_uv * ch)))
            return fail(ctx, JB_E_INVALID, "bad pitch/stride");
    } else if (pitch < W * 3 || (n_frames > 1 && frame_stride < pitch * H)) {
        return fail(ctx, JB_E_INVALID, "bad pitch/stride");
    }
    CK(cudaSetDevice(ctx->device));
    // group size: about 96 MB of RGB per group, at least one frame
    size_t frame_bytes = W * H * 3;
    size_t fpg = std::max<size_t>(1, (96u << 20) / std::max<size_t>(frame_bytes, 1));
    fpg = std::min(fpg, n_frames);
    Plan pl;
    int rc = make_plan(ctx, fpg, W, H, p, true, true, &pl);
    if (rc) return rc;
    Framing fr{};
    fr.hdr_bytes = (uint32_t)jb_header_bytes(p);
    fr.emit_eoi = 1;
    size_t n_groups = (n_frames + fpg - 1) / fpg;
    uint64_t running = 0;
    size_t harvested = 0;
    bool overflow = false;
    ctx->out_internal = true;
    for (auto& s : ctx->slot) s.busy = false;
    for (size_t gi = 0; gi < n_groups; ++gi) {
        Slot& s = ctx->slot[gi % kSlots];
        if (s.busy) {  // the slot still holds group gi - kSlots: drain it first (in order)
            if ((rc = harvest(ctx, s, out, cap, offsets, sizes, &running, &overflow))) goto drain;
            ++harvested;
        }
        CK(cudaEventSynchronize(s.ev_done));  // its previous D2H must have left the device buffer
        {
            size_t f0 = gi * fpg, nf = std::min(fpg, n_frames - f0);
            Plan gp = pl;
            if (nf != fpg && (rc = make_plan(ctx, nf, W, H, p, true, true, &gp))) goto drain;
            if (s.arena.cap == 0 && (rc = slot_prepare(ctx, s, pl))) goto drain;  // size every slot for a full group
            if ((rc = slot_prepare(ctx, s, gp))) goto drain;
            s.first_frame = f0;
            s.n_frames = nf;
            if (nvh) {
                // The planes of a group take the place of its RGB frames in the staging buffer (half the bytes): every Y
                // plane, then every CbCr plane, with 16-byte aligned pitches (what k_transform_tc_nv12 wants) unless the
                // frames are so narrow that the padding would not fit -- then tight (W H + cw ch <= 3 W H always holds).
                size_t py = align_up(W, 16), puv = align_up(cw, 16);
                if (py * H + puv * ch > gp.d_frame_stride) py = W, puv = cw;
                uint8_t* d_uv = s.d_rgb + py * H * nf;
                {
                    Timed t(ctx, s.st, 3);
                    if ((rc = upload_planes(ctx, s.st, s.d_rgb, py, py * H, rgb + f0 * frame_stride, pitch, frame_stride, W, H, nf))) goto drain;
                    if ((rc = upload_planes(ctx, s.st, d_uv, puv, puv * ch, nvh->uv + f0 * nvh->frame_stride_uv, nvh->pitch_uv,
                                            nvh->frame_stride_uv, cw, ch, nf)))
                        goto drain;
                }
                const Nv12Src nv{d_uv, puv, puv * ch};
                if ((rc = enqueue_encode(ctx, s, gp, p, s.d_rgb, py, py * H, fr, W, H, s.d_out, s.d_out_cap, s.d_frame_off,
                                         s.d_frame_size, s.d_total, nullptr, nullptr, &nv)))
                    goto drain;
            } else {
                if ((rc = upload_frames(ctx, s, gp, rgb + f0 * frame_stride, W, H, pitch, frame_stride))) goto drain;
                if ((rc = enqueue_encode(ctx, s, gp, p, s.d_rgb, gp.d_pitch, gp.d_frame_stride, fr, W, H, s.d_out,
                                         s.d_out_cap, s.d_frame_off, s.d_frame_size, s.d_total)))
                    goto drain;
            }
            CK(cudaMemcpyAsync(s.h_res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
            CK(cudaMemcpyAsync(s.h_res + 4, s.d_total, 8, cudaMemcpyDeviceToHost, s.st));
            s.h_res[5] = 0;
            CK(cudaMemcpyAsync(s.h_res + 5, s.d_scalars, 4, cudaMemcpyDeviceToHost, s.st));
            CK(cudaMemcpyAsync(s.h_res + 6, s.d_frame_off, nf * 8, cudaMemcpyDeviceToHost, s.st));
            CK(cudaMemcpyAsync(s.h_res + 6 + nf, s.d_frame_size, nf * 8, cudaMemcpyDeviceToHost, s.st));
            CK(cudaEventRecord(s.ev_scalars, s.st));
            s.busy = true;
        }
    }
    // drain the remaining groups in submission order
    for (size_t gi = harvested; gi < n_groups; ++gi) {
        Slot& s = ctx->slot[gi % kSlots];
        if (s.busy && (rc = harvest(ctx, s, out, cap, offsets, sizes, &running, &overflow))) goto drain;
    }
    rc = JB_OK;
    if (overflow) {
        ctx->required = running;
        rc = fail(ctx, JB_E_NOSPACE, "output buffer too small: the batch needs %llu bytes", (unsigned long long)running);
    }
drain:
    for (auto& s : ctx->slot) {
        cudaStreamSynchronize(s.st);
        s.busy = false;
    }
    resolve_events(ctx);
    return rc;
}

int jb_encode_batch(jb_ctx* ctx, const uint8_t* rgb, size_t n_frames, size_t W, size_t H, size_t pitch,
                    size_t frame_stride, const jb_params* p, uint8_t* out, size_t cap, uint64_t* offsets,
                    uint64_t* sizes) {
    return with_workspace_retry(ctx, [&] { return encode_batch_once(ctx, rgb, n_frames, W, H, pitch, frame_stride, p, out, cap, offsets, sizes); });
}

int jb_encode_nv12_batch(jb_ctx* ctx, const uint8_t* y, size_t pitch_y, size_t frame_stride_y, const uint8_t* uv, size_t pitch_uv,
                         size_t frame_stride_uv, size_t n_frames, size_t W, size_t H, const jb_params* p, uint8_t* out, size_t cap,
                         uint64_t* offsets, uint64_t* sizes) {
    const Nv12Src nvh{uv, pitch_uv, frame_stride_uv};
    return with_workspace_retry(ctx, [&] { return encode_batch_once(ctx, y, n_frames, W, H, pitch_y, frame_stride_y, p, out, cap, offsets, sizes, &nvh); });
}

int jb_encode_jfif(jb_ctx* ctx, const uint8_t* rgb, size_t W, size_t H, size_t pitch, const jb_params* p, uint8_t* out,
                   size_t cap, size_t* out_len) {
    uint64_t off = 0, size = 0;
    int rc = jb_encode_batch(ctx, rgb, 1, W, H, pitch, pitch * H, p, out, cap, &off, &size);
    if (out_len) *out_len = rc == JB_E_NOSPACE ? (size_t)jb_required_bytes(ctx) : (size_t)size;
    return rc;
}

int jb_encode_batch_device(jb_ctx* ctx, const uint8_t* d_rgb, size_t n_frames, size_t W, size_t H, size_t pitch,
                           size_t frame_stride, const jb_params* p, uint8_t* d_out, size_t cap, uint64_t* d_offsets,
                           uint64_t* d_sizes, uint64_t* d_total) {
    if (!ctx || !d_rgb || !d_out) return fail(ctx, JB_E_INVALID, "null argument");
    if (pitch < W * 3 || (n_frames > 1 && frame_stride < pitch * H)) return fail(ctx, JB_E_INVALID, "bad pitch/stride");
    CK(cudaSetDevice(ctx->device));
    Plan pl;
    int rc = make_plan(ctx, n_frames, W, H, p, false, false, &pl);
    if (rc) return rc;
    Slot& s = ctx->slot[0];
    if (s.arena.cap < plan_bytes(pl)) CK(cudaStreamSynchronize(s.st));  // regrowing frees the arena
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    Framing fr{};
    fr.hdr_bytes = (uint32_t)jb_header_bytes(p);
    fr.emit_eoi = 1;
    // this call's pinned staging block; the call that used it kDevStages calls ago must have finished its copies
    const unsigned k = s.dev_seq++ % kDevStages;
    if (s.dev_used[k]) CK(cudaEventSynchronize(s.dev_ev[k]));
    uint8_t* stage = s.h_dev + k * kDevStageBytes;
    uint64_t* res = reinterpret_cast<uint64_t*>(stage + 1024);
    if ((rc = enqueue_encode(ctx, s, pl, p, d_rgb, pitch, frame_stride, fr, W, H, d_out, cap, d_offsets, d_sizes,
                             d_total ? d_total : s.d_total, stage)))
        return rc;
    memset(res, 0, 64);
    CK(cudaMemcpyAsync(res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(res + 5, s.d_scalars, 4, cudaMemcpyDeviceToHost, s.st));
    CK(cudaEventRecord(s.dev_ev[k], s.st));
    s.dev_used[k] = true;
    s.busy = true;  // the status words are examined by jb_sync
    return JB_OK;
}

int jb_encode_tiles(jb_ctx* ctx, const uint8_t* rgb, size_t W, size_t H, size_t pitch, size_t tile_w, size_t tile_h,
                               const jb_params* p, uint8_t* out, size_t cap, uint64_t* offsets, uint64_t* sizes, size_t* n_tiles) {
    if (!ctx || !rgb || !out || !offsets || !sizes) return fail(ctx, JB_E_INVALID, "null argument");
    if (pitch < W * 3) return fail(ctx, JB_E_INVALID, "pitch smaller than a row");
    size_t used = 0, needed = 0;
    bool overflow = false;
    std::vector<uint64_t> o, z;
    int rc = for_each_tile_batch(ctx, W, H, tile_w, tile_h, p, [&](size_t tx, size_t ty0, size_t w, size_t h, size_t count, size_t ntx) {
        o.assign(count, 0);
        z.assign(count, 0);
        const uint8_t* base = rgb + ty0 * tile_h * pitch + tx * tile_w * 3;
        int r = jb_encode_batch(ctx, base, count, w, h, pitch, tile_h * pitch, p, out + used, overflow ? 0 : cap - used, o.data(), z.data());
        if (r == JB_E_NOSPACE) {  // keep going so that jb_required_bytes() covers the whole grid
            overflow = true;
            needed += jb_required_bytes(ctx);
            return (int)JB_OK;
        }
        if (r) return r;
        for (size_t k = 0; k < count; ++k) {
            offsets[(ty0 + k) * ntx + tx] = used + o[k];
            sizes[(ty0 + k) * ntx + tx] = z[k];
        }
        const size_t bytes = count ? (size_t)(o[count - 1] + z[count - 1]) : 0;
        used += bytes;
        needed += bytes;
        return (int)JB_OK;
    });
    if (rc) return rc;
    if (n_tiles) *n_tiles = ((W + tile_w - 1) / tile_w) * ((H + tile_h - 1) / tile_h);
    if (overflow) {
        ctx->required = needed;
        return fail(ctx, JB_E_NOSPACE, "output buffer too small: the tiles need %zu bytes", needed);
    }
    return JB_OK;
}

static int encode_strip_once(jb_ctx* ctx, const uint8_t* rgb, size_t W, size_t strip_rows, size_t pitch, const jb_params* p,
                             uint64_t first_interval, int last_strip, int device_io, uint8_t* out, size_t cap, size_t* out_len) {
    if (!ctx || !rgb || !out || !out_len) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    Plan pl;
    int rc = make_plan(ctx, 1, W, strip_rows, p, !device_io, !device_io, &pl);
    if (rc) return rc;
    if (p->restart_interval <= 0 || pl.g.ri % pl.g.mcux != 0)
        return fail(ctx, JB_E_INVALID, "strips need a restart interval that is a whole number of MCU rows");
    if (!last_strip && strip_rows % (size_t)pl.g.mcu_px)
        return fail(ctx, JB_E_INVALID, "only the last strip may have a partial MCU row");
    Slot& s = ctx->slot[0];
    CK(cudaStreamSynchronize(s.st));
    ctx->out_internal = !device_io;
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    Framing fr{};
    fr.final_rst = last_strip ? 0u : 1u;
    fr.rst_phase = (uint32_t)(first_interval & 7);
    const uint8_t* d_rgb = rgb;
    size_t d_pitch = pitch;
    if (!device_io) {
        if ((rc = upload_frames(ctx, s, pl, rgb, W, strip_rows, pitch, pitch * strip_rows))) return rc;
        d_rgb = s.d_rgb;
        d_pitch = pl.d_pitch;
    }
    uint8_t* d_out = device_io ? out : s.d_out;
    size_t d_cap = device_io ? cap : s.d_out_cap;
    if ((rc = enqueue_encode(ctx, s, pl, p, d_rgb, d_pitch, d_pitch * strip_rows, fr, W, strip_rows, d_out, d_cap,
                             nullptr, nullptr, s.d_total)))
        return rc;
    CK(cudaMemcpyAsync(s.h_res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(s.h_res + 4, s.d_total, 8, cudaMemcpyDeviceToHost, s.st));
    s.h_res[5] = 0;
    CK(cudaMemcpyAsync(s.h_res + 5, s.d_scalars, 4, cudaMemcpyDeviceToHost, s.st));
    CK(cudaStreamSynchronize(s.st));
    if ((rc = status_to_rc(ctx, s.h_res, s.h_res[5], s.tie_cap))) return rc;
    *out_len = (size_t)s.h_res[4];
    if (!device_io) {
        if (*out_len > cap) {
            ctx->required = *out_len;
            return fail(ctx, JB_E_NOSPACE, "output buffer too small: need %zu bytes", *out_len);
        }
        CK(cudaMemcpyAsync(out, s.d_out, *out_len, cudaMemcpyDeviceToHost, s.st));
        CK(cudaStreamSynchronize(s.st));
    }
    resolve_events(ctx);
    return JB_OK;
}

int jb_encode_strip(jb_ctx* ctx, const uint8_t* rgb, size_t W, size_t strip_rows, size_t pitch, const jb_params* p,
                    uint64_t first_interval, int last_strip, int device_io, uint8_t* out, size_t cap, size_t* out_len) {
    return with_workspace_retry(ctx, [&] { return encode_strip_once(ctx, rgb, W, strip_rows, pitch, p, first_interval, last_strip, device_io, out, cap, out_len); });
}

// ---- NV12-style device input -----------------------------------------------------------------------
int jb_encode_nv12_device(jb_ctx* ctx, const uint8_t* d_y, size_t pitch_y, size_t frame_stride_y, const uint8_t* d_uv, size_t pitch_uv,
                          size_t frame_stride_uv, size_t n_frames, size_t W, size_t H, const jb_params* p, uint8_t* d_out, size_t cap,
                          uint64_t* d_offsets, uint64_t* d_sizes, uint64_t* d_total) {
    if (!ctx || !d_y || !d_uv || !d_out) return fail(ctx, JB_E_INVALID, "null argument");
    if (pitch_y < W || pitch_uv < 2 * ((W + 1) / 2) || (n_frames > 1 && (frame_stride_y < pitch_y * H || frame_stride_uv < pitch_uv * ((H + 1) / 2))))
        return fail(ctx, JB_E_INVALID, "bad pitch/stride");
    CK(cudaSetDevice(ctx->device));
    Plan pl;
    int rc = make_plan(ctx, n_frames, W, H, p, false, false, &pl);
    if (rc) return rc;
    Slot& s = ctx->slot[0];
    if (s.arena.cap < plan_bytes(pl)) CK(cudaStreamSynchronize(s.st));  // regrowing frees the arena
    ctx->out_internal = false;
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    Framing fr{};
    fr.hdr_bytes = (uint32_t)jb_header_bytes(p);
    fr.emit_eoi = 1;
    const unsigned k = s.dev_seq++ % kDevStages;
    if (s.dev_used[k]) CK(cudaEventSynchronize(s.dev_ev[k]));
    uint8_t* stage = s.h_dev + k * kDevStageBytes;
    uint64_t* res = reinterpret_cast<uint64_t*>(stage + 1024);
    const Nv12Src nv{d_uv, pitch_uv, frame_stride_uv};
    if ((rc = enqueue_encode(ctx, s, pl, p, d_y, pitch_y, frame_stride_y, fr, W, H, d_out, cap, d_offsets, d_sizes,
                             d_total ? d_total : s.d_total, stage, nullptr, &nv)))
        return rc;
    memset(res, 0, 64);
    CK(cudaMemcpyAsync(res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(res + 5, s.d_scalars, 4, cudaMemcpyDeviceToHost, s.st));
    CK(cudaEventRecord(s.dev_ev[k], s.st));
    s.dev_used[k] = true;
    s.busy = true;
    return JB_OK;
}

int jb_rgb8_to_nv12_device(jb_ctx* ctx, const uint8_t* d_rgb, size_t W, size_t H, size_t pitch, uint8_t* d_y, size_t pitch_y, uint8_t* d_uv,
                           size_t pitch_uv) {
    if (!ctx || !d_rgb || !d_y || !d_uv || W == 0 || H == 0 || pitch < W * 3 || pitch_y < W || pitch_uv < 2 * ((W + 1) / 2))
        return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    ctx->tm.total_launches += launch_rgb_to_nv12(d_rgb, W, H, pitch, ctx->d_ydown, d_y, pitch_y, d_uv, pitch_uv, ctx->slot[0].st);
    CK(cudaGetLastError());
    return JB_OK;
}

// ---- decode path: baseline JFIF -> coefficients -> RGB8 on the device, PSNR ------------------------------
static int fetch_info(jb_ctx* ctx, const uint8_t* d_jfif, size_t len, JfifInfo* info, DecTables* tabs) {
    // the marker segments are parsed on the host: bring the head of the file over (4 KB, then more while a segment runs past it)
    std::vector<uint8_t> head;
    for (size_t n = 4096;; n *= 16) {
        n = std::min(n, len);
        head.resize(n);
        CK(cudaMemcpyAsync(head.data(), d_jfif, n, cudaMemcpyDeviceToHost, ctx->slot[0].st));
        CK(cudaStreamSynchronize(ctx->slot[0].st));
        int rc = parse_jfif(head.data(), n, info, tabs);
        if (rc == JB_E_NOSPACE && n < len) continue;
        if (rc) return fail(ctx, rc == JB_E_NOSPACE ? JB_E_INVALID : rc, "not a baseline 3-component JFIF file this decoder handles (4:4:4 or 4:2:0, one scan)");
        return JB_OK;
    }
}

int jb_jfif_info_host(const uint8_t* jfif, size_t len, jb_jfif_info* out) {  // host only: no device needed
    if (!jfif || !out || len < 4) return JB_E_INVALID;
    JfifInfo info;
    DecTables tabs;
    int rc = parse_jfif(jfif, len, &info, &tabs);
    if (rc) return rc == JB_E_NOSPACE ? JB_E_INVALID : rc;
    out->W = info.W;
    out->H = info.H;
    out->subsampling = info.sub;
    out->restart_interval = info.restart_interval;
    out->scan_offset = info.scan_offset;
    return JB_OK;
}

int jb_jfif_info_device(jb_ctx* ctx, const uint8_t* d_jfif, size_t len, jb_jfif_info* out) {
    if (!ctx || !d_jfif || !out || len < 4) return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    JfifInfo info;
    DecTables tabs;
    int rc = fetch_info(ctx, d_jfif, len, &info, &tabs);
    if (rc) return rc;
    out->W = info.W;
    out->H = info.H;
    out->subsampling = info.sub;
    out->restart_interval = info.restart_interval;
    out->scan_offset = info.scan_offset;
    return JB_OK;
}

int jb_decode_jfif_device(jb_ctx* ctx, const uint8_t* d_jfif, size_t len, uint8_t* d_rgb, size_t pitch, int16_t* d_coef) {
    if (!ctx || !d_jfif || len < 4 || (!d_rgb && !d_coef)) return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    JfifInfo info;
    DecTables tabs;
    int rc = fetch_info(ctx, d_jfif, len, &info, &tabs);
    if (rc) return rc;
    if (d_rgb && pitch < (size_t)info.W * 3) return fail(ctx, JB_E_INVALID, "pitch smaller than a row");
    const Geometry g = make_geometry(info.W, info.H, info.sub, (int)info.restart_interval);
    const size_t n_blocks = (size_t)g.n_mcu * g.bpm, scan_bytes = len - info.scan_offset, chunks = (scan_bytes + 255) / 256;
    const size_t pitch_y = (size_t)g.mcux * g.mcu_px, rows_y = (size_t)g.mcuy * g.mcu_px;
    const size_t pitch_c = (size_t)g.mcux * 8, rows_c = (size_t)g.mcuy * 8;
    const size_t plane_c = info.sub == JB_SUB_420 ? pitch_c * rows_c : pitch_y * rows_y;
    const size_t need = sizeof(DecTables) + (chunks + chunks / 1024 + 4) * 4 + ((size_t)g.n_int + 2) * 8 + (d_coef ? 0 : n_blocks * 128) +
                        (d_rgb ? pitch_y * rows_y + 2 * plane_c : 0) + 8 * 256;
    if ((rc = scratch(ctx, need))) return rc;
    cudaStream_t st = ctx->slot[0].st;
    Arena& A = ctx->scratch;
    DecTables* d_tabs = carve<DecTables>(A, 1);
    uint32_t* d_cnt = carve<uint32_t>(A, chunks + chunks / 1024 + 4);
    uint64_t* d_start = carve<uint64_t>(A, (size_t)g.n_int + 2);
    int16_t* coef = d_coef ? d_coef : carve<int16_t>(A, n_blocks * 64);
    CK(cudaMemcpyAsync(d_tabs, &tabs, sizeof(tabs), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_start, 0xFF, ((size_t)g.n_int + 1) * 8, st));  // intervals whose marker is missing decode nothing
    const uint8_t* d_scan = d_jfif + info.scan_offset;
    {
        Timed t(ctx, st, 2);
        ctx->tm.total_launches += launch_rst_index(d_scan, scan_bytes, d_cnt, d_cnt + chunks + 1, d_start, (uint32_t)g.n_int, st);
        ctx->tm.total_launches += launch_huff_decode(d_scan, scan_bytes, d_start, (uint32_t)g.n_int, (uint32_t)g.ri, (uint32_t)g.n_mcu, g.bpm, d_tabs,
                                                     info, coef, st);
    }
    if (d_rgb) {
        uint8_t* py = carve<uint8_t>(A, pitch_y * rows_y);
        uint8_t* pcb = carve<uint8_t>(A, plane_c);
        uint8_t* pcr = carve<uint8_t>(A, plane_c);
        Timed t(ctx, st, 0);
        ctx->tm.total_launches += launch_reconstruct(coef, (uint32_t)g.n_mcu, g.mcux, g.bpm, info, py, pcb, pcr, pitch_y,
                                                     info.sub == JB_SUB_420 ? pitch_c : pitch_y, d_rgb, pitch, st);
    }
    STAGE_END()
}

int jb_decode_jfif(jb_ctx* ctx, const uint8_t* jfif, size_t len, uint8_t* rgb, size_t cap, size_t* W, size_t* H) {
    if (!ctx || !jfif || len < 4) return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    JfifInfo info;
    DecTables tabs;
    int rc = parse_jfif(jfif, len, &info, &tabs);
    if (rc) return fail(ctx, rc == JB_E_NOSPACE ? JB_E_INVALID : rc, "not a baseline 3-component JFIF file this decoder handles");
    if (W) *W = info.W;
    if (H) *H = info.H;
    const size_t out_bytes = (size_t)info.W * info.H * 3;
    if (!rgb || cap < out_bytes) {
        ctx->required = out_bytes;
        return fail(ctx, JB_E_NOSPACE, "output buffer too small: need %zu bytes", out_bytes);
    }
    uint8_t *d_in = nullptr, *d_out = nullptr;  // (outside the scratch arena, which the device call carves)
    if (cudaMalloc(&d_in, len) != cudaSuccess || cudaMalloc(&d_out, out_bytes) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d_in);
        return fail(ctx, JB_E_NOMEM, "cudaMalloc failed");
    }
    cudaStream_t st = ctx->slot[0].st;
    cudaMemcpyAsync(d_in, jfif, len, cudaMemcpyHostToDevice, st);
    rc = jb_decode_jfif_device(ctx, d_in, len, d_out, (size_t)info.W * 3, nullptr);
    if (rc == JB_OK && (cudaMemcpyAsync(rgb, d_out, out_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess))
        rc = fail(ctx, JB_E_CUDA, "copy back failed");
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

int jb_psnr_device(jb_ctx* ctx, const uint8_t* d_a, size_t pitch_a, const uint8_t* d_b, size_t pitch_b, size_t W, size_t H, double* psnr,
                   uint64_t* sq_err) {
    if (!ctx || !d_a || !d_b || W == 0 || H == 0 || pitch_a < W * 3 || pitch_b < W * 3) return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    int rc = scratch(ctx, 256);
    if (rc) return rc;
    cudaStream_t st = ctx->slot[0].st;
    unsigned long long* d_sum = carve<unsigned long long>(ctx->scratch, 1);
    ctx->tm.total_launches += launch_sq_err(d_a, pitch_a, d_b, pitch_b, W, H, d_sum, st);
    unsigned long long sum = 0;
    CK(cudaMemcpyAsync(&sum, d_sum, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (sq_err) *sq_err = sum;
    if (psnr) *psnr = sum ? 10.0 * log10(255.0 * 255.0 * (double)(W * H * 3) / (double)sum) : INFINITY;
    return JB_OK;
}

// ---- multi-GPU strip stitch without a host round trip ------------------------------------------------------------
// begin: transform + entropy coder up to the sizes, asynchronous; *d_len (device) receives the strip's byte count.
// finish: the final placement (0xFF00 stuffing, RSTn) writes the strip to d_out + *d_off -- d_off is a device scalar
// the caller computes from the all-gathered lengths, and d_out may be another GPU's buffer mapped with jb_ipc_open:
// the kernel's coalesced 128-bit stores are then the NVLink transfer, there is no separate gather.
int jb_encode_strip_begin(jb_ctx* ctx, const uint8_t* d_rgb, size_t W, size_t strip_rows, size_t pitch, const jb_params* p,
                          uint64_t first_interval, int last_strip, uint64_t* d_len) {
    if (!ctx || !d_rgb || !d_len) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    Plan pl;
    int rc = make_plan(ctx, 1, W, strip_rows, p, false, false, &pl);
    if (rc) return rc;
    if (p->restart_interval <= 0 || pl.g.ri % pl.g.mcux != 0)
        return fail(ctx, JB_E_INVALID, "strips need a restart interval that is a whole number of MCU rows");
    if (!last_strip && strip_rows % (size_t)pl.g.mcu_px)
        return fail(ctx, JB_E_INVALID, "only the last strip may have a partial MCU row");
    if (p->flags & JB_FLAG_OPTIMIZE_HUFFMAN) return fail(ctx, JB_E_UNSUPPORTED, "strips carry no tables");
    Slot& s = ctx->slot[0];
    if (s.strip_pending) return fail(ctx, JB_E_INVALID, "jb_encode_strip_begin twice without jb_encode_strip_finish");
    if (s.arena.cap < plan_bytes(pl)) CK(cudaStreamSynchronize(s.st));  // regrowing frees the arena
    ctx->out_internal = false;
    if ((rc = slot_prepare(ctx, s, pl))) return rc;
    Framing fr{};
    fr.final_rst = last_strip ? 0u : 1u;
    fr.rst_phase = (uint32_t)(first_interval & 7);
    // d_out only has to be non-null here: nothing is written before jb_encode_strip_finish
    if ((rc = enqueue_encode(ctx, s, pl, p, d_rgb, pitch, pitch * strip_rows, fr, W, strip_rows, reinterpret_cast<uint8_t*>(d_len), 0,
                             nullptr, nullptr, d_len, nullptr, &s.strip_ea)))
        return rc;
    s.strip_pending = true;
    return JB_OK;
}

int jb_encode_strip_finish(jb_ctx* ctx, uint8_t* d_out, size_t cap, const uint64_t* d_off) {
    if (!ctx || !d_out) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[0];
    if (!s.strip_pending) return fail(ctx, JB_E_INVALID, "jb_encode_strip_finish without jb_encode_strip_begin");
    s.strip_pending = false;
    EntropyArgs ea = s.strip_ea;
    ea.out = d_out;
    ea.out_cap = cap;
    ea.out_off = d_off;
    {
        Timed t(ctx, s.st, 2);
        ctx->tm.total_launches += launch_entropy(ea, s.st, 2);
    }
    CK(cudaGetLastError());
    const unsigned k = s.dev_seq++ % kDevStages;  // status words: as for jb_encode_batch_device, examined by jb_sync
    if (s.dev_used[k]) CK(cudaEventSynchronize(s.dev_ev[k]));
    uint64_t* res = reinterpret_cast<uint64_t*>(s.h_dev + k * kDevStageBytes + 1024);
    memset(res, 0, 64);
    CK(cudaMemcpyAsync(res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
    CK(cudaMemcpyAsync(res + 5, s.d_scalars, 4, cudaMemcpyDeviceToHost, s.st));
    CK(cudaEventRecord(s.dev_ev[k], s.st));
    s.dev_used[k] = true;
    s.busy = true;
    return JB_OK;
}

int jb_copy_bytes_device(jb_ctx* ctx, uint8_t* d_dst, size_t cap, const uint64_t* d_dst_off, const uint8_t* d_src,
                         const uint64_t* d_len) {
    if (!ctx || !d_dst || !d_src || !d_len) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    ctx->tm.total_launches += launch_copy_bytes(d_dst, d_dst_off, d_src, d_len, cap, nullptr, ctx->slot[0].st);
    CK(cudaGetLastError());
    return JB_OK;
}

// The exchange step of the stitch without a collective library: lengths and completion flags are plain stores and
// polls on a 512-byte control block of the stitching rank, mapped by every rank over NVLink (jb_entropy.cu).
int jb_stitch_exchange(jb_ctx* ctx, uint64_t* d_ctl, int rank, int world, uint64_t epoch, uint64_t base, const uint64_t* d_len,
                       uint64_t* d_off) {
    if (!ctx || !d_ctl || !d_len || !d_off || world < 1 || world > 16 || rank < 0 || rank >= world || epoch == 0)
        return fail(ctx, JB_E_INVALID, "bad arguments (1 <= world <= 16, epoch >= 1)");
    CK(cudaSetDevice(ctx->device));
    ctx->tm.total_launches += launch_stitch_exchange(d_ctl, rank, world, epoch, base, d_len, d_off, ctx->slot[0].st);
    CK(cudaGetLastError());
    return JB_OK;
}
int jb_stitch_complete(jb_ctx* ctx, uint64_t* d_ctl, int rank, int world, int dst, uint64_t epoch) {
    if (!ctx || !d_ctl || world < 1 || world > 16 || rank < 0 || rank >= world || dst < 0 || dst >= world || epoch == 0)
        return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[0];
    ctx->tm.total_launches += launch_stitch_complete(d_ctl, rank, world, dst, epoch, s.w.status, s.st);
    CK(cudaGetLastError());
    if (rank == dst && s.w.status) {  // a peer that never reported shows up at jb_sync
        const unsigned k = s.dev_seq++ % kDevStages;
        if (s.dev_used[k]) CK(cudaEventSynchronize(s.dev_ev[k]));
        uint64_t* res = reinterpret_cast<uint64_t*>(s.h_dev + k * kDevStageBytes + 1024);
        memset(res, 0, 64);
        CK(cudaMemcpyAsync(res, s.w.status, 32, cudaMemcpyDeviceToHost, s.st));
        CK(cudaEventRecord(s.dev_ev[k], s.st));
        s.dev_used[k] = true;
        s.busy = true;
    }
    return JB_OK;
}

// CUDA IPC: one process per GPU (torch.distributed), so the stitching rank's buffer is shared by handle
int jb_ipc_export(jb_ctx* ctx, void* d_ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!ctx || !d_ptr || !handle) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return JB_OK;
}
int jb_ipc_open(jb_ctx* ctx, const uint8_t handle[64], void** d_ptr) {
    if (!ctx || !handle || !d_ptr) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return JB_OK;
}
int jb_ipc_close(jb_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return fail(ctx, JB_E_INVALID, "null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return JB_OK;
}

int jb_synth_rgb_device(jb_ctx* ctx, uint64_t seed, size_t W, size_t y0, size_t rows, size_t pitch, uint8_t* d_out) {
    if (!ctx || !d_out || pitch < W * 3) return fail(ctx, JB_E_INVALID, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    ctx->tm.total_launches += launch_synth(seed, W, y0, rows, pitch, d_out, ctx->slot[0].st);
    CK(cudaGetLastError());
    return JB_OK;
}

}  // extern "C"
