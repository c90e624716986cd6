#!/usr/bin/env python
"""bench.py -- throughput of the B200 JPEG encode path (BASELINE.json metric:
"encode megapixels/sec (4:2:0, q75) at 1/2/4/8 B200; fused-kernel HBM GB/s").

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic frames.  Default workload:
BASELINE.json config #4 sharded by image, weak scaling: every GPU encodes --frames
1920x1080 frames, 4:2:0, q75 (at 8 GPUs x 512 frames this is exactly the 4096-frame batch;
N = 1 is 512 of its 4096 frames).
`value` is device-resident (RGB already in HBM, JFIF bytes left in HBM); `e2e` is the same
work through the public host API jb_encode_batch (pinned host RGB in, pinned host JFIF out,
H2D/D2H inside the timed region).  After the timed regions sampled frames of the very output
that was timed are compared with the CPU oracle (`parity_check`), and the default run also
times the other single-GPU configurations briefly (`config.other_workloads`).
`--workload gigapixel|strips16k` is config #5: ONE image split into RST strips across the
ranks (strong scaling), stitched into rank 0's buffer over NVLink peer memory.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "encode megapixels/sec (4:2:0, q75)"
WORKLOADS = {
    # name: (W, H, subsampling, quality, restart interval in MCUs, default frames per GPU, seed of frame 0)
    "batch1080p": (1920, 1080, "420", 75, 0, 512, 0xF000),
    "8k": (7680, 4320, "420", 75, 480, 8, 0x4B7680),
    "4k444": (3840, 2160, "444", 90, 0, 32, 0x4B3840),
    # the reference's own mode (replicated 4:2:0 coded as 4:4:4, its q50 tables) on the batch frames
    "repl1080p": (1920, 1080, "repl420", 50, 0, 256, 0xF000),
    # one image split into strips of whole restart intervals across the GPUs (config #5; strong scaling)
    "gigapixel": (65536, 65536, "420", 75, 4096, 1, 0x65536),
    "strips16k": (16384, 16384, "420", 75, 1024, 1, 0x65536),
}
STRIP_WORKLOADS = ("gigapixel", "strips16k")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="batch1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of sampled frames after the timed region")
    ap.add_argument("--no-others", action="store_true", help="skip the brief runs of the other single-GPU workloads")
    ap.add_argument("--stitch", default="peer", choices=["peer", "peer-nccl", "nccl"],
                    help="strip workloads: peer = placement kernel stores through NVLink peer memory, lengths / completion as "
                         "flags on peer memory (default); peer-nccl = same placement, lengths by NCCL all-gather; nccl = grouped send/recv")
    ap.add_argument("--optimize-huffman", action="store_true",
                    help="JB_FLAG_OPTIMIZE_HUFFMAN: per-call optimal Huffman tables (two passes; not the headline configuration)")
    ap.add_argument("--ref-exact", action="store_true",
                    help="the reference as written: JB_FLAG_REF_INPLACE_DCT | REF_TYPO_TABLES | REF_ALWAYS_EOB (use with --workload repl1080p)")
    ap.add_argument("--tma", action="store_true", help="JB_FLAG_TMA: stage the image tiles with TMA boxes instead of per-lane cp.async (A/B)")
    ap.add_argument("--tensor-dct", type=int, default=1,
                    help="transform kernel: 1 = tcgen05 (library default), 0 = CUDA-core FMA kernel (JB_FLAG_FMA_DCT)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------
# CPU reference arm: the reference's own src/utils.cpp (oracle/_ref, compiled unmodified),
# driven in the order of JpegEncoderHost, one process per host core on distinct frames.
# ----------------------------------------------------------------------------------------

def _ref_worker(args):
    seed, W, H, o0 = args
    import numpy as np  # noqa: F401
    import oracle_lib as ol
    rgb = ol.synth(seed, W, H)
    kind = "reference" if ol.have_ref() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        ol.ref_pipeline(rgb, 0, o0=o0)  # as written: CSC, CDS, pad, shift, DCT, quant, zigzag, RLE, Huffman
    else:
        ql, qc = ol.q50()
        ol.entropy(ol.transform(rgb, ol.SUB_REPL420, ql, qc, ol.AS_WRITTEN), ol.SUB_REPL420, 0, ol.AS_WRITTEN, True)
    return time.perf_counter() - t0, kind


def cpu_reference_pass(W, H, frames_per_core, cores, pool, seed0=0xF000):
    """One bounded sample: cores x frames_per_core frames through the reference CPU path.
    Returns (MP/s over all cores, wall seconds, kind)."""
    jobs = [(seed0 + i, W, H, False) for i in range(cores * frames_per_core)]
    t0 = time.perf_counter()
    res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    return len(jobs) * W * H / 1e6 / wall, wall, res[0][1]


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    import multiprocessing as mp
    import oracle_lib as ol
    W, H, sub, q, ri, _, _ = WORKLOADS[a.workload]
    cores = os.cpu_count() or 1
    # The reference's code is loaded in THIS process too (not only in the forked workers, which the pool tears down):
    # the driver's record of loaded native libraries then shows oracle/_ref/libjpegref.so for this arm.
    have_ref = ol.have_ref() and ol.ref() is not None
    # one core, in this process: the reference is single-threaded, -O2 and as shipped (-g, no -O;
    # CMakeLists.txt:29 -> oracle/_ref/libjpegref_O0.so), on a 1920x128 strip of the first frame
    one_core = {}
    if have_ref:
        srgb = ol.synth(0xF000, min(W, 1920), 128)
        for name, o0 in (("O2", False), ("as_shipped_g_O0", True)):
            if ol.ref(o0) is None:
                continue
            ol.ref_pipeline(srgb, 0, o0=o0)
            t0 = time.perf_counter()
            ol.ref_pipeline(srgb, 0, o0=o0)
            one_core[name] = round(srgb.shape[0] * srgb.shape[1] / 1e6 / (time.perf_counter() - t0), 4)
    # a 1080p frame costs the reference ~3.5 s on one core: one frame per core per step
    sw, sh = (W, H) if W * H <= 1920 * 1080 else (W, 64)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(a.warmup, 0)):
            cpu_reference_pass(sw, sh, 1, cores, pool)
        walls, kind = [], "reference"
        for _ in range(a.steps):
            _, wall, kind = cpu_reference_pass(sw, sh, 1, cores, pool)
            walls.append(wall)
    ms = 1000.0 * sum(walls) / len(walls)
    value = cores * sw * sh / 1e6 / (ms / 1000.0)
    sample = (f"{cores} frames of {sw}x{sh} per step (one per core), reference CPU path as written "
              f"(its only mode: replicated 4:2:0 coded 4:4:4, q50 tables, -O2)"
              + ("" if (sw, sh) == (W, H) else f"; strip of the {W}x{H} image, full-image figure is extrapolated"))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "MP/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.workload, a.frames), "note": "CPU arm: rank 0 only, host cores, no GPU"},
        "cpu_baseline": {"value": round(value, 4), "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample,
                         "one_core_MPs": one_core,
                         "one_core_sample": "1920x128 strip, one thread, reference stage order (SURVEY 8d): -O2 and the flags the reference ships (-g)",
                         "library": os.path.relpath(ol.REF_SO, ROOT) if have_ref else "oracle/_build/liboracle.so (port)"},
        "e2e": {"value": round(value, 4), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(workload, frames=0):
    W, H, sub, q, ri, dflt, _ = WORKLOADS[workload]
    f = frames or dflt
    return (f"{workload}: {f} synthetic {W}x{H} RGB8 frames per GPU, {sub}, q{q}, "
            f"restart interval {ri} MCUs, sharded by image")


# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        # started BEFORE the warm-up and waited for until it streams: a timed region of a few tens of milliseconds
        # (10 steps of 2.8 ms) is shorter than nvidia-smi's start-up, and was shorter than the 100 ms period used before
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.perf_counter() + 3.0
            while not self.rows and time.perf_counter() < t_end:
                time.sleep(0.005)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        inside = [r for t, r in self.rows if t0 <= t <= t1]
        # a region shorter than the sampling period: the nearest samples on both sides (the GPU is under the same load
        # during the warm-up steps right before t0)
        near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[:3]
        rows = inside or [r for _, r in near]
        try:
            sm = [float(r[0]) for r in rows]
            mx = max(float(r[1]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                    "samples_inside_timed_region": len(inside), "power_w_max": max(float(r[2]) for r in rows)}
        except Exception:
            return None


class Dist:
    """Rank bookkeeping + the barrier / max-over-ranks helpers of the timing contract."""

    def __init__(self, torch):
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run (and first-touch pinned memory) on the CPUs NVML reports as local to the GPU.
    A no-op on single-node hosts or when NVML / the cpuset do not allow it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {i for i in allowed if (words[i // 64] >> (i % 64)) & 1}
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def transform_roofline(tm, alg_bytes, steps, kname, traffic=None):
    peak, peak_src = peaks()
    k_us = tm["transform_us"] / max(tm["transform_launches"], 1)
    achieved = alg_bytes / (k_us * 1e-6) / 1e9 if k_us > 0 else 0.0
    return {"kernel": kname, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
            "frac_of_nominal_8000": round(achieved / 8000.0, 4), "algorithmic_bytes_per_launch": int(alg_bytes),
            "kernel_us_per_launch": round(k_us, 2),
            "step_breakdown_us": {"transform": round(tm["transform_us"] / steps, 1), "edge_mcus": round(tm["edge_us"] / steps, 1),
                                  "tie_fixup": round(tm["fixup_us"] / steps, 1), "entropy": round(tm["entropy_us"] / steps, 1)}}


def profile_traffic(pname, workload, F):
    """Per-launch DRAM bytes of the transform kernel from the committed ncu --set full capture of the same command
    (profiles/r0*_transform_ncu_summary.json): a constant of the builder's profile, not measured by this run."""
    for name in ("r02_transform_ncu_summary.json", "r01_transform_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                prof = json.load(f)
            k = prof.get("kernels", {}).get(pname, {})
            if k.get("workload", prof.get("workload")) == workload and k.get("frames", prof.get("frames")) == F:
                return k.get("dram_bytes_per_launch"), "profiles/" + name
        except Exception:
            pass
    return None, None


# ----------------------------------------------------------------------------------------
# batch workloads (configs #2, #3, #4 and the reference's own mode)
# ----------------------------------------------------------------------------------------
def run_batch(a, jb, enc, torch, dd, workload, F, steps, warmup, want_e2e, want_parity, ref_exact=False, sample_clocks=True):
    """Time one batch workload on this rank's GPU.  Returns the result dict (rank 0 keeps it)."""
    import numpy as np
    W, H, subname, q, ri, _, seed0 = WORKLOADS[workload]
    rank, world, local_rank = dd.rank, dd.world, dd.local_rank
    sub = {"420": jb.SUB_420, "444": jb.SUB_444, "repl420": jb.SUB_REPL420}[subname]
    flags = (0 if a.tensor_dct else jb.FLAG_FMA_DCT) | (jb.FLAG_OPTIMIZE_HUFFMAN if a.optimize_huffman else 0) | (jb.FLAG_TMA if a.tma else 0) \
        | ((jb.FLAG_REF_INPLACE_DCT | jb.FLAG_REF_TYPO_TABLES | jb.FLAG_REF_ALWAYS_EOB) if ref_exact else 0)
    params = jb.make_params(sub, quality=q, restart_interval=ri, flags=flags)
    pitch, fstride = W * 3, W * H * 3
    px_per_step = W * H * F  # per GPU
    frame_seed = lambda f: seed0 + rank * F + f  # noqa: E731  (config #4: 0xF000 + frame index over the whole batch)

    # inputs: generated on the device (identical to the oracle's generator), copied once to pinned host
    d_rgb = torch.empty(F * fstride, dtype=torch.uint8, device="cuda")
    for f in range(F):
        enc.synth_device(frame_seed(f), W, 0, H, pitch, d_rgb.data_ptr() + f * fstride)
    enc.sync()
    cap = F * (W * H // (1 if ref_exact else 2) + 4096)  # the as-written transform yields ~4 bits/px
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_tab = torch.zeros(2 * F + 1, dtype=torch.int64, device="cuda")
    ext = torch.cuda.ExternalStream(enc.stream())

    def step_device():
        enc.encode_batch_device(d_rgb.data_ptr(), F, W, H, pitch, fstride, params, d_out.data_ptr(), cap,
                                d_tab.data_ptr(), d_tab.data_ptr() + 8 * F, d_tab.data_ptr() + 16 * F)

    # ---- device-resident timed region ------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 and sample_clocks else None  # (streaming before the warm-up starts)
    for _ in range(max(warmup, 3)):
        step_device()
    enc.sync()
    enc.set_profiling(True)
    enc.reset_counters()
    dd.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(steps):
        step_device()
    ev1.record(ext)
    enc.sync()
    dd.barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    tm = enc.timings()
    enc.set_profiling(False)
    total_bytes = int(d_tab[2 * F].item())
    ms_step = dd.max_ms(ms_total) / steps
    value = world * px_per_step / 1e6 / (ms_step / 1e3)

    # ---- roofline of the fused transform kernel (CUDA events on its own stream, timed region) ----
    g_mcu = 16 if sub == jb.SUB_420 else 8
    padded = (-(-W // g_mcu) * g_mcu) * (-(-H // g_mcu) * g_mcu)
    samples_per_px = 1.5 if sub == jb.SUB_420 else 3.0
    alg_bytes = F * (3 * W * H + 2 * samples_per_px * padded)  # read RGB8 + write int16 coefficients
    use_tc = bool(a.tensor_dct)
    kname = ("k_transform_tc" if sub == jb.SUB_420 else "k_transform_tc3") if use_tc else "k_transform"
    pname = "k_transform_tc3_repl" if (use_tc and sub == jb.SUB_REPL420) else kname  # key in the ncu summary
    traffic, traffic_src = profile_traffic(pname, workload, F)
    roofline = transform_roofline(tm, alg_bytes, steps, kname + (" (fused CSC+subsample+shift, tcgen05 FDCT+quant+zigzag)" if use_tc else
                                                                 " (fused CSC+subsample+shift+FDCT+quant+zigzag, CUDA cores)"), traffic)
    roofline["traffic_source"] = traffic_src
    launches_device = int(tm["total_launches"])

    # ---- oracle parity of the output that was just timed (outside the timing) -------------------
    parity = None
    if want_parity and rank == 0 and not a.optimize_huffman:
        import oracle_lib as ol
        tab = d_tab.cpu().numpy()
        frames = sorted({0, F - 1} if W * H > 1920 * 1080 else {0, F // 2, F - 1})
        if W * H > 1920 * 1080:
            frames = frames[:1] if W * H > 4000 * 3000 else frames[:2]
        ql, qc = ol.quality_tables(q) if not ref_exact else ol.q50()
        osub = {"420": ol.SUB_420, "444": ol.SUB_444, "repl420": ol.SUB_REPL420}[subname]
        equal, t_or = True, time.perf_counter()
        for f in frames:
            got = bytes(d_out[int(tab[f]): int(tab[f]) + int(tab[F + f])].cpu().numpy())
            want = ol.encode_jfif(ol.synth(frame_seed(f), W, H), osub, ql, qc, ri, ol.AS_WRITTEN if ref_exact else 0)
            equal = equal and got == want
        parity = {"frames": frames, "equal": bool(equal), "against": "oracle/jpeg_oracle.c (whole JFIF files, byte for byte)",
                  "oracle_seconds": round(time.perf_counter() - t_or, 2)}
        if not ref_exact:  # decoded on the GPU (jb_decode_jfif_device, libjpeg-exact reconstruction) against the source frame
            d_dec = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            enc.decode_jfif_device(d_out.data_ptr() + int(tab[0]), int(tab[F]), d_dec.data_ptr(), W * 3, None)
            parity["psnr_db_frame0_gpu_decoder"] = round(enc.psnr_device(d_dec.data_ptr(), W * 3, d_rgb.data_ptr(), pitch, W, H)[0], 3)
            del d_dec

    # ---- end to end through the public host API: pinned host in, pinned host out ----------------
    e2e = None
    if want_e2e:
        h_rgb = jb.pinned_empty((F, H, W, 3))
        enc.d2h(h_rgb, d_rgb.data_ptr())
        h_out = jb.pinned_empty((cap,))
        offs = np.zeros(F, np.uint64)
        sizes = np.zeros(F, np.uint64)

        def step_host():
            enc.encode_batch_ptr(h_rgb.ctypes.data, F, W, H, pitch, fstride, params, h_out.ctypes.data, cap, offs, sizes)

        for _ in range(max(warmup, 3)):
            step_host()
        dd.barrier()
        w0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        dd.barrier()
        e_step = dd.max_ms((w1 - w0) * 1e3) / steps
        out_bytes = int(sizes.sum())
        # host path == device path (with per-call optimal tables the 96 MB groups of the host path get their own tables)
        assert a.optimize_huffman or out_bytes == total_bytes, (out_bytes, total_bytes)
        e2e = {"value": round(world * px_per_step / 1e6 / (e_step / 1e3), 1), "unit": "MP/s",
               "ms_per_step": round(e_step, 3), "h2d_bytes_per_step": int(F * fstride),
               "d2h_bytes_per_step": int(out_bytes + 16 * F + 48),
               "h2d_GBps_per_gpu": round(F * fstride / 1e9 / (e_step / 1e3), 1),
               "api": "jb_encode_batch (pinned host RGB -> pinned host JFIF, 3 streams, 96 MB groups)"}
        del h_rgb, h_out
    del d_rgb, d_out, d_tab
    torch.cuda.empty_cache()
    return {"value": round(value, 1), "ms_per_step": round(ms_step, 4), "e2e": e2e, "roofline": roofline, "clocks": clocks,
            "parity_check": parity, "gpu_launches": launches_device, "frames": F,
            "bits_per_pixel": round(8.0 * total_bytes / px_per_step, 4), "tie_fixups_per_step": int(tm["tie_fixups"]),
            "transform_kernel": kname + (" (tcgen05)" if use_tc else " (FMA pipe)"),
            "dims": (W, H, subname, q, ri)}


def run_nv12(a, jb, enc, torch, dd, F=256, steps=3, warmup=3):
    """the batch frames as NV12-style device input (converted on the device with the reference's CSC + CDS,
    outside the timing).  Device-resident only: the point of the format is that a GPU producer left the frames in HBM.
    Parity: for these even-sized frames the files must equal the oracle's encode of the RGB frames."""
    W, H, _, q, ri, _, seed0 = WORKLOADS["batch1080p"]
    params = jb.make_params(jb.SUB_420, quality=q, restart_interval=ri)
    d_rgb = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
    d_y = torch.empty(F * W * H, dtype=torch.uint8, device="cuda")
    d_uv = torch.empty(F * W * H // 2, dtype=torch.uint8, device="cuda")
    for f in range(F):
        enc.synth_device(seed0 + f, W, 0, H, W * 3, d_rgb.data_ptr())
        enc.rgb8_to_nv12_device(d_rgb.data_ptr(), W, H, W * 3, d_y.data_ptr() + f * W * H, W, d_uv.data_ptr() + f * W * H // 2, W)
    enc.sync()
    cap = F * (W * H // 2 + 4096)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_tab = torch.zeros(2 * F + 1, dtype=torch.int64, device="cuda")
    ext = torch.cuda.ExternalStream(enc.stream())

    def step():
        enc.encode_nv12_device(d_y.data_ptr(), W, W * H, d_uv.data_ptr(), W, W * H // 2, F, W, H, params, d_out.data_ptr(), cap,
                               d_tab.data_ptr(), d_tab.data_ptr() + 8 * F, d_tab.data_ptr() + 16 * F)
    for _ in range(warmup):
        step()
    enc.sync()
    enc.set_profiling(True)
    enc.reset_counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(steps):
        step()
    ev1.record(ext)
    enc.sync()
    ms = ev0.elapsed_time(ev1) / steps
    tm = enc.timings()
    enc.set_profiling(False)
    padded = W * (-(-H // 16) * 16)
    # the complete MCU columns are the tcgen05 kernel's share (a partial last column is timed as edge_mcus): Y and CbCr
    # bytes of the image rows read, int16 coefficients of the padded rows written
    fast_w = W // 16 * 16
    roof = transform_roofline(tm, F * (1.5 * fast_w * H + 3.0 * fast_w * (-(-H // 16) * 16)), steps, "k_transform_tc_nv12 (tcgen05 FDCT+quant+zigzag, no colour conversion to do; partial MCU rows / columns: k_transform_nv12)")
    parity = None
    if not a.no_parity:
        import oracle_lib as ol
        tab = d_tab.cpu().numpy()
        ql, qc = ol.quality_tables(q)
        fr = [0, F - 1]
        ok = all(bytes(d_out[int(tab[f]): int(tab[f]) + int(tab[F + f])].cpu().numpy()) ==
                 ol.encode_jfif(ol.synth(seed0 + f, W, H), ol.SUB_420, ql, qc, ri) for f in fr)
        parity = {"frames": fr, "equal": bool(ok), "against": "oracle/jpeg_oracle.c encode of the RGB frames (even sizes: NV12 path == RGB path)"}
    # end to end through jb_encode_nv12_batch: pinned host planes in, pinned host JFIF out (1.5 B/px over the host link)
    e2e = None
    if not a.no_e2e:
        import numpy as np
        import time
        total = int(d_tab[F:2 * F].sum().item())
        h_y, h_uv = jb.pinned_empty((F, H, W)), jb.pinned_empty((F, H // 2, W))
        enc.d2h(h_y, d_y.data_ptr())
        enc.d2h(h_uv, d_uv.data_ptr())
        h_out = jb.pinned_empty((cap,))
        offs, sizes = np.zeros(F, np.uint64), np.zeros(F, np.uint64)

        def step_host():
            enc.encode_nv12_batch_ptr(h_y.ctypes.data, W, W * H, h_uv.ctypes.data, W, W * H // 2, F, W, H, params, h_out.ctypes.data, cap,
                                      offs, sizes)
        for _ in range(3):
            step_host()
        w0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        torch.cuda.synchronize()
        e_step = (time.perf_counter() - w0) * 1e3 / steps
        assert int(sizes.sum()) == total, (int(sizes.sum()), total)  # host path == device path
        e2e = {"value": round(F * W * H / 1e6 / (e_step / 1e3), 1), "unit": "MP/s", "ms_per_step": round(e_step, 3),
               "h2d_bytes_per_step": int(F * W * H * 3 // 2), "d2h_bytes_per_step": int(total + 16 * F + 48),
               "h2d_GBps_per_gpu": round(F * W * H * 1.5 / 1e9 / (e_step / 1e3), 1),
               "api": "jb_encode_nv12_batch (pinned host Y + CbCr planes -> pinned host JFIF, 3 streams)"}
        del h_y, h_uv, h_out
    del d_y, d_uv, d_out
    torch.cuda.empty_cache()
    return {"workload": f"nv12_1080p: {F} frames of 1920x1080 as NV12-style device input (Y plane + interleaved CbCr plane), 420, q{q}",
            "value": round(F * W * H / 1e6 / (ms / 1e3), 1), "ms_per_step": round(ms, 4), "e2e": e2e["value"] if e2e else None,
            "e2e_detail": e2e, "roofline_frac": roof["frac"],
            "transform_GBps": roof["achieved"], "transform_kernel": "k_transform_tc_nv12", "step_breakdown_us": roof["step_breakdown_us"],
            "parity_check": parity, "steps": steps}


# ----------------------------------------------------------------------------------------
# config #5: one image as RST strips across the ranks, stitched over NVLink
# ----------------------------------------------------------------------------------------
def run_strips(a, jb, enc, torch, dd):
    """One large image, split into horizontal strips of whole restart intervals (one MCU row each).
    Rank r encodes rows [row0, row1) from HBM; the only exchange is the stitch into rank 0's buffer:
      peer (default)  all-gather of the strip lengths (8 bytes per rank), offsets by a device-side cumsum, then the
                      entropy coder's placement kernel stores the strip through NVLink peer memory at its final
                      offset (jb_encode_strip_begin / _finish, dist.PeerStitch): no host round trip, no gather;
      nccl            jb_encode_strip into local memory, lengths to the host, one grouped send/recv (gather_stitch).
    Timed with CUDA events on the encoder's stream over K back-to-back steps, max over ranks."""
    import importlib
    import numpy as np
    D = importlib.import_module("jpegb200.dist")
    rank, world, local_rank, dist = dd.rank, dd.world, dd.local_rank, dd.dist
    W, H, subname, q, ri, _, seed = WORKLOADS[a.workload]
    params = jb.make_params(jb.SUB_420, quality=q, restart_interval=ri,
                            flags=jb.FLAG_CLAMP_SOF | (0 if a.tensor_dct else jb.FLAG_FMA_DCT))
    plan = D.plan_strips(H, 16, 1, world)
    row0, row1, first, is_last = plan[rank]
    rows, pitch = row1 - row0, W * 3
    chunk = 16384  # rows per call (keeps every call below 2^26 blocks at W = 65536)
    chunks = [(y, min(chunk, rows - y)) for y in range(0, rows, chunk)]
    d_rgb = torch.empty(max(rows, 1) * pitch, dtype=torch.uint8, device="cuda")
    for y in range(0, rows, 1024):
        enc.synth_device(seed, W, row0 + y, min(1024, rows - y), pitch, d_rgb.data_ptr() + y * pitch)
    enc.sync()
    header = torch.frombuffer(bytearray(enc.write_header(params, W, H)), dtype=torch.uint8).cuda()
    hdr_n = header.numel()
    eoi = torch.tensor([0xFF, 0xD9], dtype=torch.uint8, device="cuda")
    two = torch.arange(2, dtype=torch.int64, device="cuda")
    cap_file = W * H // 2 + (1 << 20)          # the stitched file (0.18 B/px at q75 on this content)
    cap_local = rows * W // 2 + (1 << 20)
    local = torch.empty(cap_local, dtype=torch.uint8, device="cuda")  # this rank's strips (warm-up, nccl mode, multi-call ranks)
    run = torch.zeros(len(chunks) + 1, dtype=torch.int64, device="cuda")
    ext = torch.cuda.ExternalStream(enc.stream())
    mode = a.stitch if world > 1 else "single"
    ps = None
    if world > 1 and mode.startswith("peer"):
        try:
            ps = D.PeerStitch(enc, cap_file, dst=0)
        except Exception as e:  # no CUDA IPC between the ranks (containers without a shared IPC namespace)
            mode = f"nccl (peer mapping failed: {e})"
            ps = None
        flag = torch.tensor([1 if ps is not None else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and ps is not None:
            ps.close()
            ps, mode = None, "nccl (peer mapping failed on another rank)"
    own = None
    if world == 1:
        own = torch.empty(cap_file, dtype=torch.uint8, device="cuda")
        own[:hdr_n] = header
    elif ps is not None and rank == 0:
        ps.view()[:hdr_n] = header
    torch.cuda.synchronize()

    def encode_chunks_sync():
        """classic synchronous calls into `local` (warm-up: sizes the workspaces; nccl mode: the step's encode)"""
        off = 0
        for y, n in chunks:
            off += enc.encode_strip(d_rgb.data_ptr() + y * pitch, params, first + y // 16, is_last and y + n == rows, W=W, rows=n,
                                    pitch=pitch, device_io=True, out=local.data_ptr() + off, cap=cap_local - off)
        return off

    def step_peer():
        """all stream-ordered on the encoder's stream; returns (device tensor holding the end offset of the data)"""
        with torch.cuda.stream(ext):
            if world == 1:
                run[0] = hdr_n
                for c, (y, n) in enumerate(chunks):
                    enc.encode_strip_begin(d_rgb.data_ptr() + y * pitch, params, first + y // 16, is_last and y + n == rows, W, n, pitch,
                                           run.data_ptr() + 8 * (c + 1))       # run[c+1] <- length
                    enc.encode_strip_finish(own.data_ptr(), cap_file, run.data_ptr() + 8 * c)
                    run[c + 1] += run[c]                                      # -> end offset
                end = run[len(chunks):]
                own.index_copy_(0, end + two, eoi)
                return end
            def offsets():  # -> (pointer to this rank's offset, device tensor [end of the data])
                if mode == "peer":
                    off2 = ps.exchange(hdr_n)
                    return off2.data_ptr(), off2[1:]
                offs = ps.exchange_offsets(hdr_n)
                return offs.data_ptr() + 8 * rank, offs[world:]
            if len(chunks) == 1:      # the fused form: placement kernel -> peer memory
                y, n = chunks[0]
                enc.encode_strip_begin(d_rgb.data_ptr(), params, first, is_last, W, n, pitch, ps.mine.data_ptr())
                p_off, end = offsets()
                enc.encode_strip_finish(ps.base, cap_file, p_off)
            else:                     # several calls per rank: local stitch with device-side running offsets, one push
                run[0] = 0
                for c, (y, n) in enumerate(chunks):
                    enc.encode_strip_begin(d_rgb.data_ptr() + y * pitch, params, first + y // 16, is_last and y + n == rows, W, n, pitch,
                                           run.data_ptr() + 8 * (c + 1))
                    enc.encode_strip_finish(local.data_ptr(), cap_local, run.data_ptr() + 8 * c)
                    run[c + 1] += run[c]
                ps.mine.copy_(run[len(chunks):])
                p_off, end = offsets()
                enc.copy_bytes_device(ps.base, cap_file, p_off, local.data_ptr(), run.data_ptr() + 8 * len(chunks))
            if mode == "peer":
                ps.complete()
            else:
                ps.fence()
            if rank == 0:
                ps.view().index_copy_(0, end + two, eoi)
            return end

    def step_nccl():
        off = encode_chunks_sync()
        return D.gather_stitch(local[:off], header, eoi, dst=0)

    peer_like = world == 1 or ps is not None
    step = step_peer if peer_like else step_nccl
    if rows:
        encode_chunks_sync()  # sizes the entropy workspace (sticky growth happens in the synchronous entry point)
    # (the sampler waits until nvidia-smi streams: started BEFORE the warm-up and the barrier -- started after the barrier
    # on rank 0 only, the other ranks' timed regions included the wait for rank 0's first exchange)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        step()
    enc.sync()
    enc.set_profiling(True)
    enc.reset_counters()
    dd.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(a.steps):
        res = step()
    ev1.record(ext)
    enc.sync()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dd.barrier()
    clocks = sampler.stop(t0, t1) if sampler else None
    tm = enc.timings()
    enc.set_profiling(False)
    # peer / single: device time of the K steps on the encoder's stream (the collectives order the ranks);
    # nccl mode has host synchronisations inside a step: wall clock
    ms_step = dd.max_ms(ev0.elapsed_time(ev1) if peer_like else (t1 - t0) * 1e3) / a.steps

    # ---- the stitched file on rank 0 --------------------------------------------------------
    final = None
    if rank == 0:
        if peer_like:
            total = int(res.cpu()[0]) + 2
            final = (own if world == 1 else ps.view())[:total]
        else:
            final = res[0]
            total = int(final.numel())
    lens = None
    if world > 1 and ps is not None and rank == 0:
        lens = ps.lengths() if mode == "peer" else [int(v) for v in ps.lens.cpu().tolist()]

    # ---- stage times of one step on this rank (event-timed, outside the timed region) -----------------------
    stage_ms = None
    if world > 1 and ps is not None and len(chunks) == 1:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        dd.barrier()
        with torch.cuda.stream(ext):
            evs[0].record()
            enc.encode_strip_begin(d_rgb.data_ptr(), params, first, is_last, W, rows, pitch, ps.mine.data_ptr())
            evs[1].record()
            if mode == "peer":
                p_off = ps.exchange(hdr_n).data_ptr()
            else:
                p_off = ps.exchange_offsets(hdr_n).data_ptr() + 8 * rank
            evs[2].record()
            enc.encode_strip_finish(ps.base, cap_file, p_off)
            if mode == "peer":
                ps.complete()
            else:
                ps.fence()
            evs[3].record()
        enc.sync()
        torch.cuda.synchronize()
        mine_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(3)]
        t = torch.tensor(mine_ms, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage_ms = dict(zip(["encode_to_sizes", "lengths_exchange", "placement_over_nvlink_and_completion"], [round(float(v), 4) for v in t.tolist()]))

    # ---- parity of the stitched file (outside the timing) ----------------------------------------------------
    parity = None
    if not a.no_parity and rank == 0:
        import oracle_lib as ol
        n_int = -(-H // 16)
        buf = final
        # RSTn markers delimit the restart intervals (stuffing guarantees FF Dx never occurs in the data)
        m = (buf[:-1] == 0xFF) & (buf[1:] >= 0xD0) & (buf[1:] <= 0xD7)
        pos = m.nonzero().flatten().cpu().numpy()
        markers_ok = len(pos) == n_int - 1 and bool(np.all(buf[torch.from_numpy(pos + 1).cuda()].cpu().numpy() == 0xD0 + (np.arange(n_int - 1) & 7)))
        ql, qc = ol.quality_tables(q)
        ks = sorted({0, n_int // world - 1, n_int // world, n_int // 2 + 3, n_int - 1} & set(range(n_int)))
        equal = markers_ok
        for k in ks if markers_ok else []:
            b0 = hdr_n if k == 0 else int(pos[k - 1]) + 2
            b1 = int(pos[k]) + 2 if k < n_int - 1 else total - 2
            got = bytes(buf[b0:b1].cpu().numpy())
            r0 = k * 16
            strip = ol.synth(seed, W, min(16, H - r0), y0=r0)
            want, _ = ol.entropy(ol.transform(strip, ol.SUB_420, ql, qc), ol.SUB_420, ri, rst_phase=k, final_rst=k != n_int - 1)
            equal = equal and got == bytes(want)
        equal = equal and bytes(buf[:hdr_n].cpu().numpy()) == bytes(ol.jfif_header(min(W, 65535), min(H, 65535), ol.SUB_420, ql, qc, ri)) \
            and bytes(buf[-2:].cpu().numpy()) == b"\xff\xd9"
        single_equal = None
        if W * H <= 16384 * 16384 and world > 1:  # fits one call: the stitched file must equal the single-GPU encode
            whole = torch.empty(H * pitch, dtype=torch.uint8, device="cuda")
            for y in range(0, H, 1024):
                enc.synth_device(seed, W, y, min(1024, H - y), pitch, whole.data_ptr() + y * pitch)
            o1 = torch.empty(cap_file, dtype=torch.uint8, device="cuda")
            t1d = torch.zeros(3, dtype=torch.int64, device="cuda")
            enc.encode_batch_device(whole.data_ptr(), 1, W, H, pitch, H * pitch, params, o1.data_ptr(), cap_file, t1d.data_ptr(),
                                    t1d.data_ptr() + 8, t1d.data_ptr() + 16)
            enc.sync()
            n1 = int(t1d[2].item())
            single_equal = n1 == total and bool(torch.equal(o1[:n1], buf))
            del whole, o1
        psnr_db = None
        try:  # the whole stitched file decoded on this GPU (one thread per restart interval) against the regenerated source
            dw, dh = min(W, 65535), min(H, 65535)  # (what SOF0 declares)
            src = torch.empty(H * pitch, dtype=torch.uint8, device="cuda")
            for y in range(0, H, 1024):
                enc.synth_device(seed, W, y, min(1024, H - y), pitch, src.data_ptr() + y * pitch)
            dec = torch.empty(dh * dw * 3, dtype=torch.uint8, device="cuda")
            enc.sync()
            torch.cuda.synchronize()
            t_dec = time.perf_counter()
            enc.decode_jfif_device(buf.data_ptr(), total, dec.data_ptr(), dw * 3, None)
            t_dec = time.perf_counter() - t_dec
            psnr_db = {"psnr_db": round(enc.psnr_device(dec.data_ptr(), dw * 3, src.data_ptr(), pitch, dw, dh)[0], 3),
                       "decoder": "jb_decode_jfif_device on rank 0 (libjpeg-exact reconstruction)", "decode_seconds": round(t_dec, 3)}
            del src, dec
        except Exception as e:
            psnr_db = {"error": str(e)[:200]}
        parity = {"restart_intervals": ks, "rst_markers_in_sequence": bool(markers_ok), "equal": bool(equal), "decoded": psnr_db,
                  "against": "oracle/jpeg_oracle.c on the same synthetic rows (strip identity: interval k == the oracle's encode of MCU row k), header and EOI",
                  "stitched_equals_single_gpu_file": single_equal}

    # ---- end to end: pinned host strip -> H2D -> encode + stitch -> D2H of the file on rank 0 -----------------
    e2e = None
    if not a.no_e2e and peer_like:
        try:
            h_rgb = torch.empty(rows * pitch, dtype=torch.uint8, pin_memory=True)
            h_rgb.copy_(d_rgb[: rows * pitch])
            h_out = torch.empty(cap_file if rank == 0 else 1, dtype=torch.uint8, pin_memory=True)
            torch.cuda.synchronize()

            def step_e2e():
                with torch.cuda.stream(ext):
                    d_rgb[: rows * pitch].copy_(h_rgb, non_blocking=True)
                end = step()
                if rank == 0:
                    enc.sync()                 # the length decides the size of the D2H: one host round trip on rank 0
                    n = int(end.cpu()[0]) + 2
                    with torch.cuda.stream(ext):
                        h_out[:n].copy_((own if world == 1 else ps.view())[:n], non_blocking=True)
                enc.sync()

            dd.barrier()  # pinning gigabytes takes the ranks different times: line them up before the first exchange
            step_e2e()
            dd.barrier()
            w0 = time.perf_counter()
            for _ in range(max(2, a.steps // 2)):
                step_e2e()
            torch.cuda.synchronize()
            w1 = time.perf_counter()
            dd.barrier()
            e_step = dd.max_ms((w1 - w0) * 1e3) / max(2, a.steps // 2)
            e2e = {"value": round(W * H / 1e6 / (e_step / 1e3), 1), "unit": "MP/s", "ms_per_step": round(e_step, 3),
                   "h2d_bytes_per_step": int(rows * pitch), "d2h_bytes_per_step": int(total) if rank == 0 else 0,
                   "api": "pinned host strip per rank -> H2D -> jb_encode_strip_begin/_finish + stitch -> D2H of the JFIF file on rank 0"}
            del h_rgb, h_out
        except Exception as e:
            e2e = {"value": None, "unit": "MP/s", "error": str(e)[:200]}

    if rank == 0:
        alg = rows * W * 6 if len(chunks) == 0 else chunks[0][1] * W * 6  # per launch: 3 B/px read + 2 B x 1.5 samples written
        roofline = transform_roofline(tm, alg, a.steps, "k_transform_tc (inside the strip call)" if a.tensor_dct else "k_transform")
        payload = (total - (lens[0] if lens else total)) if world > 1 else 0
        stitch = {"mode": mode, "bytes_crossing_nvlink_per_step": int(payload), "stage_ms_max_over_ranks": stage_ms}
        if stage_ms and payload:
            stitch["placement_GBps_into_rank0"] = round(payload / 1e9 / (stage_ms["placement_over_nvlink_and_completion"] / 1e3), 1)
            stitch["exchange_share_of_step"] = round((stage_ms["lengths_exchange"] + stage_ms["placement_over_nvlink_and_completion"])
                                                     / sum(stage_ms.values()), 4)
        print(json.dumps({
            "metric": METRIC, "value": round(W * H / 1e6 / (ms_step / 1e3), 1), "unit": "MP/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.workload}: one synthetic {W}x{H} RGB8 image, 420, q{q}, restart interval = one MCU row "
                                   f"({ri} MCUs), split into {world} RST strips, stitched on rank 0",
                       "strip_rows": rows, "calls_per_strip": len(chunks), "strip_bytes": lens, "jfif_bytes": total,
                       "bits_per_pixel": round(8.0 * total / (W * H), 4), "sof_clamped_to_65535": W > 65535 or H > 65535,
                       "l2": "strip inputs far exceed the 126 MB L2",
                       "timing": "CUDA events on the encoder's stream around K back-to-back steps incl. the stitch, max over ranks"
                                 if peer_like else "wall clock incl. gather and stitch, max over ranks"},
            "stitch": stitch, "parity_check": parity, "e2e": e2e, "roofline": roofline, "cpu_baseline": None, "clocks": clocks,
            "gpu_launches": int(tm["total_launches"])}), flush=True)
    if ps is not None:
        ps.close()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import __graft_entry__ as entry

    dd = Dist(torch)
    bind_to_gpu_numa_node(dd.local_rank)  # before any pinned allocation: host buffers land next to the GPU's PCIe root
    jb = entry.load()
    enc = jb.Encoder(dd.local_rank)  # raises if the CUDA library or a GPU is missing: no fallback

    if a.workload in STRIP_WORKLOADS:
        run_strips(a, jb, enc, torch, dd)
        dd.close()
        return
    W, H, subname, q, ri, dflt, _ = WORKLOADS[a.workload]
    F = a.frames or dflt
    r = run_batch(a, jb, enc, torch, dd, a.workload, F, a.steps, a.warmup, not a.no_e2e, not a.no_parity, ref_exact=a.ref_exact)

    # ---- the other single-GPU configurations, briefly, so that they are timed in the driver's run too -----------
    others = None
    if dd.world == 1 and a.workload == "batch1080p" and not a.no_others and not a.optimize_huffman and not a.ref_exact and not a.frames:
        others = {}
        for name, exact in (("4k444", False), ("8k", False), ("repl1080p", False), ("repl1080p", True)):
            try:
                o = run_batch(a, jb, enc, torch, dd, name, WORKLOADS[name][5], 3, 3, not a.no_e2e, not a.no_parity, ref_exact=exact,
                              sample_clocks=False)
                others[name + (" --ref-exact" if exact else "")] = {
                    "workload": workload_name(name), "value": o["value"], "ms_per_step": o["ms_per_step"],
                    "e2e": o["e2e"]["value"] if o["e2e"] else None, "roofline_frac": o["roofline"]["frac"],
                    "transform_GBps": o["roofline"]["achieved"], "transform_kernel": o["transform_kernel"],
                    "bits_per_pixel": o["bits_per_pixel"], "tie_fixups_per_step": o["tie_fixups_per_step"],
                    "step_breakdown_us": o["roofline"]["step_breakdown_us"], "parity_check": o["parity_check"], "steps": 3}
            except Exception as e:
                others[name] = {"error": str(e)[:200]}
        try:
            others["nv12_1080p"] = run_nv12(a, jb, enc, torch, dd)
        except Exception as e:
            others["nv12_1080p"] = {"error": str(e)[:200]}

    # ---- CPU baseline (rank 0, N=1 only): the reference's own code on the host cores ---------------
    cpu = None
    if dd.rank == 0 and dd.world == 1 and not a.no_cpu_baseline:
        # in a fresh process (no fork after CUDA initialisation): one bounded sample of the reference arm
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                "--warmup", "0", "--workload", a.workload], capture_output=True, text=True, timeout=600)
            cpu = json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the GPU numbers stand on their own; say why the baseline is missing
            cpu = {"value": None, "unit": "MP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}

    if dd.rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": "MP/s", "n_gpus": dd.world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.workload, F), "frames_per_gpu": F, "width": W, "height": H,
                       "subsampling": subname, "quality": q, "restart_interval": ri, "optimize_huffman": bool(a.optimize_huffman),
                       "ref_exact": bool(a.ref_exact),
                       "share_of_config_4": f"{dd.world * F} of BASELINE.json config #4's 4096 frames ({F} per GPU, weak scaling)"
                       if a.workload == "batch1080p" else None,
                       "l2": "inputs per step (%.2f GB) far exceed the 126 MB L2" % (F * W * H * 3 / 1e9),
                       "bits_per_pixel": r["bits_per_pixel"], "tie_fixups_per_step": r["tie_fixups_per_step"],
                       "transform_kernel": r["transform_kernel"], "other_workloads": others},
            "e2e": r["e2e"], "roofline": r["roofline"], "cpu_baseline": cpu, "clocks": r["clocks"],
            "parity_check": r["parity_check"], "gpu_launches": r["gpu_launches"],
        }
        print(json.dumps(line), flush=True)
    dd.close()


if __name__ == "__main__":
    main()
