#!/usr/bin/env python
"""bench.py -- throughput of the B200 JPEG encode path (BASELINE.json metric:
"encode megapixels/sec (4:2:0, q75) at 1/2/4/8 B200; fused-kernel HBM GB/s").

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic frames.  Default workload:
BASELINE.json config #4 sharded by image, weak scaling: every GPU encodes --frames
1920x1080 frames, 4:2:0, q75 (at 8 GPUs x 512 frames this is exactly the 4096-frame batch).
`value` is device-resident (RGB already in HBM, JFIF bytes left in HBM); `e2e` is the same
work through the public host API jb_encode_batch (pinned host RGB in, pinned host JFIF out,
H2D/D2H inside the timed region).  One JSON line is printed by rank 0.
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
    # name: (W, H, subsampling, quality, restart interval in MCUs, default frames per GPU)
    "batch1080p": (1920, 1080, "420", 75, 0, 512),
    "8k": (7680, 4320, "420", 75, 480, 8),
    "4k444": (3840, 2160, "444", 90, 0, 32),
    # the reference's own mode (replicated 4:2:0 coded as 4:4:4, its q50 tables) on the batch frames
    "repl1080p": (1920, 1080, "repl420", 50, 0, 256),
    # one image split into strips of whole restart intervals across the GPUs (config #5; strong scaling)
    "gigapixel": (65536, 65536, "420", 75, 4096, 1),
    "strips16k": (16384, 16384, "420", 75, 1024, 1),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="batch1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--optimize-huffman", action="store_true",
                    help="JB_FLAG_OPTIMIZE_HUFFMAN: per-call optimal Huffman tables (two passes; not the headline configuration)")
    ap.add_argument("--ref-exact", action="store_true",
                    help="the reference as written: JB_FLAG_REF_INPLACE_DCT | REF_TYPO_TABLES | REF_ALWAYS_EOB (use with --workload repl1080p)")
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
    W, H, sub, q, ri, _ = WORKLOADS[a.workload]
    cores = os.cpu_count() or 1
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
        "config": {"workload": workload_name(a, 0), "note": "CPU arm: rank 0 only, host cores, no GPU"},
        "cpu_baseline": {"value": round(value, 4), "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 4), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(a, frames):
    W, H, sub, q, ri, dflt = WORKLOADS[a.workload]
    f = frames or a.frames or dflt
    return (f"{a.workload}: {f} synthetic {W}x{H} RGB8 frames per GPU, {sub}, q{q}, "
            f"restart interval {ri} MCUs, sharded by image")


# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        try:
            sm = [float(r[0]) for r in rows]
            mx = max(float(r[1]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                    "power_w_max": max(float(r[2]) for r in rows)}
        except Exception:
            return None


def run_strips(a, jb, enc, torch, dist, rank, world):
    """One large image, split into horizontal strips of whole restart intervals (one MCU row each);
    rank r encodes its strip from HBM, then the only exchange: all-gather of the strip lengths and
    a gather of the compressed bytes to rank 0 over NCCL, where header + strips + EOI are stitched."""
    import importlib
    D = importlib.import_module("jpegb200.dist")
    W, H, subname, q, ri, _ = WORKLOADS[a.workload]
    params = jb.make_params(jb.SUB_420, quality=q, restart_interval=ri,
                            flags=jb.FLAG_CLAMP_SOF | (0 if a.tensor_dct else jb.FLAG_FMA_DCT))
    row0, row1, first = D.plan_strips(H, 16, 1, world)[rank]
    rows, pitch = row1 - row0, W * 3
    chunk = 8192  # rows per jb_encode_strip call (keeps every call below 2^26 blocks)
    d_rgb = torch.empty(rows * pitch, dtype=torch.uint8, device="cuda")
    for y in range(0, rows, 1024):
        enc.synth_device(0x65536, W, row0 + y, min(1024, rows - y), pitch, d_rgb.data_ptr() + y * pitch)
    enc.sync()
    cap = rows * W // 2 + (1 << 20)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    header = torch.frombuffer(bytearray(enc.write_header(params, W, H)), dtype=torch.uint8).cuda()
    eoi = torch.tensor([0xFF, 0xD9], dtype=torch.uint8, device="cuda")
    single = torch.empty(header.numel() + cap + 2, dtype=torch.uint8, device="cuda")
    single[: header.numel()] = header

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        off = 0
        for y in range(0, rows, chunk):
            n_rows = min(chunk, rows - y)
            last = rank == world - 1 and y + n_rows == rows
            off += enc.encode_strip(d_rgb.data_ptr() + y * pitch, params, first + y // 16, last, W=W, rows=n_rows, pitch=pitch,
                                    device_io=True, out=d_out.data_ptr() + off, cap=cap - off)
        if world > 1:  # the one exchange step: strips land at their final offsets on rank 0
            return D.gather_stitch(d_out[:off], header, eoi, dst=0)
        nh = header.numel()  # one GPU: header + strip + EOI into a buffer allocated once (no allocator traffic per step)
        single[nh: nh + off] = d_out[:off]
        single[nh + off: nh + off + 2] = eoi
        return single[: nh + off + 2], [off]

    for _ in range(max(a.warmup, 5)):  # NCCL channels and the caching allocator settle over the first few steps
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        final, lengths = step()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    t_ms = torch.tensor([(t1 - t0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / a.steps
    if rank == 0:
        total = int(final.numel())
        print(json.dumps({
            "metric": METRIC, "value": round(W * H / 1e6 / (ms_step / 1e3), 1), "unit": "MP/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 5), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.workload}: one synthetic {W}x{H} RGB8 image, 420, q{q}, restart interval = one MCU row "
                                   f"({ri} MCUs), split into {world} RST strips, NCCL gather + stitch on rank 0",
                       "strip_bytes": lengths, "jfif_bytes": total, "bits_per_pixel": round(8.0 * total / (W * H), 4),
                       "l2": "strip inputs far exceed the 126 MB L2", "timing": "wall clock incl. gather and stitch, max over ranks"},
            "e2e": None, "roofline": None, "cpu_baseline": None, "clocks": None,
            "gpu_launches": int(enc.timings()["total_launches"])}), flush=True)


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


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: host buffers land next to the GPU's PCIe root
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    jb = entry.load()
    enc = jb.Encoder(local_rank)  # raises if the CUDA library or a GPU is missing: no fallback

    W, H, subname, q, ri, dflt = WORKLOADS[a.workload]
    if a.workload in ("gigapixel", "strips16k"):
        run_strips(a, jb, enc, torch, dist, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return
    F = a.frames or dflt
    sub = {"420": jb.SUB_420, "444": jb.SUB_444, "repl420": jb.SUB_REPL420}[subname]
    params = jb.make_params(sub, quality=q, restart_interval=ri,
                            flags=(0 if a.tensor_dct else jb.FLAG_FMA_DCT) | (jb.FLAG_OPTIMIZE_HUFFMAN if a.optimize_huffman else 0)
                            | ((jb.FLAG_REF_INPLACE_DCT | jb.FLAG_REF_TYPO_TABLES | jb.FLAG_REF_ALWAYS_EOB) if a.ref_exact else 0))
    pitch, fstride = W * 3, W * H * 3
    px_per_step = W * H * F  # per GPU

    # inputs: generated on the device (identical to the oracle's generator), copied once to pinned host
    d_rgb = torch.empty(F * fstride, dtype=torch.uint8, device="cuda")
    for f in range(F):
        enc.synth_device(0xF000 + rank * F + f, W, 0, H, pitch, d_rgb.data_ptr() + f * fstride)
    enc.sync()
    cap = F * (W * H // (1 if a.ref_exact else 2) + 4096)  # the as-written transform yields ~4 bits/px
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_tab = torch.zeros(2 * F + 1, dtype=torch.int64, device="cuda")
    ext = torch.cuda.ExternalStream(enc.stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        enc.encode_batch_device(d_rgb.data_ptr(), F, W, H, pitch, fstride, params, d_out.data_ptr(), cap,
                                d_tab.data_ptr(), d_tab.data_ptr() + 8 * F, d_tab.data_ptr() + 16 * F)

    # ---- device-resident timed region ------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        step_device()
    enc.sync()
    enc.set_profiling(True)
    enc.reset_counters()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(a.steps):
        step_device()
    ev1.record(ext)
    enc.sync()
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    tm = enc.timings()
    enc.set_profiling(False)
    total_bytes = int(d_tab[2 * F].item())
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / a.steps
    value = world * px_per_step / 1e6 / (ms_step / 1e3)

    # ---- roofline of the fused transform kernel (CUDA events on its own stream, timed region) ----
    peak, peak_src = peaks()
    g_mcu = 16 if sub == jb.SUB_420 else 8
    padded = (-(-W // g_mcu) * g_mcu) * (-(-H // g_mcu) * g_mcu)
    samples_per_px = 1.5 if sub == jb.SUB_420 else 3.0
    alg_bytes = F * (3 * W * H + 2 * samples_per_px * padded)  # read RGB8 + write int16 coefficients
    k_us = tm["transform_us"] / max(tm["transform_launches"], 1)
    achieved = alg_bytes / (k_us * 1e-6) / 1e9
    use_tc = bool(a.tensor_dct)
    kname = ("k_transform_tc" if sub == jb.SUB_420 else "k_transform_tc3") if use_tc else "k_transform"
    pname = "k_transform_tc3_repl" if (use_tc and sub == jb.SUB_REPL420) else kname  # key in the ncu summary
    traffic = None
    try:  # per-launch DRAM bytes from the committed ncu capture of the same command, if present
        with open(os.path.join(ROOT, "profiles", "r01_transform_ncu_summary.json")) as f:
            prof = json.load(f)
        k = prof.get("kernels", {}).get(pname, {})
        if k.get("workload", prof.get("workload")) == a.workload and k.get("frames", prof.get("frames")) == F:
            traffic = k.get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": kname + (" (fused CSC+subsample+shift, tcgen05 FDCT+quant+zigzag)" if use_tc else
                                   " (fused CSC+subsample+shift+FDCT+quant+zigzag, CUDA cores)"), "bound": "hbm",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "frac_of_nominal_8000": round(achieved / 8000.0, 4), "algorithmic_bytes_per_launch": int(alg_bytes),
                "kernel_us_per_launch": round(k_us, 2),
                "step_breakdown_us": {"transform": round(tm["transform_us"] / a.steps, 1),
                                      "edge_mcus": round(tm["edge_us"] / a.steps, 1),
                                      "tie_fixup": round(tm["fixup_us"] / a.steps, 1),
                                      "entropy": round(tm["entropy_us"] / a.steps, 1)}}
    launches_device = int(tm["total_launches"])

    # ---- end to end through the public host API: pinned host in, pinned host out ----------------
    e2e = None
    if not a.no_e2e:
        h_rgb = jb.pinned_empty((F, H, W, 3))
        enc.d2h(h_rgb, d_rgb.data_ptr())
        h_out = jb.pinned_empty((cap,))
        offs = np.zeros(F, np.uint64)
        sizes = np.zeros(F, np.uint64)

        def step_host():
            enc.encode_batch_ptr(h_rgb.ctypes.data, F, W, H, pitch, fstride, params, h_out.ctypes.data, cap, offs, sizes)

        for _ in range(max(a.warmup, 3)):
            step_host()
        barrier()
        w0 = time.perf_counter()
        for _ in range(a.steps):
            step_host()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        barrier()
        e_ms = torch.tensor([(w1 - w0) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        e_step = float(e_ms.item()) / a.steps
        out_bytes = int(sizes.sum())
        # host path == device path (with per-call optimal tables the 96 MB groups of the host path get their own tables)
        assert a.optimize_huffman or out_bytes == total_bytes, (out_bytes, total_bytes)
        e2e = {"value": round(world * px_per_step / 1e6 / (e_step / 1e3), 1), "unit": "MP/s",
               "ms_per_step": round(e_step, 3), "h2d_bytes_per_step": int(F * fstride),
               "d2h_bytes_per_step": int(out_bytes + 16 * F + 48),
               "api": "jb_encode_batch (pinned host RGB -> pinned host JFIF, 3 streams, 96 MB groups)"}

    # ---- CPU baseline (rank 0, N=1 only): the reference's own code on the host cores ---------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        # in a fresh process (no fork after CUDA initialisation): one bounded sample of the reference arm
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                "--warmup", "0", "--workload", a.workload], capture_output=True, text=True, timeout=600)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the GPU numbers stand on their own; say why the baseline is missing
            cpu = {"value": None, "unit": "MP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, F), "frames_per_gpu": F, "width": W, "height": H,
                       "subsampling": subname, "quality": q, "restart_interval": ri, "optimize_huffman": bool(a.optimize_huffman), "ref_exact": bool(a.ref_exact),
                       "l2": "inputs per step (%.2f GB) far exceed the 126 MB L2" % (F * fstride / 1e9),
                       "bits_per_pixel": round(8.0 * total_bytes / px_per_step, 4),
                       "tie_fixups_per_step": int(tm["tie_fixups"]),
                       "transform_kernel": kname + (" (tcgen05)" if use_tc else " (FMA pipe)")},
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "gpu_launches": launches_device,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
