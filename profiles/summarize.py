"""Turn the round's ncu captures (gpurun_out/) into the committed summaries under profiles/.

  python profiles/summarize.py r01
reads  gpurun_out/<r>_launches.csv        (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/<r>_prof_tc.ncu-rep     (ncu --set full of the step's main kernels, default build)
       gpurun_out/<r>_prof_full.ncu-rep   (ncu --set full of the CUDA-core transform kernel, --tensor-dct 0)
       gpurun_out/<r>_prof_tc3.ncu-rep    (ncu --set full of the 8x8-MCU tcgen05 transform kernel, --workload 4k444)
writes profiles/<r>_launches.md, profiles/<r>_kernels.md, profiles/<r>_transform_ncu_summary.json
"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "gpurun_out")

def launches():
    rows = [r for r in csv.reader(open(os.path.join(OUT, f"{R}_launches.csv"))) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    vals = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if len(r) > vi]
    starts = [i for i, (k, _) in enumerate(vals) if "k_transform" in k]
    step = vals[starts[-2]:starts[-1]]  # one complete step
    tot = sum(v for _, v in step)
    lines = ["# Launch list of one bench step (512 x 1920x1080, 4:2:0, q75), ncu gpu__time_duration.sum",
             "", "Cold-cache, serialised launches (compare shares, not absolutes).", "",
             "| kernel | us | share |", "|---|---:|---:|"]
    for k, v in step:
        lines.append(f"| `{k.split('(')[0]}` | {v/1000:.1f} | {100*v/tot:.1f} % |")
    lines.append(f"| **total** | {tot/1000:.1f} | |")
    open(os.path.join(ROOT, "profiles", f"{R}_launches.md"), "w").write("\n".join(lines) + "\n")
    return step

def raw(rep):
    out = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput % of peak"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("launch__registers_per_thread", "regs/thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor (hmma) pipe %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]

def kernels():
    md = [f"# ncu --set full summaries ({R})", "",
          "Workload: `python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline` (512 x 1920x1080, 4:2:0, q75);",
          "`_prof_full`: the same with `--tensor-dct 0` (CUDA-core transform); `_prof_tc3`: `--workload 4k444` (32 x 3840x2160, 4:4:4, q90);",
          "`_prof_tc3r`: `--workload repl1080p` (256 x 1920x1080, replicated 4:2:0, q50: K = 16 chroma contraction).",
          "`_prof_tma`: the default workload with `--tma` (JB_FLAG_TMA: TMA-staged tiles, `k_transform_tma`).",
          "`_prof_nv12`: the NV12-style device input of the default run's `other_workloads` (256 x 1920x1080 as Y + CbCr planes, `k_transform_tc_nv12`).",
          "Captured by `profiles/capture.sh` (each ncu pass after a plain run of the same command that exited 0).",
          "Numbers taken under the profiler are diagnostics only; bench values come from CUDA events.", ""]
    summary = {}
    for rep in (f"{R}_prof_tc.ncu-rep", f"{R}_prof_tma.ncu-rep", f"{R}_prof_full.ncu-rep", f"{R}_prof_tc3.ncu-rep", f"{R}_prof_tc3r.ncu-rep", f"{R}_prof_nv12.ncu-rep"):
        if not os.path.exists(os.path.join(OUT, rep)):
            continue
        hdr, units, rows = raw(rep)
        ki = hdr.index("Kernel Name")
        for r in rows:
            name = r[ki].split("(")[0]
            md += [f"## `{name}`  ({rep})", "", "| metric | value |", "|---|---:|"]
            rec = {}
            for key, label in WANT:
                if key in hdr:
                    v = r[hdr.index(key)]; u = units[hdr.index(key)]
                    md.append(f"| {label} | {v} {u} |")
                    rec[label] = (v, u)
            md.append("")
            summary[name] = rec
    # overview: what bounds each kernel.  All are HBM-bound by construction (byte / integer work); "frac" = DRAM bytes moved /
    # duration / 6544 GB/s (MEASURED_PEAKS.json); the limiter is read off the issue utilisation and the top stall.
    def num(rec, label):
        return float(rec[label][0].replace(",", "")) if label in rec else 0.0
    over = ["## Overview (one 1.06 Gpx step of the default workload unless noted; durations under ncu)", "",
            "| kernel | duration | DRAM read + write | GB/s | frac of 6 544 | issue active | warps active | top stall | limiter |",
            "|---|---:|---:|---:|---:|---:|---:|---|---|"]
    for name, rec in summary.items():
        dur = num(rec, "duration")
        unit = rec["duration"][1]
        us = dur * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)
        by = to_bytes(*rec["dram read"]) + to_bytes(*rec["dram write"])
        gbs = by / (us * 1e-6) / 1e9 if us else 0
        stalls = {k: num(rec, k) for k in rec if k.startswith("stall ")}
        top = max(stalls, key=stalls.get) if stalls else ""
        issue = num(rec, "issue active %")
        lim = "instruction issue" if issue >= 55 else ("latency: " + top.replace("stall ", "") if issue < 50 else "issue + " + top.replace("stall ", ""))
        over.append(f"| `{name}` | {us:.0f} us | {by/1e9:.2f} GB | {gbs:.0f} | {gbs/6544:.2f} | {issue:.0f} % | {num(rec, 'warps active %'):.0f} % | {top} {stalls.get(top, 0):.1f} | {lim} |")
    md = md[:10] + over + [""] + md[10:]
    open(os.path.join(ROOT, "profiles", f"{R}_kernels.md"), "w").write("\n".join(md) + "\n")
    return summary

def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

if __name__ == "__main__":
    step = launches()
    summ = kernels()
    ks = {}
    for key, pred in (("k_transform_tc", lambda k: "k_transform_tc<" in k),
                      ("k_transform_tc3", lambda k: "k_transform_tc3<0" in k),
                      ("k_transform_tc3_repl", lambda k: "k_transform_tc3<1" in k),
                      ("k_transform_tc_nv12", lambda k: "k_transform_tc_nv12" in k),
                      ("k_transform", lambda k: "k_transform<" in k)):
        t = next((v for k, v in summ.items() if pred(k)), None)
        if t:
            ks[key] = {"dram_bytes_per_launch": int(to_bytes(*t["dram read"]) + to_bytes(*t["dram write"])),
                       "dram_read": t["dram read"], "dram_write": t["dram write"], "duration_under_ncu": t["duration"],
                       "workload": {"k_transform_tc3": "4k444", "k_transform_tc3_repl": "repl1080p", "k_transform_tc_nv12": "nv12_1080p"}.get(key, "batch1080p"),
                       "frames": {"k_transform_tc3": 32, "k_transform_tc3_repl": 256, "k_transform_tc_nv12": 256}.get(key, 512)}
    traffic = {k: v["dram_bytes_per_launch"] for k, v in ks.items()}
    json.dump({"round": R, "workload": "batch1080p", "frames": 512, "kernels": ks,
               "note": "k_transform_tc3 was captured on the 4k444 workload (32 frames of 3840x2160, 4:4:4, q90)",
               "source": f"gpurun_out/{R}_prof_tc.ncu-rep, {R}_prof_full.ncu-rep, {R}_prof_tc3.ncu-rep (ncu --set full --clock-control none)"},
              open(os.path.join(ROOT, "profiles", f"{R}_transform_ncu_summary.json"), "w"), indent=1)
    print("ok", traffic)
