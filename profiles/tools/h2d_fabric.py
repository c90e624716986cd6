#!/usr/bin/env python
"""Raw pinned-host -> device copy bandwidth with 1, 2, 4, ... N GPUs copying at the same time (one process per GPU,
launched under torchrun like bench.py).  No library code involved: plain cudaMemcpyAsync from cudaHostAlloc'd memory
(torch pinned tensors), so that bench.py's end-to-end H2D rate can be compared with what the host's PCIe fabric gives.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      profiles/tools/h2d_fabric.py > gpurun_out/h2d_fabric.jsonl

One JSON line per (concurrency, streams-per-GPU, direction): aggregate and per-GPU GB/s, max over ranks of the wall time.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 1 << 30
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank + 1)  # first touch by this process
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(4)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(active, n_streams, h2d, reps=8):
        part = nbytes // n_streams
        for it in range(2 + reps):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            if active:
                for s in range(n_streams):
                    with torch.cuda.stream(streams[s]):
                        a, b = (dev, host) if h2d else (host, dev)
                        a[s * part:(s + 1) * part].copy_(b[s * part:(s + 1) * part], non_blocking=True)
                torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / reps

    conc = [n for n in (1, 2, 4, 8, 16) if n <= world]
    for n in conc:
        for n_streams in (1, 3):
            for h2d in (True, False):
                dt = run(rank < n, n_streams, h2d)
                if rank == 0:
                    print(json.dumps({"gpus_copying": n, "streams_per_gpu": n_streams, "direction": "h2d" if h2d else "d2h",
                                      "bytes_per_gpu": nbytes, "seconds_max_over_ranks": round(dt, 5),
                                      "GBps_per_gpu": round(nbytes / 1e9 / dt, 2), "GBps_aggregate": round(n * nbytes / 1e9 / dt, 2),
                                      "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
