"""2-GPU timing breakdown of the strip workload: encode / length all-gather / payload gather / stitch."""
import os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import __graft_entry__ as g
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
jb = g.load(); D = importlib.import_module("jpegb200.dist")
enc = jb.Encoder(rank)
W = H = 16384
p = jb.make_params(jb.SUB_420, quality=75, restart_interval=1024, flags=jb.FLAG_CLAMP_SOF)
row0, row1, first = D.plan_strips(H, 16, 1, world)[rank]
rows, pitch = row1 - row0, W * 3
d = torch.empty(rows * pitch, dtype=torch.uint8, device="cuda")
for y in range(0, rows, 1024):
    enc.synth_device(0x65536, W, row0 + y, min(1024, rows - y), pitch, d.data_ptr() + y * pitch)
enc.sync()
cap = rows * W // 2 + (1 << 20)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(5):
    t0 = T()
    n = enc.encode_strip(d.data_ptr(), p, first, rank == world - 1, W=W, rows=rows, pitch=pitch, device_io=True, out=out.data_ptr(), cap=cap)
    t1 = T()
    nt = torch.tensor([n], dtype=torch.int64, device="cuda")
    sizes = [torch.zeros_like(nt) for _ in range(world)]
    dist.all_gather(sizes, nt)
    lengths = [int(s.item()) for s in sizes]
    t2 = T()
    parts, _ = D.gather_bytes(out[:n], dst=0)
    t3 = T()
    if rank == 0:
        final = torch.cat(list(parts))
    t4 = T()
    print(f"rank {rank} it {it}: encode {1e3*(t1-t0):.2f} ms, len all-gather {1e3*(t2-t1):.2f}, gather_bytes {1e3*(t3-t2):.2f}, cat {1e3*(t4-t3):.2f}", flush=True)
dist.destroy_process_group()
