"""Two-GPU test (skipped with fewer devices): one image encoded as RST strips on two B200s,
compressed strips gathered to rank 0 over NCCL and stitched; the result must equal the
single-GPU encode of the whole image byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import __graft_entry__ as entry
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        jb = entry.load()
        D = importlib.import_module("jpegb200.dist")
        enc = jb.Encoder(rank)
        W, H = 2048, 1000  # the last strip ends in a partial MCU row (mirror padding)
        img = enc.synth(0xABCD, W, H)
        p = jb.make_params(jb.SUB_420, quality=75, restart_interval=W // 16)
        row0, row1, first = D.plan_strips(H, 16, 1, world)[rank]
        seg = enc.encode_strip(np.ascontiguousarray(img[row0:row1]), p, first, rank == world - 1)
        parts, lengths = D.gather_bytes(torch.from_numpy(seg).cuda(), dst=0)
        ok = True
        if rank == 0:
            got = D.stitch(enc.write_header(p, W, H), [x.cpu().numpy() for x in parts])
            ok = got == enc.encode_jfif(img, p)
        hdr = torch.from_numpy(np.frombuffer(enc.write_header(p, W, H), np.uint8).copy()).cuda()
        eoi = torch.tensor([0xFF, 0xD9], dtype=torch.uint8, device="cuda")
        whole, lengths2 = D.gather_stitch(torch.from_numpy(seg).cuda(), hdr, eoi, dst=0)  # received in place
        if rank == 0:
            ok = ok and lengths2 == lengths and bytes(whole.cpu().numpy()) == enc.encode_jfif(img, p)
        q.put((rank, bool(ok), lengths))
        enc.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_strip_stitch_over_nccl():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(2)]
    assert all(ok for _, ok, _ in res)
