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
        row0, row1, first, is_last = D.plan_strips(H, 16, 1, world)[rank]
        seg = enc.encode_strip(np.ascontiguousarray(img[row0:row1]), p, first, is_last)
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
        # ---- the same stitch over NVLink peer memory: no host round trip, the placement kernel of every rank stores
        # its strip straight into rank 0's buffer at the offset computed on the device (dist.PeerStitch)
        want = enc.encode_jfif(img, p) if rank == 0 else None
        ext = torch.cuda.ExternalStream(enc.stream())
        d_img = torch.from_numpy(np.ascontiguousarray(img[row0:row1])).cuda()
        ps = D.PeerStitch(enc, cap=W * H + 65536, dst=0)
        hdr_n = hdr.numel()
        for rep in range(4):  # repeated: the pinned staging ring, the workspace and both flag parities are reused
            with torch.cuda.stream(ext):
                enc.encode_strip_begin(d_img.data_ptr(), p, first, is_last, W, row1 - row0, W * 3, ps.mine.data_ptr())
                if rep == 3:  # the exchange through NCCL instead (all-gather + cumsum, all-reduce as the fence)
                    offs = ps.exchange_offsets(hdr_n)
                    enc.encode_strip_finish(ps.base, ps.cap, offs[rank:].data_ptr())
                    ps.fence()
                    end = offs[world:]
                else:         # lengths and completion flags as stores / polls on rank 0's memory over NVLink
                    off2 = ps.exchange(hdr_n)
                    enc.encode_strip_finish(ps.base, ps.cap, off2.data_ptr())
                    ps.complete()
                    end = off2[1:]
            enc.sync()
            if rank == 0:
                assert rep == 3 or ps.lengths() == lengths
                total = int(end[0].item())
                out = ps.view()
                out[:hdr_n] = hdr
                out[total: total + 2] = eoi
                torch.cuda.synchronize()
                ok = ok and bytes(out[: total + 2].cpu().numpy()) == want
        # ---- a rank that codes its strip in two calls: local stitch with device-side running offsets, then one push
        half = ((row1 - row0) // 32) * 16
        if half and row1 - row0 - half:
            local = torch.empty(W * H, dtype=torch.uint8, device="cuda")
            run = torch.zeros(3, dtype=torch.int64, device="cuda")
            with torch.cuda.stream(ext):
                enc.encode_strip_begin(d_img.data_ptr(), p, first, False, W, half, W * 3, ps.mine.data_ptr())
                enc.encode_strip_finish(local.data_ptr(), local.numel(), run[0:].data_ptr())
                torch.add(run[0], ps.mine[0], out=run[1])
                enc.encode_strip_begin(d_img.data_ptr() + half * W * 3, p, first + half // 16, is_last, W, row1 - row0 - half, W * 3,
                                       ps.mine.data_ptr())
                enc.encode_strip_finish(local.data_ptr(), local.numel(), run[1:].data_ptr())
                torch.add(run[1], ps.mine[0], out=run[2])
                ps.mine.copy_(run[2:3])
                off2 = ps.exchange(hdr_n)
                enc.copy_bytes_device(ps.base, ps.cap, off2.data_ptr(), local.data_ptr(), run[2:].data_ptr())
                ps.complete()
            enc.sync()
            if rank == 0:
                total = int(off2[1].item())
                out = ps.view()
                out[total: total + 2] = eoi
                torch.cuda.synchronize()
                ok = ok and bytes(out[: total + 2].cpu().numpy()) == want
        # ---- fewer restart intervals than ranks: the strip that ends the image is rank 0's; rank 1 has nothing to code
        small = enc.synth(0x51, 256, 16)
        p1 = jb.make_params(jb.SUB_420, quality=75, restart_interval=16)
        r0, r1, f1, last1 = D.plan_strips(16, 16, 1, world)[rank]
        seg1 = enc.encode_strip(np.ascontiguousarray(small[r0:r1]), p1, f1, last1) if r1 > r0 else np.zeros(0, np.uint8)
        hdr1 = torch.from_numpy(np.frombuffer(enc.write_header(p1, 256, 16), np.uint8).copy()).cuda()
        whole1, len1 = D.gather_stitch(torch.from_numpy(seg1).cuda(), hdr1, eoi, dst=0)
        if rank == 0:
            ok = ok and len1[1] == 0 and bytes(whole1.cpu().numpy()) == enc.encode_jfif(small, p1)
        ps.close()
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
