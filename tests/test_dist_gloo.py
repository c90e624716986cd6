"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in jpeg-encoder-opencl_b200/dist.py:
frame sharding, strip planning, the gather of variable-length compressed strips and the final
stitch.  The per-rank "encode" is played by the oracle here (no GPU in this container); on the
GPU box the same functions carry jb_encode_strip / jb_encode_batch outputs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _load_dist():
    import __graft_entry__ as entry
    entry.load()
    import importlib
    return importlib.import_module("jpegb200.dist")


def _worker(rank, world, port, q):
    import oracle_lib as ol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D = _load_dist()
        # ---- one image split into RST strips across ranks, stitched on rank 0 -----------------
        W, H = 176, 200
        img = ol.synth(77, W, H)
        ql, qc = ol.quality_tables(75)
        mcux = W // 16
        plan = D.plan_strips(H, 16, 1, world)
        row0, row1, first, is_last = plan[rank]
        coef = ol.transform(img[row0:row1], ol.SUB_420, ql, qc)
        seg, _ = ol.entropy(coef, ol.SUB_420, mcux, rst_phase=first, final_rst=not is_last)
        parts, lengths = D.gather_bytes(torch.from_numpy(seg.copy()), dst=0)
        assert lengths[rank] == len(seg)
        ok = True
        if rank == 0:
            jf = D.stitch(ol.jfif_header(W, H, ol.SUB_420, ql, qc, mcux), [p.numpy() for p in parts])
            ok = jf == ol.encode_jfif(img, ol.SUB_420, ql, qc, mcux)
        # the same stitch as one exchange step: strips received at their final offsets on rank 0
        hdr = torch.from_numpy(np.frombuffer(ol.jfif_header(W, H, ol.SUB_420, ql, qc, mcux), np.uint8).copy())
        whole, lengths2 = D.gather_stitch(torch.from_numpy(seg.copy()), hdr, torch.tensor([0xFF, 0xD9], dtype=torch.uint8), dst=0)
        assert lengths2 == lengths
        if rank == 0:
            ok = ok and bytes(whole.numpy()) == ol.encode_jfif(img, ol.SUB_420, ql, qc, mcux)
        else:
            assert whole is None
        # ---- fewer restart intervals than ranks: the last rank's strip is empty and sends nothing ----
        small = ol.synth(78, 64, 16)
        plan1 = D.plan_strips(16, 16, 1, world)
        r0, r1, f1, last1 = plan1[rank]
        assert last1 == (rank == 0)  # the strip that ends the image is rank 0's, not rank world-1's
        if r1 > r0:
            c1 = ol.transform(small[r0:r1], ol.SUB_420, ql, qc)
            seg1, _ = ol.entropy(c1, ol.SUB_420, 4, rst_phase=f1, final_rst=not last1)
        else:
            seg1 = np.zeros(0, np.uint8)
        hdr1 = torch.from_numpy(np.frombuffer(ol.jfif_header(64, 16, ol.SUB_420, ql, qc, 4), np.uint8).copy())
        whole1, len1 = D.gather_stitch(torch.from_numpy(seg1.copy()), hdr1, torch.tensor([0xFF, 0xD9], dtype=torch.uint8), dst=0)
        assert len1[1:] == [0] * (world - 1)
        if rank == 0:
            ok = ok and bytes(whole1.numpy()) == ol.encode_jfif(small, ol.SUB_420, ql, qc, 4)
        # ---- a batch sharded by image: every rank encodes its range, sizes all-gathered --------
        N = 5
        lo, hi = D.shard_range(N, world, rank)
        mine = [ol.encode_jfif(ol.synth(100 + f, 48, 32), ol.SUB_420, ql, qc) for f in range(lo, hi)]
        blob = np.frombuffer(b"".join(mine), np.uint8)
        parts, lengths = D.gather_bytes(torch.from_numpy(blob.copy()), dst=0)
        if rank == 0:
            whole = b"".join(bytes(p.numpy()) for p in parts)
            want = b"".join(ol.encode_jfif(ol.synth(100 + f, 48, 32), ol.SUB_420, ql, qc) for f in range(N))
            ok = ok and whole == want
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_strip_stitch_and_batch_shard_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {0: True, 1: True}


def test_shard_and_strip_plans():
    D = _load_dist()
    for n, w in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        spans = [D.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    # config #5: 65536 rows, 16-px MCUs, restart interval = one MCU row, 8 GPUs -> 512 MCU rows each
    plan = D.plan_strips(65536, 16, 1, 8)
    assert [p[1] - p[0] for p in plan] == [8192] * 8 and [p[2] for p in plan] == [512 * r for r in range(8)]
    # a height that is not a multiple of the MCU: the last strip carries the partial MCU row
    plan = D.plan_strips(1080, 16, 1, 4)
    assert plan[-1][1] == 1080 and sum(p[1] - p[0] for p in plan) == 1080
    assert all(p[0] % 16 == 0 for p in plan)
    assert [p[3] for p in plan] == [False, False, False, True]
    # fewer restart intervals than ranks: empty strips at the end, and the LAST NON-EMPTY strip ends the image
    plan = D.plan_strips(40, 16, 1, 8)  # 3 MCU rows on 8 ranks
    assert [p[1] - p[0] for p in plan] == [16, 16, 8, 0, 0, 0, 0, 0]
    assert [p[3] for p in plan] == [False, False, True, False, False, False, False, False]
