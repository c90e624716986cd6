"""Multi-GPU sharding of the encode path (one process per GPU, torch.distributed).

The path shards without any data-path collective:
  * a batch of frames is split by image -- rank r encodes a contiguous range of frames;
  * one large image is split into horizontal strips that are whole numbers of restart
    intervals -- every strip is an independent, byte-aligned entropy segment (DC predictors
    reset, RSTn numbering known a priori), so the only exchange is the final stitch:
    an all-gather of the strip lengths and one gather of the compressed bytes to rank 0
    (NCCL over NVLink on the GPUs; gloo in the CPU tests).
The reference itself is single-device (src/OpenCLProject_JpegEncoder.cpp:280-286).
"""
import numpy as np


def shard_range(n_items, world, rank):
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def plan_strips(height, mcu_px, restart_mcu_rows, world):
    """Split `height` pixel rows into `world` strips of whole restart intervals.

    restart_mcu_rows = MCU rows per restart interval.  Returns a list of
    (row0, row1, first_interval) per rank; ranks beyond the available intervals get empty strips.
    """
    mcu_rows = -(-height // mcu_px)
    n_int = -(-mcu_rows // restart_mcu_rows)
    out = []
    for r in range(world):
        i0, i1 = shard_range(n_int, world, r)
        row0 = min(i0 * restart_mcu_rows * mcu_px, height)
        row1 = min(i1 * restart_mcu_rows * mcu_px, height)
        out.append((row0, row1, i0))
    return out


def gather_bytes(payload, dst=0, group=None):
    """Gather variable-length byte strings to rank `dst`.

    payload: 1-D uint8 torch tensor (CPU for gloo, CUDA for nccl).  Returns (list of tensors on dst,
    lengths) -- on other ranks the list is None.  One all-gather of lengths + one gather of payloads.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([payload.numel()], dtype=torch.int64, device=payload.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    lengths = [int(s.item()) for s in sizes]
    cap = max(max(lengths), 1)
    padded = torch.zeros(cap, dtype=torch.uint8, device=payload.device)
    padded[: payload.numel()] = payload
    if rank == dst:
        bufs = [torch.empty(cap, dtype=torch.uint8, device=payload.device) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst, group=group)
        return [b[:l] for b, l in zip(bufs, lengths)], lengths
    dist.gather(padded, None, dst=dst, group=group)
    return None, lengths


def gather_stitch(payload, header=None, trailer=None, dst=0, group=None):
    """The stitch as one exchange step: every rank's byte string lands at its final offset of the output on
    rank `dst` -- header + payload_0 + ... + payload_{n-1} + trailer -- without padding or a second copy.

    payload/header/trailer: 1-D uint8 torch tensors on the collective's device (header/trailer only matter on
    `dst`).  One all-gather of the lengths, then point-to-point sends (NCCL over NVLink on GPUs, gloo on CPU)
    received in place.  Returns (stitched tensor on dst / None elsewhere, lengths).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([payload.numel()], dtype=torch.int64, device=payload.device)
    sizes = torch.zeros(world, dtype=torch.int64, device=payload.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    lengths = [int(v) for v in sizes.tolist()]
    if rank != dst:
        if lengths[rank]:
            dist.send(payload.contiguous(), dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        return None, lengths
    nh = header.numel() if header is not None else 0
    nt = trailer.numel() if trailer is not None else 0
    out = torch.empty(nh + sum(lengths) + nt, dtype=torch.uint8, device=payload.device)
    if nh:
        out[:nh] = header
    off = nh
    for r in range(world):
        if lengths[r]:
            if r == rank:
                out[off: off + lengths[r]] = payload
            else:
                dist.recv(out[off: off + lengths[r]], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        off += lengths[r]
    if nt:
        out[off:] = trailer
    return out, lengths


def stitch(header, strips, eoi=b"\xff\xd9"):
    """header + strip_0 + ... + strip_{n-1} + EOI (strips already carry their RSTn separators)."""
    return b"".join([bytes(header)] + [bytes(np.asarray(s, np.uint8)) for s in strips] + [eoi])
