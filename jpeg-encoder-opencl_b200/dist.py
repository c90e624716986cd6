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
    (row0, row1, first_interval, is_last) per rank; ranks beyond the available intervals get empty strips
    (row0 == row1: nothing to encode, nothing to send).  is_last marks the strip that ends the image -- the one
    that is coded without a trailing RSTn -- which is not rank world-1 when there are fewer intervals than ranks.
    """
    mcu_rows = -(-height // mcu_px)
    n_int = -(-mcu_rows // restart_mcu_rows)
    out = []
    for r in range(world):
        i0, i1 = shard_range(n_int, world, r)
        row0 = min(i0 * restart_mcu_rows * mcu_px, height)
        row1 = min(i1 * restart_mcu_rows * mcu_px, height)
        out.append((row0, row1, i0, row1 == height and row1 > row0))
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
    `dst`).  One all-gather of the lengths, then ONE grouped exchange: all receives are posted together
    (batch_isend_irecv = ncclGroupStart ... ncclGroupEnd on GPUs, so the strips of all peers stream into `dst`
    concurrently over NVLink; plain isend/irecv on gloo) and every strip is received in place.  This is the
    library-collective form of the stitch; PeerStitch below is the one that needs no host round trip.
    Returns (stitched tensor on dst / None elsewhere, lengths).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([payload.numel()], dtype=torch.int64, device=payload.device)
    sizes = torch.zeros(world, dtype=torch.int64, device=payload.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    lengths = [int(v) for v in sizes.tolist()]  # the receive sizes must be known on the host: one sync per stitch
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    if rank != dst:
        if lengths[rank]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload.contiguous(), peer(dst), group)]):
                w.wait()
        return None, lengths
    nh = header.numel() if header is not None else 0
    nt = trailer.numel() if trailer is not None else 0
    out = torch.empty(nh + sum(lengths) + nt, dtype=torch.uint8, device=payload.device)
    if nh:
        out[:nh] = header
    ops, off = [], nh
    for r in range(world):
        if lengths[r]:
            if r == rank:
                out[off: off + lengths[r]] = payload
            else:
                ops.append(dist.P2POp(dist.irecv, out[off: off + lengths[r]], peer(r), group))
        off += lengths[r]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if nt:
        out[off:] = trailer
    return out, lengths


class PeerStitch:
    """The strip stitch over NVLink peer memory: no host round trip, no gather, no collective library in the data path.

    Rank `dst` owns the output file's buffer and a 512-byte control block (plain cudaMalloc through the library); every
    other rank maps both with CUDA IPC.  One step, entirely stream-ordered on the encoder's stream:
        jb_encode_strip_begin   transform + entropy coder up to the sizes; the strip's length stays on the device
        jb_stitch_exchange      one warp: store the length into dst's control block, poll until all N lengths of this
                                step are there, prefix-sum them -> this rank's offset and the end of the data
        jb_encode_strip_finish  the final placement kernel stores the strip at dst's buffer + offset: on ranks other
                                than dst these coalesced 128-bit stores ARE the NVLink transfer
        jb_stitch_complete      others: release-store a completion flag; dst: poll the N-1 flags
    A rank that codes its strip in several calls (more than 2^26 blocks) stitches them locally with device-side running
    offsets and pushes the result with jb_copy_bytes_device.  `cap` bytes are allocated on dst.
    exchange_offsets() / fence() are the same two steps through NCCL (all-gather + cumsum, all-reduce), kept for
    comparison (bench.py --stitch peer-nccl).
    """

    def __init__(self, enc, cap, dst=0, group=None):
        import numpy as np
        import torch
        import torch.distributed as dist
        self.enc, self.cap, self.dst, self.group = enc, int(cap), dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("PeerStitch: at most 16 ranks (one NVLink domain)")
        dev = torch.device("cuda", torch.cuda.current_device())
        handles = torch.zeros(128, dtype=torch.uint8, device=dev)
        self.local = self.local_ctl = None
        if self.rank == dst:
            self.local = enc.device_alloc(self.cap)
            self.local_ctl = enc.device_alloc(512)
            enc.h2d(self.local_ctl, np.zeros(512, np.uint8))
            handles.copy_(torch.from_numpy(np.concatenate([enc.ipc_export(self.local), enc.ipc_export(self.local_ctl)])))
        dist.broadcast(handles, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        h = handles.cpu().numpy()
        self.base = self.local if self.rank == dst else enc.ipc_open(h[:64])
        self.ctl = self.local_ctl if self.rank == dst else enc.ipc_open(h[64:])
        self.epoch = 0
        self.lens = torch.zeros(self.world, dtype=torch.int64, device=dev)
        self.mine = torch.zeros(1, dtype=torch.int64, device=dev)
        self.offs = torch.zeros(self.world + 1, dtype=torch.int64, device=dev)  # NCCL variant
        self.off2 = torch.zeros(2, dtype=torch.int64, device=dev)               # [this rank's offset, end of the data]
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()  # the tensors above are used on the encoder's (non-blocking) stream from here on

    # ---- the exchange as stores / polls on peer memory (default) -----------------------------------------------------
    def exchange(self, header_bytes):
        """Enqueue: publish self.mine, wait for every rank's length of this step; returns the device tensor
        [offset of this rank's strip, end of the data] (both include header_bytes)."""
        self.epoch += 1
        self.enc.stitch_exchange(self.ctl, self.rank, self.world, self.epoch, int(header_bytes), self.mine.data_ptr(), self.off2.data_ptr())
        return self.off2

    def complete(self):
        """Enqueue after the placement: ranks other than dst signal, dst waits for all of them."""
        self.enc.stitch_complete(self.ctl, self.rank, self.world, self.dst, self.epoch)

    def lengths(self):
        """The strip lengths of the last step (dst only; synchronises)."""
        import numpy as np
        w = np.zeros(64, np.uint64)
        self.enc.sync()
        self.enc.d2h(w, self.local_ctl)
        return [int(v) for v in w[(self.epoch & 1) * 16: (self.epoch & 1) * 16 + self.world]]

    # ---- the same through NCCL ------------------------------------------------------------------------------------------
    def exchange_offsets(self, header_bytes):
        """lens[r] <- every rank's self.mine; offs[r] = header_bytes + sum(lens[:r]); offs[world] = end of the data."""
        import torch
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.lens, self.mine, group=self.group)
        self.offs[0] = header_bytes
        torch.cumsum(self.lens, 0, out=self.offs[1:])
        self.offs[1:] += header_bytes
        return self.offs

    def fence(self):
        """After this (stream-ordered) collective every peer's stores into dst's buffer have been issued and completed."""
        import torch.distributed as dist
        dist.all_reduce(self.flag, group=self.group)

    def view(self, nbytes=None):
        """dst's buffer as a torch uint8 tensor (dst only)."""
        import torch

        class _Buf:
            pass
        b = _Buf()
        n = self.cap if nbytes is None else int(nbytes)
        b.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (self.local, False), "version": 2}
        t = torch.as_tensor(b, device=torch.device("cuda", torch.cuda.current_device()))
        t._jb_keepalive = self
        return t

    def close(self):
        """Collective: importers unmap first, then the owner frees."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self.rank != self.dst and self.base:
            self.enc.ipc_close(self.base)
            self.enc.ipc_close(self.ctl)
        dist.barrier(group=self.group)
        if self.local:
            self.enc.device_free(self.local)
            self.enc.device_free(self.local_ctl)
        self.base = self.local = self.ctl = self.local_ctl = None


def stitch(header, strips, eoi=b"\xff\xd9"):
    """header + strip_0 + ... + strip_{n-1} + EOI (strips already carry their RSTn separators)."""
    return b"".join([bytes(header)] + [bytes(np.asarray(s, np.uint8)) for s in strips] + [eoi])
