"""Multi-GPU encode: contiguous shards, one process per GPU (the reference is
single-GPU only, hist.cu:67 hard-codes device 0).

The data path needs NO bulk collective: every rank encodes its own contiguous shard.  Two tiny
exchanges put the shards in global phase:
  1. all-reduce of the 256-bin histogram (256 x int64)  -> every rank builds the identical codebook
     on its host (hb_build_codebook is deterministic);
  2. all-gather of one int64 per rank, the shard's bit total sum_s hist_r[s]*len[s], known BEFORE
     encoding -> exclusive prefix = the shard's global start bit.
Rank r then encodes with start_bit = off_r % 32 so that its words are already in global phase.
The optional stitch gathers the shards on rank 0 (NCCL send/recv, NVLink P2P): shard r's words land
at global word off_r // 32; the seam word shared with shard r-1 is OR-ed (hb_stitch_seam).

Two implementations of the same plan:
  * ShardComm -- the C ABI (hb_comm_* / hb_shard_* / hb_stitch_*, csrc/hb_comm.cu): NCCL called from C on the
    caller's stream, the stitch as concurrent peer stores over NVLink into an IPC-mapped buffer on the root GPU.
    This is the product path on GPUs (bench.py, the -m gpu tests).
  * make_plan / stitch_on_rank0 -- the same arithmetic over any torch.distributed backend; used with gloo on CPU
    to test the host logic with world sizes > 1, and as a cross-check of ShardComm on GPUs.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .encoder import _check, bits_from_hist, build_codebook, lib, shard_offsets


def shard_bounds(n_words, world, tile_words=None):
    """Contiguous word ranges, boundaries on encode-tile multiples (last shard takes the ragged end)."""
    if tile_words is None:
        from . import capi
        from .encoder import lib
        lib()
        tile_words = capi.TILE_BYTES // 4
    tiles = (n_words + tile_words - 1) // tile_words
    per = (tiles + world - 1) // world
    bounds = []
    for r in range(world):
        lo = min(n_words, r * per * tile_words)
        hi = min(n_words, (r + 1) * per * tile_words)
        bounds.append((lo, hi))
    return bounds


class ShardPlan:
    def __init__(self, hist_global, codewords, codewordlens, max_len, shard_bits, start_bits, total_bits,
                 rank):
        self.hist_global = hist_global
        self.codewords = codewords
        self.codewordlens = codewordlens
        self.max_len = max_len
        self.shard_bits = shard_bits          # per rank
        self.start_bits = start_bits          # per rank, global bit offsets
        self.total_bits = total_bits
        self.rank = rank

    @property
    def my_start(self):
        return int(self.start_bits[self.rank])

    @property
    def my_bits(self):
        return int(self.shard_bits[self.rank])

    @property
    def my_phase(self):
        return self.my_start % 32

    @property
    def my_words(self):
        """words the local encode writes: ceil((phase + bits) / 32), at least 1"""
        return max(1, (self.my_phase + self.my_bits + 31) // 32)


def make_plan(local_hist, group=None, device="cpu"):
    """local_hist: numpy uint64[256] of this rank's shard.  Collective over `group`."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    h = torch.from_numpy(np.ascontiguousarray(local_hist, dtype=np.uint64).view(np.int64).copy()).to(device)
    dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)                  # 2 KiB
    hist_global = h.cpu().numpy().view(np.uint64)
    cw, cl, max_len = build_codebook(hist_global)                          # identical on every rank
    mine = torch.tensor([bits_from_hist(local_hist, cl)], dtype=torch.int64, device=device)
    allbits = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allbits, mine, group=group)                # 8 B per rank
    shard_bits = allbits.cpu().numpy().view(np.uint64)
    start_bits, total = shard_offsets(shard_bits)
    return ShardPlan(hist_global, cw, cl, max_len, shard_bits, start_bits, total, rank)


def stitch_on_rank0(plan, local_words, group=None, or_fn=None):
    """Gather every shard's words on rank 0 into one stream (int32 tensor of ceil(total/32)+1 words).
    local_words: this rank's encode output (tensor, >= plan.my_words words, same device/dtype on
    every rank).  Returns the stitched tensor on rank 0, None elsewhere.
    or_fn(dst_slice, src_slice): in-place OR for the seam word (hb_stitch_seam on GPUs)."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if or_fn is None:
        def or_fn(dst, src):
            dst.bitwise_or_(src)
    if rank != 0:
        dist.send(local_words[: plan.my_words].contiguous(), dst=0, group=group)
        return None
    n_out = int(plan.total_bits) // 32 + 1
    out = torch.zeros(n_out, dtype=local_words.dtype, device=local_words.device)
    for r in range(world):
        phase = int(plan.start_bits[r]) % 32
        bits = int(plan.shard_bits[r])
        n = max(1, (phase + bits + 31) // 32)
        if r == 0:
            part = local_words[:n]
        else:
            part = torch.empty(n, dtype=local_words.dtype, device=local_words.device)
            dist.recv(part, src=r, group=group)
        if bits == 0:
            continue
        w0 = int(plan.start_bits[r]) // 32
        if phase:
            or_fn(out[w0:w0 + 1], part[:1])          # seam word: r-1's tail bits | r's head bits
            if n > 1:
                out[w0 + 1:w0 + n].copy_(part[1:n])
        else:
            out[w0:w0 + n].copy_(part[:n])
    return out


# ---- the C-ABI path ---------------------------------------------------------------------------------------------
class ShardComm:
    """hb_comm over an Encoder (one per rank).  The 128-byte NCCL unique id is created on rank 0 and handed to the
    other ranks through `bcast(bytes_or_None) -> bytes`, any channel the caller has (default: torch.distributed)."""

    def __init__(self, enc, rank, world, bcast=None):
        self.enc, self.rank, self.world = enc, int(rank), int(world)
        uid = (C.c_uint8 * capi.HB_UNIQUE_ID_BYTES)()
        if self.rank == 0:
            _check(lib().hb_comm_unique_id(uid), "hb_comm_unique_id")
        if bcast is None:
            t = torch.tensor(list(uid), dtype=torch.uint8)
            dev = torch.device("cuda", enc.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = t.to(dev)
            dist.broadcast(t, src=0)
            payload = bytes(t.cpu().tolist())
        else:
            payload = bcast(bytes(uid) if self.rank == 0 else None)
        uid = (C.c_uint8 * capi.HB_UNIQUE_ID_BYTES).from_buffer_copy(payload)
        self._comm = capi.vp()
        _check(lib().hb_comm_init(C.byref(self._comm), enc._ctx, self.rank, self.world, uid), "hb_comm_init", enc._ctx)
        self.plan = None
        self.stitched_ptr = None

    def close(self):
        if self._comm:
            lib().hb_comm_free(self._comm)
            self._comm = capi.vp()

    def _ck(self, status, where):
        if status == capi.HB_ERR_NCCL:
            raise capi.HBError(lib(), status, "%s (ncclResult %d)" % (where, lib().hb_comm_last_nccl_error(self._comm)))
        return _check(status, where, self.enc._ctx)

    def plan_build(self, d_in):
        """collective: histogram -> all-reduce -> codebook -> all-gather.  -> (codewords, codewordlens, plan, hist_global)"""
        ptr, n_words = self.enc._words(d_in)
        cw = np.zeros(256, dtype=np.uint32)
        cl = np.zeros(256, dtype=np.uint32)
        hist = np.zeros(256, dtype=np.uint64)
        plan = capi.ShardPlan()
        self._ck(lib().hb_shard_plan_build(self._comm, ptr, n_words, cw.ctypes.data_as(capi.u32p),
                                           cl.ctypes.data_as(capi.u32p), hist.ctypes.data_as(capi.u64p),
                                           C.byref(plan), self.enc._stream()), "hb_shard_plan_build")
        self.plan = plan
        return cw, cl, plan, hist

    def offsets(self):
        starts = np.zeros(self.world, dtype=np.uint64)
        bits = np.zeros(self.world, dtype=np.uint64)
        self._ck(lib().hb_comm_plan_offsets(self._comm, starts.ctypes.data_as(capi.u64p),
                                            bits.ctypes.data_as(capi.u64p)), "hb_comm_plan_offsets")
        return starts, bits

    def local_buffer(self, plan=None):
        """a device buffer for this rank's words (int32), sized as hb_shard_plan asks"""
        plan = plan or self.plan
        return torch.empty(int(plan.local_offset_words + plan.local_words + 1), dtype=torch.int32,
                           device=torch.device("cuda", self.enc.device))

    def encode_async(self, d_in, cw, cl, d_local, plan=None):
        plan = plan or self.plan
        ptr, n_words = self.enc._words(d_in)
        optr, cap = self.enc._words(d_local)
        cw = np.ascontiguousarray(cw, dtype=np.uint32)
        cl = np.ascontiguousarray(cl, dtype=np.uint32)
        self._ck(lib().hb_shard_encode_async(self._comm, ptr, n_words, cw.ctypes.data_as(capi.u32p),
                                             cl.ctypes.data_as(capi.u32p), optr, cap, C.byref(plan),
                                             self.enc._stream()), "hb_shard_encode_async")

    def encode_result(self):
        bits = C.c_uint64(0)
        self._ck(lib().hb_shard_encode_result(self._comm, C.byref(bits), self.enc._stream()), "hb_shard_encode_result")
        return int(bits.value)

    def local_words_view(self, d_local, plan=None):
        plan = plan or self.plan
        o = int(plan.local_offset_words)
        return d_local[o:o + int(plan.local_words) + 1]

    def stitch_open(self, capacity_words, root=0):
        p = capi.vp()
        self._ck(lib().hb_stitch_open(self._comm, int(capacity_words), int(root), C.byref(p), self.enc._stream()),
                 "hb_stitch_open")
        self.stitched_ptr, self.stitch_root, self.stitch_cap = p.value, int(root), int(capacity_words)
        return p.value

    def stitch_push(self, d_local, plan=None):
        plan = plan or self.plan
        self._ck(lib().hb_stitch_push(self._comm, d_local.data_ptr(), C.byref(plan), self.enc._stream()),
                 "hb_stitch_push")

    def encode_direct_async(self, d_in, cw, cl, plan=None):
        """fused encode + stitch: the shard goes straight into the root's stream (peer stores from the encode kernel)"""
        plan = plan or self.plan
        ptr, n_words = self.enc._words(d_in)
        cw = np.ascontiguousarray(cw, dtype=np.uint32)
        cl = np.ascontiguousarray(cl, dtype=np.uint32)
        self._ck(lib().hb_shard_encode_direct_async(self._comm, ptr, n_words, cw.ctypes.data_as(capi.u32p),
                                                    cl.ctypes.data_as(capi.u32p), C.byref(plan), self.enc._stream()),
                 "hb_shard_encode_direct_async")

    def stitched_view(self, n_words):
        """root only: the first n_words of the stitched stream as an int32 torch tensor (a view, no copy)"""
        assert self.rank == self.stitch_root and n_words <= self.stitch_cap

        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4",
                                        "data": (int(self.stitched_ptr), False), "version": 2}
        return torch.as_tensor(_Raw(), device=torch.device("cuda", self.enc.device))

    def stitch_close(self):
        self._ck(lib().hb_stitch_close(self._comm), "hb_stitch_close")
        self.stitched_ptr = None
