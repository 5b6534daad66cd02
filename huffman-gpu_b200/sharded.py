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

The plumbing takes any torch.distributed backend (NCCL on GPUs; gloo on CPU for the host-logic tests)
and an `encode_fn`, which on a GPU box is Encoder.encode.
"""
import numpy as np
import torch
import torch.distributed as dist

from .encoder import bits_from_hist, build_codebook, shard_offsets


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
