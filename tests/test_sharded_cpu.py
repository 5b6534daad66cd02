"""world_size-2/3 gloo runs of the multi-GPU plumbing (huffman-gpu_b200/sharded.py) on CPU.
The device encode is replaced by the CPU oracle with an explicit start phase -- allowed here because
this is a test of the host logic (plan, offsets, stitch), not of the product path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _encode_with_phase(orc, words, cw, cl, phase):
    """oracle stream shifted right by `phase` zero bits (what hb_encode does with start_bit=phase)."""
    out, bits, _ = orc.encode(words, cw, cl)
    n = (phase + bits + 31) // 32
    bitarr = np.unpackbits(out.byteswap().view(np.uint8))[:bits]
    full = np.zeros(max(n, 1) * 32, dtype=np.uint8)
    full[phase:phase + bits] = bitarr
    return np.packbits(full).view(np.uint32).byteswap(), bits


def _worker(rank, world, port, n_words, seed, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pyoracle
    import huffman_gpu_b200 as hb
    from huffman_gpu_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = pyoracle.Oracle()
        rng = np.random.default_rng(seed)
        p = 0.6 ** np.arange(40)
        p /= p.sum()
        data = rng.choice(40, size=n_words * 4, p=p).astype(np.uint8)     # same on every rank
        words = data.view(np.uint32)
        lo, hi = sharded.shard_bounds(n_words, world, tile_words=64)[rank]
        mine = words[lo:hi]
        plan = sharded.make_plan(orc.histogram(mine.view(np.uint8)) if mine.size else np.zeros(256, np.uint64))
        assert int(plan.hist_global.sum()) == n_words * 4
        local, bits = _encode_with_phase(orc, mine, plan.codewords, plan.codewordlens, plan.my_phase)
        assert bits == plan.my_bits
        t = torch.from_numpy(local.view(np.int32).copy())
        out = sharded.stitch_on_rank0(plan, t)
        if rank == 0:
            ref_words, ref_bits, _ = orc.encode(words, plan.codewords, plan.codewordlens)
            assert ref_bits == plan.total_bits
            got = out.numpy().view(np.uint32)
            assert np.array_equal(got[: ref_words.size], ref_words)
            ret.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_words", [(2, 1000), (3, 777), (2, 64), (2, 1)])
def test_sharded_plan_and_stitch_gloo(native_built, world, n_words):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + n_words % 5
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_words, 5, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"


def test_shard_bounds():
    from huffman_gpu_b200 import sharded
    b = sharded.shard_bounds(1000000, 4)
    assert b[0][0] == 0 and b[-1][1] == 1000000
    assert all(b[i][1] == b[i + 1][0] for i in range(3))
    from huffman_gpu_b200 import capi
    assert all(lo % (capi.TILE_BYTES // 4) == 0 for lo, hi in b if hi > lo)
