"""Sharded encode on real GPUs: every shard is encoded in global bit phase
(start_bit = shard offset mod 32) and the stitched stream must equal the single-GPU stream and cpu_vlc_encode.

  * one GPU: the shards are encoded one after the other on the same device (same kernels, same phases,
    same stitch), which is what a rank does;
  * two or more GPUs, torch path: torch.distributed over NCCL, one process per GPU (spawned here);
  * two or more GPUs, C-ABI path (hb_comm_* / hb_shard_* / hb_stitch_*): NCCL called from C, the unique id handed
    over through plain multiprocessing queues (no torch.distributed at all), the stitch as peer stores over NVLink
    into the root's IPC-mapped buffer.  Small cases are compared word by word with the CPU oracle; the full C4
    (2 GiB on 2 GPUs) and C5 (8 GiB on 2/4/8 GPUs) streams with the golden checksums of the unmodified reference."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _data(orc, hb, n_bytes):
    wl = hb.workloads.get("c5")
    return orc.synth_fill(0, n_bytes, wl.seed, wl.mode, wl.nbits, wl.thr)


@pytest.mark.parametrize("world,n_bytes", [(2, 3 * 32768 + 4096), (4, 2_000_000), (8, 8 << 20), (3, 4096)])
def test_shards_in_global_phase_one_device(hb, orc, world, n_bytes):
    import torch
    from huffman_gpu_b200 import sharded
    torch.cuda.set_device(0)
    data = _data(orc, hb, n_bytes)
    words = data.view(np.uint32)
    enc = hb.Encoder(device=0, max_bytes=n_bytes)
    hist = np.zeros(256, dtype=np.uint64)
    bounds = sharded.shard_bounds(words.size, world)
    d_shards, local_hists = [], []
    for lo, hi in bounds:
        d = torch.from_numpy(words[lo:hi].copy()).cuda() if hi > lo else torch.empty(0, dtype=torch.int32, device="cuda")
        h = enc.histogram(d) if hi > lo else np.zeros(256, dtype=np.uint64)
        d_shards.append(d)
        local_hists.append(h)
        hist += h
    assert np.array_equal(hist, orc.histogram(data))                       # what the all-reduce would produce
    cw, cl, _ = hb.build_codebook(hist)
    shard_bits = np.array([hb.bits_from_hist(h, cl) for h in local_hists], dtype=np.uint64)   # the all-gather
    starts, total = hb.shard_offsets(shard_bits)
    ref_words, ref_bits, _ = orc.encode(words, cw, cl)
    assert total == ref_bits
    out = torch.zeros(total // 32 + 2, dtype=torch.int32, device="cuda")
    for r in range(world):
        phase = int(starts[r]) % 32
        nw = max(1, (phase + int(shard_bits[r]) + 31) // 32)
        part = torch.full((nw + 2,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        bits = enc.encode(d_shards[r], cw, cl, part[: nw + 1], start_bit=phase)
        assert bits == int(shard_bits[r])
        if bits == 0:
            continue
        w0 = int(starts[r]) // 32
        if phase:
            enc.stitch_seam(out[w0:w0 + 1], part[:1], 1)                  # r-1's tail bits | r's head bits
            out[w0 + 1:w0 + nw].copy_(part[1:nw])
        else:
            out[w0:w0 + nw].copy_(part[:nw])
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[: ref_words.size], ref_words)
    one = torch.full((total // 32 + 2,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    d_all = torch.from_numpy(words.copy()).cuda()
    assert enc.encode(d_all, cw, cl, one) == ref_bits
    assert np.array_equal(one.cpu().numpy().view(np.uint32)[: ref_words.size], ref_words)
    enc.close()


def _nccl_worker(rank, world, port, n_bytes, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import pyoracle
    import huffman_gpu_b200 as hb
    from huffman_gpu_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        orc = pyoracle.Oracle()
        data = _data(orc, hb, n_bytes)                                    # same bytes on every rank
        words = data.view(np.uint32)
        lo, hi = sharded.shard_bounds(words.size, world)[rank]
        enc = hb.Encoder(device=rank, max_bytes=max(4, (hi - lo) * 4))
        d_in = torch.from_numpy(words[lo:hi].copy()).cuda()
        plan = sharded.make_plan(enc.histogram(d_in), device="cuda")       # NCCL all-reduce + all-gather
        d_out = torch.empty(plan.my_words + 2, dtype=torch.int32, device="cuda")
        bits = enc.encode(d_in, plan.codewords, plan.codewordlens, d_out, start_bit=plan.my_phase)
        assert bits == plan.my_bits
        stitched = sharded.stitch_on_rank0(plan, d_out, or_fn=lambda dst, src: enc.stitch_seam(dst, src, 1))
        if rank == 0:
            torch.cuda.synchronize()
            ref_words, ref_bits, _ = orc.encode(words, plan.codewords, plan.codewordlens)
            assert ref_bits == plan.total_bits
            got = stitched.cpu().numpy().view(np.uint32)
            assert np.array_equal(got[: ref_words.size], ref_words)
            ret.put("ok")
        enc.close()
    finally:
        dist.destroy_process_group()


def test_shards_nccl_two_gpus(native_built):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the one-device test above covers the same kernels and phases)")
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, 6 << 20, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"


# ---- the C-ABI multi-GPU path -----------------------------------------------------------------------------------
def _comm_worker(rank, world, queues, name, n_bytes, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import json
    import torch
    import pyoracle
    import huffman_gpu_b200 as hb
    from huffman_gpu_b200 import sharded
    from huffman_gpu_b200.streamsum import stream_sums
    try:
        torch.cuda.set_device(rank)
        wl = hb.workloads.get(name) if n_bytes is None else hb.workloads.get(name, n_bytes=n_bytes)
        n_words = wl.n_bytes // 4
        lo, hi = sharded.shard_bounds(n_words, world)[rank]
        enc = hb.Encoder(device=rank, max_bytes=max(4, (hi - lo) * 4))

        def bcast(payload):                                  # the caller's own channel for the 128-byte NCCL id
            if rank == 0:
                for q in queues[1:]:
                    q.put(payload)
                return payload
            return queues[rank].get(timeout=120)
        comm = sharded.ShardComm(enc, rank, world, bcast=bcast)
        d_in = torch.empty((hi - lo) * 4, dtype=torch.uint8, device="cuda")
        if hi > lo:
            enc.synth_fill(d_in, wl, first=lo * 4)           # bytes [lo*4, hi*4) of ONE logical stream
        cw, cl, plan, hist = comm.plan_build(d_in)            # hist kernel -> ncclAllReduce -> codebook -> ncclAllGather
        d_local = comm.local_buffer()
        d_local.fill_(0x5A5A5A5A)
        comm.encode_async(d_in, cw, cl, d_local)
        assert comm.encode_result() == plan.shard_bits
        cap = plan.total_bits // 32 + 2
        comm.stitch_open(cap, root=0)
        comm.stitch_push(d_local)
        torch.cuda.synchronize()
        want_words = want_sums = None
        if rank == 0:
            n = plan.total_bits // 32 + 1
            got_t = comm.stitched_view(n)
            if wl.n_bytes <= (64 << 20):
                orc = pyoracle.Oracle()
                data = orc.synth_fill(0, wl.n_bytes, wl.seed, wl.mode, wl.nbits, wl.thr, wl.symmap)
                assert np.array_equal(hist, orc.histogram(data))
                ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
                assert ref_bits == plan.total_bits
                got = got_t.cpu().numpy().view(np.uint32)
                bad = np.nonzero(got != ref_words[:n])[0]
                assert bad.size == 0, "first mismatch at word %d of %d" % (bad[0], n)
                want_words = ref_words
            else:
                g = json.load(open(os.path.join(ROOT, "tests", "golden", "streams.json")))[name]
                assert plan.total_bits == g["total_bits"]
                assert np.array_equal(hist, np.array(g["hist"], dtype=np.uint64))
                assert np.array_equal(cl, np.array(g["codewordlens"], dtype=np.uint32))
                sums = stream_sums(got_t, n)
                assert ["0x%016x" % x for x in sums] == g["sums"], "stitched stream differs from cpu_vlc_encode's"
                want_sums = g["sums"]
        # ---- the fused form: every shard encoded STRAIGHT into the root's stream (peer stores from the encode kernel,
        #      seam words OR-ed into zeroed words); the root's buffer is poisoned first, the result must be the same stream
        if rank == 0:
            comm.stitched_view(cap).fill_(0x5A5A5A5A)
        comm.encode_direct_async(d_in, cw, cl)
        assert comm.encode_result() == plan.shard_bits
        torch.cuda.synchronize()
        if rank == 0:
            n = plan.total_bits // 32 + 1
            again = comm.stitched_view(cap)
            if want_words is not None:
                got = again[:n].cpu().numpy().view(np.uint32)
                bad = np.nonzero(got != want_words[:n])[0]
                assert bad.size == 0, "direct: first mismatch at word %d of %d" % (bad[0], n)
            else:
                sums = stream_sums(again, n)
                assert ["0x%016x" % x for x in sums] == want_sums, "direct: stream differs from cpu_vlc_encode's"
            assert int(again[n].item()) == 0x5A5A5A5A, "direct: wrote past floor(bits/32)+1 words"
            ret.put("ok")
        comm.stitch_close()
        comm.close()
        enc.close()
    except BaseException as exc:                              # a dead rank must not leave the others in a collective
        ret.put("rank %d: %r" % (rank, exc))
        raise


def _run_comm(world, name, n_bytes, timeout=600):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    queues = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_comm_worker, args=(r, world, queues, name, n_bytes, ret)) for r in range(world)]
    for p in procs:
        p.start()
    msg = ret.get(timeout=timeout)
    for p in procs:
        p.join(60)
        if p.is_alive():
            p.kill()
    assert msg == "ok", msg
    assert all(p.exitcode == 0 for p in procs)


@pytest.mark.parametrize("name,n_bytes", [("c5", 6 << 20), ("c5", 3 * 32768 + 4096), ("c5", 4096), ("c2", 8 << 20),
                                          ("c3", (5 << 20) + 32768 + 8)])
def test_comm_two_gpus_small(native_built, name, n_bytes):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_comm(2, name, n_bytes)


def test_comm_c4_on_two_gpus_full(native_built):
    """BASELINE config 4: 2 GiB Fibonacci-skewed (code lengths 1..31) on 2 B200, stitched stream == cpu_vlc_encode's"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_comm(2, "c4", None)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_comm_c5_full(native_built, world):
    """BASELINE config 5: 8 GiB at H~4 sharded over 2/4/8 B200, stitched stream == cpu_vlc_encode's"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _run_comm(world, "c5", None)


@pytest.mark.parametrize("world,n_bytes", [(4, 2_000_000), (8, 8 << 20), (3, 32768 + 4096)])
def test_comm_many_gpus_small(native_built, world, n_bytes):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _run_comm(world, "c5", n_bytes)
