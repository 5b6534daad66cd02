#!/usr/bin/env python
"""bench.py -- Huffman encode GB/s (input) on B200, with HBM roofline and the reference's CPU encoder.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|t1g|c3|c4|c5|c1] [--impl reference]

A "step" is one pass of the hot path (hb_encode: the single-pass encode kernel) over one batch of
synthetic input that is already resident in HBM; histogram and codebook are built once, outside the
timed region (as in the reference, where only the encode is bracketed by events,
main_test_cu.cu:136-156).  `value` = input bytes of all ranks / device time (CUDA events on the
launching stream, max over ranks).  `e2e` = the same metric through the host-buffer C-ABI call
(hb_vlc_encode_host: pinned host input -> H2D -> encode -> D2H of the packed stream), wall clock.
`roofline` = algorithmic bytes (input + ceil(bits/8)) / mean kernel time against the measured HBM
copy peak.  `cpu_baseline` = the reference's cpu_vlc_encode (oracle/_ref, unmodified) on this box.

Under torchrun (N > 1) every rank encodes its own contiguous shard in global bit phase (weak scaling:
fixed bytes per GPU); the only collectives are the histogram all-reduce and the bit-total all-gather,
both outside the timed region because they depend on the histogram only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "huffman_encode_input_GBps"
UNIT = "GB/s"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(workload, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the encode kernel, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json), or None when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(workload)
        if e and e.get("kernel_variant") in (None, variant):
            return e.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """Polls NVML (SM clock, event reasons) while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.trace = []                       # (time, sm clock, reasons mask, inside the timed region?)
        self.sm_max = None
        self.active = threading.Event()
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def _one(self):
        nv = self.nv
        clk = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        return clk, mask

    def run(self):
        """polls from start() to stop: every sample is time-stamped, the summary keeps those inside the timed region"""
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                t = time.perf_counter()
                clk, mask = self._one()
                self.trace.append((t, clk, mask, self.active.is_set()))
            except Exception:
                pass
            time.sleep(0.0002)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"], "samples": 0}
        inside = [(c, m) for (_, c, m, act) in self.trace if act]
        note = None
        if not inside:
            # the region was shorter than one poll: take the polls right around it (the last before the region ended
            # is still inside the busy phase that started with the warm-up)
            busy = [(c, m) for (_, c, m, _) in self.trace[-3:]]
            try:
                busy.append(self._one())
            except Exception:
                pass
            inside = busy
            note = "timed region shorter than one NVML poll; nearest polls around it"
        reasons = set()
        for _, mask in inside:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        clocks = [c for c, _ in inside]
        out = {"sm_mhz": float(np.median(clocks)) if clocks else None, "sm_max_mhz": self.sm_max,
               "reasons": sorted(reasons), "samples": len(clocks)}
        if note:
            out["note"] = note
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_reference_encoder():
    """(callable(words, cw, cl, bits) -> (out_words, out_bytes), kind).  oracle/ is used here ONLY as the
    timed CPU baseline / checker, never as the product path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    ref = pyoracle.try_ref()
    if ref is not None:
        def run(words, cw, cl, bits):
            out, size = ref.encode(words, cw, cl, bits // 32 + 2)
            return out, size
        return run, "reference"
    orc = pyoracle.Oracle()

    def run(words, cw, cl, bits):
        out, b, size = orc.encode(words, cw, cl, total_bits_hint=bits)
        return out, size
    return run, "port"


def time_cpu(run, words, cw, cl, bits, repeats):
    ts = []
    out = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out, size = run(words, cw, cl, bits)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


def workload_for(args, hb):
    if args.workload == "c1":
        return None
    return hb.workloads.get(args.workload, n_bytes=args.bytes)


def host_sample(hb, wl, n_bytes):
    """Host copy of the first n_bytes of the workload via the CPU generator (reference arm only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    orc = pyoracle.Oracle()
    if wl is None:
        d = hb.workloads.c1_fixture_bytes()
        return d[: min(n_bytes, d.size)].copy()
    return orc.synth_fill(0, n_bytes, wl.seed, wl.mode, wl.nbits, wl.thr, wl.symmap)


def describe(wl, args):
    if wl is None:
        return "c1: data/test1024_H2.206587175259.in (1 MiB fixture, H~2.21)"
    return "%s: %s" % (wl.name, wl.note)


# ------------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """The reference's own CPU implementation of the path (cpu_vlc_encode, single-threaded by
    construction: loop-carried startbit, cpuencode.cpp:18,38), on a bounded sample of the workload."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import huffman_gpu_b200 as hb
    wl = workload_for(args, hb)
    total = wl.n_bytes if wl is not None else 1 << 20
    run, kind = cpu_reference_encoder()
    # a bounded sample per step: calibrate on 8 MiB, then size the step so that K + W steps take ~budget seconds
    cal = host_sample(hb, wl, min(total, 8 << 20))
    chist = np.bincount(cal, minlength=256).astype(np.uint64)
    ccw, ccl, _ = hb.build_codebook(chist)
    cbits = hb.bits_from_hist(chist, ccl)
    run(cal.view(np.uint32), ccw, ccl, cbits)
    c0 = time.perf_counter()
    run(cal.view(np.uint32), ccw, ccl, cbits)
    rate = cal.size / (time.perf_counter() - c0)                         # bytes/s of this box, one core
    want = int(args.cpu_budget_s * rate / max(1, args.steps + args.warmup))
    sample = max(16 << 20, min(total, args.cpu_sample_mib << 20, want))     # never a cache-resident sample
    sample -= sample % hb.capi.TILE_BYTES if sample >= hb.capi.TILE_BYTES else 0
    data = host_sample(hb, wl, sample)
    hist = np.bincount(data, minlength=256).astype(np.uint64)
    cw, cl, max_len = hb.build_codebook(hist)
    bits = hb.bits_from_hist(hist, cl)
    words = data.view(np.uint32)
    for _ in range(args.warmup):
        run(words, cw, cl, bits)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(words, cw, cl, bits)
    dt = time.perf_counter() - t0
    gbps = sample * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": describe(wl, args), "input_bytes_per_step": sample,
                   "sample": "first %.1f MiB of the workload per step" % (sample / 2.0 ** 20)},
        "cpu_baseline": {"value": gbps, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": "first %.1f MiB of %s per step, %d steps; cpu_vlc_encode is serial "
                                   "(loop-carried startbit, cpuencode.cpp:18,38): it can use 1 of the %d "
                                   "host cores" % (sample / 2.0 ** 20, args.workload, args.steps, os.cpu_count())},
        "e2e": {"value": gbps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "t1g", "c3", "c4", "c5"])
    ap.add_argument("--bytes", type=int, default=None, help="override the per-GPU input size")
    ap.add_argument("--cpu-sample-mib", type=int, default=256)
    ap.add_argument("--cpu-budget-s", type=float, default=60.0, help="--impl reference: CPU seconds for all steps")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stitch", action="store_true", help="N > 1: skip the optional gather of the shards on rank 0")
    ap.add_argument("--e2e-steps", type=int, default=None, help="default: min(steps, 20)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                            # timing hygiene: never fewer than 3 warm-up steps
    if args.e2e_steps is None:
        args.e2e_steps = max(1, min(args.steps, 20))

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import huffman_gpu_b200 as hb
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the encode path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = workload_for(args, hb)
    n_bytes = wl.n_bytes if wl is not None else 1 << 20
    enc = hb.Encoder(device=local, max_bytes=n_bytes)

    # ---- this rank's shard: bytes [rank*n_bytes, (rank+1)*n_bytes) of one logical stream -----------------
    d_in = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    if wl is None:
        d_in.copy_(torch.from_numpy(hb.workloads.c1_fixture_bytes().copy()))
    elif wl.mode == 1:
        enc.synth_fill(d_in, wl, first=0)          # exact-count stream: every rank takes the same 2^nbits positions
    else:
        enc.synth_fill(d_in, wl, first=rank * n_bytes)
    torch.cuda.synchronize()

    # ---- histogram -> (all-reduce) -> codebook -> shard start bit  (outside the timed region) ------------
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    hist = enc.histogram(d_in)                     # warm; the host copy feeds the codebook
    d_hist = torch.zeros(256, dtype=torch.int64, device="cuda")
    t0.record()
    for _ in range(10):
        enc.histogram_device(d_in, d_hist)         # the kernel alone (no host round trip between launches)
    t1.record()
    torch.cuda.synchronize()
    hist_ms = t0.elapsed_time(t1) / 10
    assert np.array_equal(d_hist.cpu().numpy().astype(np.uint64), hist * np.uint64(10))
    if world > 1:
        from huffman_gpu_b200 import sharded
        plan = sharded.make_plan(hist, device="cuda")
        cw, cl, max_len = plan.codewords, plan.codewordlens, plan.max_len
        my_bits, start_bit = plan.my_bits, plan.my_phase
    else:
        cw, cl, max_len = hb.build_codebook(hist)
        my_bits, start_bit = hb.bits_from_hist(hist, cl), 0
    out_words = (start_bit + my_bits) // 32 + 2
    d_out = torch.empty(out_words, dtype=torch.int32, device="cuda")

    # ---- warm-up + timed region ----------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        bits = enc.encode(d_in, cw, cl, d_out, start_bit=start_bit)
    assert bits == my_bits, (bits, my_bits)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    launches0 = enc.launches
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.active.set()
    ev0.record()
    for _ in range(args.steps):
        enc.encode_async(d_in, cw, cl, d_out, start_bit=start_bit)      # one kernel launch per step
    ev1.record()
    torch.cuda.synchronize()
    sampler.active.clear()
    if dist is not None:
        dist.barrier()
    region_ms = ev0.elapsed_time(ev1)
    kernel_ms = region_ms / args.steps             # mean launch-to-launch time of the one kernel in the region
    bits = enc.encode_result()
    assert bits == my_bits
    launches = enc.launches - launches0
    sampler.stop_flag.set()
    if dist is not None:
        t = torch.tensor([region_ms, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_ms, kernel_ms = float(t[0]), float(t[1])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt[0])

    total_in = n_bytes * world
    value = total_in * args.steps / (region_ms * 1e-3) / 1e9
    out_bytes = (my_bits + 7) // 8
    algo_bytes = n_bytes + out_bytes                                   # per launch, per GPU
    peak, peak_src = measured_hbm_peak()
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    variant = hb.encode_variant(cl)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": region_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {
            "workload": describe(wl, args), "input_bytes_per_gpu": n_bytes,
            "output_bytes_per_gpu": out_bytes, "mean_code_len_bits": my_bits / n_bytes,
            "max_code_len": int(max_len), "kernel_variant": variant,
            "l2": "no flush: input+output per step (%d MB) exceeds the 126 MB L2" % (algo_bytes // 10 ** 6)
                  if algo_bytes > 130e6 else "WARNING: working set fits L2; numbers are L2-assisted",
            "parallelism": "1 process per GPU, contiguous shards, no data-path collective",
        },
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": known_traffic(args.workload, variant),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                     "kernel_ms": kernel_ms, "frac_of_nominal_8TBps": achieved / 8000.0},
        "histogram": {"GBps": n_bytes / (hist_ms * 1e-3) / 1e9, "ms": hist_ms},
        "clocks": sampler.summary(),
    }

    # ---- optional stitch (N > 1): the shards gathered on rank 0 over NVLink, seam words OR-ed; reported separately ----
    if dist is not None and not args.no_stitch:
        try:
            from huffman_gpu_b200 import sharded
            or_fn = lambda dst, src: enc.stitch_seam(dst, src, 1)          # noqa: E731
            stitched = sharded.stitch_on_rank0(plan, d_out, or_fn=or_fn)   # warm: NCCL sets its P2P channels up lazily
            del stitched
            dist.barrier()
            torch.cuda.synchronize()
            s0 = time.perf_counter()
            stitched = sharded.stitch_on_rank0(plan, d_out, or_fn=or_fn)
            torch.cuda.synchronize()
            dist.barrier()
            st = torch.tensor([time.perf_counter() - s0], dtype=torch.float64, device="cuda")
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
            out_total = int(plan.total_bits) // 8
            line["stitch"] = {"ms": float(st[0]) * 1e3, "stream_bytes": out_total,
                              "GBps_of_stream": out_total / float(st[0]) / 1e9,
                              "what": "NCCL send/recv of every shard's words to rank 0 + one OR per seam word; "
                                      "not part of `value`"}
            del stitched
        except Exception as exc:                                   # never let the optional leg break the line
            line["stitch"] = {"error": repr(exc)}

    # ---- e2e through the host-buffer C-ABI call (rank-local, wall clock, pinned host buffers) ----------------
    if not args.no_e2e:
        pin_in = hb.PinnedBuffer(n_bytes)
        pin_out = hb.PinnedBuffer((my_bits // 32 + 2) * 4)
        torch.cuda.synchronize()
        pin_in.u8[:] = d_in.cpu().numpy()
        h_in = pin_in.u8.view(np.uint32)
        h_out = pin_out.u8.view(np.uint32)
        enc.encode_host(h_in, cw, cl, h_out)                           # warm (allocates device buffers)
        if dist is not None:
            dist.barrier()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eb, _ = enc.encode_host(h_in, cw, cl, h_out)
        wall = time.perf_counter() - w0
        assert eb == my_bits
        if dist is not None:
            t = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t[0])
        line["e2e"] = {"value": total_in * args.e2e_steps / wall / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": n_bytes * world,
                       "d2h_bytes_per_step": (my_bits // 32 + 1) * 4 * world + 16 * world,
                       "steps": args.e2e_steps, "api": "hb_vlc_encode_host (pinned host in/out)"}
        e2e_words = h_out.copy()
        pin_in.free()
        pin_out.free()
    else:
        e2e_words = None

    # ---- cpu_baseline: the reference's cpu_vlc_encode on this box, rank 0, N=1 only; doubles as parity check --
    if rank == 0 and world == 1 and not args.no_cpu:
        run, kind = cpu_reference_encoder()
        sample = min(n_bytes, args.cpu_sample_mib << 20)
        sample -= sample % hb.capi.TILE_BYTES if sample >= hb.capi.TILE_BYTES else 0
        host = d_in[:sample].cpu().numpy()
        shist = np.bincount(host, minlength=256).astype(np.uint64)
        sbits = hb.bits_from_hist(shist, cl)
        secs, cpu_out = time_cpu(run, host.view(np.uint32), cw, cl, sbits, args.cpu_repeats)
        got = d_out[: sbits // 32].cpu().numpy().view(np.uint32)       # the prefix is position independent
        exact = bool(np.array_equal(got, cpu_out[: sbits // 32]))
        if e2e_words is not None:
            exact = exact and bool(np.array_equal(e2e_words[: sbits // 32], cpu_out[: sbits // 32]))
        line["cpu_baseline"] = {
            "value": sample / secs / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d MiB of the workload, median of %d runs; cpu_vlc_encode is serial "
                      "(1 of %d host cores)" % (sample >> 20, args.cpu_repeats, os.cpu_count()),
            "bit_exact_vs_gpu": exact}
        if not exact:
            line["parity_error"] = "GPU stream differs from cpu_vlc_encode"
    # ---- the reference's own 3-pass GPU pipeline on the same device buffer (N=1; inside its validity limits only) --
    if rank == 0 and world == 1 and not args.no_cpu and args.workload in ("c1", "c2") and n_bytes % 16384 == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        rg = pyoracle.try_ref_gpu()
        if rg is not None:
            d_ref = torch.empty(n_bytes // 4, dtype=torch.int32, device="cuda")
            rbits, ms_e, ms_s, ms_p = rg.run(d_in.data_ptr(), n_bytes // 4, cw, cl, d_ref.data_ptr(), n_bytes, repeats=5)
            torch.cuda.synchronize()
            nw = my_bits // 32
            same = bool(rbits == my_bits and torch.equal(d_ref[:nw], d_out[:nw]))
            line["reference_gpu"] = {
                "what": "vlc_encode_kernel_sm64huff + prescanArray + cudaMemset + pack2 (unmodified kernels, sm_100a, "
                        "oracle/ref_gpu_shim.cu), same device buffer",
                "ms_encode": ms_e, "ms_scan": ms_s, "ms_memset_pack": ms_p, "ms_total": ms_e + ms_s + ms_p,
                "value": n_bytes / ((ms_e + ms_s + ms_p) * 1e-3) / 1e9, "unit": UNIT, "bit_exact_vs_ours": same}
            del d_ref
    if rank == 0:
        print(json.dumps(line))
    enc.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
