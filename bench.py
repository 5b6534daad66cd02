#!/usr/bin/env python
"""bench.py -- Huffman encode GB/s (input) on 1/2/4/8 B200, with HBM roofline and the reference's CPU encoder.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4|c3|t1g|c2|c1] [--impl reference]

Default workload at EVERY N: BASELINE config 5 -- ONE 8 GiB stream at H~4.0, split into N contiguous shards (strong
scaling; N = 1 is the whole stream on one GPU).  A "step" is one pass of the hot path over the stream with inputs already
resident in HBM:
  value     input bytes of all ranks / device time of K back-to-back encode launches per rank (CUDA events on the
            launching stream, barrier + synchronize on both sides, MAX over ranks).  Histogram and codebook are outside,
            as in the reference, where only the encode is bracketed by events (main_test_cu.cu:136-156).
  pipeline  the same stream through the whole sharded sequence, every step: hist_kernel -> ncclAllReduce(256 bins) ->
            hb_build_codebook on the host -> ncclAllGather(bit totals) -> encode in global phase  (C ABI: hb_shard_*).
  stitch    (N > 1) the optional gather of the shards into one stream on GPU 0: concurrent peer stores over NVLink into
            an IPC-mapped buffer (hb_stitch_push), timed separately, not part of `value`.
  e2e       the same metric through the host-buffer C-ABI call (hb_vlc_encode_host: pinned host input -> H2D -> encode
            -> D2H of the packed stream), wall clock, every rank its shard.
  roofline  algorithmic bytes (input + ceil(bits/8)) / mean kernel time against the measured HBM copy peak.
  cpu_baseline  the reference's cpu_vlc_encode (oracle/_ref, unmodified) on this box's host, bounded sample.
PARITY INSIDE THE RUN: the stream (N = 1) or the stitched stream (N > 1) must have the checksums of the stream the
unmodified cpu_vlc_encode produced for the same input (tests/golden/streams.json), and three windows of every shard
(both seams and the middle) must equal the CPU oracle bit for bit.  Any mismatch: `parity_error` in the line, exit 1.
At N = 1 the line also carries `per_config`: 1 GiB H2.2 (the north-star target case), C2, C3, C4 -- value, roofline
fraction, kernel variant and the same whole-stream parity check each.
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
TILE_BYTES = 32768           # hb_tile_bytes() of this build; the reference arm must not load the product to ask


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(workload, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the encode kernel from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json) -- NOT measured in this run; (None, why) when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(workload)
        if e and e.get("kernel_variant") in (None, variant):
            return e.get("dram_bytes_per_launch"), "profiles/traffic.json: %s" % e.get("source", "ncu --set full capture")
    except Exception:
        pass
    return None, "no ncu capture of this workload/variant committed"


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
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def load_workloads():
    """the workload DEFINITIONS (sizes, seeds, thresholds) without importing the product package"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hb_workloads", os.path.join(ROOT, "huffman-gpu_b200", "workloads.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["hb_workloads"] = m
    spec.loader.exec_module(m)
    return m


def describe(wl):
    if wl is None:
        return "c1: data/test1024_H2.206587175259.in (1 MiB fixture, H~2.21)"
    return "%s: %s" % (wl.name, wl.note)


def make_config(wl, total_bytes):
    """identical in both arms (the driver compares it)"""
    return {"workload": describe(wl), "input_bytes": int(total_bytes),
            "split": "one stream, N contiguous shards on encode-tile boundaries (strong scaling)"}


def golden_stream(name):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "streams.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    return pyoracle


def cpu_reference_encoder():
    """(callable(words, cw, cl, bits) -> (out_words, out_bytes), kind).  oracle/ is used here ONLY as the timed CPU
    baseline / checker, never as the product path."""
    pyoracle = oracle_module()
    ref = pyoracle.try_ref()
    if ref is not None:
        return (lambda words, cw, cl, bits: ref.encode(words, cw, cl, bits // 32 + 2)), "reference"
    orc = pyoracle.Oracle()

    def run(words, cw, cl, bits):
        out, b, size = orc.encode(words, cw, cl, total_bits_hint=bits)
        return out, size
    return run, "port"


# ------------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """The reference's own CPU implementation of the path (cpu_vlc_encode, single-threaded by construction: loop-carried
    startbit, cpuencode.cpp:18,38), on a bounded sample of the workload.  Nothing of the product is loaded here: the
    sample comes from the CPU generator, the codebook from the restated tree builder (both oracle/)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    workloads = load_workloads()
    pyoracle = oracle_module()
    orc = pyoracle.Oracle()
    wl = None if args.workload == "c1" else workloads.get(args.workload, n_bytes=args.bytes)
    total = wl.n_bytes if wl is not None else 1 << 20
    run, kind = cpu_reference_encoder()

    def host_sample(n):
        if wl is None:
            d = workloads.c1_fixture_bytes()
            return d[: min(n, d.size)].copy()
        return orc.synth_fill(0, n, wl.seed, wl.mode, wl.nbits, wl.thr, wl.symmap)

    def tables(data):
        hist = np.bincount(data, minlength=256).astype(np.uint64)
        rc, cw, cl = orc.build_codebook(hist)
        return cw, cl, int((hist * cl.astype(np.uint64)).sum())

    # a bounded sample per step: calibrate on 8 MiB, then size the step so that K + W steps take ~budget seconds
    cal = host_sample(min(total, 8 << 20))
    ccw, ccl, cbits = tables(cal)
    run(cal.view(np.uint32), ccw, ccl, cbits)
    c0 = time.perf_counter()
    run(cal.view(np.uint32), ccw, ccl, cbits)
    rate = cal.size / (time.perf_counter() - c0)                         # bytes/s of this box, one core
    want = int(args.cpu_budget_s * rate / max(1, args.steps + args.warmup))
    sample = max(16 << 20, min(total, args.cpu_sample_mib << 20, want))     # never a cache-resident sample
    sample = min(sample, total)
    sample -= sample % TILE_BYTES if sample >= TILE_BYTES else 0
    data = host_sample(sample)
    cw, cl, bits = tables(data)
    words = data.view(np.uint32)
    for _ in range(args.warmup):
        run(words, cw, cl, bits)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(words, cw, cl, bits)
    dt = time.perf_counter() - t0
    gbps = sample * args.steps / dt / 1e9
    what = "first %.1f MiB of the workload per step" % (sample / 2.0 ** 20)
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": make_config(wl, total),
        "cpu_baseline": {"value": gbps, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": "%s, %d steps; cpu_vlc_encode is serial (loop-carried startbit, "
                                   "cpuencode.cpp:18,38): it can use 1 of the %d host cores"
                                   % (what, args.steps, os.cpu_count())},
        "e2e": {"value": gbps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
class Bench:
    """One rank's view of one workload."""

    def __init__(self, args, hb, torch, dist, rank, world, local):
        self.args, self.hb, self.torch, self.dist = args, hb, torch, dist
        self.rank, self.world, self.local = rank, world, local
        self.errors = []

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(self, x):
        if self.dist is None:
            return int(x)
        t = self.torch.tensor([x], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return int(t[0])

    # ---- this rank's shard of the ONE logical stream ---------------------------------------------------------
    def load_shard(self, enc, wl):
        torch, hb = self.torch, self.hb
        from huffman_gpu_b200 import sharded
        total = wl.n_bytes if wl is not None else 1 << 20
        lo_w, hi_w = sharded.shard_bounds(total // 4, self.world)[self.rank]
        lo, hi = lo_w * 4, hi_w * 4
        d_in = torch.empty(hi - lo, dtype=torch.uint8, device="cuda")
        if wl is None:
            d_in.copy_(torch.from_numpy(hb.workloads.c1_fixture_bytes()[lo:hi].copy()))
        elif hi > lo:
            enc.synth_fill(d_in, wl, first=lo)         # bytes [lo, hi) of the stream (mode 1: positions lo.. of the bijection)
        torch.cuda.synchronize()
        return d_in, lo, hi

    def time_encode(self, enc, d_in, cw, cl, d_out, start_bit, steps, warmup, sampler=None):
        torch = self.torch
        for _ in range(warmup):
            bits = enc.encode(d_in, cw, cl, d_out, start_bit=start_bit)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = enc.launches
        self.barrier()
        if sampler:
            sampler.active.set()
        ev0.record()
        for _ in range(steps):
            enc.encode_async(d_in, cw, cl, d_out, start_bit=start_bit)      # one kernel launch per step
        ev1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.active.clear()
        self.barrier()
        region_ms = ev0.elapsed_time(ev1)
        assert enc.encode_result() == bits
        return self.max_over_ranks(region_ms), bits, enc.launches - launches0

    # ---- parity: windows of this rank's shard against the CPU oracle -----------------------------------------
    def check_windows(self, enc, d_in, d_out, cw, cl, phase, n_windows=3):
        """encode(window) is position independent: the window's start bit inside the shard is recomputed from the
        histogram of the bytes before it.  Windows: the shard's first and last MiB (the seams) and its middle."""
        hb = self.hb
        orc = oracle_module().Oracle()
        n = d_in.numel()
        if n == 0:
            return True
        win = min(n, 1 << 20)
        win -= win % 4
        starts = sorted(set([0, max(0, (n // 2) // TILE_BYTES * TILE_BYTES), max(0, n - win)]))[:n_windows]
        for a in starts:
            b = min(n, a + win)
            pre = enc.histogram(d_in[:a]) if a else np.zeros(256, np.uint64)
            start = phase + hb.bits_from_hist(pre, cl)
            window = d_in[a:b].cpu().numpy()
            o_out, o_bits, _ = orc.encode(window.view(np.uint32), cw, cl)
            w0, w1 = (start + 31) // 32, (start + o_bits) // 32           # fully covered words of the shard's stream
            if w1 <= w0:
                continue
            got = d_out[w0:w1].cpu().numpy().view(np.uint32)
            obits = np.unpackbits(o_out.byteswap().view(np.uint8))
            lo = w0 * 32 - start
            want = np.packbits(obits[lo:lo + (w1 - w0) * 32]).view(np.uint32).byteswap()
            if not np.array_equal(got, want):
                self.errors.append("rank %d: window at byte %d differs from the CPU oracle" % (self.rank, a))
                return False
        return True

    def check_sums(self, name, words_t, total_bits, what):
        """-> True / False / None (no golden for this workload and size)"""
        from huffman_gpu_b200.streamsum import stream_sums
        g = golden_stream(name)
        if g is None:
            return None
        if total_bits != g["total_bits"]:
            self.errors.append("%s: %d bits, the reference stream has %d" % (what, total_bits, g["total_bits"]))
            return False
        sums = ["0x%016x" % s for s in stream_sums(words_t, g["n_words"])]
        if sums != g["sums"]:
            self.errors.append("%s: stream checksums %s differ from the reference stream's %s" % (what, sums, g["sums"]))
            return False
        return True

    # ---- one single-GPU workload: value, roofline, whole-stream parity (per_config records, N = 1) -------------
    def single_config(self, name, steps, warmup):
        torch, hb = self.torch, self.hb
        wl = hb.workloads.get(name)
        enc = hb.Encoder(device=self.local, max_bytes=wl.n_bytes)
        try:
            d_in, _, _ = self.load_shard(enc, wl)
            hist = enc.histogram(d_in)
            cw, cl, max_len = hb.build_codebook(hist)
            my_bits = hb.bits_from_hist(hist, cl)
            d_out = torch.empty(my_bits // 32 + 2, dtype=torch.int32, device="cuda")
            region_ms, bits, _ = self.time_encode(enc, d_in, cw, cl, d_out, 0, steps, warmup)
            kernel_ms = region_ms / steps
            algo = wl.n_bytes + (bits + 7) // 8
            peak, _ = measured_hbm_peak()
            variant = hb.encode_variant(cl)
            rec = {"workload": describe(wl), "value": wl.n_bytes / (kernel_ms * 1e-3) / 1e9, "unit": UNIT,
                   "ms_per_step": kernel_ms, "steps": steps, "kernel_variant": variant, "max_code_len": int(max_len),
                   "mean_code_len_bits": bits / wl.n_bytes,
                   "roofline": {"achieved": algo / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": algo / (kernel_ms * 1e-3) / 1e9 / peak,
                                "frac_of_nominal_8TBps": algo / (kernel_ms * 1e-3) / 1e9 / 8000.0,
                                "algorithmic_bytes_per_launch": algo},
                   "bit_exact_whole_stream": self.check_sums(name, d_out, bits, name)}
            if name == "c2" and not self.args.no_cpu:
                rg = oracle_module().try_ref_gpu()
                if rg is not None:
                    d_ref = torch.empty(wl.n_bytes // 4, dtype=torch.int32, device="cuda")
                    rbits, ms_e, ms_s, ms_p = rg.run(d_in.data_ptr(), wl.n_bytes // 4, cw, cl, d_ref.data_ptr(),
                                                     wl.n_bytes, repeats=5)
                    torch.cuda.synchronize()
                    nw = bits // 32
                    rec["reference_gpu"] = {
                        "what": "vlc_encode_kernel_sm64huff + prescanArray + cudaMemset + pack2 (unmodified kernels, "
                                "sm_100a, oracle/ref_gpu_shim.cu), same device buffer",
                        "ms_encode": ms_e, "ms_scan": ms_s, "ms_memset_pack": ms_p, "ms_total": ms_e + ms_s + ms_p,
                        "value": wl.n_bytes / ((ms_e + ms_s + ms_p) * 1e-3) / 1e9, "unit": UNIT,
                        "bit_exact_vs_ours": bool(rbits == bits and torch.equal(d_ref[:nw], d_out[:nw]))}
                    del d_ref
            del d_in, d_out
            return rec
        finally:
            enc.close()
            torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "t1g", "c3", "c4", "c5"])
    ap.add_argument("--bytes", type=int, default=None, help="override the TOTAL input size (iid workloads)")
    ap.add_argument("--cpu-sample-mib", type=int, default=256)
    ap.add_argument("--cpu-budget-s", type=float, default=60.0, help="--impl reference: CPU seconds for all steps")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the reference GPU comparator")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stitch", action="store_true", help="N > 1: skip the gather of the shards on GPU 0 (and its check)")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="N = 1: skip the t1g / c2 / c3 / c4 sub-records")
    ap.add_argument("--per-config-steps", type=int, default=50)
    ap.add_argument("--pipeline-steps", type=int, default=10)
    ap.add_argument("--e2e-steps", type=int, default=None, help="default: 3 for >= 4 GiB per rank, else min(steps, 10)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                            # timing hygiene: never fewer than 3 warm-up steps

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
    from huffman_gpu_b200 import sharded
    B = Bench(args, hb, torch, dist, rank, world, local)

    wl = None if args.workload == "c1" else hb.workloads.get(args.workload, n_bytes=args.bytes)
    total_bytes = wl.n_bytes if wl is not None else 1 << 20
    if os.environ.get("HB_LIB") is None:
        assert hb.lib().hb_tile_bytes() == TILE_BYTES      # (A/B builds selected with $HB_LIB may differ)
    lo_w, hi_w = sharded.shard_bounds(total_bytes // 4, world)[rank]
    n_bytes = (hi_w - lo_w) * 4                                       # this rank's shard
    enc = hb.Encoder(device=local, max_bytes=max(4, n_bytes))
    d_in, lo, hi = B.load_shard(enc, wl)

    # ---- histogram -> all-reduce -> codebook -> all-gather (C ABI, NCCL from C) ----------------------------------
    comm = None
    if world > 1:
        # NCCL prints its version banner to stdout when $NCCL_DEBUG asks for it: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            comm = sharded.ShardComm(enc, rank, world)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    if comm is not None:
        cw, cl, plan, hist_global = comm.plan_build(d_in)
        max_len, my_bits, start_bit, total_bits = plan.max_len, int(plan.shard_bits), int(plan.phase), int(plan.total_bits)
        d_local = comm.local_buffer()
        d_out = comm.local_words_view(d_local)
    else:
        hist_global = enc.histogram(d_in)
        cw, cl, max_len = hb.build_codebook(hist_global)
        my_bits, start_bit = hb.bits_from_hist(hist_global, cl), 0
        total_bits = my_bits
        d_out = torch.empty(my_bits // 32 + 2, dtype=torch.int32, device="cuda")

    # the histogram kernel alone (no host round trip between launches)
    d_hist = torch.zeros(256, dtype=torch.int64, device="cuda")
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc.histogram_device(d_in, d_hist)
    d_hist.zero_()
    torch.cuda.synchronize()
    h0.record()
    for _ in range(10):
        enc.histogram_device(d_in, d_hist)
    h1.record()
    torch.cuda.synchronize()
    hist_ms = B.max_over_ranks(h0.elapsed_time(h1) / 10)

    # ---- warm-up + timed region: the encode -------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    region_ms, bits, launches = B.time_encode(enc, d_in, cw, cl, d_out, start_bit, args.steps, args.warmup, sampler)
    sampler.stop_flag.set()
    assert bits == my_bits, (bits, my_bits)
    kernel_ms = region_ms / args.steps             # mean launch-to-launch time of the one kernel in the region
    launches = B.sum_over_ranks(launches)
    value = total_bytes * args.steps / (region_ms * 1e-3) / 1e9
    out_bytes = (my_bits + 7) // 8
    algo_bytes = n_bytes + out_bytes                                   # per launch, this GPU
    algo_max = B.max_over_ranks(algo_bytes)
    peak, peak_src = measured_hbm_peak()
    achieved = algo_max / (kernel_ms * 1e-3) / 1e9
    variant = hb.encode_variant(cl)
    traffic, traffic_src = known_traffic(args.workload if world == 1 else "%s/%d" % (args.workload, world), variant)

    config = make_config(wl, total_bytes)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": region_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config,
        "detail": {
            "input_bytes_per_gpu": n_bytes, "output_bytes_per_gpu": out_bytes,
            "total_bits": total_bits, "mean_code_len_bits": total_bits / total_bytes,
            "max_code_len": int(max_len), "kernel_variant": variant,
            "l2": "no flush: input+output per step and GPU (%d MB) exceeds the 126 MB L2" % (algo_bytes // 10 ** 6)
                  if algo_bytes > 130e6 else "WARNING: working set fits L2; numbers are L2-assisted",
            "parallelism": "1 process per GPU, contiguous shards, no data-path collective; NCCL from C "
                           "(hb_shard_plan_build) for the 2 KiB histogram all-reduce and the 8 B/rank all-gather",
        },
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_max),
                     "kernel_ms": kernel_ms, "frac_of_nominal_8TBps": achieved / 8000.0,
                     "what": "per GPU: the rank with the most bytes / the slowest rank's mean kernel time"},
        "histogram": {"GBps": total_bytes / (hist_ms * 1e-3) / 1e9, "ms": hist_ms},
        "clocks": sampler.summary(),
    }

    # ---- parity inside the run ---------------------------------------------------------------------------------------
    windows_ok = B.check_windows(enc, d_in, d_out, cw, cl, start_bit)
    ok_all = B.sum_over_ranks(0 if windows_ok else 1) == 0
    line["parity"] = {"shard_windows_vs_cpu_oracle": ok_all,
                      "windows": "first MiB, middle MiB and last MiB of every shard (both seams), bit-exact"}
    if world == 1:
        whole = B.check_sums(args.workload, d_out, my_bits, args.workload) if args.bytes is None else None
        line["parity"]["whole_stream_vs_cpu_vlc_encode"] = whole
        line["parity"]["how"] = ("checksums S1,S2,S3 + bit count of the whole stream == those of the stream the unmodified "
                                 "cpu_vlc_encode produced (tests/golden/streams.json)")

    # ---- pipeline: hist -> all-reduce -> codebook -> all-gather -> encode, every step ---------------------------------
    if not args.no_pipeline:
        def one():
            if comm is not None:
                pcw, pcl, pplan, _ = comm.plan_build(d_in)
                comm.encode_async(d_in, pcw, pcl, d_local, pplan)
                return comm.encode_result()
            h = enc.histogram(d_in)
            pcw, pcl, _ = hb.build_codebook(h)
            return enc.encode(d_in, pcw, pcl, d_out)
        assert one() == my_bits
        B.barrier()
        p0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.pipeline_steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        B.barrier()
        pms = B.max_over_ranks(e0.elapsed_time(e1)) / args.pipeline_steps
        wall = B.max_over_ranks(time.perf_counter() - p0) / args.pipeline_steps
        line["pipeline"] = {"value": total_bytes / (pms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": pms,
                            "wall_ms_per_step": wall * 1e3, "steps": args.pipeline_steps,
                            "what": "hist_kernel -> ncclAllReduce(256 x u64) -> hb_build_codebook (host) -> "
                                    "ncclAllGather(1 x u64 per rank) -> encode in global phase; two host syncs per step "
                                    "(the codebook is built on the host, as in the reference)"
                                    if comm is not None else
                                    "hist_kernel -> D2H -> hb_build_codebook (host) -> encode; one GPU: no collective"}

    # ---- stitch (N > 1): every shard pushed to its place in ONE stream on GPU 0, and that stream checked -------------
    if comm is not None and not args.no_stitch:
        try:
            cap = total_bits // 32 + 2
            comm.stitch_open(cap, root=0)
            comm.stitch_push(d_local)                                  # warm (peer mappings, NCCL channels)
            B.barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            s0.record()
            for _ in range(reps):
                comm.stitch_push(d_local)
            s1.record()
            torch.cuda.synchronize()
            B.barrier()
            sms = B.max_over_ranks(s0.elapsed_time(s1)) / reps
            stream_bytes = (total_bits + 7) // 8
            moved = stream_bytes - B.max_over_ranks(out_bytes if rank == 0 else 0)      # what crosses NVLink
            exact = None
            if rank == 0 and args.bytes is None:
                exact = B.check_sums(args.workload, comm.stitched_view(total_bits // 32 + 1), total_bits, "stitched stream")
            flag = torch.tensor([-1 if exact is None else int(exact)], dtype=torch.int64, device="cuda")
            dist.broadcast(flag, src=0)
            exact = None if int(flag[0]) < 0 else bool(int(flag[0]))
            line["stitch"] = {"ms": sms, "stream_bytes": stream_bytes, "GBps_of_stream": stream_bytes / (sms * 1e-3) / 1e9,
                              "GBps_over_nvlink": moved / (sms * 1e-3) / 1e9,
                              "what": "hb_stitch_push: all ranks store their words straight into GPU 0's IPC-mapped "
                                      "buffer at once (peer stores over NVLink), seam words OR-ed by their owner; "
                                      "not part of `value`"}
            line["stitched_bit_exact"] = exact
            line["parity"]["stitched_stream_vs_cpu_vlc_encode"] = exact
            # the fused form: every rank encodes STRAIGHT into GPU 0's stream (the encode kernel's copy-out stores go to
            # peer memory, seam words OR-ed): encode + stitch in one kernel per rank, no local output, no second pass
            if rank == 0:
                comm.stitched_view(cap).fill_(0x5A5A5A5A)
            comm.encode_direct_async(d_in, cw, cl)                     # warm
            assert comm.encode_result() == my_bits
            B.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                comm.encode_direct_async(d_in, cw, cl)
            f1.record()
            assert comm.encode_result() == my_bits
            torch.cuda.synchronize()
            B.barrier()
            fms = B.max_over_ranks(f0.elapsed_time(f1)) / reps
            fexact = None
            if rank == 0 and args.bytes is None:
                fexact = B.check_sums(args.workload, comm.stitched_view(total_bits // 32 + 1), total_bits, "fused encode+stitch stream")
            flag = torch.tensor([-1 if fexact is None else int(fexact)], dtype=torch.int64, device="cuda")
            dist.broadcast(flag, src=0)
            fexact = None if int(flag[0]) < 0 else bool(int(flag[0]))
            line["stitch_fused"] = {"ms": fms, "value": total_bytes / (fms * 1e-3) / 1e9, "unit": UNIT,
                                    "GBps_over_nvlink": moved / (fms * 1e-3) / 1e9, "bit_exact": fexact,
                                    "what": "hb_shard_encode_direct_async: encode + stitch as ONE kernel per rank (copy-out "
                                            "stores straight into GPU 0's stream over NVLink); compare with ms_per_step + "
                                            "stitch.ms for the two-step form"}
            line["parity"]["fused_stream_vs_cpu_vlc_encode"] = fexact
            comm.stitch_close()
        except Exception as exc:
            line["stitch"] = {"error": repr(exc)}
            B.errors.append("stitch: %r" % (exc,))

    # ---- e2e through the host-buffer C-ABI call (rank-local, wall clock, pinned host buffers) ----------------
    e2e_words = None
    if not args.no_e2e and B.sum_over_ranks(1 if n_bytes else 0) == world:
        e2e_steps = args.e2e_steps or (3 if n_bytes >= (4 << 30) else max(1, min(args.steps, 10)))
        pin_in = hb.PinnedBuffer(n_bytes)
        pin_out = hb.PinnedBuffer((my_bits // 32 + 2) * 4)
        torch.cuda.synchronize()
        step_b = 1 << 30
        for a in range(0, n_bytes, step_b):
            pin_in.u8[a:a + step_b] = d_in[a:a + step_b].cpu().numpy()
        h_in = pin_in.u8.view(np.uint32)
        h_out = pin_out.u8.view(np.uint32)
        enc.encode_host(h_in, cw, cl, h_out)                           # warm (allocates device buffers)
        B.barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            eb, _ = enc.encode_host(h_in, cw, cl, h_out)
        wall = B.max_over_ranks(time.perf_counter() - w0)
        assert eb == my_bits
        line["e2e"] = {"value": total_bytes * e2e_steps / wall / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": total_bytes,
                       "d2h_bytes_per_step": B.sum_over_ranks((my_bits // 32 + 1) * 4 + 16),
                       "steps": e2e_steps, "api": "hb_vlc_encode_host (pinned host in/out), every rank its shard"}
        # the host path's stream (phase 0) against the device path's: same words when this shard starts on a word boundary
        same = True
        if start_bit == 0:
            n_cmp = my_bits // 32
            for a in range(0, n_cmp, 1 << 28):
                b = min(n_cmp, a + (1 << 28))
                same = same and bool(torch.equal(torch.from_numpy(h_out[a:b].view(np.int32)).cuda(), d_out[a:b]))
            if not same:
                B.errors.append("rank %d: hb_vlc_encode_host stream differs from hb_encode's" % rank)
        line["parity"]["e2e_stream_equals_device_stream"] = bool(B.sum_over_ranks(0 if same else 1) == 0)
        e2e_words = h_out[: min(h_out.size, (64 << 20))].copy()
        pin_in.free()
        pin_out.free()

    # ---- cpu_baseline: the reference's cpu_vlc_encode on this box, rank 0, N=1 only; doubles as a parity check --
    if rank == 0 and world == 1 and not args.no_cpu:
        run, kind = cpu_reference_encoder()
        sample = min(n_bytes, args.cpu_sample_mib << 20)
        sample -= sample % TILE_BYTES if sample >= TILE_BYTES else 0
        host = d_in[:sample].cpu().numpy()
        shist = np.bincount(host, minlength=256).astype(np.uint64)
        sbits = hb.bits_from_hist(shist, cl)
        ts, cpu_out = [], None
        for _ in range(args.cpu_repeats):
            c0 = time.perf_counter()
            cpu_out, _ = run(host.view(np.uint32), cw, cl, sbits)
            ts.append(time.perf_counter() - c0)
        secs = float(np.median(ts))
        got = d_out[: sbits // 32].cpu().numpy().view(np.uint32)       # the prefix is position independent
        exact = bool(np.array_equal(got, cpu_out[: sbits // 32]))
        if e2e_words is not None:
            n_cmp = min(sbits // 32, e2e_words.size)
            exact = exact and bool(np.array_equal(e2e_words[:n_cmp], cpu_out[:n_cmp]))
        line["cpu_baseline"] = {
            "value": sample / secs / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d MiB of the workload, median of %d runs; cpu_vlc_encode is serial "
                      "(1 of %d host cores)" % (sample >> 20, args.cpu_repeats, os.cpu_count()),
            "bit_exact_vs_gpu": exact}
        if not exact:
            B.errors.append("GPU stream differs from cpu_vlc_encode on the first %d MiB" % (sample >> 20))

    # ---- per_config (N = 1): the other BASELINE configurations, the 1 GiB H2.2 target case first -----------------
    launches_extra = 0
    if world == 1 and not args.no_per_config and args.workload == "c5" and args.bytes is None:
        del d_in, d_out
        enc.close()
        enc = None
        torch.cuda.empty_cache()
        line["per_config"] = {}
        for name in ("t1g", "c2", "c3", "c4"):
            try:
                line["per_config"][name] = B.single_config(name, args.per_config_steps, args.warmup)
            except Exception as exc:
                line["per_config"][name] = {"error": repr(exc)}
                B.errors.append("per_config %s: %r" % (name, exc))

    if B.errors:
        line["parity_error"] = B.errors
    n_err = B.sum_over_ranks(len(B.errors))
    if rank == 0:
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if enc is not None:
        enc.close()
    if dist is not None:
        dist.destroy_process_group()
    return 1 if n_err else 0


if __name__ == "__main__":
    sys.exit(main())
