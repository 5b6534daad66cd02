"""encode time vs input size (same distribution): the intercept is the fixed cost of a launch.
usage: tools/size_sweep.py [workload] [tiles-per-CTA list] [series-dir]
With a series directory the samples are also written in the reference's stats_logger format (LogStats2: time over
data size, and the derived data-rate series; huffman_gpu_b200.stats)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import huffman_gpu_b200 as hb
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
tiles = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 4, 8, 16, 32, 55]
series_dir = sys.argv[3] if len(sys.argv) > 3 else None
sm = torch.cuda.get_device_properties(0).multi_processor_count
for t in tiles:
    n = t * sm * hb.capi.TILE_BYTES if hasattr(hb.capi, 'TILE_BYTES') and hb.capi.TILE_BYTES else t * sm * 32768
    wl = hb.workloads.get(name, n)
    enc = hb.Encoder(0, wl.n_bytes)
    d = torch.empty(wl.n_bytes, dtype=torch.uint8, device='cuda')
    enc.synth_fill(d, wl)
    hist = enc.histogram(d)
    cw, cl, ml = hb.build_codebook(hist)
    bits = hb.bits_from_hist(hist, cl)
    out = torch.empty(bits // 32 + 2, dtype=torch.int32, device='cuda')
    for _ in range(5):
        enc.encode(d, cw, cl, out)
    steps = 50
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        enc.encode_async(d, cw, cl, out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print("%s %3d tiles/CTA %8.2f MiB: %8.2f us/launch  %7.1f GB/s" % (name, t, n / 2**20, ms * 1e3, n / ms / 1e6))
    if series_dir:
        hb.stats.log_stats2(series_dir, "encode_" + name, "B200_single_pass", ms, n / 2.0 ** 20,
                            description="hb_encode, device-resident, mean of %d launches" % steps)
    enc.encode_result(); enc.close()
