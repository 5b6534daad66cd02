"""times hb_histogram_device on a synthetic workload: tools/hist_run.py [c2|c3|t1g|...] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import huffman_gpu_b200 as hb
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wl = hb.workloads.get(name)
enc = hb.Encoder(0, wl.n_bytes)
d = torch.empty(wl.n_bytes, dtype=torch.uint8, device='cuda')
enc.synth_fill(d, wl)
ref = torch.bincount(d[: 1 << 26].to(torch.int64), minlength=256).cpu().numpy()
h = torch.zeros(256, dtype=torch.int64, device='cuda')
enc.histogram_device(d[: 1 << 26], h)
torch.cuda.synchronize()
assert np.array_equal(h.cpu().numpy(), ref), "histogram mismatch"
for _ in range(3):
    enc.histogram_device(d, h)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    enc.histogram_device(d, h)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print("%s histogram: %.4f ms/step, %.1f GB/s" % (name, ms, wl.n_bytes / ms / 1e6))
enc.close()
