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
hist = enc.histogram(d)
cw, cl, ml = hb.build_codebook(hist)
bits = hb.bits_from_hist(hist, cl)
out = torch.empty(bits // 32 + 2, dtype=torch.int32, device='cuda')
for _ in range(3):
    enc.encode(d, cw, cl, out)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    enc.encode_async(d, cw, cl, out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print("%s variant %s: %.4f ms/step, %.1f GB/s input, %.1f GB/s traffic" % (name, hb.encode_variant(cl), ms, wl.n_bytes / ms / 1e6, (wl.n_bytes + bits / 8) / ms / 1e6))
enc.encode_result()
enc.close()
