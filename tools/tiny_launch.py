import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import huffman_gpu_b200 as hb
for n in [32768, 32768*8, 32768*148]:
    wl = hb.workloads.get('c2', n)
    enc = hb.Encoder(0, wl.n_bytes)
    d = torch.empty(wl.n_bytes, dtype=torch.uint8, device='cuda')
    enc.synth_fill(d, wl)
    hist = enc.histogram(d)
    cw, cl, ml = hb.build_codebook(hist)
    bits = hb.bits_from_hist(hist, cl)
    out = torch.empty(bits // 32 + 2, dtype=torch.int32, device='cuda')
    for _ in range(5): enc.encode(d, cw, cl, out)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): enc.encode_async(d, cw, cl, out)
    b.record(); torch.cuda.synchronize()
    print(n, "bytes: %.2f us per launch" % (a.elapsed_time(b) / 200 * 1e3))
    enc.encode_result(); enc.close()
# an empty torch kernel back to back for comparison
x = torch.zeros(1, device='cuda')
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): x.add_(1)
b.record(); torch.cuda.synchronize()
print("tiny torch kernel: %.2f us per launch" % (a.elapsed_time(b) / 200 * 1e3))
