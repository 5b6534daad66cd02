"""times the tile-parallel decoder on a synthetic workload: tools/decode_run.py [c2|t1g|c3|c4|c5]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import huffman_gpu_b200 as hb
name = sys.argv[1] if len(sys.argv) > 1 else 't1g'
wl = hb.workloads.get(name)
enc = hb.Encoder(0, wl.n_bytes)
d = torch.empty(wl.n_bytes, dtype=torch.uint8, device='cuda')
enc.synth_fill(d, wl)
cw, cl, ml = hb.build_codebook(enc.histogram(d))
bits = hb.bits_from_hist(enc.histogram(d), cl)
out = torch.empty(bits // 32 + 2, dtype=torch.int32, device='cuda')
assert enc.encode(d, cw, cl, out) == bits
n_tiles = (wl.n_bytes + hb.capi.TILE_BYTES - 1) // hb.capi.TILE_BYTES
idx = torch.empty(n_tiles + 1, dtype=torch.int64, device='cuda')
enc.tile_index(bits, idx)
back = torch.empty_like(d)
enc.decode(out, idx, cw, cl, back)
assert torch.equal(back, d)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    enc.decode(out, idx, cw, cl, back)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("%s decode: %.3f ms, %.1f GB/s of symbols (round trip exact)" % (name, ms, wl.n_bytes / ms / 1e6))
enc.close()
