"""Small multi-tile jobs through every kernel family, for compute-sanitizer (tools/sanitize.sh): hist_kernel, one
packed encode variant without and with the over-long-group check, one wide variant, a ragged end, a sharded phase
(start_bit != 0) and the chunked host path.  Every result is compared with the CPU oracle, so a run that is clean under
the sanitizer is also a correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import pyoracle
import huffman_gpu_b200 as hb

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 24
orc = pyoracle.Oracle()
torch.cuda.set_device(0)
enc = hb.Encoder(0, 64 << 20)
for name, extra in (("c2", 0), ("c3", 0), ("c4", 0), ("c5", 4 * 777)):
    wl = hb.workloads.get(name)
    n = tiles * 32768 + extra
    data = orc.synth_fill(0, n, wl.seed, wl.mode, wl.nbits, wl.thr, wl.symmap)
    d_in = torch.from_numpy(data).cuda()
    hist = enc.histogram(d_in)
    assert np.array_equal(hist, orc.histogram(data)), name
    if name == "c4":
        hist = np.zeros(256, dtype=np.uint64)
        hist[:32] = hb.workloads.fibonacci_counts()                  # the full-size counts: code lengths 1..31 (wide kernel)
    cw, cl, _ = hb.build_codebook(hist)
    for start_bit in (0, 13):
        ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
        d_out = torch.full(((start_bit + ref_bits) // 32 + 2,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        bits = enc.encode(d_in, cw, cl, d_out, start_bit=start_bit)
        torch.cuda.synchronize()
        assert bits == ref_bits
        if start_bit == 0:
            got = d_out.cpu().numpy().view(np.uint32)
            assert np.array_equal(got[: ref_words.size], ref_words), name
    print(name, hb.encode_variant(cl), "ok", flush=True)
# the chunked host path (several launches of one job)
os.environ["HB_CHUNK_MIB"] = "1"
wl = hb.workloads.get("c2")
data = orc.synth_fill(0, 3 << 20, wl.seed, wl.mode, wl.nbits, wl.thr)
cw, cl, _ = hb.build_codebook(orc.histogram(data))
ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
h_out = np.zeros(ref_words.size + 1, dtype=np.uint32)
bits, _ = enc.encode_host(data.view(np.uint32), cw, cl, h_out)
assert bits == ref_bits and np.array_equal(h_out[: ref_words.size], ref_words)
print("host path ok", flush=True)
enc.close()
