#!/usr/bin/env python
"""Golden facts of the FULL-SIZE BASELINE.json streams (C2, 1 GiB H2.2, C3, C4, C5), from the UNMODIFIED reference.

  python oracle/make_golden_streams.py [c2 t1g c3 c4 c5]        (build container; ~10 min, ~14 GB of RAM for c5)

For every workload: the synthetic input is produced by the CPU generator (oracle.c orc_synth_fill, the independent
restatement of the device generator), histogrammed on the CPU, the codebook built by the restated tree builder
(checked against the unmodified huffTree.h whenever the total fits its `int` weights), and the whole stream encoded by
the unmodified cpu_vlc_encode (oracle/_ref/libref.so).  tests/golden/streams.json then holds, per workload: sizes, the
histogram, the code lengths, total bits, the word FNV and the parallel checksums of
huffman-gpu_b200/streamsum.py over words [0, total_bits/32] -- what bench.py and the -m gpu tests compare the CUDA
streams (single-GPU and stitched multi-GPU) against at full size.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import pyoracle  # noqa: E402


def generate(orc, wl, threads):
    n = wl.n_bytes
    out = np.empty(n, dtype=np.uint8)
    piece = 1 << 24

    def fill(lo):
        hi = min(n, lo + piece)
        out[lo:hi] = orc.synth_fill(lo, hi - lo, wl.seed, wl.mode, wl.nbits, wl.thr, wl.symmap)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(fill, range(0, n, piece)))
    return out


def main():
    # the workload DEFINITIONS (sizes, seeds, thresholds) are data, not code under test: import the module alone
    import importlib.util
    spec = importlib.util.spec_from_file_location("hb_workloads", os.path.join(ROOT, "huffman-gpu_b200", "workloads.py"))
    workloads = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(workloads)
    spec = importlib.util.spec_from_file_location("hb_streamsum", os.path.join(ROOT, "huffman-gpu_b200", "streamsum.py"))
    streamsum = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(streamsum)

    names = sys.argv[1:] or ["c2", "t1g", "c3", "c4", "c5"]
    path = os.path.join(ROOT, "tests", "golden", "streams.json")
    golden = json.load(open(path)) if os.path.exists(path) else {}
    pyoracle.build()
    orc, ref = pyoracle.Oracle(), pyoracle.Ref()
    threads = os.cpu_count()
    for name in names:
        wl = workloads.get(name)
        t0 = time.time()
        data = generate(orc, wl, threads)
        hist = np.bincount(data[: 1 << 30], minlength=256).astype(np.uint64)
        for lo in range(1 << 30, data.size, 1 << 30):
            hist += np.bincount(data[lo:lo + (1 << 30)], minlength=256).astype(np.uint64)
        rc, cw, cl = orc.build_codebook(hist)
        assert rc >= 0
        if int(hist.sum()) <= 2 ** 31 - 1:                      # huffTree.h keeps weights in `int`
            rc2, cw2, cl2 = ref.build_codebook(hist.astype(np.uint32))
            assert np.array_equal(cw, cw2) and np.array_equal(cl, cl2), "restated tree builder != huffTree.h"
            tree = "huffTree.h (unmodified) == oracle.c"
        else:
            tree = "oracle.c (64-bit weights; huffTree.h's int weights overflow at this size)"
        bits = int((hist * cl.astype(np.uint64)).sum())
        t1 = time.time()
        words = data.view(np.uint32)
        out, outsize = ref.encode(words, cw, cl, bits // 32 + 2)       # the unmodified cpu_vlc_encode
        t2 = time.time()
        assert outsize == ((bits + 7) // 8) % (1 << 32)                # (uint32 bytes: wraps at 4 GiB, cpuencode.cpp:45)
        n_words = bits // 32 + 1
        assert out[n_words:].sum() == 0
        s1, s2, s3 = streamsum.stream_sums(out, n_words)
        golden[name] = {
            "note": wl.note, "n_bytes": int(wl.n_bytes), "seed": int(wl.seed), "mode": int(wl.mode),
            "hist": [int(x) for x in hist], "codewordlens": [int(x) for x in cl],
            "codewords": [int(x) for x in cw], "codebook_from": tree, "max_len": int(cl.max()),
            "total_bits": bits, "n_words": n_words, "word_fnv": "0x%016x" % orc.word_fnv(out[:n_words]),
            "sums": ["0x%016x" % s1, "0x%016x" % s2, "0x%016x" % s3],
            "first_words": ["%08x" % int(x) for x in out[:4]],
            "last_words": ["%08x" % int(x) for x in out[n_words - 4:n_words]],
            "encoder": "cpu_vlc_encode (cpuencode.cpp:12-46, unmodified, oracle/_ref/libref.so)",
            "cpu_seconds": {"generate+hist": round(t1 - t0, 1), "cpu_vlc_encode": round(t2 - t1, 1)},
        }
        print(name, bits, golden[name]["word_fnv"], golden[name]["sums"], golden[name]["cpu_seconds"], flush=True)
        del data, words, out
        with open(path, "w") as f:
            json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()
