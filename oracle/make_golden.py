#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Outputs (all small, committed):
  tests/golden/c1_period.bin.xz   one 262,144-byte period of data/test1024_H2.206587175259.in
                                  (the fixture is exactly 4 repetitions; sha256 checked below)
  tests/golden/c1_fixture.json    reference codebook + stream facts for the fixture
  tests/golden/kat.json           hand-made-table known-answer tests
  tests/golden/codebooks.json     random histograms -> huffTree.h codewords/lengths
  tests/golden/encode_cases.json  small random inputs/tables -> cpu_vlc_encode words/outsize
TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pyoracle  # noqa: E402

REF_DIR = "/root/reference"
FIXTURE = os.path.join(REF_DIR, "data", "test1024_H2.206587175259.in")
OUT = os.path.join(HERE, "..", "tests", "golden")
FIXTURE_SHA256 = "813b542f2dfabf07500689b17553a2eff0bb882dce3c56ba28e03496b6301599"


def fnv(words):
    h = 1469598103934665603
    for w in words.tolist():
        h ^= w
        h = (h * 1099511628211) & (2 ** 64 - 1)
    return h


def main():
    os.makedirs(OUT, exist_ok=True)
    pyoracle.build(quiet=False)
    ref = pyoracle.Ref()

    # ---- C1 fixture -------------------------------------------------------------------
    raw = open(FIXTURE, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == FIXTURE_SHA256
    period = 262144
    assert raw == raw[:period] * 4
    with open(os.path.join(OUT, "c1_period.bin.xz"), "wb") as f:
        f.write(lzma.compress(raw[:period], preset=9 | lzma.PRESET_EXTREME))
    data = np.frombuffer(raw, dtype=np.uint8)
    freqs = np.bincount(data, minlength=256).astype(np.uint32)      # true CPU byte histogram
    rc, cw, cl = ref.build_codebook(freqs)
    words = data.view(np.uint32)
    out, outsize = ref.encode(words, cw, cl, words.size + 2)
    total_bits = int((freqs.astype(np.uint64) * cl).sum())
    assert (total_bits + 7) // 8 == outsize
    nw = (outsize + 3) // 4
    c1 = {
        "file": "data/test1024_H2.206587175259.in", "sha256": FIXTURE_SHA256,
        "n_bytes": len(raw), "period": period, "repeats": 4,
        "max_len": rc, "total_bits": total_bits, "outsize_bytes": outsize, "n_words": nw,
        "first_words": [int(x) for x in out[:8]], "last_words": [int(x) for x in out[nw - 4:nw]],
        "fnv": fnv(out[:nw]), "freqs": freqs.tolist(), "codewords": cw.tolist(),
        "codewordlens": cl.tolist(),
    }
    assert c1["total_bits"] == 2330672 and c1["outsize_bytes"] == 291334 and nw == 72834
    assert c1["fnv"] == 0x6774223E44CA33FB, hex(c1["fnv"])
    json.dump(c1, open(os.path.join(OUT, "c1_fixture.json"), "w"))

    # ---- KATs -----------------------------------------------------
    cwt = np.zeros(256, dtype=np.uint32)
    clt = np.zeros(256, dtype=np.uint32)
    cwt[1], clt[1] = 0xFFFFFFFF, 32
    cwt[2], clt[2] = 1, 1
    cwt[3], clt[3] = 0x7FFFFFFF, 31
    cwt[4], clt[4] = 0b101, 3
    cwt[5], clt[5] = 0xF5, 3
    kats = []
    for name, inp, in_domain in [
        ("len32_aligned_oracle_defect", [0x01010101], False),
        ("len32_unaligned", [0x02010101], False),
        ("len31", [0x03030303], True),
        ("len0_skipped", [0x04000400, 0x00000004], True),
        ("dirty_high_bits_oracle_defect", [0x05050505], False),
        ("empty", [], True),
        ("len1_x32_word_aligned_end", [0x02020202] * 8, True),
    ]:
        w = np.array(inp, dtype=np.uint32)
        o, sz = ref.encode(w, cwt, clt, 64)
        kats.append({"name": name, "in": [int(x) for x in w], "outsize_bytes": sz,
                     "out_words": [int(x) for x in o[: sz // 4 + 1]], "parity_domain": in_domain})
    json.dump({"codewords": cwt.tolist(), "codewordlens": clt.tolist(), "cases": kats},
              open(os.path.join(OUT, "kat.json"), "w"))

    # ---- random histograms -> huffTree.h codebooks -------------------------------------
    rng = np.random.default_rng(0xB200)
    books = []
    for case in range(48):
        nsym = int(rng.integers(1, 257))
        syms = rng.choice(256, size=nsym, replace=False)
        kind = case % 4
        if kind == 0:    # heavy ties
            f = rng.integers(1, 4, size=nsym)
        elif kind == 1:  # geometric
            f = np.maximum(1, (2.0 ** 24 * rng.uniform(0.4, 0.95) ** np.arange(nsym))).astype(np.int64)
        elif kind == 2:  # wide uniform
            f = rng.integers(1, 1 << 20, size=nsym)
        else:            # fibonacci-like prefix (deep trees), capped so lengths <= 32
            fib = [1, 1]
            while len(fib) < min(nsym, 30):
                fib.append(fib[-1] + fib[-2])
            f = np.array((fib * 9)[:nsym])
        h = np.zeros(256, dtype=np.uint32)
        h[syms] = f
        rc, cwb, clb = ref.build_codebook(h)
        books.append({"hist": h.tolist(), "rc": rc, "codewords": cwb.tolist(),
                      "codewordlens": clb.tolist()})
    # the C4 histogram: 376*Fib(k), k=1..32, remainder on the most frequent symbol
    fib = [1, 1]
    while len(fib) < 32:
        fib.append(fib[-1] + fib[-2])
    h = np.zeros(256, dtype=np.uint32)
    h[:32] = 376 * np.array(fib, dtype=np.uint64)
    h[31] += 2 ** 31 - int(h.astype(np.uint64).sum())
    assert int(h.astype(np.uint64).sum()) == 2 ** 31
    # total == 2^31 overflows the reference's `int f` at the ROOT only (never compared again),
    # so huffTree.h still yields the intended tree.
    rc, cwb, clb = ref.build_codebook(h)
    books.append({"hist": h.tolist(), "rc": rc, "codewords": cwb.tolist(),
                  "codewordlens": clb.tolist(), "name": "c4_fibonacci"})
    json.dump(books, open(os.path.join(OUT, "codebooks.json"), "w"))

    # ---- small encode cases -> cpu_vlc_encode ------------------------------------------
    cases = []
    for case in range(24):
        n_words = int(rng.integers(0, 600))
        maxlen = [3, 8, 12, 16, 24, 31][case % 6]
        clr = rng.integers(0 if case % 5 == 0 else 1, maxlen + 1, size=256).astype(np.uint32)
        cwr = np.array([int(rng.integers(0, 1 << int(l))) if l else 0 for l in clr], dtype=np.uint32)
        w = rng.integers(0, 2 ** 32, size=n_words, dtype=np.uint64).astype(np.uint32)
        if case % 3 == 0:   # skewed bytes
            b = (rng.geometric(0.4, size=n_words * 4) - 1).clip(0, 255).astype(np.uint8)
            w = b.view(np.uint32).copy()
        bits = int(clr.astype(np.uint64)[w.view(np.uint8)].sum()) if n_words else 0
        o, sz = ref.encode(w, cwr, clr, bits // 32 + 2)
        assert sz == (bits + 7) // 8
        cases.append({"in": [int(x) for x in w], "codewords": cwr.tolist(),
                      "codewordlens": clr.tolist(), "total_bits": bits, "outsize_bytes": sz,
                      "out_words": [int(x) for x in o[: bits // 32 + 1]]})
    json.dump(cases, open(os.path.join(OUT, "encode_cases.json"), "w"))
    print("golden vectors written to", os.path.normpath(OUT))
    for fn in sorted(os.listdir(OUT)):
        print("  %-24s %8d B" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
