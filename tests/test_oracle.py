"""The CPU oracle (oracle/oracle.c) pinned against the reference's own outputs: the golden vectors
generated from the unmodified cpuencode.cpp / huffTree.h (oracle/make_golden.py), and -- when the
prebuilt oracle/_ref/libref.so is present -- the reference itself on fresh random inputs."""
import hashlib
import os

import numpy as np
import pytest

from conftest import load_golden


def fnv(words):
    h = 1469598103934665603
    for w in words.tolist():
        h = ((h ^ w) * 1099511628211) & (2 ** 64 - 1)
    return h


def test_c1_fixture_reassembles(hb, c1):
    data = hb.workloads.c1_fixture_bytes()
    assert data.size == c1["n_bytes"] == 1048576
    assert hashlib.sha256(data.tobytes()).hexdigest() == c1["sha256"]


def test_oracle_c1_golden(hb, orc, c1):
    """total_bits 2,330,672; outsize 291,334 B; 72,834 words; hash 0x6774223e44ca33fb."""
    data = hb.workloads.c1_fixture_bytes()
    hist = orc.histogram(data)
    assert np.array_equal(hist, c1["freqs"])
    rc, cw, cl = orc.build_codebook(hist)
    assert rc == c1["max_len"] == 21
    assert np.array_equal(cw, c1["codewords"]) and np.array_equal(cl, c1["codewordlens"])
    words, bits, nbytes = orc.encode(data.view(np.uint32), cw, cl)
    assert bits == 2330672 and nbytes == 291334
    nw = (nbytes + 3) // 4
    assert nw == 72834
    assert [int(x) for x in words[:4]] == [0xB3DF4F4A, 0x53CBF8F2, 0x064F5E2D, 0xB63C6746]
    assert [int(x) for x in words[:8]] == c1["first_words"]
    assert [int(x) for x in words[nw - 4:nw]] == c1["last_words"]
    assert orc.word_fnv(words[:nw]) == 0x6774223E44CA33FB == c1["fnv"]
    assert fnv(words[:nw]) == c1["fnv"]


def test_oracle_kats(orc):
    g = load_golden("kat.json")
    cw = np.array(g["codewords"], dtype=np.uint32)
    cl = np.array(g["codewordlens"], dtype=np.uint32)
    for case in g["cases"]:
        if not case["parity_domain"]:
            continue       # documents a defect of the reference (length 32 / dirty codewords)
        w = np.array(case["in"], dtype=np.uint32)
        words, bits, nbytes = orc.encode(w, cw, cl)
        assert nbytes == case["outsize_bytes"], case["name"]
        gold = case["out_words"]          # sz//4+1 words of the reference's (zeroed) output buffer
        n = min(len(gold), words.size)
        assert [int(x) for x in words[:n]] == gold[:n], case["name"]
        assert not any(gold[n:]) and not words[n:].any(), case["name"]


def test_oracle_rejects_len32(orc):
    cw = np.zeros(256, dtype=np.uint32)
    cl = np.zeros(256, dtype=np.uint32)
    cl[1] = 32
    with pytest.raises(ValueError):
        orc.encode(np.array([0x01010101], dtype=np.uint32), cw, cl, total_bits_hint=128)


def test_oracle_encode_cases(orc):
    for i, case in enumerate(load_golden("encode_cases.json")):
        w = np.array(case["in"], dtype=np.uint32)
        cw = np.array(case["codewords"], dtype=np.uint32)
        cl = np.array(case["codewordlens"], dtype=np.uint32)
        words, bits, nbytes = orc.encode(w, cw, cl)
        assert bits == case["total_bits"] and nbytes == case["outsize_bytes"], i
        assert [int(x) for x in words] == case["out_words"], i


def test_oracle_codebooks(orc):
    for i, case in enumerate(load_golden("codebooks.json")):
        rc, cw, cl = orc.build_codebook(np.array(case["hist"], dtype=np.uint64))
        assert rc == case["rc"], i
        assert cw.tolist() == case["codewords"] and cl.tolist() == case["codewordlens"], i


def test_oracle_vs_reference_live(orc, ref):
    """Fresh random inputs through the unmodified reference (skipped when _ref is not shipped)."""
    if ref is None:
        pytest.skip("oracle/_ref/libref.so not present")
    rng = np.random.default_rng(1234)
    for it in range(200):
        nsym = int(rng.integers(1, 257))
        h = np.zeros(256, dtype=np.uint32)
        syms = rng.choice(256, size=nsym, replace=False)
        h[syms] = rng.integers(1, [4, 100, 1 << 16, 1 << 24][it % 4], size=nsym)
        rc_r, cw_r, cl_r = ref.build_codebook(h)
        rc_o, cw_o, cl_o = orc.build_codebook(h.astype(np.uint64))
        assert rc_r == rc_o and np.array_equal(cw_r, cw_o) and np.array_equal(cl_r, cl_o), it
        if rc_r > 31:
            continue
        n_words = int(rng.integers(0, 3000))
        p = h / h.sum()
        data = rng.choice(256, size=n_words * 4, p=p).astype(np.uint8)
        w = data.view(np.uint32)
        words, bits, nbytes = orc.encode(w, cw_o, cl_o)
        out_r, size_r = ref.encode(w, cw_r, cl_r, bits // 32 + 2)
        assert size_r == nbytes
        assert np.array_equal(out_r[: bits // 32 + 1], words), it


def test_oracle_decode_roundtrip(orc):
    rng = np.random.default_rng(7)
    for it in range(20):
        h = np.zeros(256, dtype=np.uint64)
        nsym = int(rng.integers(1, 200))
        h[rng.choice(256, size=nsym, replace=False)] = rng.integers(1, 1000, size=nsym)
        rc, cw, cl = orc.build_codebook(h)
        n_words = int(rng.integers(1, 500))
        data = rng.choice(256, size=n_words * 4, p=h / h.sum()).astype(np.uint8)
        words, bits, _ = orc.encode(data.view(np.uint32), cw, cl)
        back, end = orc.decode(words, 0, n_words * 4, cw, cl)
        assert end == bits
        assert np.array_equal(back, data)


def test_synth_generator_statistics(hb, orc):
    for name, H in (("c2", 2.2), ("c3", 7.9), ("c5", 4.0)):
        w = hb.workloads.get(name)
        assert abs(w.entropy_bits() - H) < 0.01
        a = orc.synth_fill(0, 1 << 20, w.seed, w.mode, w.nbits, w.thr)
        b = orc.synth_fill(1 << 19, 1 << 19, w.seed, w.mode, w.nbits, w.thr)
        assert np.array_equal(a[1 << 19:], b)          # counter-based: any window reproduces
        p = np.bincount(a, minlength=256) / a.size
        nz = p > 0
        Hs = -(p[nz] * np.log2(p[nz])).sum()
        assert abs(Hs - H) < 0.02, (name, Hs)
        hist = orc.histogram(a)
        rc, cw, cl = orc.build_codebook(hist)
        assert 0 < rc <= 24


def test_synth_c4_exact_counts(hb, orc):
    """mode 1 is a bijection: over a full power-of-two range every symbol count is exact."""
    counts = np.array([5, 3, 6, 2], dtype=np.uint64)       # sums to 16 = 2^4
    thr = np.cumsum(counts).astype(np.uint32)
    a = orc.synth_fill(0, 16, 99, 1, 4, thr)
    assert np.bincount(a, minlength=4).tolist() == counts.tolist()
    w = hb.workloads.get("c4")
    assert int(hb.workloads.fibonacci_counts().sum()) == 2 ** 31
    g = [c for c in load_golden("codebooks.json") if c.get("name") == "c4_fibonacci"][0]
    assert g["rc"] == 31 and sorted(set(g["codewordlens"])) == list(range(0, 32))
    assert w.nbits == 31 and w.thr.size == 32


def test_golden_stream_c2_restatement_at_full_size(orc):
    """tests/golden/streams.json (made by oracle/make_golden_streams.py from the UNMODIFIED cpu_vlc_encode): the
    restatement reproduces the C2 entry -- generator, histogram, tree builder, encoder, and the numpy side of the
    parallel checksums the GPU tests and bench.py use at full size."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mods = {}
    for name in ("workloads", "streamsum"):
        spec = importlib.util.spec_from_file_location("hb_" + name, os.path.join(root, "huffman-gpu_b200", name + ".py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    g = load_golden("streams.json")
    assert set(g) >= {"c2", "t1g", "c3", "c4", "c5"}
    for e in g.values():
        assert e["encoder"].startswith("cpu_vlc_encode (cpuencode.cpp:12-46, unmodified")
        assert sum(e["hist"]) == e["n_bytes"]
        assert sum(h * l for h, l in zip(e["hist"], e["codewordlens"])) == e["total_bits"]
    wl = mods["workloads"].get("c2")
    data = orc.synth_fill(0, wl.n_bytes, wl.seed, wl.mode, wl.nbits, wl.thr)
    hist = orc.histogram(data)
    assert hist.tolist() == g["c2"]["hist"]
    rc, cw, cl = orc.build_codebook(hist)
    assert cl.tolist() == g["c2"]["codewordlens"] and cw.tolist() == g["c2"]["codewords"]
    out, bits, _ = orc.encode(data.view(np.uint32), cw, cl, total_bits_hint=g["c2"]["total_bits"])
    assert bits == g["c2"]["total_bits"] and out.size == g["c2"]["n_words"]
    assert "0x%016x" % orc.word_fnv(out) == g["c2"]["word_fnv"]
    assert ["0x%016x" % s for s in mods["streamsum"].stream_sums(out)] == g["c2"]["sums"]
