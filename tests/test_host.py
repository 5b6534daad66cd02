"""CPU-side checks of the product library: the ABI, the host codebook builder, the host helpers.
No device compute is attempted here (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_abi_exports_every_declared_symbol(hb):
    header = open(os.path.join(ROOT, "include", "huffman_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = C.CDLL(hb.capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libhuffb200.so does not export " + name
    assert declared == set(hb.capi.SIGNATURES), declared ^ set(hb.capi.SIGNATURES)


def test_abi_only_sm100a_code(hb):
    """the shared library carries exactly one device target: sm_100a (no multi-arch fallback)"""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", hb.capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(hb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hb.HBError) as e:
        hb.Encoder(device=0, max_bytes=1 << 20)
    assert e.value.status == hb.capi.HB_ERR_CUDA
    out = np.zeros(8, dtype=np.uint32)
    with pytest.raises(hb.HBError):
        hb.vlc_encode(np.zeros(4, dtype=np.uint32), 4, out, np.zeros(256, np.uint32),
                      np.ones(256, np.uint32))


def test_cli_driver_builds_and_fails_loudly_without_gpu(hb, tmp_path):
    """pavle_b200 (the reference's command-line driver on top of the C ABI)"""
    import subprocess
    import torch
    cli = os.path.join(ROOT, "huffman-gpu_b200", "pavle_b200")
    assert os.path.exists(cli), "the driver is built by `make -C huffman-gpu_b200/csrc`"
    out = subprocess.run([cli], capture_output=True, text=True)
    assert out.returncode == 2 and "No input file" in out.stdout           # load_data.h:27
    if not torch.cuda.is_available():
        p = tmp_path / "x.in"
        p.write_bytes(bytes(range(256)) * 16)
        out = subprocess.run([cli, str(p)], capture_output=True, text=True)
        assert out.returncode == 3 and "PASS" not in out.stdout              # no device: no fallback, no verdict


def test_product_never_references_oracle():
    """the product path may not import/link/execute anything under oracle/"""
    pkg = os.path.join(ROOT, "huffman-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                for line in text.splitlines():
                    code = line.split("//")[0]
                    if re.search(r"(import|include|dlopen|CDLL).*(pyoracle|liboracle|libref|oracle\.h)", code):
                        raise AssertionError("%s references the oracle: %s" % (fn, line))
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libhuffb200.so")], capture_output=True, text=True)
    assert "oracle" not in out.stdout and "libref" not in out.stdout


def test_codebook_golden(hb):
    """hb_build_codebook == huffTree.h + load_data.h:40-47 on the reference-generated vectors."""
    for i, case in enumerate(load_golden("codebooks.json")):
        cw, cl, rc = hb.build_codebook(np.array(case["hist"], dtype=np.uint64))
        assert rc == case["rc"], i
        assert cw.tolist() == case["codewords"] and cl.tolist() == case["codewordlens"], i


def test_codebook_c1(hb, c1):
    cw, cl, rc = hb.build_codebook(c1["freqs"])
    assert rc == 21
    assert np.array_equal(cw, c1["codewords"]) and np.array_equal(cl, c1["codewordlens"])
    assert hb.bits_from_hist(c1["freqs"], cl) == c1["total_bits"] == 2330672
    assert hb.encode_variant(cl) == "packed_g4c"


def test_codebook_random_vs_oracle_and_reference(hb, orc, ref):
    rng = np.random.default_rng(42)
    for it in range(2000):
        nsym = int(rng.integers(1, 257))
        h = np.zeros(256, dtype=np.uint64)
        hi = [3, 17, 1 << 12, 1 << 30][it % 4]
        h[rng.choice(256, size=nsym, replace=False)] = rng.integers(1, hi, size=nsym)
        rc_o, cw_o, cl_o = orc.build_codebook(h)
        if rc_o > 31 or rc_o < 0:
            with pytest.raises(hb.HBError):
                hb.build_codebook(h)
            continue
        cw, cl, rc = hb.build_codebook(h)
        assert rc == rc_o and np.array_equal(cw, cw_o) and np.array_equal(cl, cl_o), it
        if ref is not None and it % 10 == 0 and int(h.sum()) < 2 ** 31:
            rc_r, cw_r, cl_r = ref.build_codebook(h.astype(np.uint32))
            assert rc == rc_r and np.array_equal(cw, cw_r) and np.array_equal(cl, cl_r), it


def test_codebook_edges(hb):
    z = np.zeros(256, dtype=np.uint64)
    cw, cl, rc = hb.build_codebook(z)
    assert rc == 0 and not cw.any() and not cl.any()          # empty input
    z[200] = 12345
    cw, cl, rc = hb.build_codebook(z)
    assert rc == 0 and not cw.any() and not cl.any()          # one symbol -> length 0 (huffTree.h root leaf)
    # 64-bit weights: counts above INT_MAX (the reference's `int f` wraps)
    big = np.zeros(256, dtype=np.uint64)
    big[0], big[1], big[2] = 0x90000000, 5, 7
    cw, cl, rc = hb.build_codebook(big)
    assert cl[0] == 1 and cl[1] == 2 and cl[2] == 2
    # a 33-symbol Fibonacci chain needs a 32-bit code: refused
    fib = [1, 1]
    while len(fib) < 34:
        fib.append(fib[-1] + fib[-2])
    deep = np.zeros(256, dtype=np.uint64)
    deep[:34] = fib
    with pytest.raises(hb.HBError) as e:
        hb.build_codebook(deep)
    assert e.value.status == hb.capi.HB_ERR_CODELEN


def test_shard_offsets(hb):
    starts, total = hb.shard_offsets([10, 0, 33, 2 ** 40])
    assert starts.tolist() == [0, 10, 10, 43] and total == 43 + 2 ** 40


def test_workload_thresholds(hb):
    for name in ("c2", "c3", "c5", "t1g"):
        w = hb.workloads.get(name)
        assert np.all(np.diff(w.thr.astype(np.int64)) >= 0)
        assert w.thr[-1] == 0xFFFFFFFF
    w = hb.workloads.get("c2", n_bytes=4096)
    assert w.n_words == 1024


def test_stats_series_sink_format(hb, tmp_path):
    """stats_logger's file format (stats_logger.cpp:13-44): header once, then one "x y" line per sample; LogStats2
    also derives the data-rate series MB * 1000 / (ms * 1024)."""
    d = str(tmp_path)
    hb.stats.log_stats2(d, "encode_t", "B200", 2.0, 256.0, series_number=3, description="d")
    hb.stats.log_stats2(d, "encode_t", "B200", 4.0, 1024.0)
    lines = open(os.path.join(d, "encode_t__3_B200.txt")).read().splitlines()
    assert lines[:16] == ["SERIES_NAME", "B200", "X_AXIS_QUANTITY", "Data size", "Y_AXIS_QUANTITY", "Time",
                          "X_AXIS_UNIT", "MB", "Y_AXIS_UNIT", "ms", "X_AXIS_SCALE_TYPE", "log",
                          "Y_AXIS_SCALE_TYPE", "lin", "DESCRIPTION", "d"]
    assert lines[16:] == ["__DATA__", "256.000000 2.000000", "1024.000000 4.000000"]
    rate = open(os.path.join(d, "encode_t_datarate__3_B200.txt")).read().splitlines()
    assert rate[-2:] == ["256.000000 125.000000", "1024.000000 250.000000"] and "GB/s" in rate
