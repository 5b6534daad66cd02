"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle / the unmodified
reference / the committed golden vectors.  Bit-exact: this is integer work, tolerance zero.
Run on the B200 box with `pytest -m gpu`."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

TILE = 32768          # about one encode tile (hb_tile_bytes(): worker warps x 2 KiB chunks)


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def enc(hb, torch_mod):
    e = hb.Encoder(device=0, max_bytes=(1 << 30) + (1 << 20))
    yield e
    e.close()


def gpu_encode(enc, torch, data_u8, cw, cl, start_bit=0, cap_words=None, poison=True):
    """-> (out uint32 numpy (whole buffer), total_bits)"""
    assert data_u8.size % 4 == 0
    d_in = torch.from_numpy(np.ascontiguousarray(data_u8)).cuda() if data_u8.size else \
        torch.empty(0, dtype=torch.uint8, device="cuda")
    bits_max = int(np.asarray(cl, dtype=np.uint64)[data_u8].sum()) if data_u8.size else 0
    if cap_words is None:
        cap_words = (start_bit + bits_max) // 32 + 2
    d_out = torch.empty(cap_words + 4, dtype=torch.int32, device="cuda")
    if poison:
        d_out.fill_(0x5A5A5A5A)        # the kernel must not rely on a pre-zeroed output (no memset)
    bits = enc.encode(d_in, cw, cl, d_out[:cap_words], start_bit=start_bit)
    torch.cuda.synchronize()
    return d_out.cpu().numpy().view(np.uint32), bits


def check_against_oracle(orc, enc, torch, data_u8, cw, cl):
    ref_words, ref_bits, _ = orc.encode(data_u8.view(np.uint32), cw, cl)
    out, bits = gpu_encode(enc, torch, data_u8, cw, cl)
    assert bits == ref_bits
    n = ref_words.size                                  # floor(bits/32)+1, incl. the courtesy zero word
    bad = np.nonzero(out[:n] != ref_words)[0]
    assert bad.size == 0, "first mismatch at word %d of %d: got %08x want %08x" % (
        bad[0], n, out[bad[0]], ref_words[bad[0]])
    assert np.all(out[n:] == 0x5A5A5A5A), "wrote past floor(bits/32)+1 words"
    return bits


# ---- histogram ------------------------------------------------------------------------------------
def test_histogram_c1(hb, enc, orc, torch_mod, c1):
    data = hb.workloads.c1_fixture_bytes()
    d_in = torch_mod.from_numpy(data.copy()).cuda()
    hist = enc.histogram(d_in)
    assert np.array_equal(hist, c1["freqs"])
    assert np.array_equal(hist, orc.histogram(data))


@pytest.mark.parametrize("n_bytes", [0, 4, 12, 16, 36, 4096 + 4, (1 << 20) + 20, 5_000_004])
def test_histogram_sizes_and_alignment(enc, orc, torch_mod, n_bytes):
    rng = np.random.default_rng(n_bytes)
    data = rng.integers(0, 256, size=n_bytes + 16, dtype=np.uint8)
    data[: n_bytes // 2] = 7                                    # a long run: the merge-equal path
    buf = torch_mod.from_numpy(data).cuda()
    for off in (0, 4, 8, 12):                                   # 4-byte aligned, not 16-byte aligned
        view = buf[off:off + n_bytes]
        assert np.array_equal(enc.histogram(view), orc.histogram(data[off:off + n_bytes]))


# ---- config 1: the reference's own fixture -----------------------------------------------------------
def test_c1_fixture_bit_exact(hb, enc, orc, ref, torch_mod, c1):
    data = hb.workloads.c1_fixture_bytes()
    d_in = torch_mod.from_numpy(data.copy()).cuda()
    hist = enc.histogram(d_in)
    cw, cl, max_len = enc.build_codebook(hist)
    assert np.array_equal(cw, c1["codewords"]) and np.array_equal(cl, c1["codewordlens"])
    out, bits = gpu_encode(enc, torch_mod, data, cw, cl)
    assert bits == c1["total_bits"] == 2330672
    nw = c1["n_words"]
    assert [int(x) for x in out[:8]] == c1["first_words"]
    assert [int(x) for x in out[nw - 4:nw]] == c1["last_words"]
    assert orc.word_fnv(out[:nw]) == c1["fnv"] == 0x6774223E44CA33FB
    check_against_oracle(orc, enc, torch_mod, data, cw, cl)
    if ref is not None:                                         # the unmodified cpu_vlc_encode
        r_out, r_size = ref.encode(data.view(np.uint32), cw, cl, data.size // 4 + 2)
        assert r_size == c1["outsize_bytes"] == (bits + 7) // 8
        assert np.array_equal(r_out[:nw], out[:nw])


# ---- golden small cases and KATs ----------------------------------------------------------------------
def test_golden_encode_cases(enc, orc, torch_mod):
    for i, case in enumerate(load_golden("encode_cases.json")):
        w = np.array(case["in"], dtype=np.uint32)
        cw = np.array(case["codewords"], dtype=np.uint32)
        cl = np.array(case["codewordlens"], dtype=np.uint32)
        out, bits = gpu_encode(enc, torch_mod, w.view(np.uint8), cw, cl)
        assert bits == case["total_bits"], i
        gold = np.array(case["out_words"], dtype=np.uint32)
        assert np.array_equal(out[: gold.size], gold), i


def test_kats(hb, enc, torch_mod):
    g = load_golden("kat.json")
    cw = np.array(g["codewords"], dtype=np.uint32)
    cl = np.array(g["codewordlens"], dtype=np.uint32)
    # length 32 and dirty codewords are outside the parity domain: rejected, never mis-encoded
    with pytest.raises(hb.HBError) as e:
        gpu_encode(enc, torch_mod, np.zeros(4, np.uint8), cw, cl)
    assert e.value.status == hb.capi.HB_ERR_CODELEN
    cw2, cl2 = cw.copy(), cl.copy()
    cw2[1], cl2[1] = 0, 0
    with pytest.raises(hb.HBError) as e:
        gpu_encode(enc, torch_mod, np.zeros(4, np.uint8), cw2, cl2)
    assert e.value.status == hb.capi.HB_ERR_CODEWORD
    cw2[5], cl2[5] = 0, 0
    for case in g["cases"]:
        if not case["parity_domain"]:
            continue
        w = np.array(case["in"], dtype=np.uint32)
        if np.isin(w.view(np.uint8), [1, 5]).any():
            continue
        out, bits = gpu_encode(enc, torch_mod, w.view(np.uint8), cw2, cl2)
        assert (bits + 7) // 8 == case["outsize_bytes"], case["name"]
        gold = case["out_words"]
        n = min(len(gold), bits // 32 + 1)
        assert [int(x) for x in out[:n]] == gold[:n], case["name"]


def test_empty_input(enc, torch_mod):
    cw = np.zeros(256, np.uint32)
    cl = np.ones(256, np.uint32)
    out, bits = gpu_encode(enc, torch_mod, np.zeros(0, np.uint8), cw, cl, cap_words=2)
    assert bits == 0 and out[0] == 0                             # cpuencode.cpp:17


# ---- property tests over random codebooks ---------------------------------------------------------------
def random_prefix_code(rng, nsym, skew):
    h = np.zeros(256, dtype=np.uint64)
    syms = rng.choice(256, size=nsym, replace=False)
    if skew == "geo":
        # ratio floor keeps the deepest code at <= 31 bits (the parity domain)
        r = rng.uniform(max(0.3, 2.0 ** (-26.0 / nsym)), 0.97)
        h[syms] = np.maximum(1, (2.0 ** 30 * r ** np.arange(nsym))).astype(np.uint64)
    elif skew == "flat":
        h[syms] = rng.integers(1, 1000, size=nsym)
    else:                                                        # fibonacci: the deepest trees
        fib = [1, 1]
        while len(fib) < nsym:
            fib.append(fib[-1] + fib[-2])
        h[syms] = np.array(fib[:nsym], dtype=np.uint64)
    return h


@pytest.mark.parametrize("skew,nsym", [("geo", 2), ("geo", 22), ("geo", 64), ("geo", 256), ("flat", 256),
                                       ("flat", 3), ("fib", 17), ("fib", 25), ("fib", 32)])
@pytest.mark.parametrize("n_bytes", [4, 1020, 1024, 1028, 2048, 16380, 32764, 32768, 32772, 3 * 32768 + 20, 1_000_000,
                                     5 * 1024 * 1024 + 4])
def test_random_codebooks(hb, enc, orc, torch_mod, skew, nsym, n_bytes):
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((skew, nsym, n_bytes)).encode()))
    h = random_prefix_code(rng, nsym, skew)
    cw, cl, max_len = hb.build_codebook(h)
    p = h.astype(np.float64) / float(h.sum())
    floor = 1.0 / 4096                                           # make the rare, long codes show up
    p = np.where(h > 0, np.maximum(p, floor), 0)
    p /= p.sum()
    data = rng.choice(256, size=n_bytes, p=p).astype(np.uint8)
    check_against_oracle(orc, enc, torch_mod, data, cw, cl)


def test_variant_choice(hb, c1):
    """the kernel variant follows from the code lengths alone (group size G, packed/wide table, run-time check)"""
    want = {1: "packed_g8", 3: "packed_g8", 5: "packed_g6", 7: "packed_g4", 10: "packed_g3", 15: "packed_g2",
            16: "packed_g1", 24: "packed_g1", 25: "wide_g1", 31: "wide_g1"}
    for max_len, name in want.items():
        cl = np.zeros(256, np.uint32)
        cl[0] = cl[1] = max_len
        assert hb.encode_variant(cl) == name, max_len
    assert hb.encode_variant(c1["codewordlens"]) == "packed_g4c"      # skewed: long codes are rare


@pytest.mark.parametrize("group", [1, 2, 3, 4, 6, 8])
def test_every_kernel_variant_vs_oracle(hb, enc, orc, torch_mod, monkeypatch, group):
    """$HB_FORCE_GROUP pins G; each (G, packed|wide, check|nocheck) kernel must equal cpu_vlc_encode, on data
    that follows the codebook (fast path) and on data that does not (over-long groups -> symbol-by-symbol path)"""
    monkeypatch.setenv("HB_FORCE_GROUP", str(group))
    rng = np.random.default_rng(group)
    seen = set()
    for skew, nsym in (("flat", 6), ("geo", 22), ("flat", 256), ("fib", 32)):
        h = random_prefix_code(rng, nsym, skew)
        cw, cl, max_len = hb.build_codebook(h)
        seen.add(hb.encode_variant(cl))
        p = h.astype(np.float64) / float(h.sum())
        matched = rng.choice(256, size=6 * TILE + 1000, p=p).astype(np.uint8)
        check_against_oracle(orc, enc, torch_mod, matched, cw, cl)
        uniform = rng.choice(np.nonzero(h)[0], size=2 * TILE + 36).astype(np.uint8)   # rare symbols everywhere
        check_against_oracle(orc, enc, torch_mod, uniform, cw, cl)
    assert len(seen) >= 2, seen


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_nonstationary_many_rounds(hb, enc, orc, torch_mod, seed):
    """Ten and more tiles per CTA with chunk sizes that swing between a few words and a full ring: segments of
    very different entropy under ONE codebook (ring wrap and skip, ring-full and depth-limit retirements, records that
    arrive early and late), ragged end, a shard phase."""
    rng = np.random.default_rng(1000 + seed)
    nsym = [256, 64, 32][seed - 1]
    h = random_prefix_code(rng, nsym, ["flat", "geo", "fib"][seed - 1])
    cw, cl, max_len = hb.build_codebook(h)
    syms = np.nonzero(h)[0]
    order = syms[np.argsort(cl[syms])]                       # shortest codes first
    parts = []
    total = 0
    while total < 40 * 1024 * 1024:
        n = int(rng.integers(1, 64)) * 4096 + 4 * int(rng.integers(0, 64))
        kind = int(rng.integers(0, 4))
        if kind == 0:                                        # the shortest code only: a chunk of a few words
            seg = np.full(n, order[0], dtype=np.uint8)
        elif kind == 1:                                      # the longest codes only: the largest chunks this book allows
            seg = rng.choice(order[-max(1, nsym // 8):], size=n).astype(np.uint8)
        elif kind == 2:                                      # matched to the codebook
            seg = rng.choice(256, size=n, p=h / h.sum()).astype(np.uint8)
        else:                                                # uniform over the used symbols
            seg = rng.choice(syms, size=n).astype(np.uint8)
        parts.append(seg)
        total += n
    data = np.concatenate(parts)
    data = data[: data.size - data.size % 4]
    ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
    out, bits = gpu_encode(enc, torch_mod, data, cw, cl)
    assert bits == ref_bits
    assert np.array_equal(out[: ref_words.size], ref_words)
    # the same shard in a global phase: shifted by start_bit % 32, first word's leading bits zero
    sb = 64 * 7 + 13
    out2, bits2 = gpu_encode(enc, torch_mod, data[: 9 * TILE + 8], cw, cl, start_bit=sb)
    r2, rb2, _ = orc.encode(data[: 9 * TILE + 8].view(np.uint32), cw, cl)
    assert bits2 == rb2
    got_bits = np.unpackbits(out2[sb // 32: sb // 32 + rb2 // 32 + 2].byteswap().view(np.uint8))
    want_bits = np.unpackbits(r2.byteswap().view(np.uint8))
    assert np.array_equal(got_bits[sb % 32: sb % 32 + rb2], want_bits[:rb2]) and not got_bits[: sb % 32].any()


@pytest.mark.parametrize("group", [4, 6, 8])
@pytest.mark.parametrize("long_len", [17, 24, 31])
def test_isolated_overlong_groups(enc, orc, torch_mod, monkeypatch, group, long_len):
    """One over-long group (>= 32 bits) in an otherwise ordinary lane: the group-level detour of the G >= 4 kernels
    (the normal pass skips the words completed inside the group, the group is re-encoded from its input words).
    Bursts of 2..4 long codewords are dropped at every position of a lane, its first and last group included, sparse
    enough that most lanes have a single such group and dense enough that some have two (the lane redo)."""
    monkeypatch.setenv("HB_FORCE_GROUP", str(group))
    rng = np.random.default_rng(100 * group + long_len)
    cl = np.zeros(256, np.uint32)
    cw = np.zeros(256, np.uint32)
    cl[0], cw[0] = 1, 0
    cl[1], cw[1] = 2, 0b10
    cl[2], cw[2] = 3, 0b110
    cl[255], cw[255] = long_len, (1 << long_len) - 2
    cl[254], cw[254] = long_len, ((1 << long_len) - 1) & 0x55555555 | (1 << (long_len - 1))
    n = 12 * TILE + 40
    data = rng.choice([0, 1, 2], size=n, p=[0.6, 0.3, 0.1]).astype(np.uint8)
    for burst in (2, 3, 4):
        starts = rng.choice(n - 8, size=n // 400, replace=False)
        for b in range(burst):
            data[starts + b] = rng.choice([254, 255], size=starts.size)
    # and the corners of a lane's 64 symbols explicitly
    for lane_start in range(0, 6 * 64 * 32, 64 * 7):
        data[lane_start:lane_start + 2] = 255
        data[lane_start + 62 + 64:lane_start + 64 + 64] = 254
    check_against_oracle(orc, enc, torch_mod, data, cw, cl)


@pytest.mark.parametrize("max_len", [1, 5, 8, 10, 13, 16, 20, 24, 27, 31])
def test_arbitrary_tables_lengths_0_to_31(enc, orc, torch_mod, max_len):
    """cpu_vlc_encode accepts ANY table (not only prefix codes), including zero-length symbols."""
    rng = np.random.default_rng(max_len)
    for trial in range(3):
        lo = 0 if trial == 0 else 1
        cl = rng.integers(lo, max_len + 1, size=256).astype(np.uint32)
        cl[rng.integers(0, 256)] = max_len
        cw = np.array([int(rng.integers(0, 1 << int(l))) if l else 0 for l in cl], dtype=np.uint32)
        n_bytes = [40, TILE * 3, 200_004][trial]
        data = rng.integers(0, 256, size=n_bytes, dtype=np.uint8)
        check_against_oracle(orc, enc, torch_mod, data, cw, cl)


def test_mostly_zero_length_codes(enc, orc, torch_mod):
    """threads and whole tiles that emit < 32 bits (the atomicOr path), tiles that emit nothing"""
    rng = np.random.default_rng(3)
    cl = np.zeros(256, np.uint32)
    cw = np.zeros(256, np.uint32)
    cl[9], cw[9] = 3, 0b101
    cl[200], cw[200] = 31, 0x5EADBEEF & 0x7FFFFFFF
    for density in (0.0, 0.0005, 0.05):
        data = np.zeros(TILE * 6 + 64, dtype=np.uint8)
        hits = rng.random(data.size) < density
        data[hits] = rng.choice([9, 200], size=int(hits.sum()))
        check_against_oracle(orc, enc, torch_mod, data, cw, cl)


def test_single_symbol_input(hb, enc, orc, torch_mod):
    data = np.full(TILE * 2 + 16, 77, dtype=np.uint8)
    h = orc.histogram(data)
    cw, cl, max_len = hb.build_codebook(h)
    assert max_len == 0
    out, bits = gpu_encode(enc, torch_mod, data, cw, cl, cap_words=2)
    assert bits == 0 and out[0] == 0


# ---- start_bit phase, capacity, re-entrancy --------------------------------------------------------------
@pytest.mark.parametrize("start_bit", [1, 31, 32, 45, 64 * 1000 + 17])
def test_start_bit_phase(hb, enc, orc, torch_mod, start_bit):
    rng = np.random.default_rng(start_bit)
    w = hb.workloads.get("c2")
    data = orc.synth_fill(0, TILE * 5 + 400, w.seed, w.mode, w.nbits, w.thr)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
    out, bits = gpu_encode(enc, torch_mod, data, cw, cl, start_bit=start_bit)
    assert bits == ref_bits
    # expected: the oracle stream shifted by start_bit zero bits
    refbits = np.unpackbits(ref_words.byteswap().view(np.uint8))[:ref_bits]
    n = (start_bit + ref_bits + 31) // 32
    full = np.zeros(n * 32, dtype=np.uint8)
    full[start_bit:start_bit + ref_bits] = refbits
    want = np.packbits(full).view(np.uint32).byteswap()
    w0 = start_bit // 32
    assert np.array_equal(out[w0:n], want[w0:n])
    assert np.all(out[:w0] == 0x5A5A5A5A)                          # words before the phase are untouched


def test_capacity_error(hb, enc, orc, torch_mod):
    data = np.random.default_rng(0).integers(0, 256, size=TILE * 4, dtype=np.uint8)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    with pytest.raises(hb.HBError) as e:
        gpu_encode(enc, torch_mod, data, cw, cl, cap_words=100)
    assert e.value.status == hb.capi.HB_ERR_CAPACITY
    check_against_oracle(orc, enc, torch_mod, data, cw, cl)       # the context is still usable
    big = hb.Encoder(device=0, max_bytes=TILE)
    with pytest.raises(hb.HBError) as e:
        gpu_encode(big, torch_mod, data, cw, cl)
    assert e.value.status == hb.capi.HB_ERR_CAPACITY
    big.close()


def test_repeated_calls_and_two_contexts(hb, enc, orc, torch_mod):
    """no state leaks between calls (epoch-tagged descriptors, monotonic tickets); contexts are independent"""
    other = hb.Encoder(device=0, max_bytes=1 << 22)
    rng = np.random.default_rng(11)
    for it in range(12):
        n = int(rng.integers(1, 300)) * 4096
        data = rng.integers(0, [2, 16, 256][it % 3], size=n, dtype=np.uint8)
        cw, cl, _ = hb.build_codebook(orc.histogram(data))
        check_against_oracle(orc, enc if it % 2 else other, torch_mod, data, cw, cl)
    other.close()


def test_non_default_stream(hb, enc, orc, torch_mod):
    data = np.random.default_rng(5).integers(0, 64, size=TILE * 9 + 12, dtype=np.uint8)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    s = torch_mod.cuda.Stream()
    with torch_mod.cuda.stream(s):
        check_against_oracle(orc, enc, torch_mod, data, cw, cl)


def test_jobs_on_alternating_streams_share_one_context(hb, enc, orc, torch_mod):
    """A context serialises its jobs (they share the device table, the result block and the two alternating look-back
    trees): asynchronous jobs handed alternately to two streams, with different codebooks, must each equal the oracle
    (nothing ordered job B's table upload and tree reset behind job A's kernel on another stream)."""
    rng = np.random.default_rng(11)
    s1, s2 = torch_mod.cuda.Stream(), torch_mod.cuda.Stream()
    jobs = []
    for i in range(8):
        nsym = [4, 64, 200, 17][i % 4]
        data = rng.integers(0, nsym, size=TILE * (20 + i) + 4 * i, dtype=np.uint8)
        cw, cl, _ = hb.build_codebook(orc.histogram(data))
        ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
        d_in = torch_mod.from_numpy(data).cuda()
        d_out = torch_mod.full((ref_words.size + 2,), 0x5A5A5A5A, dtype=torch_mod.int32, device="cuda")
        jobs.append((d_in, d_out, cw, cl, ref_words, ref_bits))
    torch_mod.cuda.synchronize()
    for i, (d_in, d_out, cw, cl, ref_words, ref_bits) in enumerate(jobs):
        with torch_mod.cuda.stream(s1 if i % 2 == 0 else s2):
            enc.encode_async(d_in, cw, cl, d_out)                      # no fetch in between: the jobs queue up
    torch_mod.cuda.synchronize()
    with torch_mod.cuda.stream(s2):
        assert enc.encode_result() == jobs[-1][5]
    for d_in, d_out, cw, cl, ref_words, ref_bits in jobs:
        got = d_out.cpu().numpy().view(np.uint32)
        assert np.array_equal(got[: ref_words.size], ref_words)


def test_job_size_limit(hb):
    """a look-back tree node counts tiles in 22 bits: contexts for more than 2^21 tiles (64 GiB) are refused"""
    import ctypes as C
    ctx = hb.capi.vp()
    rc = hb.lib().hb_init(C.byref(ctx), 0, (1 << 21) * 8192 + 1)
    assert rc == hb.capi.HB_ERR_CAPACITY and not ctx.value


def test_repeated_launches_are_deterministic(hb, enc, orc, torch_mod):
    """200 back-to-back launches of one job (no host synchronisation in between), each into a freshly poisoned buffer:
    every stream must have the oracle's checksums.  The kernel's hand-offs are relaxed atomics, mbarriers and plain
    shared-memory stores ordered by them; a race would show up as a sporadic mismatch (compute-sanitizer's racecheck
    is not available on this pool)."""
    from huffman_gpu_b200.streamsum import stream_sums
    wl = hb.workloads.get("c2", n_bytes=48 << 20)
    data = orc.synth_fill(0, wl.n_bytes, wl.seed, wl.mode, wl.nbits, wl.thr)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
    want = stream_sums(ref_words)
    d_in = torch_mod.from_numpy(data).cuda()
    outs = [torch_mod.empty(ref_words.size + 2, dtype=torch_mod.int32, device="cuda") for _ in range(8)]
    for rep in range(25):
        for o in outs:
            o.fill_(0x5A5A5A5A)
        for o in outs:
            enc.encode_async(d_in, cw, cl, o)
        assert enc.encode_result() == ref_bits
        for o in outs:
            assert stream_sums(o, ref_words.size) == want, rep
            assert int(o[ref_words.size].item()) == 0x5A5A5A5A


# ---- decoder (the reference has none): encode -> decode must give the input back ----------
@pytest.mark.parametrize("skew,nsym", [("geo", 2), ("geo", 22), ("flat", 256), ("fib", 32)])
@pytest.mark.parametrize("n_bytes", [4, 1028, 32768, 32772, 3 * 32768 + 20, 5 * 1024 * 1024 + 4])
def test_decode_round_trip(hb, enc, orc, torch_mod, skew, nsym, n_bytes):
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr(("dec", skew, nsym, n_bytes)).encode()))
    h = random_prefix_code(rng, nsym, skew)
    cw, cl, max_len = hb.build_codebook(h)
    p = h.astype(np.float64) / float(h.sum())
    data = rng.choice(256, size=n_bytes, p=p).astype(np.uint8)
    ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
    d_in = torch_mod.from_numpy(data).cuda()
    d_out = torch_mod.full((ref_words.size + 2,), 0x5A5A5A5A, dtype=torch_mod.int32, device="cuda")
    assert enc.encode(d_in, cw, cl, d_out) == ref_bits
    n_tiles = (n_bytes + TILE - 1) // TILE
    d_idx = torch_mod.empty(n_tiles + 1, dtype=torch_mod.int64, device="cuda")
    enc.tile_index(ref_bits, d_idx)
    idx = d_idx.cpu().numpy()
    assert idx[0] == 0 and idx[-1] == ref_bits and np.all(np.diff(idx) >= 0)
    # every tile offset is the bit count of the symbols before it
    for t in (1, n_tiles // 2, n_tiles - 1):
        if 0 < t < n_tiles:
            assert idx[t] == int(cl.astype(np.uint64)[data[: t * TILE]].sum())
    d_back = torch_mod.full((n_bytes,), 0xEE, dtype=torch_mod.uint8, device="cuda")
    enc.decode(d_out[: ref_words.size + 1], d_idx, cw, cl, d_back)
    assert np.array_equal(d_back.cpu().numpy(), data)
    # the CPU oracle's bit-serial decoder agrees (an independent decoder of the same stream)
    if n_bytes <= 200_000:
        back, end = orc.decode(d_out.cpu().numpy().view(np.uint32), 0, n_bytes, cw, cl)
        assert end == ref_bits and np.array_equal(back[:n_bytes], data)


def test_decode_rejects_foreign_tables(hb, enc, orc, torch_mod):
    """a stream decoded with another codebook does not end tile by tile at the indexed offsets: HB_ERR_CODEWORD"""
    rng = np.random.default_rng(3)
    data = rng.integers(0, 16, size=TILE * 3, dtype=np.uint8)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    bits_expected = hb.bits_from_hist(orc.histogram(data), cl)
    d_in = torch_mod.from_numpy(data).cuda()
    d_out = torch_mod.zeros(bits_expected // 32 + 3, dtype=torch_mod.int32, device="cuda")
    bits = enc.encode(d_in, cw, cl, d_out)
    d_idx = torch_mod.empty(4, dtype=torch_mod.int64, device="cuda")
    enc.tile_index(bits, d_idx)
    other = rng.integers(0, 16, size=TILE * 3, dtype=np.uint8)
    other[: TILE] = 3
    cw2, cl2, _ = hb.build_codebook(orc.histogram(other))
    assert not np.array_equal(cl, cl2)
    d_back = torch_mod.empty(data.size, dtype=torch_mod.uint8, device="cuda")
    with pytest.raises(hb.HBError) as e:
        enc.decode(d_out, d_idx, cw2, cl2, d_back)
    assert e.value.status == hb.capi.HB_ERR_CODEWORD


# ---- host-buffer entry points (the reference-facing call) ---------------------------------------------------
def test_vlc_encode_drop_in_signature(hb, orc, ref, c1):
    """hb_vlc_encode(indata, num_elements, outdata, &outsize, codewords, codewordlens), host pointers."""
    data = hb.workloads.c1_fixture_bytes()
    words = data.view(np.uint32).copy()
    out = np.full(words.size + 2, 0xAAAAAAAA, dtype=np.uint32)
    outsize = hb.vlc_encode(words, words.size, out, c1["codewords"], c1["codewordlens"])
    assert outsize == c1["outsize_bytes"] == 291334
    nw = c1["n_words"]
    assert orc.word_fnv(out[:nw]) == c1["fnv"]
    if ref is not None:
        r_out, r_size = ref.encode(words, c1["codewords"], c1["codewordlens"], words.size + 2)
        assert r_size == outsize and np.array_equal(r_out[:nw], out[:nw])


@pytest.mark.parametrize("n_bytes,chunk_mib", [(64 << 20, None), ((3 << 20) + 32772, 1), (32768, 1), (4, 1)])
def test_encode_host_chunked(hb, enc, orc, torch_mod, monkeypatch, n_bytes, chunk_mib):
    """H2D -> chunked launches over ONE job -> D2H must equal the single-launch stream"""
    w = hb.workloads.get("c5")
    data = orc.synth_fill(0, n_bytes, w.seed, w.mode, w.nbits, w.thr)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    ref_words, ref_bits, ref_bytes = orc.encode(data.view(np.uint32), cw, cl)
    pin_in = hb.PinnedBuffer(n_bytes)
    pin_in.u8[:] = data
    pin_out = hb.PinnedBuffer((ref_words.size + 1) * 4)
    h_out = pin_out.u8.view(np.uint32)
    h_out[:] = 0x77777777
    e2 = hb.Encoder(device=0, max_bytes=n_bytes)
    bits, nbytes = e2.encode_host(pin_in.u8.view(np.uint32), cw, cl, h_out)
    assert bits == ref_bits and nbytes == ref_bytes
    assert np.array_equal(h_out[: ref_words.size], ref_words)
    # pageable host memory takes the same path
    out2 = np.zeros(ref_words.size + 1, dtype=np.uint32)
    bits2, _ = e2.encode_host(data.view(np.uint32).copy(), cw, cl, out2)
    assert bits2 == ref_bits and np.array_equal(out2[: ref_words.size], ref_words)
    e2.close()
    pin_in.free()
    pin_out.free()


# ---- BASELINE.json configurations at full size ----------------------------------------------------------------
def synth_on_device(hb, enc, torch, wl, n_bytes=None):
    n = wl.n_bytes if n_bytes is None else n_bytes
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    enc.synth_fill(d, wl)
    return d


def test_encode_host_error_exits_leave_the_context_clean(hb, enc, orc, torch_mod, monkeypatch):
    """hb_vlc_encode_host with an output that is too small (several chunk launches in flight when the overflow is seen):
    HB_ERR_CAPACITY, nothing left running or dirty -- the next device-path and host-path jobs on the same context are
    exact (early returns left launches queued, D2H copies in flight and a stale overflow flag behind)."""
    monkeypatch.setenv("HB_CHUNK_MIB", "1")
    rng = np.random.default_rng(21)
    data = rng.integers(0, 200, size=6 << 20, dtype=np.uint8)
    cw, cl, _ = hb.build_codebook(orc.histogram(data))
    ref_words, ref_bits, _ = orc.encode(data.view(np.uint32), cw, cl)
    small = np.zeros(ref_words.size // 3, dtype=np.uint32)
    for _ in range(3):
        with pytest.raises(hb.HBError) as e:
            enc.encode_host(data.view(np.uint32), cw, cl, small)
        assert e.value.status == hb.capi.HB_ERR_CAPACITY
        check_against_oracle(orc, enc, torch_mod, data[: 5 * TILE + 8], cw, cl)      # device path, same context
        h_out = np.zeros(ref_words.size + 1, dtype=np.uint32)
        bits, _ = enc.encode_host(data.view(np.uint32), cw, cl, h_out)                # host path, same context
        assert bits == ref_bits and np.array_equal(h_out[: ref_words.size], ref_words)


def test_synth_device_matches_host(hb, enc, orc, torch_mod):
    for name in ("c2", "c3", "c5"):
        wl = hb.workloads.get(name)
        d = synth_on_device(hb, enc, torch_mod, wl, n_bytes=1 << 20)
        assert np.array_equal(d.cpu().numpy(), orc.synth_fill(0, 1 << 20, wl.seed, wl.mode, wl.nbits, wl.thr))
    wl = hb.workloads.get("c4")
    d = torch_mod.empty(1 << 16, dtype=torch_mod.uint8, device="cuda")
    enc.synth_fill(d, wl, first=12345678)
    assert np.array_equal(d.cpu().numpy(),
                          orc.synth_fill(12345678, 1 << 16, wl.seed, wl.mode, wl.nbits, wl.thr))


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_full_size_config_vs_oracle(hb, enc, orc, ref, torch_mod, name):
    """C2 (256 MiB, H~2.2) and C3 (1 GiB, H~7.9): the whole stream against the CPU encoder."""
    wl = hb.workloads.get(name)
    d_in = synth_on_device(hb, enc, torch_mod, wl)
    hist = enc.histogram(d_in)
    assert int(hist.sum()) == wl.n_bytes
    cw, cl, max_len = hb.build_codebook(hist)
    bits_expected = hb.bits_from_hist(hist, cl)
    d_out = torch_mod.empty(bits_expected // 32 + 2, dtype=torch_mod.int32, device="cuda")
    d_out.fill_(0x5A5A5A5A)
    bits = enc.encode(d_in, cw, cl, d_out)
    assert bits == bits_expected
    got = d_out.cpu().numpy().view(np.uint32)
    words = d_in.cpu().numpy().view(np.uint32)
    assert np.array_equal(orc.histogram(words.view(np.uint8)), hist)
    if ref is not None and bits_expected < 2 ** 35:
        r_out, r_size = ref.encode(words, cw, cl, bits_expected // 32 + 2)   # unmodified cpu_vlc_encode
        assert r_size == ((bits + 7) // 8) % 2 ** 32
        assert np.array_equal(r_out[: bits // 32 + 1], got[: bits // 32 + 1])
    else:
        o_out, o_bits, _ = orc.encode(words, cw, cl, total_bits_hint=bits_expected)
        assert np.array_equal(o_out, got[: o_out.size])


def test_c4_fibonacci_sample_and_properties(hb, enc, orc, torch_mod):
    """C4 (2 GiB, code lengths 1..31, the wide kernel).  Full-size run checked through size-independent
    properties: exact histogram, bit total = sum hist*len, and oracle equality on random windows whose
    start bit is recomputed from the prefix histogram (encode(window, start_bit) is position independent)."""
    wl = hb.workloads.get("c4")
    big = hb.Encoder(device=0, max_bytes=wl.n_bytes)
    d_in = synth_on_device(hb, big, torch_mod, wl)
    hist = big.histogram(d_in)
    assert np.array_equal(hist[:32], hb.workloads.fibonacci_counts())
    cw, cl, max_len = hb.build_codebook(hist)
    assert max_len == 31 and hb.encode_variant(cl).startswith("wide_")
    g = [c for c in load_golden("codebooks.json") if c.get("name") == "c4_fibonacci"][0]
    assert cw.tolist() == g["codewords"] and cl.tolist() == g["codewordlens"]
    bits_expected = hb.bits_from_hist(hist, cl)
    d_out = torch_mod.empty(bits_expected // 32 + 2, dtype=torch_mod.int32, device="cuda")
    bits = big.encode(d_in, cw, cl, d_out)
    assert bits == bits_expected
    rng = np.random.default_rng(4)
    for _ in range(6):
        t0 = int(rng.integers(0, wl.n_bytes // TILE - 40))
        a, b = t0 * TILE, (t0 + 33) * TILE + 4 * int(rng.integers(0, 2048))
        pre = big.histogram(d_in[:a]) if a else np.zeros(256, np.uint64)
        start = hb.bits_from_hist(pre, cl)
        window = d_in[a:b].cpu().numpy()
        o_out, o_bits, _ = orc.encode(window.view(np.uint32), cw, cl)
        w0, w1 = (start + 31) // 32, (start + o_bits) // 32           # fully covered global words
        got = d_out[w0:w1].cpu().numpy().view(np.uint32)
        obits = np.unpackbits(o_out.byteswap().view(np.uint8))
        lo = w0 * 32 - start
        want = np.packbits(obits[lo:lo + (w1 - w0) * 32]).view(np.uint32).byteswap()
        assert np.array_equal(got, want)
    big.close()


def test_c5_8gib_properties(hb, orc, torch_mod):
    """C5 (8 GiB, H~4.0): exact histogram mass, bit total = sum hist*len (the 64-bit oracle where
    cpu_vlc_encode's uint32 outsize wraps), and oracle equality on random windows whose start bit is recomputed
    from the prefix histogram.  Output offsets exceed 2^32 bits here (the reference's scan is uint32)."""
    wl = hb.workloads.get("c5")
    free, _ = torch_mod.cuda.mem_get_info()
    if free < wl.n_bytes * 1.7:
        pytest.skip("not enough device memory for the 8 GiB configuration")
    big = hb.Encoder(device=0, max_bytes=wl.n_bytes)
    d_in = synth_on_device(hb, big, torch_mod, wl)
    hist = big.histogram(d_in)
    assert int(hist.sum()) == wl.n_bytes
    cw, cl, max_len = hb.build_codebook(hist)
    bits_expected = hb.bits_from_hist(hist, cl)
    assert bits_expected > 2 ** 32
    d_out = torch_mod.empty(bits_expected // 32 + 2, dtype=torch_mod.int32, device="cuda")
    assert big.encode(d_in, cw, cl, d_out) == bits_expected
    rng = np.random.default_rng(5)
    starts = [0, wl.n_bytes - 40 * TILE] + [int(x) * TILE for x in rng.integers(1, wl.n_bytes // TILE - 40, size=4)]
    for a in starts:
        b = a + 33 * TILE + 4 * int(rng.integers(0, 2048))
        pre = big.histogram(d_in[:a]) if a else np.zeros(256, np.uint64)
        start = hb.bits_from_hist(pre, cl)
        window = d_in[a:b].cpu().numpy()
        o_out, o_bits, _ = orc.encode(window.view(np.uint32), cw, cl)
        w0, w1 = (start + 31) // 32, (start + o_bits) // 32           # fully covered global words
        got = d_out[w0:w1].cpu().numpy().view(np.uint32)
        obits = np.unpackbits(o_out.byteswap().view(np.uint8))
        lo = w0 * 32 - start
        want = np.packbits(obits[lo:lo + (w1 - w0) * 32]).view(np.uint32).byteswap()
        assert np.array_equal(got, want), a
    big.close()


def test_reference_gpu_pipeline_agrees_on_fixture(hb, enc, orc, torch_mod, c1):
    """the reference's own 3-pass GPU path (unmodified kernels, oracle/ref_gpu_shim.cu), the
    single-pass kernel and cpu_vlc_encode produce the same stream on the reference's fixture (config 1)."""
    import pyoracle
    rg = pyoracle.try_ref_gpu()
    if rg is None:
        pytest.skip("oracle/_ref/libref_gpu.so was not shipped")
    data = hb.workloads.c1_fixture_bytes()
    d_in = torch_mod.from_numpy(data.copy()).cuda()
    d_ref = torch_mod.zeros(data.size // 4, dtype=torch_mod.int32, device="cuda")
    bits, ms_enc, ms_scan, ms_pack = rg.run(d_in.data_ptr(), data.size // 4, c1["codewords"], c1["codewordlens"],
                                            d_ref.data_ptr(), data.size)
    torch_mod.cuda.synchronize()
    assert bits == c1["total_bits"]
    ours, our_bits = gpu_encode(enc, torch_mod, data, c1["codewords"], c1["codewordlens"])
    nw = c1["n_words"]
    ref_words = d_ref.cpu().numpy().view(np.uint32)[:nw]
    assert our_bits == bits and np.array_equal(ref_words, ours[:nw])
    assert orc.word_fnv(ref_words) == c1["fnv"]


def test_cli_driver_on_the_fixture_and_a_ragged_file(hb, tmp_path):
    """pavle_b200 <file>: the reference driver's sequence and output fields (main_test_cu.cu:41-180); PASS! means the
    bit count equals sum(hist * len) and the GPU stream decodes back to the file."""
    import subprocess
    cli = os.path.join(os.path.dirname(hb.capi.LIB_PATH), "pavle_b200")
    fixture = tmp_path / "test1024.in"
    fixture.write_bytes(hb.workloads.c1_fixture_bytes().tobytes())
    out = subprocess.run([cli, str(fixture)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "GPU Encoded to 291334 [B]" in out.stdout and "PASS!" in out.stdout       # SURVEY section 8c golden value
    assert "entropy 2.2065" in out.stdout
    rng = np.random.default_rng(7)
    ragged = tmp_path / "ragged.in"
    ragged.write_bytes(rng.choice(256, size=3 * TILE + 4 * 37 + 3, p=np.r_[0.7, np.full(255, 0.3 / 255)]).astype(np.uint8).tobytes())
    out = subprocess.run([cli, str(ragged), "--repeats", "3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "PASS!" in out.stdout, out.stdout + out.stderr
    # the reference's CPU fields and its word-by-word compare (main_test_cu.cu:120-126,171), with the unmodified
    # cpu_vlc_encode handed in as a shared library (test infrastructure: oracle/_ref); the driver links no CPU encoder
    from conftest import ROOT
    cpu_lib = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    if os.path.exists(cpu_lib):
        out = subprocess.run([cli, str(fixture), "--cpu-lib", cpu_lib], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "CPU Encoded to 291334 [B]" in out.stdout and "CPU Encoding time (CPU):" in out.stdout
        assert "PASS! vectors are matching!" in out.stdout
