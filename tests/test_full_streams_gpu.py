"""Full-size parity on every BASELINE.json configuration.

  * test_stream_sums_match_reference: the WHOLE stream of c2 / 1 GiB H2.2 / c3 / c4 / c5 from one GPU, compared
    through the position-sensitive checksums of huffman-gpu_b200/streamsum.py with the stream the UNMODIFIED
    cpu_vlc_encode produced for the same input (tests/golden/streams.json, made by oracle/make_golden_streams.py
    from the CPU generator + oracle/_ref); histogram, codebook and bit count are compared too.
  * test_round_trip_at_full_size: encode -> hb_decode (the tile-parallel GPU decoder) gives the input back, on the device,
    for every configuration: a size-independent property that needs no CPU at all.
  * test_whole_stream_word_by_word: C4 (2 GiB) and C5 (8 GiB) once more, this time literally word by word against
    cpu_vlc_encode run live on this box's host (5 s / 21 s of one core), as main_test_cu.cu:170-171 compares.
Integer work: tolerance zero."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return load_golden("streams.json")


def _encode_whole(hb, torch, name):
    wl = hb.workloads.get(name)
    enc = hb.Encoder(device=0, max_bytes=wl.n_bytes)
    d_in = torch.empty(wl.n_bytes, dtype=torch.uint8, device="cuda")
    enc.synth_fill(d_in, wl, first=0)
    hist = enc.histogram(d_in)
    cw, cl, max_len = hb.build_codebook(hist)
    bits_expected = hb.bits_from_hist(hist, cl)
    d_out = torch.full((bits_expected // 32 + 2,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    bits = enc.encode(d_in, cw, cl, d_out)
    torch.cuda.synchronize()
    return enc, d_in, d_out, hist, cw, cl, bits


@pytest.mark.parametrize("name", ["c2", "t1g", "c3", "c4", "c5"])
def test_stream_sums_match_reference(hb, golden, name):
    import torch
    from huffman_gpu_b200.streamsum import stream_sums
    torch.cuda.set_device(0)
    g = golden[name]
    enc, d_in, d_out, hist, cw, cl, bits = _encode_whole(hb, torch, name)
    try:
        assert np.array_equal(hist, np.array(g["hist"], dtype=np.uint64)), "device generator / histogram differ"
        assert np.array_equal(cl, np.array(g["codewordlens"], dtype=np.uint32))
        assert np.array_equal(cw, np.array(g["codewords"], dtype=np.uint32))
        assert bits == g["total_bits"]
        n = g["n_words"]                                   # floor(bits/32) + 1: incl. padding / the courtesy zero word
        sums = stream_sums(d_out, n)
        assert ["0x%016x" % s for s in sums] == g["sums"], "stream differs from cpu_vlc_encode's"
        head = d_out[:4].cpu().numpy().view(np.uint32)
        assert ["%08x" % int(x) for x in head] == g["first_words"]
        assert int(d_out[n].item()) == 0x5A5A5A5A, "wrote past floor(bits/32)+1 words"
    finally:
        enc.close()
        del d_in, d_out
        torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_whole_stream_word_by_word(hb, ref, name):
    import torch
    if ref is None:
        pytest.skip("oracle/_ref/libref.so (the unmodified reference) was not shipped")
    torch.cuda.set_device(0)
    enc, d_in, d_out, hist, cw, cl, bits = _encode_whole(hb, torch, name)
    try:
        host = d_in.cpu().numpy()
        del d_in
        torch.cuda.empty_cache()
        n = bits // 32 + 1
        want, _ = ref.encode(host.view(np.uint32), cw, cl, n + 1)           # cpu_vlc_encode, whole input, one core
        del host
        want_t = torch.from_numpy(want[:n].view(np.int32))
        # compare in slices of 256 Mi words on the device
        step = 1 << 28
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            same = torch.equal(d_out[lo:hi], want_t[lo:hi].cuda())
            if not same:
                got = d_out[lo:hi].cpu().numpy().view(np.uint32)
                bad = np.nonzero(got != want[lo:hi])[0]
                raise AssertionError("first mismatch at word %d of %d: got %08x want %08x"
                                     % (lo + bad[0], n, got[bad[0]], want[lo + bad[0]]))
    finally:
        enc.close()
        del d_out
        torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["c2", "t1g", "c3", "c4", "c5"])
def test_round_trip_at_full_size(hb, name):
    import torch
    torch.cuda.set_device(0)
    enc, d_in, d_out, hist, cw, cl, bits = _encode_whole(hb, torch, name)
    try:
        tile = hb.capi.TILE_BYTES
        n_tiles = (d_in.numel() + tile - 1) // tile
        d_idx = torch.empty(n_tiles + 1, dtype=torch.int64, device="cuda")
        enc.tile_index(bits, d_idx)
        d_back = torch.full((d_in.numel(),), 0xEE, dtype=torch.uint8, device="cuda")
        enc.decode(d_out[: bits // 32 + 1], d_idx, cw, cl, d_back)
        step = 1 << 30
        for lo in range(0, d_in.numel(), step):
            assert torch.equal(d_back[lo:lo + step], d_in[lo:lo + step]), "decode(encode(x)) != x in GiB %d" % (lo >> 30)
    finally:
        enc.close()
        del d_in, d_out
        torch.cuda.empty_cache()
