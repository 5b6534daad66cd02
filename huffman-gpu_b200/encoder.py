"""Host-side mirror of the reference's call sequence for the encode hot path.

runVLCTest (main_test_cu.cu:52-180) does, in order:
    loadData -> runHisto (hist.cu:54)                       -> Encoder.histogram
             -> BuildTree / GenerateCodes / flatten         -> Encoder.build_codebook
    vlc_encode_kernel_sm64huff + prescanArray + pack2       -> Encoder.encode   (one kernel here)
    cpu_vlc_encode(indata, n, outdata, &outsize, cw, cwl)   -> vlc_encode(...)  (same argument list,
                                                               host arrays, runs on the GPU)
Everything below is a thin veneer over the C ABI (include/huffman_b200.h); torch is used only for
device memory and streams.  No CPU fallback: without libhuffb200.so / a B200 these calls raise.
"""
import ctypes as C

import numpy as np

from . import capi

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = capi.load()
    return _lib


def _np_u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(capi.u32p)


def _check(status, where, ctx=None):
    if status < 0:
        cuda = lib().hb_last_cuda_error(ctx) if ctx else 0
        raise capi.HBError(lib(), status, where, cuda)
    return status


def build_codebook(hist):
    """hist: 256 counts -> (codewords uint32[256], codewordlens uint32[256], max_len).
    Same tables as huffTree.h + load_data.h:40-47 (tie-exact)."""
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    assert h.size == 256
    cw = np.zeros(256, dtype=np.uint32)
    cl = np.zeros(256, dtype=np.uint32)
    rc = lib().hb_build_codebook(h.ctypes.data_as(capi.u64p), cw.ctypes.data_as(capi.u32p),
                                 cl.ctypes.data_as(capi.u32p))
    _check(rc, "hb_build_codebook")
    return cw, cl, rc


def bits_from_hist(hist, codewordlens):
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    _, clp = _np_u32(codewordlens)
    return int(lib().hb_bits_from_hist(h.ctypes.data_as(capi.u64p), clp))


def shard_offsets(shard_bits):
    b = np.ascontiguousarray(shard_bits, dtype=np.uint64)
    starts = np.zeros(b.size, dtype=np.uint64)
    total = C.c_uint64(0)
    _check(lib().hb_shard_offsets(b.ctypes.data_as(capi.u64p), int(b.size),
                                  starts.ctypes.data_as(capi.u64p), C.byref(total)), "hb_shard_offsets")
    return starts, int(total.value)


def encode_variant(codewordlens):
    _, clp = _np_u32(codewordlens)
    return lib().hb_encode_variant(clp).decode()


def vlc_encode(indata, num_elements, outdata, codewords, codewordlens):
    """cpu_vlc_encode's argument list (cpuencode.h:4-7) on host numpy arrays; returns outsize in BYTES."""
    indata = np.ascontiguousarray(indata, dtype=np.uint32)
    assert outdata.dtype == np.uint32 and outdata.flags["C_CONTIGUOUS"]
    cw, cwp = _np_u32(codewords)
    cl, clp = _np_u32(codewordlens)
    outsize = C.c_uint32(0)
    rc = lib().hb_vlc_encode(indata.ctypes.data, int(num_elements), outdata.ctypes.data,
                             C.byref(outsize), cwp, clp)
    _check(rc, "hb_vlc_encode")
    return int(outsize.value)


class PinnedBuffer:
    """Pinned host memory from hb_host_alloc, viewed as a numpy array."""

    def __init__(self, n_bytes):
        p = capi.vp()
        _check(lib().hb_host_alloc(C.byref(p), int(n_bytes)), "hb_host_alloc")
        self.ptr = p.value
        self.n_bytes = int(n_bytes)
        buf = (C.c_uint8 * self.n_bytes).from_address(self.ptr)
        self.u8 = np.frombuffer(buf, dtype=np.uint8)

    def free(self):
        if self.ptr:
            self.u8 = None
            lib().hb_host_free(self.ptr)
            self.ptr = None


class Encoder:
    """One hb_ctx: scratch for inputs up to max_bytes on one device."""

    def __init__(self, device=0, max_bytes=1 << 30):
        self._ctx = capi.vp()
        self.device = int(device)
        self.max_words = (int(max_bytes) + 3) // 4
        _check(lib().hb_init(C.byref(self._ctx), self.device, self.max_words), "hb_init")

    def close(self):
        if self._ctx:
            lib().hb_free(self._ctx)
            self._ctx = capi.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----------------------------------------------------------------------------
    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _words(t):
        """device tensor (uint8 / int32 / uint32 ...) -> (data_ptr, n_words)."""
        nbytes = t.numel() * t.element_size()
        if nbytes % 4:
            raise ValueError("input must be a whole number of uint32 words (cpuencode.h:4-7)")
        if not t.is_contiguous():
            raise ValueError("input tensor must be contiguous")
        return t.data_ptr(), nbytes // 4

    @property
    def launches(self):
        return int(lib().hb_launch_count(self._ctx))

    # ---- the five steps ----------------------------------------------------------------------
    def histogram(self, d_in):
        ptr, n_words = self._words(d_in)
        hist = np.zeros(256, dtype=np.uint64)
        _check(lib().hb_histogram(self._ctx, ptr, n_words, hist.ctypes.data_as(capi.u64p),
                                  self._stream()), "hb_histogram", self._ctx)
        return hist

    def histogram_device(self, d_in, d_hist):
        """adds into d_hist (torch int64[256] on the same device); asynchronous."""
        ptr, n_words = self._words(d_in)
        assert d_hist.numel() == 256 and d_hist.element_size() == 8
        _check(lib().hb_histogram_device(self._ctx, ptr, n_words, d_hist.data_ptr(), self._stream()),
               "hb_histogram_device", self._ctx)

    build_codebook = staticmethod(build_codebook)

    def encode(self, d_in, codewords, codewordlens, d_out, start_bit=0):
        """-> total_bits.  d_out: device tensor of 4-byte elements (capacity = numel)."""
        ptr, n_words = self._words(d_in)
        optr, cap = self._words(d_out)
        cw, cwp = _np_u32(codewords)
        cl, clp = _np_u32(codewordlens)
        bits = C.c_uint64(0)
        _check(lib().hb_encode(self._ctx, ptr, n_words, cwp, clp, optr, cap, int(start_bit),
                               C.byref(bits), self._stream()), "hb_encode", self._ctx)
        return int(bits.value)

    def encode_async(self, d_in, codewords, codewordlens, d_out, start_bit=0):
        ptr, n_words = self._words(d_in)
        optr, cap = self._words(d_out)
        cw, cwp = _np_u32(codewords)
        cl, clp = _np_u32(codewordlens)
        _check(lib().hb_encode_async(self._ctx, ptr, n_words, cwp, clp, optr, cap, int(start_bit),
                                     self._stream()), "hb_encode_async", self._ctx)

    def encode_result(self):
        bits = C.c_uint64(0)
        _check(lib().hb_encode_result(self._ctx, C.byref(bits), self._stream()), "hb_encode_result",
               self._ctx)
        return int(bits.value)

    def encode_host(self, h_in, codewords, codewordlens, h_out):
        """Host arrays (numpy uint32, ideally views of PinnedBuffer) -> (total_bits, out_bytes)."""
        assert h_in.dtype == np.uint32 and h_out.dtype == np.uint32
        cw, cwp = _np_u32(codewords)
        cl, clp = _np_u32(codewordlens)
        ob, tb = C.c_uint64(0), C.c_uint64(0)
        _check(lib().hb_vlc_encode_host(self._ctx, h_in.ctypes.data, int(h_in.size), h_out.ctypes.data,
                                        int(h_out.size), cwp, clp, C.byref(ob), C.byref(tb)),
               "hb_vlc_encode_host", self._ctx)
        return int(tb.value), int(ob.value)

    def tile_index(self, total_bits, d_tile_bits):
        """bit offsets of the tiles of the last encode job -> d_tile_bits (torch int64[n_tiles + 1])"""
        _check(lib().hb_encode_tile_index(self._ctx, int(total_bits), d_tile_bits.data_ptr(), self._stream()),
               "hb_encode_tile_index", self._ctx)

    def decode(self, d_stream, d_tile_bits, codewords, codewordlens, d_out):
        """the decoder (no reference equivalent): d_stream (int32 words) -> d_out (the encoder's input, any 4-byte
        or byte tensor of the right size)"""
        sptr, swords = self._words(d_stream)
        optr, n_words = self._words(d_out)
        cw, cwp = _np_u32(codewords)
        cl, clp = _np_u32(codewordlens)
        _check(lib().hb_decode(self._ctx, sptr, swords, d_tile_bits.data_ptr(), n_words, cwp, clp, optr,
                               self._stream()), "hb_decode", self._ctx)

    def stitch_seam(self, d_dst, d_src, n_words):
        _check(lib().hb_stitch_seam(self._ctx, d_dst.data_ptr(), d_src.data_ptr(), int(n_words),
                                    self._stream()), "hb_stitch_seam", self._ctx)

    def synth_fill(self, d_out, workload, first=0, n=None):
        """Fill a device uint8 tensor with bytes [first, first+n) of a workloads.Workload."""
        n = d_out.numel() if n is None else int(n)
        thr = np.ascontiguousarray(workload.thr, dtype=np.uint32)
        sm = None
        if workload.symmap is not None:
            sm = np.ascontiguousarray(workload.symmap, dtype=np.uint8).ctypes.data_as(capi.u8p)
        _check(lib().hb_synth_fill(self._ctx, d_out.data_ptr(), int(first), n, int(workload.seed),
                                   int(workload.mode), int(workload.nbits),
                                   thr.ctypes.data_as(capi.u32p), int(thr.size), sm, self._stream()),
               "hb_synth_fill", self._ctx)
