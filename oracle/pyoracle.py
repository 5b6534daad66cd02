"""ctypes front-end for the CPU checker (oracle/liboracle.so and oracle/_ref/libref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (huffman-gpu_b200/) never imports it.

`Oracle`  = our plain-C restatement (oracle.c), cites cpuencode.cpp / huffTree.h / load_data.h.
`Ref`     = the UNMODIFIED reference CPU path compiled by oracle/Makefile (None when the
            prebuilt oracle/_ref/libref.so is absent).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def build(quiet=True):
    """make liboracle.so (+ _ref/libref.so when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_vlc_encode.argtypes = [_u32p, C.c_uint64, _u32p, _u64p, _u64p, _u32p, _u32p]
        L.orc_vlc_encode.restype = C.c_int
        L.orc_cpu_vlc_encode.argtypes = [_u32p, C.c_uint, _u32p, _u32p, _u32p, _u32p]
        L.orc_cpu_vlc_encode.restype = None
        L.orc_histogram.argtypes = [_u8p, C.c_uint64, _u64p]
        L.orc_histogram.restype = None
        L.orc_build_codebook.argtypes = [_u64p, _u32p, _u32p]
        L.orc_build_codebook.restype = C.c_int
        L.orc_word_fnv.argtypes = [_u32p, C.c_uint64]
        L.orc_word_fnv.restype = C.c_uint64
        L.orc_vlc_decode.argtypes = [_u32p, C.c_uint64, C.c_uint64, _u8p, _u32p, _u32p]
        L.orc_vlc_decode.restype = C.c_uint64
        L.orc_synth_fill.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                     _u32p, C.c_int, _u8p]
        L.orc_synth_fill.restype = None
        self.L = L

    # -- encode -------------------------------------------------------------------------
    def encode(self, words, cw, cl, total_bits_hint=None):
        """words: uint32[n]; returns (out_words uint32[floor(bits/32)+1], total_bits, out_bytes)."""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        cw = np.ascontiguousarray(cw, dtype=np.uint32)
        cl = np.ascontiguousarray(cl, dtype=np.uint32)
        if total_bits_hint is None:
            sym = words.view(np.uint8)
            total_bits_hint = int(cl.astype(np.uint64)[sym].sum()) if sym.size else 0
        out = np.zeros(total_bits_hint // 32 + 2, dtype=np.uint32)
        ob, tb = C.c_uint64(0), C.c_uint64(0)
        rc = self.L.orc_vlc_encode(_ptr(words, _u32p), words.size, _ptr(out, _u32p),
                                   C.byref(ob), C.byref(tb), _ptr(cw, _u32p), _ptr(cl, _u32p))
        if rc != 0:
            raise ValueError("orc_vlc_encode: codeword length > 31 (outside parity domain)")
        assert tb.value == total_bits_hint, (tb.value, total_bits_hint)
        return out[: tb.value // 32 + 1], tb.value, ob.value

    def histogram(self, data):
        data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        h = np.zeros(256, dtype=np.uint64)
        self.L.orc_histogram(_ptr(data, _u8p), data.size, _ptr(h, _u64p))
        return h

    def build_codebook(self, hist):
        hist = np.ascontiguousarray(hist, dtype=np.uint64)
        cw = np.zeros(256, dtype=np.uint32)
        cl = np.zeros(256, dtype=np.uint32)
        rc = self.L.orc_build_codebook(_ptr(hist, _u64p), _ptr(cw, _u32p), _ptr(cl, _u32p))
        return rc, cw, cl

    def word_fnv(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        return int(self.L.orc_word_fnv(_ptr(words, _u32p), words.size))

    def decode(self, stream, bit0, n_symbols, cw, cl):
        stream = np.ascontiguousarray(stream, dtype=np.uint32)
        cw = np.ascontiguousarray(cw, dtype=np.uint32)
        cl = np.ascontiguousarray(cl, dtype=np.uint32)
        out = np.zeros((n_symbols + 3) // 4 * 4, dtype=np.uint8)
        end = self.L.orc_vlc_decode(_ptr(stream, _u32p), bit0, n_symbols, _ptr(out, _u8p),
                                    _ptr(cw, _u32p), _ptr(cl, _u32p))
        if end == 2 ** 64 - 1:
            raise ValueError("orc_vlc_decode: dead prefix")
        return out, int(end)

    def synth_fill(self, first, n, seed, mode, nbits, thr, symmap=None):
        thr = np.ascontiguousarray(thr, dtype=np.uint32)
        out = np.empty(n, dtype=np.uint8)
        sm = None
        if symmap is not None:
            symmap = np.ascontiguousarray(symmap, dtype=np.uint8)
            sm = _ptr(symmap, _u8p)
        self.L.orc_synth_fill(_ptr(out, _u8p), first, n, seed, mode, nbits, _ptr(thr, _u32p),
                              int(thr.size), sm)
        return out


class Ref:
    """The unmodified reference CPU path (cpuencode.cpp cpu_vlc_encode; huffTree.h via ref_shim.cpp)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/cpuencode.cpp"):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        L.cpu_vlc_encode.argtypes = [_u32p, C.c_uint, _u32p, _u32p, _u32p, _u32p]
        L.cpu_vlc_encode.restype = None
        L.ref_build_codebook.argtypes = [_u32p, _u32p, _u32p]
        L.ref_build_codebook.restype = C.c_int
        self.L = L

    def encode(self, words, cw, cl, cap_words):
        """cpu_vlc_encode(indata, num_elements, outdata, &outsize, codewords, codewordlens)."""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        cw = np.ascontiguousarray(cw, dtype=np.uint32)
        cl = np.ascontiguousarray(cl, dtype=np.uint32)
        out = np.zeros(cap_words, dtype=np.uint32)
        outsize = C.c_uint32(0)
        self.L.cpu_vlc_encode(_ptr(words, _u32p), words.size, _ptr(out, _u32p),
                              C.byref(outsize), _ptr(cw, _u32p), _ptr(cl, _u32p))
        return out, int(outsize.value)

    def build_codebook(self, freqs):
        freqs = np.ascontiguousarray(freqs, dtype=np.uint32)
        cw = np.zeros(256, dtype=np.uint32)
        cl = np.zeros(256, dtype=np.uint32)
        rc = self.L.ref_build_codebook(_ptr(freqs, _u32p), _ptr(cw, _u32p), _ptr(cl, _u32p))
        return rc, cw, cl


class RefGpu:
    """The unmodified reference GPU pipeline (encode -> scan -> memset + pack2) behind oracle/ref_gpu_shim.cu.
    Valid only inside the reference's limits: 4 codewords <= 64 bits, <= 8192 bits per 1024-symbol block,
    < 2^32 output bits, word count a multiple of 4096."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libref_gpu.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/pack_kernels.cu"):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        L.ref_gpu_pipeline.argtypes = [C.c_void_p, C.c_uint, _u32p, _u32p, C.c_void_p, C.c_uint64, _u32p,
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
        L.ref_gpu_pipeline.restype = C.c_int
        self.L = L

    def run(self, d_in_ptr, n_words, cw, cl, d_packed_ptr, packed_bytes, repeats=1):
        """-> (total_bits, ms_encode, ms_scan, ms_memset_pack); device pointers are plain integers"""
        cw = np.ascontiguousarray(cw, dtype=np.uint32)
        cl = np.ascontiguousarray(cl, dtype=np.uint32)
        bits = C.c_uint32(0)
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        rc = self.L.ref_gpu_pipeline(d_in_ptr, n_words, _ptr(cw, _u32p), _ptr(cl, _u32p), d_packed_ptr,
                                     packed_bytes, C.byref(bits), C.byref(a), C.byref(b), C.byref(c), repeats)
        if rc != 0:
            raise RuntimeError("ref_gpu_pipeline failed: %d" % rc)
        return int(bits.value), a.value, b.value, c.value


def try_ref_gpu():
    try:
        return RefGpu()
    except (FileNotFoundError, OSError):
        return None


def try_ref():
    try:
        return Ref()
    except (FileNotFoundError, OSError):
        return None
