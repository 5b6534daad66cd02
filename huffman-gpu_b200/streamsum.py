"""Position-sensitive checksums of a packed word stream, computable in parallel (numpy on the host, torch on a
GPU, identical int64 wrap-around arithmetic): the size-independent parity check for streams too large to keep
as fixtures (C4: 0.7 GB, C5: 4.3 GB of output).  tests/golden/streams.json holds the sums of the streams the
UNMODIFIED reference cpu_vlc_encode produces (oracle/make_golden_streams.py); bench.py and the -m gpu tests
compare the CUDA path's streams -- single-GPU and stitched -- against them.

  S1 = sum w_i,  S2 = sum (i+1) * w_i,  S3 = sum m(i+1) * w_i   (mod 2^64; m = an odd multiplicative hash)
"""
import numpy as np

_K = 0x9E3779B97F4A7C15 - (1 << 64)          # as int64
_CHUNK = 1 << 24


def _sums(xp, w_i64, first_index, arange, asint):
    i1 = arange(first_index + 1, first_index + 1 + w_i64.shape[0])
    m = (((i1 * _K) >> 20) ^ i1) | 1
    return asint((w_i64).sum()), asint((w_i64 * i1).sum()), asint((w_i64 * m).sum())


def stream_sums(words, n_words=None):
    """words: numpy uint32/int32 array or torch int32 tensor (host or device).  -> (S1, S2, S3) as ints mod 2^64."""
    n = int(words.shape[0] if n_words is None else n_words)
    s = [0, 0, 0]
    is_np = isinstance(words, np.ndarray)
    if not is_np:
        import torch
    for lo in range(0, n, _CHUNK):
        hi = min(n, lo + _CHUNK)
        if is_np:
            w = words[lo:hi].view(np.uint32).astype(np.int64)
            with np.errstate(over="ignore"):
                part = _sums(np, w, lo, lambda a, b: np.arange(a, b, dtype=np.int64), int)
        else:
            w = words[lo:hi].to(torch.int64) & 0xFFFFFFFF
            part = _sums(torch, w, lo, lambda a, b: torch.arange(a, b, dtype=torch.int64, device=words.device),
                         lambda t: int(t.item()))
        for k in range(3):
            s[k] = (s[k] + part[k]) & (2 ** 64 - 1)
    return tuple(s)
