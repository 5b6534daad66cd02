"""The five BASELINE.json configurations as deterministic synthetic inputs.

The reference ships one fixture (data/test1024_H2.206587175259.in) and no usable generator
(testdatagen.h:62-67 draws uniform words only), so C2..C5 are defined here as
(distribution, size, seed) and realised by the same counter-based arithmetic on the device
(hb_synth_fill, csrc/hb_misc.cu) and on the host (oracle/oracle.c orc_synth_fill, tests only).
"""
import lzma
import os
from dataclasses import dataclass, field

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


@dataclass
class Workload:
    name: str
    n_bytes: int
    mode: int                 # 0 = iid draws from thresholds, 1 = exact counts via a bijection
    seed: int
    thr: np.ndarray           # uint32 thresholds: first k with u < thr[k] (last symbol implicit)
    nbits: int = 0            # mode 1: positions live in [0, 2^nbits)
    symmap: np.ndarray = None
    note: str = ""
    probs: np.ndarray = field(default=None, repr=False)

    @property
    def n_words(self):
        return self.n_bytes // 4

    def entropy_bits(self):
        p = self.probs[self.probs > 0]
        return float(-(p * np.log2(p)).sum())


def _geometric(r, K):
    p = r ** np.arange(K, dtype=np.float64)
    return p / p.sum()


def _thresholds_from_probs(p):
    cum = np.cumsum(p)
    thr = np.minimum(np.floor(cum * 4294967296.0), 4294967295.0).astype(np.uint64)
    return thr[:-1].astype(np.uint32) if len(thr) > 1 else np.array([0xFFFFFFFF], dtype=np.uint32)


def _iid(name, n_bytes, r, K, seed, note):
    p = _geometric(r, K)
    thr = _thresholds_from_probs(p)
    # hb_synth_fill searches k in [0, K-1) and falls through to K-1: pass K-1 thresholds + sentinel
    thr = np.concatenate([thr, np.array([0xFFFFFFFF], dtype=np.uint32)])
    return Workload(name, n_bytes, 0, seed, thr, 0, None, note, p)


def fibonacci_counts():
    """C4: 32 symbols, counts 376*Fib(k) (k=1..32), remainder on the most frequent -> 2^31 bytes."""
    fib = [1, 1]
    while len(fib) < 32:
        fib.append(fib[-1] + fib[-2])
    counts = 376 * np.array(fib, dtype=np.uint64)
    counts[31] += np.uint64(2 ** 31) - counts.sum()
    return counts


def get(name, n_bytes=None):
    """name in {c2, t1g, c3, c4, c5}; n_bytes overrides the size (mode 0 only)."""
    if name == "c2":
        w = _iid("c2", 1 << 28, 0.5486, 22, 0xB2000002, "256 MiB, truncated geometric r=0.5486 K=22, H~2.2")
    elif name == "t1g":
        w = _iid("t1g", 1 << 30, 0.5486, 22, 0xB2000002, "1 GiB, C2's distribution (north_star target case)")
    elif name == "c3":
        w = _iid("c3", 1 << 30, 0.994867, 256, 0xB2000003, "1 GiB, near-uniform r=0.994867 K=256, H~7.9")
    elif name == "c5":
        w = _iid("c5", 1 << 33, 0.843583, 71, 0xB2000005, "8 GiB, geometric r=0.843583 K=71, H~4.0")
    elif name == "c4":
        counts = fibonacci_counts()
        cum = np.cumsum(counts)
        thr = cum.astype(np.uint64)
        thr = np.minimum(thr, 0xFFFFFFFF).astype(np.uint32)
        w = Workload("c4", 1 << 31, 1, 0xB2000004, thr, 31, None,
                     "2 GiB, exact Fibonacci counts (code lengths 1..31)",
                     counts.astype(np.float64) / float(counts.sum()))
        if n_bytes is not None and n_bytes != w.n_bytes:
            raise ValueError("c4 has exact counts over 2^31 positions; size is fixed")
        return w
    else:
        raise KeyError(name)
    if n_bytes is not None:
        w.n_bytes = int(n_bytes)
    return w


def c1_fixture_bytes():
    """The reference fixture data/test1024_H2.206587175259.in: 4 repetitions of a 256 KiB period
    (tests/golden/c1_period.bin.xz; sha256 of the reassembled file is checked by the tests)."""
    with open(os.path.join(_GOLDEN, "c1_period.bin.xz"), "rb") as f:
        period = lzma.decompress(f.read())
    return np.frombuffer(period * 4, dtype=np.uint8)
