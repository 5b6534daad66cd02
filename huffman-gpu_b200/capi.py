"""ctypes binding of libhuffb200.so (include/huffman_b200.h).  No fallback: if the shared library
is missing or cannot be loaded the import of this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HB_LIB") or os.path.join(_HERE, "libhuffb200.so")    # $HB_LIB: A/B builds

HB_OK = 0
HB_ERR_ARG, HB_ERR_CAPACITY, HB_ERR_CODELEN, HB_ERR_CODEWORD = -1, -2, -3, -4
HB_ERR_CUDA, HB_ERR_NOMEM, HB_ERR_STATE, HB_ERR_NCCL = -5, -6, -7, -8
HB_UNIQUE_ID_BYTES = 128
HB_MAX_RANKS = 64

TILE_BYTES = None           # hb::kTileBytes, filled in by load()

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p



class ShardPlan(C.Structure):
    """hb_shard_plan (include/huffman_b200.h)"""
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("max_len", C.c_int32), ("phase", C.c_uint32),
                ("shard_bits", C.c_uint64), ("start_bit", C.c_uint64), ("total_bits", C.c_uint64),
                ("first_word", C.c_uint64), ("local_words", C.c_uint64), ("local_offset_words", C.c_uint32),
                ("reserved", C.c_uint32)]


planp = C.POINTER(ShardPlan)

# name -> (restype, argtypes); mirrors include/huffman_b200.h declaration by declaration
SIGNATURES = {
    "hb_init": (C.c_int, [C.POINTER(vp), C.c_int, C.c_uint64]),
    "hb_free": (None, [vp]),
    "hb_histogram": (C.c_int, [vp, vp, C.c_uint64, u64p, vp]),
    "hb_histogram_device": (C.c_int, [vp, vp, C.c_uint64, vp, vp]),
    "hb_build_codebook": (C.c_int, [u64p, u32p, u32p]),
    "hb_bits_from_hist": (C.c_uint64, [u64p, u32p]),
    "hb_encode": (C.c_int, [vp, vp, C.c_uint64, u32p, u32p, vp, C.c_uint64, C.c_uint64, u64p, vp]),
    "hb_encode_async": (C.c_int, [vp, vp, C.c_uint64, u32p, u32p, vp, C.c_uint64, C.c_uint64, vp]),
    "hb_encode_result": (C.c_int, [vp, u64p, vp]),
    "hb_vlc_encode": (C.c_int, [vp, C.c_uint, vp, u32p, u32p, u32p]),
    "hb_vlc_encode_host": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, u32p, u32p, u64p, u64p]),
    "hb_host_alloc": (C.c_int, [C.POINTER(vp), C.c_uint64]),
    "hb_host_free": (None, [vp]),
    "hb_comm_unique_id": (C.c_int, [u8p]),
    "hb_comm_init": (C.c_int, [C.POINTER(vp), vp, C.c_int, C.c_int, u8p]),
    "hb_comm_adopt": (C.c_int, [C.POINTER(vp), vp, vp, C.c_int, C.c_int]),
    "hb_comm_free": (None, [vp]),
    "hb_comm_last_nccl_error": (C.c_int, [vp]),
    "hb_shard_plan_build": (C.c_int, [vp, vp, C.c_uint64, u32p, u32p, u64p, planp, vp]),
    "hb_comm_plan_offsets": (C.c_int, [vp, u64p, u64p]),
    "hb_shard_encode_async": (C.c_int, [vp, vp, C.c_uint64, u32p, u32p, vp, C.c_uint64, planp, vp]),
    "hb_shard_encode_result": (C.c_int, [vp, u64p, vp]),
    "hb_stitch_open": (C.c_int, [vp, C.c_uint64, C.c_int, C.POINTER(vp), vp]),
    "hb_stitch_push": (C.c_int, [vp, vp, planp, vp]),
    "hb_stitch_close": (C.c_int, [vp]),
    "hb_shard_encode_direct_async": (C.c_int, [vp, vp, C.c_uint64, u32p, u32p, planp, vp]),
    "hb_shard_offsets": (C.c_int, [u64p, C.c_int, u64p, u64p]),
    "hb_stitch_seam": (C.c_int, [vp, vp, vp, C.c_uint64, vp]),
    "hb_encode_tile_index": (C.c_int, [vp, C.c_uint64, vp, vp]),
    "hb_decode": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, u32p, u32p, vp, vp]),
    "hb_synth_fill": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, u32p,
                                C.c_int, u8p, vp]),
    "hb_tile_bytes": (C.c_uint32, []),
    "hb_launch_count": (C.c_uint64, [vp]),
    "hb_encode_variant": (C.c_char_p, [u32p]),
    "hb_strerror": (C.c_char_p, [C.c_int]),
    "hb_last_cuda_error": (C.c_int, [vp]),
    "hb_version": (C.c_char_p, []),
}


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    global TILE_BYTES
    TILE_BYTES = int(lib.hb_tile_bytes())
    return lib


class HBError(RuntimeError):
    def __init__(self, lib, status, where, cuda=0):
        self.status = status
        self.cuda_error = cuda
        msg = lib.hb_strerror(status).decode()
        super().__init__("%s: %s (status %d%s)" % (where, msg, status,
                                                    ", cudaError %d" % cuda if cuda else ""))
