"""ctypes binding of librd_b200.so (C ABI: include/rd_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The
library is built in-tree by ``__graft_entry__.build()`` / ``ribodetector_b200/csrc/build.sh``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RD_B200_LIB") or os.path.join(_HERE, "librd_b200.so")     # (RD_B200_LIB: A/B builds, tools/)

RD_OK, RD_ERR_INVALID, RD_ERR_CUDA, RD_ERR_EMPTY_READ, RD_ERR_NOMEM, RD_ERR_UNSUPPORTED, RD_ERR_PARSE = range(7)
FMT = {"fastq": 0, "fasta": 1}
SEM = {"packed": 0, "padded": 1}
PREC = {"fp32": 0, "tc_exact": 1, "tc_fast": 2, "tc_auto": 3, "tc_mixed": 4, "tc_mixed_raw": 5}
PAIR = {"none": 0, "rrna": 1, "norrna": 2, "both": 3}
ONEHOT = {"ragged": 0, "padded": 1}
RD_MAX_LEN = 4096

_c = ctypes
_vp, _i, _i64 = _c.c_void_p, _c.c_int, _c.c_int64

# name → (restype, argtypes); every symbol include/rd_b200.h declares
SIGNATURES = {
    "rd_abi_version": (_i, []),
    "rd_create": (_i, [_i] + [_vp] * 10 + [_i, _c.POINTER(_vp)]),
    "rd_destroy": (None, [_vp]),
    "rd_last_error": (_c.c_char_p, [_vp]),
    "rd_reserve": (_i, [_vp, _i64, _i]),
    "rd_encode_onehot": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "rd_classify": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "rd_pair_combine": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "rd_classify_host": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rd_classify_pairs_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rd_kernel_launches": (_i64, [_vp]),
    "rd_reverse_lut": (_i, [_vp, _i, _vp]),
    "rd_set_timing": (_i, [_vp, _i]),
    "rd_get_timing": (_i, [_vp, _vp, _vp, _i]),
    "rd_scan_fastx": (_i64, [_vp, _i64, _i, _i, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i]),
    "rd_partition_records": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "rd_fastx_last_error": (_c.c_char_p, []),
    "rd_bgzf_inflate": (_i64, [_vp, _i64, _vp, _i64, _vp, _i]),
    "rd_gz_open": (_vp, []),
    "rd_gz_inflate": (_i64, [_vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "rd_gz_close": (None, [_vp]),
    "rd_scan_fastq_device": (_i, [_vp, _vp, _i64, _i, _i64, _vp, _vp, _vp]),
    "rd_classify_records": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "rd_partition_records_device": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "rd_fastq_submit": (_i, [_vp, _i, _i, _vp, _i64, _vp, _i64, _i, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rd_fastq_collect": (_i, [_vp, _i, _vp, _vp]),
    "rd_scan_fasta_device": (_i, [_vp, _vp, _i64, _i64, _i, _i64, _vp, _vp, _vp]),
    "rd_partition_fasta_device": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "rd_fasta_submit": (_i, [_vp, _i, _i, _vp, _i64, _vp, _i64, _i, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "librd_b200.so is not built (%s). Build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"` or ribodetector_b200/csrc/build.sh — there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RdError(RuntimeError):
    pass


def check(lib, handle, rc, what):
    if rc == RD_OK:
        return
    msg = lib.rd_last_error(handle)
    msg = msg.decode("utf-8", "replace") if msg else ""
    text = "%s failed (code %d): %s" % (what, rc, msg)
    if rc == RD_ERR_INVALID:
        raise ValueError(text)
    raise RdError(text)
