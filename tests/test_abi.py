"""CPU: the C-ABI library loads, exports every symbol include/rd_b200.h declares, and fails
loudly (no fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from ribodetector_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rd_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load_library()
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
        assert s in _lib.SIGNATURES, "no ctypes signature for " + s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.rd_abi_version() == 1


def test_enum_values_match_header():
    text = open(os.path.join(ROOT, "include", "rd_b200.h")).read()
    d = dict(re.findall(r"#define\s+(RD_[A-Z_0-9]+)\s+(-?\d+)", text))
    assert int(d["RD_SEM_PACKED"]) == _lib.SEM["packed"] and int(d["RD_SEM_PADDED"]) == _lib.SEM["padded"]
    assert [int(d["RD_PREC_" + k]) for k in ("FP32", "TC_EXACT", "TC_FAST")] == [0, 1, 2]
    assert [int(d["RD_PAIR_" + k.upper()]) for k in ("none", "rrna", "norrna", "both")] == \
        [_lib.PAIR[k] for k in ("none", "rrna", "norrna", "both")]
    assert int(d["RD_MAX_LEN"]) == _lib.RD_MAX_LEN


def test_create_without_gpu_fails_loudly(weights):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ribodetector_b200.model import SeqModel
    m = SeqModel(4, 128, 1, 2)
    m.load_state_dict(weights)
    with pytest.raises(RuntimeError):
        m.to("cuda")
    with pytest.raises(RuntimeError):
        m.to("cpu")                        # no CPU path exists
    lib = _lib.load_library()
    h = ctypes.c_void_p()
    import numpy as np
    w = [np.zeros(4, np.float32)] * 10
    rc = lib.rd_create(0, *[ctypes.c_void_p(a.ctypes.data) for a in w], 128, ctypes.byref(h))
    assert rc == _lib.RD_ERR_CUDA and h.value is None
    assert b"CUDA" in lib.rd_last_error(None) or b"device" in lib.rd_last_error(None)


def test_model_argument_validation(weights):
    from ribodetector_b200.model import SeqModel
    for bad_h in (16, 100, 288, 0):        # hidden_size: multiples of 32 between 32 and 256
        with pytest.raises(ValueError):
            SeqModel(4, bad_h, 1, 2)
    SeqModel(4, 64, 1, 2)
    SeqModel(4, 256, 1, 2)
    with pytest.raises(ValueError):
        SeqModel(4, 128, 2, 2)
    with pytest.raises(RuntimeError):
        SeqModel(4, 64, 1, 2).load_state_dict(weights)      # a 128-unit checkpoint into a 64-unit model: size mismatch
    m = SeqModel(4, 128, 1, 2)
    bad = dict(weights)
    bad.pop("out.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)
    bad = dict(weights)
    bad["out.weight"] = bad["out.weight"][:, :10]
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)
    with pytest.raises(RuntimeError):
        m.classify_host(b"", [0], 100)     # no handle yet
