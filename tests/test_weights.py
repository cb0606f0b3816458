"""CPU: checkpoint ingestion — the committed .npz, a reference-style .pth, and the reference's .onnx
export read without the onnx package (protobuf walked by hand)."""
import os

import numpy as np
import pytest

from ribodetector_b200.utils.weights import STATE_KEYS, load_weights
from ribodetector_b200.utils.onnx_weights import load_onnx_state_dict

REF_DATA = "/root/reference/ribodetector/data"


def _varint(x):
    out = b""
    while True:
        c = x & 0x7F
        x >>= 7
        out += bytes([c | (0x80 if x else 0)])
        if not x:
            return out


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _tensor(name, arr, packed_dims):
    dims = b"".join(_varint(d) for d in arr.shape)
    body = _ld(1, dims) if packed_dims else b"".join(_varint(1 << 3) + _varint(d) for d in arr.shape)
    body += _varint(2 << 3) + _varint(1) + _ld(8, name.encode()) + _ld(9, arr.astype("<f4").tobytes())
    return body


def _fake_onnx(sd, path, packed_dims):
    """Minimal ModelProto{graph{initializer x5}} in ONNX gate order i, o, f, c."""
    H = 128
    inv = np.concatenate([np.arange(0, H), np.arange(3 * H, 4 * H), np.arange(H, 2 * H), np.arange(2 * H, 3 * H)])
    W = np.stack([sd["rnn.weight_ih_l0"][inv], sd["rnn.weight_ih_l0_reverse"][inv]])
    R = np.stack([sd["rnn.weight_hh_l0"][inv], sd["rnn.weight_hh_l0_reverse"][inv]])
    B = np.stack([np.concatenate([sd["rnn.bias_ih_l0"][inv], sd["rnn.bias_hh_l0"][inv]]),
                  np.concatenate([sd["rnn.bias_ih_l0_reverse"][inv], sd["rnn.bias_hh_l0_reverse"][inv]])])
    graph = b"".join(_ld(5, _tensor(n, a, packed_dims)) for n, a in
                     (("out.weight", sd["out.weight"]), ("out.bias", sd["out.bias"]), ("85", W), ("86", R), ("87", B)))
    graph += _ld(2, b"torch-jit-export")                                   # GraphProto.name, ignored
    model = _varint(1 << 3) + _varint(6) + _ld(2, b"pytorch") + _ld(7, graph)
    with open(path, "wb") as f:
        f.write(model)


def test_npz_has_reference_keys_and_shapes(weights):
    assert tuple(weights) == STATE_KEYS
    assert weights["rnn.weight_hh_l0"].shape == (512, 128) and weights["out.weight"].shape == (2, 256)
    n_params = sum(v.size for v in weights.values())
    assert n_params == 137730


@pytest.mark.parametrize("packed_dims", [True, False])
def test_onnx_round_trip_restores_pytorch_gate_order(weights, tmp_path, packed_dims):
    p = tmp_path / "m.onnx"
    _fake_onnx(weights, p, packed_dims)
    got = load_weights(str(p))
    for k in STATE_KEYS:
        assert np.array_equal(got[k], weights[k]), k


def test_onnx_without_the_initializers_is_rejected(tmp_path):
    p = tmp_path / "empty.onnx"
    p.write_bytes(_ld(7, _ld(2, b"g")))
    with pytest.raises(KeyError):
        load_onnx_state_dict(str(p))


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not mounted (GPU box)")
def test_shipped_onnx_and_pth_equal_the_committed_npz(weights):
    base = os.path.join(REF_DATA, "ribodetector_600k_variable_len70_101_epoch47")
    for ext in (".onnx", ".pth"):
        got = load_weights(base + ext)
        for k in STATE_KEYS:
            assert np.array_equal(got[k], weights[k]), (ext, k)
