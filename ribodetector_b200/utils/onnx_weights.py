"""Read the reference's ``.onnx`` checkpoint without the onnx package.

``ribodetector_cpu`` loads ``state_file.replace('.pth', '.onnx')`` (``detect_cpu.py:74-75``); the file
is the export of ``model_cpu.SeqModel`` (``convert_onnx.py:45-54``) and carries five initializers:
LSTM ``W [2, 4H, 4]``, ``R [2, 4H, H]``, ``B [2, 8H]`` in ONNX gate order i, o, f, c, plus
``out.weight [2, 2H]`` and ``out.bias [2]``.  This module walks the protobuf wire format just far
enough to pull those tensors out (ModelProto.graph = 7, GraphProto.initializer = 5, TensorProto dims = 1,
data_type = 2, float_data = 4, name = 8, raw_data = 9) and returns them under the ``.pth`` key names
in PyTorch gate order i, f, g, o."""
import numpy as np


def _varint(b, i):
    x = s = 0
    while True:
        c = b[i]
        i += 1
        x |= (c & 0x7F) << s
        if c < 0x80:
            return x, i
        s += 7


def _fields(b):
    """Yield (field_number, wire_type, value) of one protobuf message; bytes for length-delimited."""
    i, n = 0, len(b)
    while i < n:
        key, i = _varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            ln, i = _varint(b, i)
            v, i = b[i:i + ln], i + ln
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % w)
        yield f, w, v


def _tensor(b):
    dims, dtype, name, raw, floats = [], None, "", None, None
    for f, w, v in _fields(b):
        if f == 1:
            if w == 0:
                dims.append(v)
            else:                                   # packed repeated int64
                j = 0
                while j < len(v):
                    d, j = _varint(v, j)
                    dims.append(d)
        elif f == 2:
            dtype = v
        elif f == 8:
            name = bytes(v).decode("utf-8", "replace")
        elif f == 9:
            raw = bytes(v)
        elif f == 4:
            floats = np.frombuffer(bytes(v), "<f4") if w == 2 else None
    if dtype != 1:                                  # TensorProto.FLOAT
        return name, None
    data = np.frombuffer(raw, "<f4") if raw is not None else floats
    if data is None or data.size != int(np.prod(dims, dtype=np.int64)):
        return name, None
    return name, data.reshape(dims).astype(np.float32)


def read_initializers(path):
    with open(path, "rb") as f:
        model = memoryview(f.read())
    out = {}
    for f1, w1, graph in _fields(model):
        if f1 == 7 and w1 == 2:
            for f2, w2, t in _fields(graph):
                if f2 == 5 and w2 == 2:
                    name, arr = _tensor(t)
                    if arr is not None:
                        out[name] = arr
    return out


def load_onnx_state_dict(path, hidden=128):
    """→ dict with the reference ``.pth`` keys (float32, PyTorch gate order)."""
    H = hidden
    by_shape = {}
    for name, a in read_initializers(path).items():
        by_shape.setdefault(a.shape, []).append(a)
    try:
        (W,), (R,), (B,) = by_shape[(2, 4 * H, 4)], by_shape[(2, 4 * H, H)], by_shape[(2, 8 * H)]
        (ow,), (ob,) = by_shape[(2, 2 * H)], by_shape[(2,)]
    except (KeyError, ValueError):
        raise KeyError("%s does not hold the five RiboDetector initializers (found shapes %s)"
                       % (path, sorted(by_shape)))
    order = np.concatenate([np.arange(0, H), np.arange(2 * H, 3 * H), np.arange(3 * H, 4 * H), np.arange(H, 2 * H)])
    sd = {}
    for d, suffix in ((0, ""), (1, "_reverse")):
        sd["rnn.weight_ih_l0" + suffix] = W[d][order]
        sd["rnn.weight_hh_l0" + suffix] = R[d][order]
        sd["rnn.bias_ih_l0" + suffix] = B[d][:4 * H][order]
        sd["rnn.bias_hh_l0" + suffix] = B[d][4 * H:][order]
    sd["out.weight"], sd["out.bias"] = ow, ob
    return {k: np.ascontiguousarray(v, np.float32) for k, v in sd.items()}
