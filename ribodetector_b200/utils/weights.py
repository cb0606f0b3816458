"""Checkpoint ingestion.  The reference loads ``torch.load(state_file)['state_dict']`` and calls
``load_state_dict`` (``ribodetector/detect.py:101,115-116``); the keys are kept verbatim.  The
shipped checkpoint's ten fp32 tensors are carried in this repo as an ``.npz`` (data, 0.55 MB)
because the GPU box only receives the repo; a reference ``.pth`` is read through torch."""
import os
import numpy as np

STATE_KEYS = (
    "rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0",
    "rnn.weight_ih_l0_reverse", "rnn.weight_hh_l0_reverse",
    "rnn.bias_ih_l0_reverse", "rnn.bias_hh_l0_reverse",
    "out.weight", "out.bias",
)

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def default_weights_path():
    return os.path.join(_HERE, "data", "ribodetector_600k_variable_len70_101_epoch47.npz")


def load_weights(path=None):
    """→ dict key → contiguous float32 ndarray.  Accepts .npz, a reference .pth/.pt, or the reference's
    .onnx export (what ribodetector_cpu loads, detect_cpu.py:74-75)."""
    path = path or default_weights_path()
    if path.endswith(".onnx"):
        from .onnx_weights import load_onnx_state_dict
        sd = load_onnx_state_dict(path)
    elif path.endswith(".npz"):
        with np.load(path) as z:
            sd = {k: z[k] for k in z.files}
    else:
        import torch
        st = torch.load(path, map_location="cpu")
        sd = st["state_dict"] if "state_dict" in st else st
        sd = {k[7:] if k.startswith("module.") else k: v.detach().cpu().numpy()
              for k, v in sd.items()}
    missing = [k for k in STATE_KEYS if k not in sd]
    if missing:
        raise KeyError("checkpoint lacks keys: %s" % ", ".join(missing))
    return {k: np.ascontiguousarray(sd[k], dtype=np.float32) for k in STATE_KEYS}
