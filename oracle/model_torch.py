"""torch-CPU restatement of the reference model (test infrastructure; see oracle/__init__.py).

The reference delegates all arithmetic to PyTorch ``nn.LSTM`` + ``nn.Linear``
(``ribodetector/model/model.py:16-24``); this file calls the same library on CPU fp32 and
restates only the glue around it:

* packed semantics (``ribodetector`` GPU CLI): ``detect.py:681-685`` (encode + ``pack_sequence``
  with ``enforce_sorted=False``) → ``model.py:32-37`` (``forward1``) → ``model.py:114-119``
  (``last_items``: output row of each read's last VALID step, restored to input order).
* padded semantics (``ribodetector_cpu``): ``detect_cpu.py:699-700`` (zero-pad to ``-l``) →
  ``model_cpu.py:29-37`` (``forward_last``) → ``model_cpu.py:57-62`` (row of the last step whose
  one-hot row is non-zero; all-zero read → step T-1).

Outputs are RAW LOGITS [B,2] (no softmax in the reference, ``model.py:36-37``).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_sequence, pad_packed_sequence

from . import encoders

STATE_KEYS = (
    "rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0",
    "rnn.weight_ih_l0_reverse", "rnn.weight_hh_l0_reverse",
    "rnn.bias_ih_l0_reverse", "rnn.bias_hh_l0_reverse",
    "out.weight", "out.bias",
)


class TorchOracle:
    def __init__(self, weights, hidden_size=128, num_classes=2):
        """weights: mapping of the reference state_dict keys → float32 arrays."""
        self.rnn = nn.LSTM(input_size=4, hidden_size=hidden_size, num_layers=1,
                           batch_first=True, bidirectional=True)
        self.out = nn.Linear(hidden_size * 2, num_classes)
        sd = {k: torch.as_tensor(np.asarray(weights[k], dtype=np.float32)) for k in STATE_KEYS}
        self.rnn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("rnn.")})
        self.out.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("out.")})
        self.rnn.eval()
        self.out.eval()

    @torch.no_grad()
    def logits_packed(self, reads, max_len, batch=4096):
        outs = []
        for s in range(0, len(reads), batch):
            chunk = reads[s:s + batch]
            xs = [torch.from_numpy(encoders.encode_read(encoders.as_bytes(r)[:max_len].tobytes()))
                  for r in chunk]
            if any(x.shape[0] == 0 for x in xs):
                # pack_sequence raises on zero-length sequences, as does the reference
                raise RuntimeError("zero-length read cannot be packed (reference behaviour)")
            packed = pack_sequence(xs, enforce_sorted=False)
            r_out, _ = self.rnn(packed, None)
            padded, lens = pad_packed_sequence(r_out, batch_first=True)
            last = padded[torch.arange(len(chunk)), lens - 1]
            outs.append(self.out(last).numpy())
        return np.concatenate(outs, 0) if outs else np.zeros((0, 2), np.float32)

    @torch.no_grad()
    def logits_padded(self, reads, max_len, batch=4096):
        outs = []
        for s in range(0, len(reads), batch):
            chunk = reads[s:s + batch]
            x = torch.from_numpy(np.stack(
                [encoders.encode_variable_len_read(r, max_len) for r in chunk]))
            r_out, _ = self.rnn(x, None)
            t = x.shape[1]
            idx = t - 1 - torch.flip(x.sum(2), [1]).argmax(1)
            last = r_out[torch.arange(len(chunk)), idx]
            outs.append(self.out(last).numpy())
        return np.concatenate(outs, 0) if outs else np.zeros((0, 2), np.float32)

    def logits(self, reads, max_len, semantics="packed", **kw):
        if semantics == "packed":
            return self.logits_packed(reads, max_len, **kw)
        if semantics == "padded":
            return self.logits_padded(reads, max_len, **kw)
        raise ValueError(semantics)
