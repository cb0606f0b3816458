"""Independent NumPy restatement of the forward pass (test infrastructure; see
oracle/__init__.py).  Arbiter between torch-CPU and the CUDA kernels: run it in float64.

Restates the published LSTM cell (PyTorch ``nn.LSTM`` docs; gate row order i, f, g, o in
``weight_*``), as used by ``ribodetector/model/model.py:16-37``:

    z   = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh
    i,f,o = sigmoid(z_i), sigmoid(z_f), sigmoid(z_o);  g = tanh(z_g)
    c_t = f*c_{t-1} + i*g ;  h_t = o*tanh(c_t)

and the two "last valid step" conventions:

* packed (``model.py:114-119``): row n-1, n = min(len, L).  The reverse direction's output at
  that row is its FIRST step (zero state, input x_{n-1}).
* padded (``model_cpu.py:57-62``): row p = last position of the zero-padded [L,4] input whose
  one-hot row is non-zero (L-1 if none).  The reverse direction has first consumed the
  k = L-1-p trailing zero rows, so its output is a function of (k, x_p) only.

x_t is one-hot or zero, so W_ih x_t + b_ih + b_hh is a row select from a 5-row table.
"""
import numpy as np

from . import encoders


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class NumpyOracle:
    def __init__(self, weights, dtype=np.float64):
        self.dtype = dtype
        w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
        self.H = w["rnn.weight_hh_l0"].shape[1]
        H = self.H

        def table(wih, bih, bhh):
            t = np.zeros((5, 4 * H), dtype=dtype)
            t[:4] = wih.T
            return t + (bih + bhh)[None, :]

        self.tab_f = table(w["rnn.weight_ih_l0"], w["rnn.bias_ih_l0"], w["rnn.bias_hh_l0"])
        self.tab_r = table(w["rnn.weight_ih_l0_reverse"], w["rnn.bias_ih_l0_reverse"],
                           w["rnn.bias_hh_l0_reverse"])
        self.whh_f_t = np.ascontiguousarray(w["rnn.weight_hh_l0"].T)          # [H,4H]
        self.whh_r_t = np.ascontiguousarray(w["rnn.weight_hh_l0_reverse"].T)
        self.w_out = w["out.weight"]                                           # [2,2H]
        self.b_out = w["out.bias"]

    # one LSTM cell step on a batch; z_in = table rows [B,4H]
    def _cell(self, z_in, h, c, whh_t):
        H = self.H
        z = z_in + h @ whh_t
        i = _sigmoid(z[:, :H])
        f = _sigmoid(z[:, H:2 * H])
        g = np.tanh(z[:, 2 * H:3 * H])
        o = _sigmoid(z[:, 3 * H:])
        c2 = f * c + i * g
        return o * np.tanh(c2), c2

    def reverse_lut(self, kmax):
        """h_rev[k, code] after k zero-input steps from the zero state followed by one step
        on `code` (0..3 bases, 4 = zero row).  Shape [kmax+1, 5, H]."""
        H = self.H
        out = np.zeros((kmax + 1, 5, H), dtype=self.dtype)
        h = np.zeros((1, H), dtype=self.dtype)
        c = np.zeros((1, H), dtype=self.dtype)
        for k in range(kmax + 1):
            hh, _ = self._cell(self.tab_r, np.repeat(h, 5, 0), np.repeat(c, 5, 0), self.whh_r_t)
            out[k] = hh
            h, c = self._cell(self.tab_r[4:5], h, c, self.whh_r_t)
        return out

    @staticmethod
    def plan(reads, max_len, semantics):
        """Per-read step plan shared by both conventions:
        codes [B,T] uint8 (4 beyond the read), nfwd [B] forward steps to run,
        krev [B] trailing zero steps the reverse direction consumes first, crev [B] the code
        it then sees."""
        B = len(reads)
        cs = [encoders.base_codes(r)[:max_len] for r in reads]
        lens = np.array([len(c) for c in cs], dtype=np.int64)
        T = max_len if semantics == "padded" else int(lens.max(initial=0))
        codes = np.full((B, max(T, 1)), 4, dtype=np.uint8)
        for b, c in enumerate(cs):
            codes[b, : len(c)] = c
        if semantics == "packed":
            if (lens == 0).any():
                raise RuntimeError("zero-length read cannot be packed (reference behaviour)")
            nfwd = lens.copy()
            krev = np.zeros(B, dtype=np.int64)
            crev = codes[np.arange(B), lens - 1]
        elif semantics == "padded":
            nz = codes != 4
            has = nz.any(1)
            p = np.where(has, T - 1 - np.argmax(nz[:, ::-1], axis=1), T - 1)
            nfwd = p + 1
            krev = T - 1 - p
            crev = codes[np.arange(B), p]
        else:
            raise ValueError(semantics)
        return codes, nfwd, krev, crev

    def logits(self, reads, max_len, semantics="packed", return_hidden=False):
        H = self.H
        B = len(reads)
        if B == 0:
            return np.zeros((0, 2), dtype=self.dtype)
        codes, nfwd, krev, crev = self.plan(reads, max_len, semantics)
        h = np.zeros((B, H), dtype=self.dtype)
        c = np.zeros((B, H), dtype=self.dtype)
        for t in range(int(nfwd.max())):
            act = nfwd > t
            if not act.any():
                break
            h2, c2 = self._cell(self.tab_f[codes[act, t]], h[act], c[act], self.whh_f_t)
            h[act] = h2
            c[act] = c2
        lut = self.reverse_lut(int(krev.max()))
        h_rev = lut[krev, crev]
        feat = np.concatenate([h, h_rev], axis=1)
        out = feat @ self.w_out.T + self.b_out
        return (out, feat) if return_hidden else out


def softmax2(logits):
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)
