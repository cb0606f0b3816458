#!/usr/bin/env python
"""CPU emulation (numpy, fp64 arithmetic around quantised OPERANDS) of the operand-precision schemes the tensor-core LSTM
kernels were chosen from — the numbers DESIGN.md §10/§11 quote, reproducible without a GPU:

    python tests/prec_emulate.py [n_reads] [read_len]      -> table of max / p99.9 |dlogit| and |dp| vs the fp64 recurrence
    RD_EMU_GATES=1 adds the e5m2 scheme with the correction products restricted to the columns of some gates

Per step the gate pre-activations are  z = tab[code] + W_hh . h  with the product evaluated from ROUNDED operands as the
scheme prescribes; activations, cell update and the FC are exact (fp64), so the table isolates operand rounding (what the
MMA passes decide) from the activation approximations (tanh.approx is 2^-11, ex2/rcp a few ulp) and from the fp32
accumulation order of the tensor core (second order; rd_lstm_tc.cu issues the small correction products first).

Schemes (W = W_hh, k = 128):
  fp16x1      W16 . h16                                   one fp16 pass            = tc_fast's operands,  9 MMAs per chunk
  split3      W_hi.h_hi + W_lo.h_hi + W_hi.h_lo  (fp16)   three fp16 passes        = tc_exact,           25 MMAs
  w_only      W_hi.h_hi + W_lo.h_hi                        weight residual only                           17 MMAs (fp16)
  h_only      W_hi.h_hi + W_hi.h_lo                        state residual only                            17 MMAs (fp16)
  e5m2        W_hi.h_hi + e5m2(W_lo).e5m2(h_hi) + e5m2(W_hi).e5m2(h_lo)   kind::f8f6f4, K = 32 = tc_mixed_raw, 17 MMAs
  e4m3        the same with e4m3 operands (ideal scaling; on the chip it needs a 2^16 rescale between the passes)
  mxf4        the same with e2m1 operands and one power-of-two scale per 32 elements along k   (kind::mxf4, K = 64: 13 MMAs)
  mxf4_w      e2m1 for the W_lo.h_hi half only, e5m2 for W_hi.h_lo                                         15 MMAs
Test infrastructure: imports oracle/ (reverse-direction LUT and the exact recurrence), never imported by the product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.model_numpy import NumpyOracle, softmax2     # noqa: E402
from ribodetector_b200.utils import synth               # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402


def quant(x, mbits, emin, emax_val):
    """Round-to-nearest-even onto a binary float grid: `mbits` stored mantissa bits, smallest normal exponent `emin`
    (subnormals below), saturating at +-emax_val."""
    x = np.asarray(x, np.float64)
    a = np.abs(x)
    e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    e = np.maximum(e, emin)
    ulp = np.exp2(e - mbits)
    q = np.rint(a / ulp) * ulp
    return np.sign(x) * np.minimum(q, emax_val)


def fp16(x):
    return quant(x, 10, -14, 65504.0)


def e5m2(x):
    return quant(x, 2, -14, 57344.0)


def e4m3(x):
    return quant(x, 3, -6, 448.0)


def e2m1_block(x, axis, block=32):
    """mxf4: e2m1 elements {0, .5, 1, 1.5, 2, 3, 4, 6} times one power-of-two scale per `block` elements along `axis`,
    the scale chosen so that the block's largest magnitude lands in [4, 8) -> representable up to 6 (saturating)."""
    x = np.moveaxis(np.asarray(x, np.float64), axis, -1)
    shp = x.shape
    xb = x.reshape(shp[:-1] + (shp[-1] // block, block))
    amax = np.abs(xb).max(-1, keepdims=True)
    scale = np.exp2(np.floor(np.log2(np.where(amax > 0, amax, 1.0))) - 2.0)
    q = quant(xb / scale, 1, 0, 6.0) * scale
    return np.moveaxis(q.reshape(shp), -1, axis)


def run(scheme, tab, whh_t, codes, nsteps, gates="ifgo"):
    """h after nsteps for every read; whh_t = W_hh^T [k, 4H].  `gates`: the gate columns that receive the correction
    products (e5m2 scheme only) — would a correction pass over fewer than the 512 gate columns do?"""
    n = codes.shape[0]
    H = whh_t.shape[0]
    gmask = np.zeros(4 * H)
    for ch in gates:
        gmask["ifgo".index(ch) * H:("ifgo".index(ch) + 1) * H] = 1.0
    h = np.zeros((n, H))
    c = np.zeros((n, H))
    W_hi = fp16(whh_t)
    W_lo = whh_t - W_hi
    if scheme in ("split3", "w_only"):
        W_lo_q = fp16(W_lo)
    elif scheme in ("e5m2", "mxf4_w"):
        W_lo_q, W_hi_q = e5m2(W_lo) * gmask[None, :], e5m2(W_hi) * gmask[None, :]
    elif scheme == "e4m3":
        W_lo_q, W_hi_q = e4m3(W_lo * 2.0 ** 12) * 2.0 ** -12, e4m3(W_hi)
    if scheme in ("mxf4", "mxf4_w"):
        W_lo_4 = e2m1_block(W_lo, 0)
        W_hi_4 = e2m1_block(W_hi, 0)
    for t in range(nsteps):
        if scheme == "exact":
            z = h @ whh_t
        else:
            h_hi = fp16(h)
            h_lo = h - h_hi
            z = h_hi @ W_hi
            if scheme == "split3":
                z += h_hi @ W_lo_q + fp16(h_lo) @ W_hi
            elif scheme == "w_only":
                z += h_hi @ W_lo_q
            elif scheme == "h_only":
                z += fp16(h_lo) @ W_hi
            elif scheme == "e5m2":
                z += e5m2(h_hi) @ W_lo_q + (e5m2(h_lo * 2.0 ** 14) * 2.0 ** -14) @ W_hi_q
            elif scheme == "e4m3":
                z += e4m3(h_hi) @ W_lo_q + (e4m3(h_lo * 2.0 ** 14) * 2.0 ** -14) @ W_hi_q
            elif scheme == "mxf4":
                z += e2m1_block(h_hi, 1) @ W_lo_4 + e2m1_block(h_lo, 1) @ W_hi_4
            elif scheme == "mxf4_w":
                z += e2m1_block(h_hi, 1) @ W_lo_4 + (e5m2(h_lo * 2.0 ** 14) * 2.0 ** -14) @ W_hi_q
        z = z + tab[codes[:, t]]
        i = 1.0 / (1.0 + np.exp(-z[:, :H]))
        f = 1.0 / (1.0 + np.exp(-z[:, H:2 * H]))
        g = np.tanh(z[:, 2 * H:3 * H])
        o = 1.0 / (1.0 + np.exp(-z[:, 3 * H:]))
        c = f * c + i * g
        h = o * np.tanh(c)
    return h


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    w = load_weights()
    orc = NumpyOracle(w, np.float64)
    seq, off = synth.synth_reads_fixed(n, L, synth.SEED_BASE + 40, n_frac=0.001)
    reads = synth.to_strings(seq, off)
    codes, nfwd, krev, crev = orc.plan(reads, L, "packed")
    lut = orc.reverse_lut(0)[krev, crev]
    w_out, b_out = orc.w_out, orc.b_out
    H = orc.H

    def logits(h):
        return np.concatenate([h, lut], 1) @ w_out.T + b_out

    ref = logits(run("exact", orc.tab_f, orc.whh_f_t, codes, L))
    assert np.abs(ref - orc.logits(reads, L, "packed")).max() < 1e-9
    pref = softmax2(ref)
    print("%d reads x %d bp, hidden %d; error of the final logits / probabilities against the fp64 recurrence" % (n, L, H))
    print("%-8s %5s %12s %12s %12s" % ("scheme", "MMAs", "max|dlogit|", "p99.9", "max|dp|"))
    for scheme, mmas in (("fp16x1", 9), ("w_only", 17), ("h_only", 17), ("mxf4", 13), ("mxf4_w", 15), ("e5m2", 17), ("e4m3", 17),
                         ("split3", 25)):
        got = logits(run(scheme, orc.tab_f, orc.whh_f_t, codes, L))
        d = np.abs(got - ref).max(1)
        print("%-8s %5d %12.2e %12.2e %12.2e" % (scheme, mmas, d.max(), np.quantile(d, 0.999), np.abs(softmax2(got) - pref).max()),
              flush=True)
    if os.environ.get("RD_EMU_GATES"):          # e5m2 corrections over a subset of the gate columns only
        print("e5m2 corrections applied to the columns of some gates only (all four: the e5m2 row above)")
        for gates in ("fgo", "igo", "ifo", "ifg", "fg", "go", "if", "g", "f"):
            got = logits(run("e5m2", orc.tab_f, orc.whh_f_t, codes, L, gates))
            d = np.abs(got - ref).max(1)
            print("%-8s %5s %12.2e %12.2e %12.2e" % (gates, "", d.max(), np.quantile(d, 0.999), np.abs(softmax2(got) - pref).max()),
                  flush=True)


if __name__ == "__main__":
    main()
