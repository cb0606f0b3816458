#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference)
on seeded inputs, and pin the oracle restatement against it while doing so.

Runs only in the build container (the GPU box has no /root/reference); the fixtures it writes
are committed.  Usage:  python oracle/gen_golden.py [--ref /root/reference]

What is executed from the reference, unmodified:
  * ribodetector.data_loader.seq_encoder.encode_read / encode_variable_len_read
  * ribodetector.detect.unlabeled_read_collate_fn  (encode + pack_sequence)
  * ribodetector.model.model.SeqModel (pack_seq=True → forward1)   — `ribodetector` semantics
  * ribodetector.model.model_cpu.SeqModel (forward_last)           — `ribodetector_cpu` semantics
  * ribodetector.detect.Predictor.separate_paired_reads            — pair combination
`Bio.Seq` (imported by seq_encoder.py:3 for training helpers only) is absent from this image
and is stubbed; onnxruntime is absent, so detect_cpu.py itself cannot be imported — its model
is model_cpu.SeqModel, the module the .onnx was exported from (convert_onnx.py:16,28-54).
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import encoders, pairs                      # noqa: E402
from oracle.model_torch import TorchOracle              # noqa: E402
from oracle.model_numpy import NumpyOracle              # noqa: E402
from ribodetector_b200.utils import synth               # noqa: E402
from ribodetector_b200.utils.weights import load_weights  # noqa: E402


def import_reference(ref_root):
    sys.path.insert(0, ref_root)
    bio = types.ModuleType("Bio")
    bseq = types.ModuleType("Bio.Seq")
    bseq.Seq = object
    bio.Seq = bseq
    sys.modules.setdefault("Bio", bio)
    sys.modules.setdefault("Bio.Seq", bseq)
    from ribodetector.model import model as ref_model
    from ribodetector.model import model_cpu as ref_model_cpu
    from ribodetector.data_loader import seq_encoder as ref_enc
    from ribodetector import detect as ref_detect
    from ribodetector.parse_config import ConfigParser
    return ref_model, ref_model_cpu, ref_enc, ref_detect, ConfigParser


EDGE_READS_100 = None


def edge_reads(rng, L):
    """The edge cases of SURVEY.md §8c, all PROBE-verified behaviours of the reference."""
    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
    base = rnd(L)
    out = [
        base,                                   # exactly -l
        base.lower(),                           # lower case → all zero rows
        base.replace("T", "U"),                 # U ≡ T
        rnd(L + 37),                            # longer than -l → first -l bases
        rnd(L // 2),                            # shorter than -l (packed ≠ padded)
        rnd(L - 10) + "N" * 10,                 # trailing N (packed runs them, padded skips)
        "N" * 7 + rnd(L - 7),                   # leading N
        rnd(30) + "NNNN" + rnd(L - 34),         # interior N
        rnd(20) + "RYKMSWBDHVN-" + rnd(20),     # IUPAC / gap
        "N" * L,                                # all N
        "N" * 5,                                # short all N
        "A", "C", "G", "T", "N",                # length 1
        "AC" * (L // 2),
        "A" * L, "C" * L, "G" * L, "T" * L,
        rnd(L - 1), rnd(L + 1), rnd(2), rnd(39), rnd(40), rnd(41),
        rnd(60) + "n" * 3 + "acgt" + rnd(10),   # mixed case
        rnd(L - 1) + "N",                       # exactly one trailing N at position L-1
        rnd(L) + "N" * 20,                      # N's only beyond -l
        "N" + rnd(3) + "N" * (L - 4),           # mostly trailing N
    ]
    return out


def ref_logits_packed(ref_detect, model, reads, L):
    """Exactly detect.py:666-689 + model.py:32-37, batch by batch."""
    outs = []
    with torch.no_grad():
        for s in range(0, len(reads), 512):
            batch = [("@r%d" % i, r, "+", "I" * len(r)) for i, r in enumerate(reads[s:s + 512])]
            _txt, data = ref_detect.unlabeled_read_collate_fn(batch, max_len=L, pack_seq=True)
            outs.append(model(data).numpy())
    return np.concatenate(outs, 0)


def ref_logits_padded(ref_enc, model_cpu, reads, L):
    """Exactly detect_cpu.py:699-703 with model_cpu.SeqModel in place of the ORT session."""
    outs = []
    with torch.no_grad():
        for s in range(0, len(reads), 512):
            x = np.array([ref_enc.encode_variable_len_read(r, max_len=L)
                          for r in reads[s:s + 512]], dtype=np.float32)
            outs.append(model_cpu(torch.from_numpy(x)).numpy())
    return np.concatenate(outs, 0)


def pack_reads(reads):
    seq, off = encoders.flatten_reads(reads)
    return {"seq": seq, "off": off}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(1)        # deterministic summation order for the fixtures
    ref_model, ref_model_cpu, ref_enc, ref_detect, ConfigParser = import_reference(args.ref)

    cfg = ConfigParser.from_json(os.path.join(args.ref, "ribodetector", "config.json"))
    state = torch.load(os.path.join(args.ref, "ribodetector", cfg["state_file"]["mcc"]),
                       map_location="cpu")
    m_packed = cfg.init_obj("arch", ref_model)
    m_packed.load_state_dict(state["state_dict"])
    m_packed.eval()
    cpu_args = dict(cfg["arch"]["args"])
    cpu_args["pack_seq"] = False
    m_padded = ref_model_cpu.SeqModel(**cpu_args)
    m_padded.load_state_dict(state["state_dict"])
    m_padded.eval()

    weights = load_weights()
    for k, v in state["state_dict"].items():          # repo blob == shipped checkpoint
        assert np.array_equal(weights[k], v.numpy()), k
    o_t = TorchOracle(weights)
    o_n = NumpyOracle(weights, np.float64)
    os.makedirs(args.out, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(synth.SEED_BASE))

    # ---- 1. encoders -------------------------------------------------------------------
    enc_reads = edge_reads(rng, 100) + ["ACGTUNacgtuRYKM-*", "", "T" * 3]
    rows = [np.asarray(ref_enc.encode_read(r), dtype=np.float32).reshape(-1, 4) for r in enc_reads]
    padded16 = np.array([ref_enc.encode_variable_len_read(r, max_len=16) for r in enc_reads],
                        dtype=np.float32)
    padded100 = np.array([ref_enc.encode_variable_len_read(r, max_len=100) for r in enc_reads],
                         dtype=np.float32)
    for r, ref in zip(enc_reads, rows):
        assert np.array_equal(encoders.encode_read(r), ref)
    assert np.array_equal(np.stack([encoders.encode_variable_len_read(r, 16) for r in enc_reads]), padded16)
    np.savez_compressed(os.path.join(args.out, "encode.npz"), **pack_reads(enc_reads),
                        onehot_rows=np.concatenate(rows, 0), padded16=padded16, padded100=padded100)

    # ---- 2. single-end logits ------------------------------------------------------------
    report = []

    def se_case(name, reads, L):
        reads_pk = [r for r in reads if len(r) > 0]
        lp = ref_logits_packed(ref_detect, m_packed, reads_pk, L)
        ld = ref_logits_padded(ref_enc, m_padded, reads_pk, L)
        # pin the oracle
        d_t_p = np.abs(o_t.logits_packed(reads_pk, L) - lp).max()
        d_t_d = np.abs(o_t.logits_padded(reads_pk, L) - ld).max()
        d_n_p = np.abs(o_n.logits(reads_pk, L, "packed") - lp).max()
        d_n_d = np.abs(o_n.logits(reads_pk, L, "padded") - ld).max()
        report.append((name, len(reads_pk), d_t_p, d_t_d, d_n_p, d_n_d))
        assert d_t_p <= 2e-6 and d_t_d <= 2e-6, (name, d_t_p, d_t_d)
        assert d_n_p <= 5e-5 and d_n_d <= 5e-5, (name, d_n_p, d_n_d)
        np.savez_compressed(os.path.join(args.out, name + ".npz"), **pack_reads(reads_pk),
                            max_len=np.int64(L), logits_packed=lp, logits_padded=ld,
                            logits_packed_f64=o_n.logits(reads_pk, L, "packed"),
                            logits_padded_f64=o_n.logits(reads_pk, L, "padded"))

    s, o = synth.synth_reads_fixed(384, 100, synth.SEED_BASE + 1)
    fixed = synth.to_strings(s, o)
    s, o = synth.synth_reads(192, 1, 150, synth.SEED_BASE + 2, n_frac=0.01)
    ragged = synth.to_strings(s, o)
    se_case("se_L100", edge_reads(rng, 100) + fixed + ragged, 100)
    s, o = synth.synth_reads(160, 40, 300, synth.SEED_BASE + 5)
    se_case("se_L300", edge_reads(rng, 300)[:12] + synth.to_strings(s, o), 300)
    s, o = synth.synth_reads(96, 20, 200, synth.SEED_BASE + 4)
    se_case("se_L150", edge_reads(rng, 150)[:12] + synth.to_strings(s, o), 150)

    # ---- 3. paired-end: combine modes ------------------------------------------------------
    s, o = synth.synth_reads_fixed(6000, 100, synth.SEED_BASE + 3)
    pool = synth.to_strings(s, o)
    lp = ref_logits_packed(ref_detect, m_packed, pool, 100)
    lab = lp.argmax(1)
    pos = np.flatnonzero(lab == 1)
    neg = np.flatnonzero(lab == 0)
    assert len(pos) >= 64
    r1_idx = np.concatenate([pos[:48], pos[48:64], neg[:32], neg[32:160]])
    r2_idx = np.concatenate([pos[16:64], neg[200:216], pos[:32], neg[300:428]])
    r1 = [pool[i] for i in r1_idx]
    r2 = [pool[i] for i in r2_idx]
    l1 = ref_logits_packed(ref_detect, m_packed, r1, 100)
    l2 = ref_logits_packed(ref_detect, m_packed, r2, 100)
    pe = dict(max_len=np.int64(100), logits1=l1, logits2=l2)
    pe.update({"r1_" + k: v for k, v in pack_reads(r1).items()})
    pe.update({"r2_" + k: v for k, v in pack_reads(r2).items()})
    for mode in pairs.MODES:
        fake = types.SimpleNamespace(args=types.SimpleNamespace(ensure=mode))
        ids = list(range(len(r1)))
        d1, _d2 = ref_detect.Predictor.separate_paired_reads(
            fake, ids, torch.from_numpy(l1), ids, torch.from_numpy(l2))
        labels = np.full(len(r1), 99, dtype=np.int8)
        for lab_k, members in d1.items():
            labels[np.asarray(members, dtype=np.int64)] = lab_k
        assert (labels != 99).all()
        assert np.array_equal(labels, pairs.pair_labels(l1, l2, mode)), mode
        pe["labels_" + mode] = labels
    np.savez_compressed(os.path.join(args.out, "pe_L100.npz"), **pe)

    # ---- 4. argmax ties (detect.py:288 → torch.argmax first-max) ---------------------------
    ties = np.array([[0.5, 0.5], [1.0, -1.0], [-1.0, 1.0], [0.0, 0.0], [-0.0, 0.0],
                     [3.25, 3.25], [1e-8, 0.0]], dtype=np.float32)
    tl = torch.argmax(torch.from_numpy(ties), dim=1).numpy().astype(np.int8)
    assert np.array_equal(tl, pairs.argmax_labels(ties))
    np.savez_compressed(os.path.join(args.out, "ties.npz"), logits=ties, labels=tl)

    print("%-10s %6s %10s %10s %10s %10s" % ("case", "n", "torch/pk", "torch/pad", "np64/pk", "np64/pad"))
    for r in report:
        print("%-10s %6d %10.2e %10.2e %10.2e %10.2e" % r)
    print("pair label histogram:", {m: np.bincount(pe["labels_" + m] + 1, minlength=3).tolist()
                                    for m in pairs.MODES})


if __name__ == "__main__":
    main()
