"""CPU: the oracle restatement against the golden vectors the REAL reference produced
(oracle/gen_golden.py), plus the semantic edge cases of SURVEY.md §8c."""
import numpy as np
import pytest

from conftest import load_golden, split_reads
from oracle import encoders, pairs
from oracle.model_numpy import NumpyOracle, softmax2


def test_encoders_match_reference_golden():
    g = load_golden("encode")
    reads = split_reads(g["seq"], g["off"])
    rows = np.concatenate([encoders.encode_read(r) for r in reads], 0)
    assert np.array_equal(rows, g["onehot_rows"])
    for L, key in ((16, "padded16"), (100, "padded100")):
        got = np.stack([encoders.encode_variable_len_read(r, L) for r in reads])
        assert np.array_equal(got, g[key])


def test_encoder_table():
    x = encoders.encode_read("ACGTUNacgtRY-")
    assert x[:5].tolist() == [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0, 0, 0, 1]]
    assert not x[5:].any()                      # N, lower case, IUPAC, gap → zero rows
    assert encoders.encode_variable_len_read("ACGT", 2).shape == (2, 4)   # first max_len bases
    assert not encoders.encode_variable_len_read("AC", 5)[2:].any()


@pytest.mark.parametrize("case", ["se_L100", "se_L150", "se_L300"])
def test_torch_oracle_bit_matches_reference(case, torch_oracle):
    g = load_golden(case)
    reads = split_reads(g["seq"], g["off"])
    L = int(g["max_len"])
    import torch
    torch.set_num_threads(1)
    assert np.abs(torch_oracle.logits_packed(reads, L) - g["logits_packed"]).max() <= 2e-6
    assert np.abs(torch_oracle.logits_padded(reads, L) - g["logits_padded"]).max() <= 2e-6


@pytest.mark.parametrize("case", ["se_L100", "se_L150", "se_L300"])
def test_numpy_oracle_matches_reference(case, numpy_oracle, weights):
    g = load_golden(case)
    reads = split_reads(g["seq"], g["off"])
    L = int(g["max_len"])
    for sem in ("packed", "padded"):
        ref = g["logits_" + sem]
        got = numpy_oracle.logits(reads, L, sem)
        assert np.abs(got - ref).max() <= 5e-5
        assert np.abs(got - g["logits_%s_f64" % sem]).max() <= 1e-12
        margin = np.abs(ref[:, 1] - ref[:, 0])
        same = got.argmax(1) == ref.argmax(1)
        assert same[margin > 1e-4].all()
    got32 = NumpyOracle(weights, np.float32).logits(reads, L, "packed")
    assert np.abs(got32 - g["logits_packed"]).max() <= 2e-4


@pytest.mark.parametrize("H", [32, 64, 96, 192, 256])
def test_oracles_match_reference_golden_at_other_hidden_sizes(H):
    """The reference's own SeqModel / model_cpu.SeqModel instantiated at another hidden_size with seeded weights
    (oracle/gen_golden_arch.py → tests/golden/arch.npz): both restatements are pinned at every size the kernels take."""
    import torch
    from oracle.model_torch import TorchOracle
    from ribodetector_b200.utils import synth
    torch.set_num_threads(1)
    g = load_golden("arch")
    assert H in g["hidden_sizes"].tolist()
    w = synth.synth_weights(H, int(g["weight_seed"]))
    reads = split_reads(g["seq"], g["off"])
    L = int(g["max_len"])
    o_t, o_n = TorchOracle(w, hidden_size=H), NumpyOracle(w, np.float64)
    for sem in ("packed", "padded"):
        ref = g["logits_%s_h%d" % (sem, H)]
        assert np.abs(o_t.logits(reads, L, sem) - ref).max() <= 2e-6
        assert np.abs(o_n.logits(reads, L, sem) - ref).max() <= 5e-5
    assert np.abs(g["logits_packed_h%d" % H][:, 1] - g["logits_packed_h%d" % H][:, 0]).std() > 0.05     # not a degenerate model


def test_packed_vs_padded_semantics_differ_on_short_reads(numpy_oracle):
    g = load_golden("se_L100")
    off = g["off"]
    lens = off[1:] - off[:-1]
    d = np.abs(g["logits_packed"] - g["logits_padded"]).max(1)
    reads = split_reads(g["seq"], off)
    clean_full = np.array([l >= 100 and set(r[:100]) <= set("ACGTU") for l, r in zip(lens, reads)])
    assert d[clean_full].max() < 1e-4            # same answer on full-length N-free reads
    assert d[lens < 60].max() > 1e-2             # and a different one on short reads


def test_empty_read_rejected_under_packed_semantics(numpy_oracle, torch_oracle):
    with pytest.raises(RuntimeError):
        numpy_oracle.logits(["ACGT", ""], 100, "packed")
    with pytest.raises(RuntimeError):
        torch_oracle.logits_packed(["ACGT", ""], 100)
    out = numpy_oracle.logits(["", "NNN"], 100, "padded")     # padded: all-zero rows, index T-1
    assert np.allclose(out[0], out[1])
    g = load_golden("se_L100")                                 # reads 9, 10 are all-N reads
    reads = split_reads(g["seq"], g["off"])
    assert set(reads[9]) == {"N"} and set(reads[10]) == {"N"}
    assert np.abs(out[0] - g["logits_padded"][9]).max() < 5e-5
    assert np.abs(out[0] - g["logits_padded"][10]).max() < 5e-5


def test_pair_modes_match_reference_golden():
    g = load_golden("pe_L100")
    for mode in pairs.MODES:
        assert np.array_equal(pairs.pair_labels(g["logits1"], g["logits2"], mode), g["labels_" + mode])
    hist = {m: np.bincount(g["labels_" + m] + 1, minlength=3) for m in pairs.MODES}
    assert hist["both"][0] > 0 and hist["rrna"][2] > 0 and hist["norrna"][2] > hist["rrna"][2]
    assert np.array_equal(pairs.counts(g["labels_both"]), hist["both"][[1, 2, 0]])


def test_argmax_ties_go_to_class0():
    g = load_golden("ties")
    assert np.array_equal(pairs.argmax_labels(g["logits"]), g["labels"])


def test_softmax_rows_sum_to_one():
    g = load_golden("se_L100")
    p = softmax2(g["logits_packed"].astype(np.float64))
    assert np.allclose(p.sum(1), 1.0)


def test_plan_matches_definition():
    codes, nfwd, krev, crev = NumpyOracle.plan(["ACGTNN", "NNNN", "ACGTACGTAC"], 8, "padded")
    assert nfwd.tolist() == [4, 8, 8] and krev.tolist() == [4, 0, 0] and crev.tolist() == [3, 4, 3]
    codes, nfwd, krev, crev = NumpyOracle.plan(["ACGTNN", "ACGTACGTAC"], 8, "packed")
    assert nfwd.tolist() == [6, 8] and krev.tolist() == [0, 0] and crev.tolist() == [4, 3]
