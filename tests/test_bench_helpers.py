"""CPU: the pieces of bench.py that do not need a GPU — the parity block arithmetic, the source-hash guard of
roofline.traffic, the FLOP accounting of SURVEY.md 8d, and the reference-code CPU arm on a small sample."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flop_accounting_matches_survey():
    assert bench.flop_per_read(100) == 13108224.0           # 131 072 n + 1 024
    assert bench.flop_per_read(150) == 19661824.0
    assert set(bench.TOL) == set(bench.DTYPE) == set(bench.EXECUTED_PER_ALGORITHMIC) == set(bench.MUFU_PER_UNIT_STEP)


def test_parity_block_counts_flips_and_band():
    ref = np.array([[0.0, 1.0], [1.0, 0.0], [0.0, 1e-4], [2.0, -2.0]])
    got = ref.copy()
    got[2] = [1e-4, 0.0]                                     # a flip inside the 4e-4 band
    got[0] += 5e-4
    d = bench.parity_block(got, ref, "tc_mixed", vs="x")
    assert d["n"] == 4 and d["flips"] == 1 and d["flips_outside_band"] == 0 and d["reads_inside_band"] == 1
    assert abs(d["max_dlogit"] - 5e-4) < 1e-12 and d["ok"]
    got[3] = [-2.0, 2.0]                                     # a flip far outside the band and a huge error
    d = bench.parity_block(got, ref, "tc_mixed")
    assert d["flips_outside_band"] == 1 and not d["ok"]


def test_traffic_is_withheld_when_the_kernel_source_changed(tmp_path, monkeypatch):
    src = os.path.join(ROOT, "ribodetector_b200", "csrc", "rd_lstm_tc.cu")
    sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
    prof = tmp_path / "profiles"
    prof.mkdir()
    (tmp_path / "ribodetector_b200" / "csrc").mkdir(parents=True)
    (tmp_path / "ribodetector_b200" / "csrc" / "rd_lstm_tc.cu").write_bytes(open(src, "rb").read())
    entry = {"bytes_per_read": 126.8, "reads": 4194304, "report": "r.ncu-rep", "src_sha": sha}
    (prof / "k2_traffic.json").write_text(json.dumps({"tc_mixed_L100": entry}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    v, note = bench.k2_traffic("tc_mixed", 100)
    assert v == 126.8 and "r.ncu-rep" in note
    assert bench.k2_traffic("tc_fast", 100)[0] is None       # no capture for that precision
    (tmp_path / "ribodetector_b200" / "csrc" / "rd_lstm_tc.cu").write_bytes(b"// edited\n")
    v, note = bench.k2_traffic("tc_mixed", 100)
    assert v is None and "changed" in note


def test_committed_traffic_record_is_well_formed():
    with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
        d = json.load(f)
    assert "tc_mixed_L100" in d
    for e in d.values():
        assert e["bytes_per_read"] > 100 and len(e["src_sha"]) == 64 and e["reads"] > 0


def test_reference_code_cpu_arm_small_sample():
    """The CPU arm on the reference's own encoder + model (when importable) equals the oracle port bit for bit."""
    from oracle import ref_cpu_arm, cpu_pipeline
    from ribodetector_b200.utils import synth
    from ribodetector_b200.utils.weights import load_weights
    if ref_cpu_arm.find_reference() is None:
        pytest.skip("reference package not importable here")
    seq, off = synth.synth_reads(2048, 30, 130, 77)
    lab, logits, _ = ref_cpu_arm.classify(seq, off, 100, threads=2)
    lab2, logits2, _ = cpu_pipeline.classify(seq, off, 100, load_weights(), threads=2)
    assert np.array_equal(lab, lab2) and np.array_equal(logits, logits2)
    info, lg, _ = bench.cpu_arm(load_weights(), 2, seq, off, split=True)
    assert info["kind"] == "reference" and info["cores"] == 2 and np.array_equal(lg, logits)
    assert info["encode_only_reads_per_s_one_core"] > info["model_only_reads_per_s_one_core"] > 0
