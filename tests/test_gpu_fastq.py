"""GPU: the device-side edges of the path — K0 record scan (rd_scan_fastq_device), rd_classify_records, K4
label partition (rd_partition_records_device) and the streaming form rd_fastq_submit / rd_fastq_collect —
against (a) the golden records produced by the REFERENCE's own parser (tests/golden/fastx.json), (b) the host
scanner / writer of the same library on seeded random text, bit-exact (byte and index work)."""
import ctypes
import gzip
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from ribodetector_b200 import _lib
from ribodetector_b200.data_loader import FastxReader, partition_records

pytestmark = pytest.mark.gpu


def _records_from_index(text, rec):
    rec = rec.cpu().numpy()
    t = bytes(text)
    return [tuple(t[r[2 * k]:r[2 * k + 1]].decode("latin-1") for k in range(4)) for r in rec]


def _host_scan(text, final=True, max_records=None):
    """rd_scan_fastx on the same bytes → (n, consumed, hdr, plus, qual, seq, seq_off)."""
    lib = _lib.load_library()
    buf = np.frombuffer(bytes(text), np.uint8).copy() if len(text) else np.zeros(1, np.uint8)
    cap = max_records if max_records is not None else len(text) // 8 + 1
    hdr, plus, qual = (np.empty(2 * cap + 2, np.int64) for _ in range(3))
    seq = np.empty(len(text) + 1, np.uint8)
    seq_off = np.empty(cap + 2, np.int64)
    consumed = ctypes.c_int64(0)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)      # noqa: E731
    n = lib.rd_scan_fastx(p(buf), len(text), 0, int(final), cap, p(hdr), p(plus), p(qual), p(seq), len(text), p(seq_off),
                          ctypes.byref(consumed), 3)
    return n, consumed.value, hdr, plus, qual, seq, seq_off


def _random_fastq(n, seed, crlf_frac=0.0, space_frac=0.0, min_len=1, max_len=150, final_newline=True):
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        s = "".join(rng.choice(list("ACGTNacgu"), size=L, p=[.22, .22, .22, .22, .04, .02, .02, .02, .02]))
        eol = "\r\n" if rng.random() < crlf_frac else "\n"
        pad = " \t"[: int(rng.integers(0, 3))] if rng.random() < space_frac else ""
        q = "".join(chr(int(c)) for c in rng.integers(33, 74, size=L))      # may start with '@'
        parts.append("@read%d extra/%d%s%s%s%s%s+%s%s%s%s" % (i, i % 7, pad, eol, s, pad, eol, "" if i % 3 else "read%d" % i, eol, q, eol))
    text = "".join(parts)
    if not final_newline:
        text = text.rstrip("\r\n")
    return text.encode("latin-1")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "fastx.json")) as f:
        return json.load(f)


def test_device_scan_matches_reference_parser(gpu_model, golden):
    seen = 0
    for name, case in golden["cases"].items():
        if case["type"] != "fastq":
            continue
        text = case["text"].encode("latin-1")
        _, rec, n, consumed = gpu_model.scan_fastq(text, final_chunk=True)
        assert _records_from_index(text, rec) == [tuple(r) for r in case["records"]], name
        hn, hc = _host_scan(text)[:2]
        assert (n, consumed) == (hn, hc), name
        seen += 1
    assert seen >= 5


@pytest.mark.parametrize("variant", ["plain", "crlf_spaces", "no_final_newline", "tiny_reads"])
def test_device_scan_equals_host_scanner(gpu_model, variant):
    kw = {"plain": {}, "crlf_spaces": dict(crlf_frac=0.3, space_frac=0.3), "no_final_newline": dict(final_newline=False),
          "tiny_reads": dict(max_len=3)}[variant]
    text = _random_fastq(60000, 7, **kw)                       # several MB: hundreds of 16-KB scan tiles
    for final, cut, cap in ((True, len(text), None), (False, len(text) * 2 // 3 + 5, None), (True, len(text), 12345),
                            (False, 70001, None), (False, 17, None), (True, 0, None)):
        t = text[:cut]
        _, rec, n, consumed = gpu_model.scan_fastq(t, final_chunk=final, max_records=cap)
        hn, hc, hdr, plus, qual, seq, seq_off = _host_scan(t, final, cap)
        assert (n, consumed) == (hn, hc), (variant, final, cut, cap)
        r = rec.cpu().numpy()
        assert np.array_equal(r[:, 0:2].ravel(), hdr[:2 * n]) and np.array_equal(r[:, 4:6].ravel(), plus[:2 * n])
        assert np.array_equal(r[:, 6:8].ravel(), qual[:2 * n])
        assert np.array_equal(r[:, 3] - r[:, 2], np.diff(seq_off[:n + 1]))
        tb = np.frombuffer(t, np.uint8)
        got_seq = np.concatenate([tb[a:b] for a, b in r[:200, 2:4]]) if n else np.zeros(0, np.uint8)
        assert np.array_equal(got_seq, seq[:got_seq.size])


def test_device_scan_reports_the_first_malformed_record(gpu_model):
    good = _random_fastq(5000, 3).decode()
    recs = good.split("@read")[1:]
    bad_blank = "@read" + "@read".join(recs[:3000]) + "\n" + "@read" + "@read".join(recs[3000:])
    with pytest.raises(ValueError, match="record 3000"):
        gpu_model.scan_fastq(bad_blank.encode())
    bad_hdr = good.replace("@read1234 ", "read1234 ")
    with pytest.raises(ValueError, match="record 1234"):
        gpu_model.scan_fastq(bad_hdr.encode())


def test_device_partition_equals_host_writer(gpu_model, tmp_path):
    text = _random_fastq(50000, 11, crlf_frac=0.2, space_frac=0.2)
    d_text, rec, n, _ = gpu_model.scan_fastq(text)
    rng = np.random.default_rng(5)
    for probs in ([0.7, 0.25, 0.05], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]):
        labels = rng.choice(np.array([0, 1, -1], np.int8), size=n, p=probs)
        out, sizes = gpu_model.partition_records(d_text, rec, labels)
        p = tmp_path / "x.fq"
        p.write_bytes(text)
        ch = next(iter(FastxReader(str(p), max_records=n + 1)))
        assert ch.n == n
        outs, hsizes = partition_records(ch, labels, (True, True, True), 3)
        want = b"".join(b"" if o is None else o.tobytes() for o in outs)
        assert np.array_equal(sizes.numpy(), hsizes)
        assert out.cpu().numpy().tobytes() == want
    out0, sizes0 = gpu_model.partition_records(d_text, rec[:0], np.zeros(0, np.int8))
    assert out0.numel() == 0 and int(sizes0.sum()) == 0


def test_classify_records_is_bitwise_classify(gpu_model):
    """The record-index addressing (stride 8 into the FASTQ text) runs the same kernels as off[n+1]."""
    text = _random_fastq(20000, 13, min_len=30, max_len=140, crlf_frac=0.1)
    d_text, rec, n, _ = gpu_model.scan_fastq(text)
    hn, _, _, _, _, seq, seq_off = _host_scan(text)
    assert hn == n
    for prec in ("tc_exact", "tc_fast", "fp32", "tc_auto"):
        for sem in ("packed", "padded"):
            la, lab_a = gpu_model.classify_records(d_text, rec, 100, semantics=sem, precision=prec)
            lb, _, lab_b = gpu_model.classify(seq[:seq_off[n]].copy(), seq_off[:n + 1].copy(), 100, semantics=sem, precision=prec)
            assert torch.equal(la, lb) and torch.equal(lab_a, lab_b), (prec, sem)


def _write_pairs(tmp_path, n, gz_first=False):
    t1 = _random_fastq(n, 21, min_len=40, max_len=120)
    t2 = _random_fastq(n, 22, min_len=20, max_len=150, crlf_frac=0.1)          # different record sizes per end
    f1 = tmp_path / ("r1.fq.gz" if gz_first else "r1.fq")
    f2 = tmp_path / "r2.fq"
    if gz_first:
        with gzip.open(f1, "wb") as f:
            f.write(t1)
    else:
        f1.write_bytes(t1)
    f2.write_bytes(t2)
    return f1, f2


def _run_cli(args):
    from ribodetector_b200 import detect
    return detect.main(args)


def test_streaming_single_end_small_blocks_equal_host_path(gpu_model, tmp_path):
    """Blocks far smaller than the file (carry-over of cut records, many submit/collect rounds, slot reuse)."""
    from ribodetector_b200.data_loader.fastq_gpu import FastqGpuStream
    text = _random_fastq(30000, 17, min_len=30, max_len=140, final_newline=False)
    inp = tmp_path / "in.fq"
    inp.write_bytes(text)
    want = _run_cli(["-l", "100", "-i", str(inp), "-o", str(tmp_path / "h_non.fq"), "-r", str(tmp_path / "h_rrna.fq"), "--host_ingest"])
    for block in (1 << 16, 100003, 1 << 24):
        with open(tmp_path / "d_non.fq", "wb") as fn, open(tmp_path / "d_rrna.fq", "wb") as fr:
            st = FastqGpuStream([gpu_model], [str(inp)], 100, block_bytes=block, threads=2)
            counts = st.run({"non": [fn], "rrna": [fr], "unc": None})
        assert (tmp_path / "d_non.fq").read_bytes() == (tmp_path / "h_non.fq").read_bytes(), block
        assert (tmp_path / "d_rrna.fq").read_bytes() == (tmp_path / "h_rrna.fq").read_bytes(), block
        assert (st.num_seqs, int(counts[0]), int(counts[1])) == (want.num_seqs, want.num_nonrrna, want.num_rrna) and want.num_seqs == 30000


@pytest.mark.parametrize("mode", ["none", "rrna", "norrna", "both"])
def test_streaming_pairs_equal_host_path(gpu_model, tmp_path, mode):
    from ribodetector_b200.data_loader.fastq_gpu import FastqGpuStream
    f1, f2 = _write_pairs(tmp_path, 20000, gz_first=(mode == "both"))
    base = ["-l", "100", "-i", str(f1), str(f2), "-e", mode]
    want = _run_cli(base + ["-o", str(tmp_path / "h1.fq"), str(tmp_path / "h2.fq"), "-r", str(tmp_path / "hr1.fq"),
                            str(tmp_path / "hr2.fq"), "--host_ingest"])
    names = ["d1.fq", "d2.fq", "dr1.fq", "dr2.fq", "du1.fq", "du2.fq"]
    fhs = [open(tmp_path / x, "wb") for x in names]
    st = FastqGpuStream([gpu_model], [str(f1), str(f2)], 100, mode=mode, block_bytes=1 << 17, threads=2)
    counts = st.run({"non": fhs[0:2], "rrna": fhs[2:4], "unc": fhs[4:6]})
    for fh in fhs:
        fh.close()
    for d, h in (("d1.fq", "h1.fq"), ("d2.fq", "h2.fq"), ("dr1.fq", "hr1.fq"), ("dr2.fq", "hr2.fq")):
        assert (tmp_path / d).read_bytes() == (tmp_path / h).read_bytes(), (mode, d)
    if mode == "both":
        for d, h in (("du1.fq", "h1.fq.unclassified.gz"), ("du2.fq", "h2.fq.unclassified.gz")):
            assert (tmp_path / d).read_bytes() == gzip.open(tmp_path / h).read()
        assert int(counts[2]) == want.num_unknown > 0
    assert (st.num_seqs, int(counts[0]), int(counts[1])) == (20000, want.num_nonrrna, want.num_rrna)


def test_streaming_rejects_unequal_pair_files_and_bad_text(gpu_model, tmp_path):
    from ribodetector_b200.data_loader.fastq_gpu import FastqGpuStream
    (tmp_path / "a.fq").write_bytes(_random_fastq(3000, 1))
    (tmp_path / "b.fq").write_bytes(_random_fastq(2990, 2))
    sinks = {"non": [open(os.devnull, "wb"), open(os.devnull, "wb")], "rrna": None, "unc": None}
    for block in (1 << 15, 1 << 24):
        with pytest.raises(RuntimeError, match="different numbers"):
            FastqGpuStream([gpu_model], [str(tmp_path / "a.fq"), str(tmp_path / "b.fq")], 100, block_bytes=block).run(sinks)
    (tmp_path / "bad.fq").write_bytes(_random_fastq(3000, 1).replace(b"@read2000 ", b"\n@read2000 "))
    with pytest.raises(ValueError, match="blank line"):
        FastqGpuStream([gpu_model], [str(tmp_path / "bad.fq")], 100, block_bytes=1 << 15).run(sinks)
    # the model still works after the failed runs (no slot left pending)
    (tmp_path / "ok.fq").write_bytes(_random_fastq(100, 4))
    st = FastqGpuStream([gpu_model], [str(tmp_path / "ok.fq")], 100)
    st.run(sinks)
    assert st.num_seqs == 100


def test_cli_default_is_device_ingest_and_equals_host_ingest(gpu_model, tmp_path):
    text = _random_fastq(8000, 29, min_len=50, max_len=120)
    inp = tmp_path / "in.fq.gz"
    with gzip.open(inp, "wb") as f:
        f.write(text)
    a = _run_cli(["-l", "100", "-i", str(inp), "-o", str(tmp_path / "a.fq"), "-r", str(tmp_path / "ar.fq.gz")])
    b = _run_cli(["-l", "100", "-i", str(inp), "-o", str(tmp_path / "b.fq"), "-r", str(tmp_path / "br.fq.gz"), "--host_ingest"])
    assert set(a.stage_seconds) == {"read", "submit", "collect", "write"} and set(b.stage_seconds) == {"read", "classify", "write"}
    assert (tmp_path / "a.fq").read_bytes() == (tmp_path / "b.fq").read_bytes()
    assert gzip.open(tmp_path / "ar.fq.gz").read() == gzip.open(tmp_path / "br.fq.gz").read()
    assert (a.num_seqs, a.num_rrna) == (b.num_seqs, b.num_rrna) == (8000, b.num_rrna)


def test_streaming_round_robin_over_two_handles(gpu_model, weights, tmp_path):
    """Blocks alternate between handles (as they do between GPUs): 4 units in flight, output still in file order."""
    from ribodetector_b200.data_loader.fastq_gpu import FastqGpuStream
    from ribodetector_b200.model import SeqModel
    second = SeqModel(4, 128, 1, 2, pack_seq=True)
    second.load_state_dict(weights)
    second.to("cuda:0").eval()
    try:
        f1, f2 = _write_pairs(tmp_path, 12000)
        outs = {}
        for tag, models in (("one", [gpu_model]), ("two", [gpu_model, second])):
            names = [tmp_path / ("%s_%d.fq" % (tag, i)) for i in range(4)]
            fhs = [open(x, "wb") for x in names]
            st = FastqGpuStream(models, [str(f1), str(f2)], 100, mode="rrna", block_bytes=1 << 17, threads=2)
            counts = st.run({"non": fhs[0:2], "rrna": fhs[2:4], "unc": None})
            for fh in fhs:
                fh.close()
            outs[tag] = ([x.read_bytes() for x in names], counts.tolist(), st.num_seqs)
        assert outs["one"] == outs["two"] and outs["one"][2] == 12000
    finally:
        second.close()
