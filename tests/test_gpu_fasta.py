"""GPU: the device-side FASTA path — rd_scan_fasta_device (K0 for seq_parser's FASTA branch, fastx_parser.py:39-55),
rd_classify_records over its index, rd_partition_fasta_device and the streaming form rd_fasta_submit — against (a) the
golden records produced by the REFERENCE's own parser (tests/golden/fastx.json), (b) the host scanner / writer of the
same library (rd_scan_fastx / rd_partition_records, themselves pinned to the reference parser in tests/test_fastx.py)
on seeded random text, bit-exact (byte and index work)."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from ribodetector_b200 import _lib
from ribodetector_b200.data_loader import FastxReader, partition_records

pytestmark = pytest.mark.gpu


def _records_from_index(d_buf, rec):
    b = d_buf.cpu().numpy().tobytes()
    return [(b[r[0]:r[1]].decode("latin-1"), b[r[2]:r[3]].decode("latin-1")) for r in rec.cpu().numpy()]


def _host_scan(text, final=True, max_records=None):
    lib = _lib.load_library()
    buf = np.frombuffer(bytes(text), np.uint8).copy() if len(text) else np.zeros(1, np.uint8)
    cap = max_records if max_records is not None else len(text) // 2 + 1
    hdr = np.empty(2 * cap + 2, np.int64)
    seq = np.empty(len(text) + 1, np.uint8)
    seq_off = np.empty(cap + 2, np.int64)
    consumed = ctypes.c_int64(0)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)      # noqa: E731
    n = lib.rd_scan_fastx(p(buf), len(text), 1, int(final), cap, p(hdr), None, None, p(seq), len(text), p(seq_off),
                          ctypes.byref(consumed), 3)
    tb = bytes(text)
    recs = [(tb[hdr[2 * i]:hdr[2 * i + 1]].decode("latin-1"), seq[seq_off[i]:seq_off[i + 1]].tobytes().decode("latin-1"))
            for i in range(max(n, 0))]
    return n, consumed.value, recs


def _random_fasta(n, seed, wrap=(0, 60, 70), crlf_frac=0.0, space_frac=0.0, blank_frac=0.0, empty_frac=0.0,
                  final_newline=True, min_len=1, max_len=300):
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(n):
        L = 0 if rng.random() < empty_frac else int(rng.integers(min_len, max_len + 1))
        s = "".join(rng.choice(list("ACGTNacgtn"), size=L, p=[.2, .2, .2, .2, .02, .04, .04, .04, .04, .02]))
        w = int(rng.choice(wrap))
        lines = [s[k:k + w] for k in range(0, L, w)] if w else ([s] if L else [])
        eol = "\r\n" if rng.random() < crlf_frac else "\n"
        pad = " \t"[: int(rng.integers(0, 3))] if rng.random() < space_frac else ""
        rec = ">seq%d desc > %d%s%s" % (i, i % 11, pad, eol)
        for ln in lines:
            lead = " " if rng.random() < space_frac / 4 else ""
            rec += lead + ln + pad + eol
            if rng.random() < blank_frac:
                rec += eol
        parts.append(rec)
    text = "".join(parts)
    if not final_newline:
        text = text.rstrip("\r\n")
    return text.encode("latin-1")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "fastx.json")) as f:
        return json.load(f)


def test_device_fasta_scan_matches_reference_parser(gpu_model, golden):
    seen = 0
    for name, case in golden["cases"].items():
        if case["type"] != "fasta":
            continue
        text = case["text"].encode("latin-1")
        d_buf, rec, n, consumed = gpu_model.scan_fasta(text, final_chunk=True)
        assert _records_from_index(d_buf, rec) == [tuple(r) for r in case["records"]], name
        assert consumed == len(text), name
        seen += 1
    assert seen >= 4


@pytest.mark.parametrize("variant", ["plain", "crlf_spaces_blanks", "no_final_newline", "empty_records", "single_line"])
def test_device_fasta_scan_equals_host_scanner(gpu_model, variant):
    kw = {"plain": {}, "crlf_spaces_blanks": dict(crlf_frac=0.3, space_frac=0.3, blank_frac=0.1),
          "no_final_newline": dict(final_newline=False), "empty_records": dict(empty_frac=0.2, max_len=40),
          "single_line": dict(wrap=(0,), min_len=40, max_len=150)}[variant]
    text = _random_fasta(30000, 17, **kw)                      # several MB: hundreds of 16-KB scan tiles
    for final, cut, cap in ((True, len(text), None), (False, len(text) * 2 // 3 + 5, None), (True, len(text), 12345),
                            (False, 70001, None), (False, 9, None), (True, 0, None), (False, len(text), 777)):
        t = text[:cut]
        d_buf, rec, n, consumed = gpu_model.scan_fasta(t, final_chunk=final, max_records=cap)
        hn, hc, hrecs = _host_scan(t, final, cap)
        assert (n, consumed) == (hn, hc), (variant, final, cut, cap)
        assert _records_from_index(d_buf, rec) == hrecs, (variant, final, cut, cap)


def test_device_fasta_quirks_of_the_reference_parser(gpu_model):
    """Sequence lines before the first header stay attached to it; a header-less file is one record with an empty
    header; a last header without sequence is dropped, an empty record in the middle is kept (fastx_parser.py:39-55)."""
    for text in (b"ACGT\nacgt\n>h1\nTT\n>h2\nGG\n", b"acgtnn\nACGT", b">h1\nAC\n>h2\n>h3\nGT\n>h4\n", b"\n\n  \n", b">only\n"):
        d_buf, rec, n, consumed = gpu_model.scan_fasta(text, final_chunk=True)
        hn, hc, hrecs = _host_scan(text, True)
        assert (n, consumed) == (hn, hc), text
        assert _records_from_index(d_buf, rec) == hrecs, text


def test_device_fasta_partition_equals_host_writer(gpu_model, tmp_path):
    text = _random_fasta(40000, 23, crlf_frac=0.2, space_frac=0.2, blank_frac=0.05, empty_frac=0.02)
    d_buf, rec, n, _ = gpu_model.scan_fasta(text)
    rng = np.random.default_rng(5)
    p = tmp_path / "x.fa"
    p.write_bytes(text)
    ch = next(iter(FastxReader(str(p), max_records=n + 1)))
    assert ch.n == n
    for probs in ([0.7, 0.25, 0.05], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]):
        labels = rng.choice(np.array([0, 1, -1], np.int8), size=n, p=probs)
        out, sizes = gpu_model.partition_fasta(d_buf, rec, labels)
        outs, hsizes = partition_records(ch, labels, (True, True, True), 3)
        want = b"".join(b"" if o is None else o.tobytes() for o in outs)
        assert np.array_equal(sizes.numpy(), hsizes)
        assert out.cpu().numpy().tobytes() == want


def test_fasta_classify_records_is_bitwise_classify(gpu_model):
    text = _random_fasta(20000, 29, min_len=30, max_len=200)
    d_buf, rec, n, _ = gpu_model.scan_fasta(text)
    hn, _, hrecs = _host_scan(text)
    assert hn == n
    seq = np.frombuffer("".join(r[1] for r in hrecs).encode("latin-1"), np.uint8).copy()
    off = np.concatenate([[0], np.cumsum([len(r[1]) for r in hrecs])]).astype(np.int64)
    for prec in ("tc_mixed", "tc_exact", "fp32"):
        la, lab_a = gpu_model.classify_records(d_buf, rec, 100, semantics="padded", precision=prec)
        lb, _, lab_b = gpu_model.classify(seq, off, 100, semantics="padded", precision=prec)
        assert torch.equal(la, lb) and torch.equal(lab_a, lab_b), prec


def _cli(args):
    from ribodetector_b200 import detect
    return detect.main(args)


@pytest.mark.parametrize("paired", [False, True])
def test_cli_fasta_device_ingest_equals_host_ingest(gpu_model, tmp_path, paired):
    """The command line on FASTA input: the device path (default) writes the same files as --host_ingest."""
    n = 30000
    t1 = _random_fasta(n, 31, crlf_frac=0.1, space_frac=0.1, min_len=40, max_len=160)
    t2 = _random_fasta(n, 37, wrap=(0,), min_len=40, max_len=160)
    f1, f2 = tmp_path / "r1.fa", tmp_path / "r2.fasta"
    f1.write_bytes(t1)
    f2.write_bytes(t2)
    ins = [str(f1), str(f2)] if paired else [str(f1)]
    outs = {}
    for tag, extra in (("dev", []), ("host", ["--host_ingest"])):
        o = [str(tmp_path / ("%s_o%d.fa" % (tag, e))) for e in range(len(ins))]
        r = [str(tmp_path / ("%s_r%d.fa" % (tag, e))) for e in range(len(ins))]
        pred = _cli(["-l", "100", "-i"] + ins + ["-o"] + o + ["-r"] + r + (["-e", "both"] if paired else []) + extra)
        outs[tag] = [open(x, "rb").read() for x in o + r] + [pred.num_seqs, pred.num_rrna, pred.num_nonrrna, pred.num_unknown]
    assert outs["dev"] == outs["host"]
    assert outs["dev"][-4] == n


def test_fasta_streaming_small_blocks_equal_whole_file(gpu_model, tmp_path):
    """FastqGpuStream on FASTA with 64-KB blocks (records cut at every block edge) == one big block."""
    from ribodetector_b200.data_loader.fastq_gpu import FastqGpuStream
    text = _random_fasta(20000, 41, crlf_frac=0.1, blank_frac=0.05, min_len=30, max_len=250)
    f = tmp_path / "in.fa"
    f.write_bytes(text)
    res = {}
    for tag, blk in (("small", 1 << 16), ("big", 1 << 24)):
        sinks = {k: [open(tmp_path / ("%s_%s.fa" % (tag, k)), "wb")] for k in ("non", "rrna")}
        st = FastqGpuStream([gpu_model], [str(f)], 100, block_bytes=blk, threads=2)
        counts = st.run(sinks)
        for v in sinks.values():
            v[0].close()
        res[tag] = [open(tmp_path / ("%s_%s.fa" % (tag, k)), "rb").read() for k in ("non", "rrna")] + [counts.tolist(), st.num_seqs]
    assert res["small"] == res["big"] and res["big"][-1] == 20000
