"""CPU: the host logic of the device-ingest pipeline (data_loader/fastq_gpu.py — block cutting, carry-over of cut
records, pairing of two files with different record sizes, slot recycling, error paths) against a STAND-IN model
that implements fastq_submit / fastq_collect with the library's host scanner and writer (no GPU involved).  The
real rd_fastq_submit / rd_fastq_collect are checked against the same host functions in tests/test_gpu_fastq.py."""
import ctypes
import gzip

import numpy as np
import pytest

from ribodetector_b200 import _lib
from ribodetector_b200.data_loader.fastq_gpu import FastqGpuStream


def _label_rule(seq_bytes):
    return 1 if len(seq_bytes) % 3 == 0 else 0


class HostStandIn:
    """Same contract as SeqModel.fastq_submit / fastq_collect; labels = a rule on the R1 sequence length."""

    def __init__(self):
        self.lib = _lib.load_library()
        self.slots = {}
        self.submits = 0

    def _scan(self, buf, n_bytes, final, cap, fasta=False):
        p = lambda a: ctypes.c_void_p(a.ctypes.data)      # noqa: E731
        hdr, plus, qual = (np.empty(2 * cap + 2, np.int64) for _ in range(3))
        seq = np.empty(n_bytes + 1, np.uint8)
        off = np.empty(cap + 2, np.int64)
        consumed = ctypes.c_int64(0)
        n = self.lib.rd_scan_fastx(p(buf), n_bytes, int(fasta), int(final), cap, p(hdr), None if fasta else p(plus),
                                   None if fasta else p(qual), p(seq), n_bytes, p(off), ctypes.byref(consumed), 2)
        if n < 0:
            raise ValueError(self.lib.rd_fastx_last_error().decode())
        return n, consumed.value, hdr, plus, qual, seq, off

    def fastq_submit(self, slot, bufs, lens, final_chunk, max_records, max_len, outs, labels=None, mode="none",
                     semantics=None, precision=None, fasta=False):
        assert slot not in self.slots, "slot resubmitted before it was collected"
        self.submits += 1
        scans = [self._scan(b, int(l), final_chunk, max_records, fasta) for b, l in zip(bufs, lens)]
        n = min(s[0] for s in scans)
        consumed, sizes = [], np.zeros((2, 3), np.int64)
        lab = None
        for e, (ne, ce, hdr, plus, qual, seq, off) in enumerate(scans):
            if ne > n and fasta:                          # give the surplus records back: a rescan capped at n records
                ce = 0 if n == 0 else self._scan(bufs[e], int(lens[e]), final_chunk, n, True)[1]
            elif ne > n:                                  # FASTQ: cut after record n-1
                ce = 0 if n == 0 else int(np.flatnonzero(bufs[e][:int(lens[e])] == 10)[4 * n - 1]) + 1
            consumed.append(ce)
            if n == 0:
                continue
            if lab is None:
                lab = np.array([_label_rule(seq[off[i]:off[i + 1]]) for i in range(n)], np.int8)
            p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
            s3 = np.zeros(3, np.int64)
            args = (p(bufs[e]), int(fasta), n, p(hdr), None if fasta else p(plus), None if fasta else p(qual), p(seq), p(off), p(lab))
            assert self.lib.rd_partition_records(*args, None, None, None, p(s3), 1) == 0
            o = [np.empty(max(int(x), 1), np.uint8) for x in s3]
            assert self.lib.rd_partition_records(*args, p(o[0]), p(o[1]), p(o[2]), p(s3), 1) == 0
            cat = np.concatenate([o[c][:int(s3[c])] for c in range(3)])
            outs[e][:cat.size] = cat
            sizes[e] = s3
        if n:
            self.slots[slot] = (sizes, np.array([(lab == 0).sum(), (lab == 1).sum(), 0], np.int64))
        return n, consumed, [c + 2 for c in consumed]

    def fastq_collect(self, slot):
        if slot not in self.slots:
            raise _lib.RdError("slot not pending")
        return self.slots.pop(slot)


def _fastq(n, seed, min_len=5, max_len=90, crlf=False, final_newline=True):
    rng = np.random.default_rng(seed)
    eol = "\r\n" if crlf else "\n"
    recs = []
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        s = "".join(rng.choice(list("ACGTN"), size=L))
        recs.append(("@q%d" % i, s, "+", "".join(chr(int(c)) for c in rng.integers(33, 74, size=L))))
    text = "".join(eol.join(r) + eol for r in recs)
    if not final_newline:
        text = text.rstrip("\r\n")
    return recs, text.encode()


def _expect(recs, labels, want):
    return "".join("\n".join(r) + "\n" for r, l in zip(recs, labels) if l == want).encode()


@pytest.mark.parametrize("block", [4096, 50001, 1 << 22])
@pytest.mark.parametrize("gz", [False, True])
def test_single_end_blocks_and_carry_over(tmp_path, block, gz):
    recs, text = _fastq(3000, 1, final_newline=(block != 50001))
    p = tmp_path / ("in.fq.gz" if gz else "in.fq")
    if gz:
        with gzip.open(p, "wb") as f:
            f.write(text)
    else:
        p.write_bytes(text)
    m = HostStandIn()
    with open(tmp_path / "non.fq", "wb") as fn, open(tmp_path / "rr.fq", "wb") as fr:
        st = FastqGpuStream([m], [str(p)], 100, block_bytes=block, threads=2)
        counts = st.run({"non": [fn], "rrna": [fr], "unc": None})
    labels = [_label_rule(r[1]) for r in recs]
    assert (tmp_path / "non.fq").read_bytes() == _expect(recs, labels, 0)
    assert (tmp_path / "rr.fq").read_bytes() == _expect(recs, labels, 1)
    assert st.num_seqs == 3000 and counts.tolist() == [labels.count(0), labels.count(1), 0]
    assert not m.slots and (m.submits >= len(text) // block or block > len(text))


def test_pairs_with_different_record_sizes_two_models(tmp_path):
    r1, t1 = _fastq(2500, 2, 30, 60)
    r2, t2 = _fastq(2500, 3, 5, 120, crlf=True)
    (tmp_path / "a.fq").write_bytes(t1)
    (tmp_path / "b.fq").write_bytes(t2)
    models = [HostStandIn(), HostStandIn()]
    names = ["n1", "n2", "x1", "x2"]
    fhs = [open(tmp_path / x, "wb") for x in names]
    st = FastqGpuStream(models, [str(tmp_path / "a.fq"), str(tmp_path / "b.fq")], 100, mode="rrna", block_bytes=8192, threads=2)
    st.run({"non": fhs[:2], "rrna": fhs[2:], "unc": None})
    for fh in fhs:
        fh.close()
    labels = [_label_rule(r[1]) for r in r1]                # the stand-in labels a pair by its R1
    assert (tmp_path / "n1").read_bytes() == _expect(r1, labels, 0) and (tmp_path / "x1").read_bytes() == _expect(r1, labels, 1)
    assert (tmp_path / "n2").read_bytes() == _expect(r2, labels, 0) and (tmp_path / "x2").read_bytes() == _expect(r2, labels, 1)
    assert st.num_seqs == 2500 and models[0].submits > 5 and models[1].submits > 5


def test_error_paths(tmp_path):
    _, t1 = _fastq(500, 4)
    _, t2 = _fastq(490, 5)
    (tmp_path / "a.fq").write_bytes(t1)
    (tmp_path / "b.fq").write_bytes(t2)
    sinks = {"non": [open(tmp_path / "o1", "wb"), open(tmp_path / "o2", "wb")], "rrna": None, "unc": None}
    for block in (2048, 1 << 22):
        with pytest.raises(RuntimeError, match="different numbers"):
            FastqGpuStream([HostStandIn()], [str(tmp_path / "a.fq"), str(tmp_path / "b.fq")], 100, block_bytes=block).run(sinks)
    (tmp_path / "bad.fq").write_bytes(t1.replace(b"@q300\n", b"\n@q300\n"))
    with pytest.raises(ValueError, match="blank line"):
        FastqGpuStream([HostStandIn()], [str(tmp_path / "bad.fq")], 100, block_bytes=4096).run(sinks)
    (tmp_path / "long.fq").write_bytes(b"@r\n" + b"A" * 5000 + b"\n+\n" + b"I" * 5000 + b"\n")
    with pytest.raises(RuntimeError, match="does not fit"):
        FastqGpuStream([HostStandIn()], [str(tmp_path / "long.fq")], 100, block_bytes=1024).run(sinks)
    with pytest.raises(ValueError):                       # one kind of input per run: FASTQ and FASTA do not mix
        FastqGpuStream([HostStandIn()], [str(tmp_path / "r1.fq"), str(tmp_path / "x.fa")], 100)
    (tmp_path / "empty.fq").write_bytes(b"")
    st = FastqGpuStream([HostStandIn()], [str(tmp_path / "empty.fq")], 100)
    st.run(sinks)
    assert st.num_seqs == 0


def test_randomised_blocks_against_a_line_parser(tmp_path):
    """200 random small files (CRLF, trailing blanks, no final newline, a truncated last record) x random block sizes:
    the pipeline's output equals what the reference's four-line state machine (fastx_parser.py:15-47) would route."""
    rng = np.random.default_rng(11)
    for case in range(200):
        n = int(rng.integers(0, 40))
        recs, lines = [], []
        for i in range(n):
            L = int(rng.integers(1, 30))
            seq = "".join(rng.choice(list("ACGTN"), size=L))
            rec = ("@r%d" % i + (" x" if rng.random() < 0.3 else ""), seq, "+" + ("r%d" % i if rng.random() < 0.2 else ""),
                   "".join(chr(int(c)) for c in rng.integers(33, 74, size=L)))
            recs.append(rec)
            for ln in rec:
                lines.append(ln + (" " if rng.random() < 0.1 else "") + ("\r\n" if rng.random() < 0.2 else "\n"))
        text = "".join(lines)
        if n and rng.random() < 0.3:
            text = text.rstrip("\r\n ")                                   # no final newline (the record still counts)
            recs[-1] = tuple(x.rstrip() for x in recs[-1])
        if rng.random() < 0.3 and (not text or text.endswith("\n")):
            text += "@cut\nACGT\n+"                                        # truncated final record: dropped
        p = tmp_path / ("c%d.fq" % case)
        p.write_bytes(text.encode())
        block = int(rng.integers(64, 600))
        longest = max([len(l) for l in lines] + [1]) * 4 + 16
        block = max(block, longest)
        with open(tmp_path / "n", "wb") as fn, open(tmp_path / "r", "wb") as fr:
            st = FastqGpuStream([HostStandIn(), HostStandIn()], [str(p)], 100, block_bytes=block, threads=1)
            st.run({"non": [fn], "rrna": [fr], "unc": None})
        labels = [_label_rule(r[1]) for r in recs]
        assert (tmp_path / "n").read_bytes() == _expect(recs, labels, 0), (case, block)
        assert (tmp_path / "r").read_bytes() == _expect(recs, labels, 1), (case, block)
        assert st.num_seqs == n


def _fasta(n, seed, min_len=5, max_len=150, wrap=(0, 17, 60)):
    rng = np.random.default_rng(seed)
    recs, text = [], ""
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        s = "".join(rng.choice(list("ACGTNacgt"), size=L))
        w = int(rng.choice(wrap))
        lines = [s[k:k + w] for k in range(0, L, w)] if w else [s]
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        text += ">f%d d%s" % (i, eol) + "".join(ln + eol for ln in lines) + (eol if rng.random() < 0.1 else "")
        recs.append((">f%d d" % i, s.upper()))
    return recs, text.encode()


@pytest.mark.parametrize("block", [2048, 30011, 1 << 22])
def test_fasta_blocks_and_carry_over(tmp_path, block):
    """The same block-cutting logic on FASTA text (records end at the NEXT header, so every block leaves its last record
    for the next one): output = the reference parser's 2-line records routed by label, whatever the block size."""
    recs, text = _fasta(2500, 7)
    p = tmp_path / "in.fasta"
    p.write_bytes(text)
    m = HostStandIn()
    with open(tmp_path / "non.fa", "wb") as fn, open(tmp_path / "rr.fa", "wb") as fr:
        st = FastqGpuStream([m], [str(p)], 100, block_bytes=block, threads=2)
        counts = st.run({"non": [fn], "rrna": [fr], "unc": None})
    labels = [_label_rule(r[1].encode()) for r in recs]
    assert (tmp_path / "non.fa").read_bytes() == _expect(recs, labels, 0)
    assert (tmp_path / "rr.fa").read_bytes() == _expect(recs, labels, 1)
    assert st.num_seqs == 2500 and counts.tolist() == [labels.count(0), labels.count(1), 0]


def test_fasta_pairs_and_unequal_files(tmp_path):
    r1, t1 = _fasta(1200, 8, 30, 60)
    r2, t2 = _fasta(1200, 9, 5, 200, wrap=(0,))
    (tmp_path / "a.fa").write_bytes(t1)
    (tmp_path / "b.fa").write_bytes(t2)
    fhs = [open(tmp_path / x, "wb") for x in ("n1", "n2", "x1", "x2")]
    st = FastqGpuStream([HostStandIn(), HostStandIn()], [str(tmp_path / "a.fa"), str(tmp_path / "b.fa")], 100, mode="rrna",
                        block_bytes=8192, threads=2)
    st.run({"non": fhs[:2], "rrna": fhs[2:], "unc": None})
    for fh in fhs:
        fh.close()
    labels = [_label_rule(r[1].encode()) for r in r1]
    assert (tmp_path / "n1").read_bytes() == _expect(r1, labels, 0) and (tmp_path / "x1").read_bytes() == _expect(r1, labels, 1)
    assert (tmp_path / "n2").read_bytes() == _expect(r2, labels, 0) and (tmp_path / "x2").read_bytes() == _expect(r2, labels, 1)
    assert st.num_seqs == 1200
    _, t3 = _fasta(1190, 9, 5, 200, wrap=(0,))
    (tmp_path / "c.fa").write_bytes(t3)
    sinks = {"non": [open(tmp_path / "o1", "wb"), open(tmp_path / "o2", "wb")], "rrna": None, "unc": None}
    for block in (4096, 1 << 22):
        with pytest.raises(RuntimeError, match="different numbers"):
            FastqGpuStream([HostStandIn()], [str(tmp_path / "a.fa"), str(tmp_path / "c.fa")], 100, block_bytes=block).run(sinks)
