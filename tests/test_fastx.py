"""CPU: the host-side record scanner / writer (rd_scan_fastx, rd_partition_records — no GPU
involved) against golden vectors produced by the REFERENCE's own parser (oracle/gen_golden_fastx.py)."""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from ribodetector_b200.data_loader import FastxReader, get_seq_format, open_for_write, partition_records


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "fastx.json")) as f:
        return json.load(f)


def _read_all(path, **kw):
    out = []
    for ch in FastxReader(path, **kw):
        out += ch.records()
    return out


def test_scanner_matches_reference_parser(golden, tmp_path):
    for name, case in golden["cases"].items():
        ext = ".fq" if case["type"] == "fastq" else ".fa"
        p = tmp_path / (name + ext)
        p.write_bytes(case["text"].encode("latin-1"))
        want = [tuple(r) for r in case["records"]]
        # one block; blocks far smaller than a record (forces carry-over / growth); 1 record per chunk
        for kw in ({}, {"block_bytes": 5}, {"max_records": 1, "block_bytes": 64}):
            assert _read_all(str(p), **kw) == want, (name, kw)
        gz = tmp_path / (name + ext + ".gz")
        with gzip.open(gz, "wb") as f:
            f.write(case["text"].encode("latin-1"))
        assert _read_all(str(gz), block_bytes=16) == want, name


def test_format_sniffing_matches_reference(golden):
    for name, want in golden["formats"].items():
        if want == "ValueError":
            with pytest.raises(ValueError):
                get_seq_format(name)
        else:
            assert get_seq_format(name) == want


def test_malformed_fastq_is_an_error(tmp_path):
    for text in ("@r1\nACGT\n+\nIIII\n\n@r2\nAC\n+\nII\n",      # blank line: the reference raises IndexError
                 "r1\nACGT\n+\nIIII\n"):                          # no '@'
        p = tmp_path / "bad.fq"
        p.write_text(text)
        with pytest.raises(ValueError):
            _read_all(str(p))


def test_partition_preserves_order_and_text(golden, tmp_path):
    """Output = '\\n'.join(record) + '\\n' per record, routed by label, input order (detect.py:601-614,295-298)."""
    rng = np.random.default_rng(1)
    recs = [("@r%d x" % i, "".join(rng.choice(list("ACGTN"), size=int(rng.integers(1, 80)))), "+", None) for i in range(20000)]
    recs = [(h, s, p, "I" * len(s)) for h, s, p, _ in recs]
    path = tmp_path / "big.fq"
    path.write_text("".join("\n".join(r) + "\n" for r in recs))
    labels_all = rng.choice(np.array([0, 1, -1], np.int8), size=len(recs), p=[0.7, 0.25, 0.05])
    got = [b"", b"", b""]
    seen = 0
    for ch in FastxReader(str(path), max_records=6000):
        lab = labels_all[seen:seen + ch.n]
        for threads in (1, 5):
            outs, sizes = partition_records(ch, lab, (True, True, True), threads)
            for c, o in enumerate(outs):
                assert (0 if o is None else o.size) == sizes[c]
        for c, o in enumerate(outs):
            got[c] += b"" if o is None else o.tobytes()
        only_non, sizes2 = partition_records(ch, lab, (True, False, False), 2)
        assert only_non[1] is None and only_non[2] is None and np.array_equal(sizes, sizes2)
        seen += ch.n
    assert seen == len(recs)
    for c, want_label in enumerate((0, 1, -1)):
        want = "".join("\n".join(r) + "\n" for r, l in zip(recs, labels_all) if l == want_label)
        assert got[c].decode() == want


def test_fasta_partition_writes_uppercased_joined_sequence(tmp_path):
    p = tmp_path / "x.fa"
    p.write_text(">a\nacgt\nnn\n>b\nGG\n")
    ch = next(iter(FastxReader(str(p))))
    outs, sizes = partition_records(ch, np.array([1, 0], np.int8))
    assert outs[0].tobytes() == b">b\nGG\n" and outs[1].tobytes() == b">a\nACGTNN\n" and outs[2] is None


def test_open_for_write_gz(tmp_path):
    with open_for_write(str(tmp_path / "o.fq.gz")) as f:
        f.write(b"@r\nA\n+\nI\n")
    assert gzip.open(tmp_path / "o.fq.gz").read() == b"@r\nA\n+\nI\n"


def test_cli_argument_surface():
    """Flags and defaults of detect.py:764-798 / detect_cpu.py:777-807."""
    from ribodetector_b200 import detect
    a = detect.build_parser(True).parse_args(["-l", "100", "-i", "a.fq", "b.fq", "-o", "c.fq", "d.fq"])
    assert (a.len, a.input, a.output, a.rrna, a.ensure, a.threads, a.memory, a.chunk_size, a.log, a.deviceid, a.config) == \
        (100, ["a.fq", "b.fq"], ["c.fq", "d.fq"], None, "none", 10, 32, None, None, None, None)
    c = detect.build_parser(False).parse_args(["-l", "50", "-i", "a.fq", "-o", "c.fq", "-e", "both", "-t", "3"])
    assert c.threads == 3 and c.ensure == "both" and not hasattr(c, "memory") and not hasattr(c, "deviceid")
    assert detect.build_parser(False).parse_args(["-l", "5", "-i", "a", "-o", "b"]).threads == 20
    with pytest.raises(SystemExit):
        detect.build_parser(True).parse_args(["-i", "a.fq", "-o", "c.fq"])          # -l is required


def test_config_json_surface():
    from ribodetector_b200.parse_config import ConfigParser
    from ribodetector_b200 import detect
    from ribodetector_b200.model import model as module_arch
    cfg = ConfigParser.from_json(os.path.join(detect.cd, "config.json"))
    assert cfg["arch"]["type"] == "SeqModel" and cfg["n_gpu"] == 1
    assert set(cfg["arch"]["args"]) == {"input_size", "hidden_size", "num_layers", "num_classes", "batch_first",
                                        "bidirectional", "pack_seq"}
    assert set(cfg["state_file"]) == {"mcc", "recall"}
    for v in cfg["state_file"].values():
        assert os.path.exists(os.path.join(detect.cd, v))
    m = cfg.init_obj("arch", module_arch)                       # the plugin seam: getattr(module, arch.type)(**args)
    assert type(m).__name__ == "SeqModel" and m.pack_seq is True


REF_ROOT = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="reference checkout not mounted (GPU box)")
def test_scanner_differential_fuzz_against_reference_parser(tmp_path):
    """Random FASTQ/FASTA texts (odd whitespace, CRLF, truncated tails, multi-line FASTA, '@' qualities)
    through rd_scan_fastx at several block sizes vs the reference's seq_parser imported unmodified."""
    import io
    import sys
    import types
    sys.path.insert(0, REF_ROOT)
    bio, seqm = types.ModuleType("Bio"), types.ModuleType("Bio.Seq")
    seqm.Seq = object
    sys.modules.setdefault("Bio", bio)
    sys.modules.setdefault("Bio.Seq", seqm)
    from ribodetector.data_loader.fastx_parser import seq_parser
    rng = np.random.default_rng(7)
    alpha = list("ACGTNacgtnRYU-")

    def rseq(lo, hi):
        return "".join(rng.choice(alpha, size=int(rng.integers(lo, hi))))

    for trial in range(60):
        eol = "\r\n" if trial % 3 == 0 else "\n"
        pad = lambda: " " * int(rng.integers(0, 3)) if trial % 4 == 0 else ""     # noqa: E731
        n = int(rng.integers(0, 40))
        if trial % 2 == 0:
            typ, ext, text = "fastq", ".fq", ""
            for i in range(n):
                s = rseq(1, 120)
                q = "".join(rng.choice(list("@+I#5"), size=len(s)))
                text += "".join(["@r%d d" % i, pad(), eol, s, pad(), eol, "+", "x" * int(rng.integers(0, 2)), pad(), eol,
                                 q, pad(), eol])
            if n and rng.random() < 0.4:                       # truncated tail / missing final newline
                text = text[:len(text) - int(rng.integers(1, 30))]
        else:
            typ, ext, text = "fasta", ".fa", ""
            for i in range(n):
                text += ">s%d%s%s" % (i, pad(), eol)
                for _ in range(int(rng.integers(0, 4))):
                    text += pad() + rseq(1, 70) + pad() + eol
                if rng.random() < 0.2:
                    text += eol
            if n and rng.random() < 0.3:
                text = text.rstrip("\r\n")
        try:
            want = [tuple(r) for r in seq_parser(io.StringIO(text, newline=None), typ)]
            ref_error = False
        except IndexError:                                     # blank line inside a FASTQ record
            ref_error = True
        p = tmp_path / ("t%d%s" % (trial, ext))
        p.write_bytes(text.encode("latin-1"))
        for kw in ({}, {"block_bytes": 7}, {"max_records": 3, "block_bytes": 50}):
            if ref_error:
                with pytest.raises(ValueError):
                    _read_all(str(p), **kw)
            else:
                try:
                    got = _read_all(str(p), **kw)
                except ValueError:
                    # the only tolerated divergence: a non-'@' line where a header is expected, which the
                    # reference silently turns into a garbage record (fastx_parser.py:31-36)
                    assert typ == "fastq" and any(not l.startswith("@") for l in text.splitlines()[0::4] if l.strip())
                    continue
                assert got == want, (trial, kw)


def _bgzf_bytes(text, block=0xff00, eof_marker=True):
    """BGZF framing of `text` (what bgzip / bcl2fastq write): <= 64 KB members with a 'BC' size subfield."""
    import struct
    import zlib
    out = []
    chunks = [text[i:i + block] for i in range(0, len(text), block)] + ([b""] if eof_marker else [])
    for c in chunks:
        z = zlib.compressobj(6, zlib.DEFLATED, -15)
        d = z.compress(c) + z.flush()
        bsize = 12 + 6 + len(d) + 8 - 1
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize)
                   + d + struct.pack("<II", zlib.crc32(c), len(c)))
    return b"".join(out)


def test_bgzf_reader_inflates_members_in_parallel(tmp_path):
    from ribodetector_b200.data_loader import BgzfReader, open_text
    rng = np.random.default_rng(3)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(20, 150))).tolist()), b"I" * 10)
                    for i in range(40000))
    p = tmp_path / "x.fq.gz"
    p.write_bytes(_bgzf_bytes(text))
    assert BgzfReader.sniff(str(p)) and gzip.open(p).read() == text
    for threads, step in ((1, 1 << 20), (4, 1 << 22), (3, 70001), (2, 1000)):     # 1000 < a member: the spill path
        r = open_text(str(p), True, threads)
        assert isinstance(r, BgzfReader)
        got = bytearray()
        buf = np.empty(step, np.uint8)
        while True:
            k = r.readinto(memoryview(buf))
            if not k:
                break
            got += buf[:k].tobytes()
        r.close()
        assert bytes(got) == text, (threads, step)
    plain_gz = tmp_path / "y.fq.gz"
    with gzip.open(plain_gz, "wb") as f:
        f.write(text[:1000])
    assert not BgzfReader.sniff(str(plain_gz)) and not isinstance(open_text(str(plain_gz), True), BgzfReader)
    # corrupt member -> error, truncated file -> error
    raw = bytearray(_bgzf_bytes(text))
    raw[200] ^= 0x55
    (tmp_path / "bad.fq.gz").write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        r = BgzfReader(str(tmp_path / "bad.fq.gz"))
        r.readinto(memoryview(np.empty(1 << 20, np.uint8)))
    (tmp_path / "cut.fq.gz").write_bytes(_bgzf_bytes(text)[:-500])
    with pytest.raises(ValueError):
        r = BgzfReader(str(tmp_path / "cut.fq.gz"))
        buf = np.empty(1 << 24, np.uint8)
        while r.readinto(memoryview(buf)):
            pass


def test_fastx_reader_takes_bgzf_input(tmp_path, golden):
    case = golden["cases"]["fq_plain"]
    p = tmp_path / "g.fq.gz"
    p.write_bytes(_bgzf_bytes(case["text"].encode("latin-1"), block=37))
    assert _read_all(str(p), block_bytes=64) == [tuple(r) for r in case["records"]]


def test_gz_stream_reader_matches_gzip_module(tmp_path):
    from ribodetector_b200.data_loader import GzStreamReader, open_text
    rng = np.random.default_rng(5)
    text = bytes(rng.choice(list(b"ACGTN\n@+I"), size=3_000_000).tolist())

    def read_all(path, step):
        r = open_text(str(path), True)
        assert isinstance(r, GzStreamReader)
        got, buf = bytearray(), np.empty(step, np.uint8)
        while True:
            k = r.readinto(memoryview(buf))
            if not k:
                break
            got += buf[:k].tobytes()
        r.close()
        return bytes(got)

    one = tmp_path / "one.fq.gz"
    with gzip.open(one, "wb") as f:
        f.write(text)
    multi = tmp_path / "multi.fq.gz"                      # concatenated members (what ParallelGzipWriter writes), zero padding
    multi.write_bytes(gzip.compress(text[:1_000_000]) + gzip.compress(text[1_000_000:2_500_000]) + b"\0" * 37
                      + gzip.compress(text[2_500_000:]) + b"\0" * 512)
    with open_for_write(str(tmp_path / "ours.fq.gz"), 3) as f:
        f.write(text)
    empty = tmp_path / "empty.fq.gz"
    empty.write_bytes(gzip.compress(b""))
    for step in (1 << 22, 65537, 100):
        assert read_all(one, step) == text
        assert read_all(multi, step) == text
        assert read_all(tmp_path / "ours.fq.gz", step) == text
        assert read_all(empty, step) == b""
    (tmp_path / "cut.fq.gz").write_bytes(one.read_bytes()[:-100])
    with pytest.raises(EOFError):
        read_all(tmp_path / "cut.fq.gz", 1 << 20)
    bad = bytearray(one.read_bytes())
    bad[len(bad) // 2] ^= 0xFF
    (tmp_path / "bad.fq.gz").write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        read_all(tmp_path / "bad.fq.gz", 1 << 20)


def test_fastx_reader_memory_stays_bounded_for_small_chunks(tmp_path):
    """with max_records x record_bytes far below the block size the reader used to carry (and double) an
    ever larger tail.  Buffers must stay within block_bytes, every chunk but the last must hold exactly max_records
    records, and the records must come out unchanged — FASTA and FASTQ."""
    rng = np.random.default_rng(5)
    n = 60000
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(l))) for l in rng.integers(40, 90, size=n)]
    fa = tmp_path / "small.fa"
    fa.write_text("".join(">r%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    fq = tmp_path / "small.fq"
    fq.write_text("".join("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)) for i, s in enumerate(seqs)))
    for path, width in ((fa, 2), (fq, 4)):
        block = 1 << 20
        with FastxReader(str(path), max_records=4096, block_bytes=block, threads=2) as rd:
            got, sizes = [], []
            for chunk in rd:
                sizes.append(chunk.n)
                got.extend(r[1] for r in chunk.records())
                assert len(chunk.records()[0]) == width
                chunk.release()
            assert rd.peak_buffer_bytes <= block, rd.peak_buffer_bytes
        assert got == seqs
        assert all(s == 4096 for s in sizes[:-1]) and sum(sizes) == n
    # a record larger than the block still grows the buffer instead of looping
    big = tmp_path / "big.fa"
    big.write_text(">x\n" + "ACGT" * 100000 + "\n>y\nAC\n")
    with FastxReader(str(big), max_records=8, block_bytes=1 << 16) as rd:
        recs = [r for c in rd for r in c.records()]
    assert [len(r[1]) for r in recs] == [400000, 2]
