"""FASTQ/FASTA ingest and label-partitioned output on byte buffers.

Mirrors the reference's loaders — ``get_seq_format`` / ``load_reads`` / ``get_seq_chunks``
(``ribodetector/data_loader/seq_encoder.py:21-39,56-92``) on top of ``seq_parser``
(``fastx_parser.py:15-55``) — but never builds Python strings per record: the file is read in large
blocks, ``rd_scan_fastx`` (C ABI, host code in librd_b200.so) indexes the records, and the sequence
bytes go to the GPU path as one buffer + offsets.  ``partition_records`` replaces the
``'\\n'.join(record)`` / ``separate_reads`` / ``fh.write`` sequence (``detect.py:680,601-614,295-298``).
"""
import ctypes
import gzip
import queue
import time
from mimetypes import guess_type
from pathlib import Path

import numpy as np

from .. import _lib

FA_EXTS = (".fasta", ".fa", ".fna", ".fas")
FQ_EXTS = (".fq", ".fastq")


def get_seq_format(seq_file):
    """'fq' | 'fa' (+ 'gz') from the file name, same rules and errors as seq_encoder.py:21-39."""
    encoding = guess_type(seq_file)[1]
    if encoding is None:
        encoding = ""
    elif encoding == "gzip":
        encoding = "gz"
    else:
        raise ValueError('Unknown file encoding: "{}"'.format(encoding))
    name = Path(seq_file).stem if encoding == "gz" else Path(seq_file).name
    ext = Path(name).suffix
    if ext not in FA_EXTS + FQ_EXTS:
        raise ValueError("""Unknown extension {}. Only fastq and fasta sequence formats are supported.
And the file must end with one of ".fasta", ".fa", ".fna", ".fas", ".fq", ".fastq"
and followed by ".gz" or ".gzip" if they are gzipped.""".format(ext))
    return ("fa" if ext in FA_EXTS else "fq") + encoding


class ParallelGzipWriter:
    """gzip level 5 like the reference's open_for_write (detect.py:729-741), but every write() is cut into
    blocks that are deflated on `threads` threads (zlib releases the GIL) and appended as separate gzip
    members.  A multi-member file is a valid .gz: gzip / zcat / Python's gzip read back the identical
    text; the reference's single-threaded writer is what makes it "2 times slower to write gz files"."""

    BLOCK = 8 << 20

    def __init__(self, path, threads=4, compresslevel=5):
        from concurrent.futures import ThreadPoolExecutor
        self.fh = open(path, "wb")
        self.level = compresslevel
        self.pool = ThreadPoolExecutor(max(1, int(threads)))

    def write(self, data):
        mv = memoryview(data).cast("B")
        blocks = [mv[i:i + self.BLOCK] for i in range(0, len(mv), self.BLOCK)]
        for z in self.pool.map(lambda b: gzip.compress(b, self.level), blocks):
            self.fh.write(z)
        return len(mv)

    def close(self):
        if self.fh.tell() == 0:
            self.fh.write(gzip.compress(b"", self.level))     # an empty but valid gzip file
        self.fh.close()
        self.pool.shutdown()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_for_write(read_file, threads=4):
    """Binary twin of detect.py:729-741: gzip level 5 for names ending in 'gz', plain otherwise."""
    if read_file.endswith("gz"):
        return ParallelGzipWriter(read_file, threads)
    return open(read_file, "wb")


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class BgzfReader:
    """File-like reader (readinto / close) of BGZF-framed .gz text — bgzip, bcl2fastq / BCL Convert FASTQ.gz: every
    gzip member states its compressed size, so `rd_bgzf_inflate` inflates the members of a buffer side by side on
    `threads` host threads.  The reference reads every .gz through `gzip.open` on one thread (seq_encoder.py:43-53)."""

    COMP = 32 << 20

    def __init__(self, path, threads=8):
        self.fh = open(path, "rb", buffering=0)
        self.threads = max(1, int(threads))
        self.lib = _lib.load_library()
        self.comp = np.empty(self.COMP, np.uint8)
        self.lo = self.hi = 0                       # unread compressed bytes are comp[lo:hi]
        self.raw_eof = False
        self.spill = b""                            # text of one member that did not fit the caller's buffer

    @staticmethod
    def sniff(path):
        """True if the file starts with a BGZF member (gzip FEXTRA with a 'BC' subfield)."""
        with open(path, "rb") as f:
            h = f.read(18)
        return len(h) >= 18 and h[:4] == b"\x1f\x8b\x08\x04" and h[12:14] == b"BC"

    def _refill(self):
        if self.lo and self.lo < self.hi:
            self.comp[:self.hi - self.lo] = self.comp[self.lo:self.hi]
        self.hi -= self.lo
        self.lo = 0
        mv = memoryview(self.comp)
        while self.hi < self.COMP and not self.raw_eof:
            k = self.fh.readinto(mv[self.hi:])
            if not k:
                self.raw_eof = True
            self.hi += k or 0

    def readinto(self, b):
        out = np.frombuffer(b, np.uint8)
        pos = 0
        if self.spill:
            k = min(len(self.spill), out.size)
            out[:k] = np.frombuffer(self.spill[:k], np.uint8)
            self.spill = self.spill[k:]
            pos = k
        used = ctypes.c_int64(0)
        while pos < out.size:
            if self.hi - self.lo < (1 << 17) and not self.raw_eof:
                self._refill()
            if self.lo == self.hi:
                break                                # end of file
            n = self.lib.rd_bgzf_inflate(ctypes.c_void_p(self.comp.ctypes.data + self.lo), self.hi - self.lo,
                                         ctypes.c_void_p(out.ctypes.data + pos), out.size - pos, ctypes.byref(used),
                                         self.threads)
            if n < 0:
                raise ValueError("%s" % self.lib.rd_fastx_last_error().decode("utf-8", "replace"))
            self.lo += used.value
            pos += n
            if used.value == 0:
                if self.hi - self.lo < (1 << 17) and not self.raw_eof:
                    continue                         # an incomplete member: read more
                if self.hi - self.lo >= 18 and pos < out.size:
                    # the next member's text does not fit what is left of `out`: inflate it aside and hand out a part
                    import zlib
                    d = zlib.decompressobj(31)
                    self.spill = d.decompress(self.comp[self.lo:self.hi].tobytes())
                    if not d.eof:
                        raise ValueError("BGZF: truncated or corrupt member")
                    self.lo = self.hi - len(d.unused_data)
                    k = min(len(self.spill), out.size - pos)
                    out[pos:pos + k] = np.frombuffer(self.spill[:k], np.uint8)
                    self.spill = self.spill[k:]
                    pos += k
                    continue
                if self.raw_eof and self.lo < self.hi:
                    raise ValueError("BGZF: truncated last member")
                break
        return pos

    def close(self):
        self.fh.close()


class GzStreamReader:
    """File-like reader (readinto / close) of any other .gz: zlib driven by `rd_gz_inflate` straight into the caller's
    buffer (no intermediate bytes objects; concatenated members, zero padding and CRC-32 handled like `gzip.open`)."""

    COMP = 8 << 20

    def __init__(self, path):
        self.fh = open(path, "rb", buffering=0)
        self.lib = _lib.load_library()
        self.g = ctypes.c_void_p(self.lib.rd_gz_open())
        if not self.g:
            raise MemoryError("rd_gz_open failed")
        self.comp = np.empty(self.COMP, np.uint8)
        self.lo = self.hi = 0
        self.raw_eof = False
        self.mid = ctypes.c_int(0)

    def readinto(self, b):
        out = np.frombuffer(b, np.uint8)
        pos = 0
        used = ctypes.c_int64(0)
        while pos < out.size:
            if self.lo == self.hi and not self.raw_eof:
                k = self.fh.readinto(memoryview(self.comp))
                self.lo, self.hi = 0, k or 0
                self.raw_eof = not k
            if self.lo == self.hi:
                if self.mid.value:
                    raise EOFError("Compressed file ended before the end-of-stream marker was reached")
                break
            n = self.lib.rd_gz_inflate(self.g, ctypes.c_void_p(self.comp.ctypes.data + self.lo), self.hi - self.lo,
                                       ctypes.byref(used), ctypes.c_void_p(out.ctypes.data + pos), out.size - pos,
                                       ctypes.byref(self.mid))
            if n < 0:
                raise ValueError(self.lib.rd_fastx_last_error().decode("utf-8", "replace"))
            self.lo += used.value
            pos += n
        return pos

    def close(self):
        self.fh.close()
        if self.g:
            self.lib.rd_gz_close(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:                    # noqa: BLE001
            pass


def open_text(path, gz, threads=8):
    """Binary reader of a sequence file's TEXT: the file itself, a parallel BGZF reader, or a zlib stream reader."""
    if not gz:
        return open(path, "rb", buffering=0)
    if BgzfReader.sniff(path):
        return BgzfReader(path, threads)
    return GzStreamReader(path)


def _host_array(n, dtype, pinned):
    """numpy array, page-locked when a CUDA device is there: the sequence bytes and offsets go to the
    GPU by cudaMemcpyAsync straight from these recycled buffers."""
    if pinned:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty(n, dtype={np.uint8: torch.uint8, np.int64: torch.int64}[dtype], pin_memory=True).numpy()
        except Exception:                    # noqa: BLE001 — fall back to pageable memory
            pass
    return np.empty(n, dtype)


class RecordChunk:
    """A block of file text plus the record index rd_scan_fastx built over it.  ``seq``/``seq_off``
    are what the classifier consumes; ``hdr``/``plus``/``qual`` are [begin,end) pairs into ``buf``."""

    def __init__(self, fmt, buf, n, hdr, plus, qual, seq, seq_off, lo=0, parent=None, on_release=None):
        self.format, self.buf, self.n = fmt, buf, n
        self.hdr, self.plus, self.qual, self.seq, self.seq_off = hdr, plus, qual, seq, seq_off
        self.lo = lo
        self._parent, self._on_release, self._released = parent, on_release, 0

    def view(self, lo, hi):
        """Records [lo, hi) of this chunk, sharing its buffers."""
        return RecordChunk(self.format, self.buf, hi - lo, self.hdr[2 * lo:2 * hi],
                           None if self.plus is None else self.plus[2 * lo:2 * hi],
                           None if self.qual is None else self.qual[2 * lo:2 * hi],
                           self.seq, self.seq_off[lo:hi + 1], self.lo + lo, parent=self._parent or self)

    def release(self):
        """Hand the chunk's buffers back to its reader once every record (all views) is done with."""
        root = self._parent or self
        root._released += self.n
        if root._released >= root.n and root._on_release is not None:
            cb, root._on_release = root._on_release, None
            cb()

    def records(self):
        """The reference's record tuples (header, seq[, plus, qual]) as str — for tests / small inputs."""
        b = self.buf.tobytes()
        s = self.seq.tobytes()
        out = []
        for i in range(self.n):
            h = b[self.hdr[2 * i]:self.hdr[2 * i + 1]].decode("latin-1")
            q = s[self.seq_off[i]:self.seq_off[i + 1]].decode("latin-1")
            if self.format == "fastq":
                out.append((h, q, b[self.plus[2 * i]:self.plus[2 * i + 1]].decode("latin-1"),
                            b[self.qual[2 * i]:self.qual[2 * i + 1]].decode("latin-1")))
            else:
                out.append((h, q))
        return out


def warn_if_truncated(tail, records_read):
    """The reference's parser silently drops an unfinished final record (fastx_parser.py:15-47, PROBE in SURVEY.md
    Appendix A); so does this reader, but it says so when the dropped bytes are more than white space."""
    t = np.asarray(tail)
    if t.size and (t > 0x20).any():
        import warnings
        warnings.warn("input ends inside a record: %d trailing bytes after record %d were dropped (truncated file?)"
                      % (t.size, records_read), RuntimeWarning, stacklevel=3)


class FastxReader:
    """Iterate RecordChunks of up to ``max_records`` records over a (optionally gzipped) FASTQ/FASTA
    file, holding one block of at most ``block_bytes`` of text at a time (bounded memory: the
    reference's ``get_seq_chunks``, seq_encoder.py:75-87)."""

    def __init__(self, path, max_records=1 << 22, block_bytes=1 << 28, threads=4, pinned=False):
        fmt = get_seq_format(path)
        self.format = "fasta" if fmt.startswith("fa") else "fastq"
        self.fh = open_text(path, fmt.endswith("gz"), threads)
        self.plain = not fmt.endswith("gz")
        self.file_pos = 0
        self.pool = None
        self.max_records = int(max_records)
        self.block_bytes = int(block_bytes)
        self.threads = int(threads)
        self.pinned = bool(pinned)
        self.tail = np.zeros(0, np.uint8)
        self.eof = False
        self.lib = _lib.load_library()
        self.records_read = 0
        self.bytes_scanned = 0
        self.rec_bytes = 512.0                 # running mean of the text bytes per record (first guess)
        self.peak_buffer_bytes = 0
        self.wait_seconds = 0.0
        # buffer sets are recycled (RecordChunk.release): fresh 256 MB allocations per chunk cost more in
        # page faults than the scan itself
        self.free = queue.Queue()

    def close(self):
        if self.fh is not None:
            self.fh.close()
            self.fh = None
        if self.pool is not None:
            self.pool.shutdown()
            self.pool = None

    def _fill_parallel(self, buf, start):
        """Plain files: positional reads of the block's slices on a few threads (page-cache copies scale
        with threads; a single read() tops out near 2.5 GB/s)."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        if self.pool is None:
            self.pool = ThreadPoolExecutor(max(1, min(self.threads, 8)))
        fd = self.fh.fileno()
        want = buf.size - start
        k = max(1, min(self.threads, 8)) if want >= (1 << 24) else 1
        step = max(1, -(-want // k))
        mv = memoryview(buf)

        def rd(i):
            lo, got = i * step, 0
            hi = min(want, lo + step)
            while lo + got < hi:
                r = os.preadv(fd, [mv[start + lo + got:start + hi]], self.file_pos + lo + got)
                if r == 0:
                    break
                got += r
            return got, hi - lo

        total = 0
        for got, span in self.pool.map(rd, range(k)):
            total += got
            if got < span:
                self.eof = True
                break
        self.file_pos += total
        return start + total

    def _fill(self, buf, start):
        if self.plain:                       # (never mixed with read(): the positional reads keep their own offset)
            return self._fill_parallel(buf, start)
        pos = start
        mv = memoryview(buf)
        while pos < buf.size:
            k = self.fh.readinto(mv[pos:])
            if not k:
                self.eof = True
                break
            pos += k
        return pos

    def __iter__(self):
        return self

    def __next__(self):
        while True:
            if self.eof and self.tail.size == 0:
                raise StopIteration
            cap = self.max_records
            # Read only what `cap` records are expected to need (running mean of the record size, 25 % head room):
            # the bytes carried over to the next call then stay a fraction of a chunk, so the carry copy is linear
            # in the file size and the buffers never exceed block_bytes (a single record larger than the block
            # still grows it).  An underestimate costs one rescan with a doubled estimate.
            want = int(cap * self.rec_bytes * 1.25) + (1 << 16)
            size = min(self.block_bytes, max(want, self.tail.size + (1 << 16)))
            size = max(size, self.tail.size)
            try:
                bs = self.free.get_nowait()                   # a recycled set (RecordChunk.release), else a fresh one:
            except queue.Empty:                               # callers that never release just allocate per chunk
                bs = {}
            if bs.get("size", 0) < size or bs.get("cap", 0) < cap:
                alloc = max(size, bs.get("size", 0))
                bs.clear()
                bs.update(size=alloc, cap=cap, buf=np.empty(alloc, np.uint8), seq=_host_array(alloc + 1, np.uint8, self.pinned),
                          hdr=np.empty(2 * cap, np.int64), seq_off=_host_array(cap + 1, np.int64, self.pinned),
                          plus=np.empty(2 * cap, np.int64) if self.format == "fastq" else None,
                          qual=np.empty(2 * cap, np.int64) if self.format == "fastq" else None)
            buf, seq, hdr, plus, qual, seq_off = bs["buf"], bs["seq"], bs["hdr"], bs["plus"], bs["qual"], bs["seq_off"]
            buf[:self.tail.size] = self.tail
            fill = self._fill(buf[:size], self.tail.size) if not self.eof else self.tail.size
            self.peak_buffer_bytes = max(self.peak_buffer_bytes, int(bs["size"]))
            consumed = ctypes.c_int64(0)
            n = self.lib.rd_scan_fastx(_p(buf), fill, _lib.FMT[self.format], int(self.eof), cap, _p(hdr), _p(plus),
                                       _p(qual), _p(seq), fill, _p(seq_off), ctypes.byref(consumed), self.threads)
            if n < 0:
                self.free.put(bs)
                msg = self.lib.rd_fastx_last_error().decode("utf-8", "replace")
                raise ValueError("%s (record %d of the file)" % (msg, self.records_read))
            c = consumed.value
            if n < cap and not self.eof and size < self.block_bytes:
                # the estimate was too small for `cap` records: keep everything, read more, scan again
                self.tail = buf[:fill].copy()
                self.rec_bytes = max(2.0 * self.rec_bytes, (c / n) if n else 0.0)
                self.free.put(bs)
                continue
            self.tail = buf[c:fill].copy()
            if n == 0:
                self.free.put(bs)
                if self.eof:
                    warn_if_truncated(self.tail, self.records_read)
                    self.tail = np.zeros(0, np.uint8)      # truncated final record: dropped like the reference
                    raise StopIteration
                if c == 0:
                    self.block_bytes *= 2                    # one record larger than the block: grow and retry
                continue
            self.records_read += n
            self.bytes_scanned += c
            self.rec_bytes = max(16.0, self.bytes_scanned / self.records_read)
            return RecordChunk(self.format, buf, int(n), hdr[:2 * n], None if plus is None else plus[:2 * n],
                               None if qual is None else qual[:2 * n], seq, seq_off[:n + 1],
                               on_release=lambda bs=bs: self.free.put(bs))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def partition_records(chunk, labels, want=(True, True, True), threads=4, scratch=None):
    """labels int8[n] in {0, 1, -1} → [non-rRNA bytes, rRNA bytes, unclassified bytes] (uint8 arrays,
    None where not wanted) plus the three byte counts; record text and order as the reference.
    `scratch` (a dict the caller keeps) lets the output buffers be reused between calls."""
    lib = _lib.load_library()
    labels = np.ascontiguousarray(labels, dtype=np.int8)
    if labels.size != chunk.n:
        raise ValueError("labels/records mismatch")
    sizes = np.zeros(3, np.int64)
    args = (_p(chunk.buf), _lib.FMT[chunk.format], chunk.n, _p(chunk.hdr), _p(chunk.plus), _p(chunk.qual),
            _p(chunk.seq), _p(chunk.seq_off), _p(labels))
    rc = lib.rd_partition_records(*args, None, None, None, _p(sizes), int(threads))
    if rc:
        raise ValueError(lib.rd_fastx_last_error().decode())
    outs = []
    for c in range(3):
        if not (want[c] and sizes[c]):
            outs.append(None)
        elif scratch is None:
            outs.append(np.empty(int(sizes[c]), np.uint8))
        else:
            if scratch.get(c) is None or scratch[c].size < sizes[c]:
                scratch[c] = np.empty(int(sizes[c] * 1.25) + 4096, np.uint8)
            outs.append(scratch[c][:int(sizes[c])])
    if any(o is not None for o in outs):
        rc = lib.rd_partition_records(*args, _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(sizes), int(threads))
        if rc:
            raise ValueError(lib.rd_fastx_last_error().decode())
    return outs, sizes
