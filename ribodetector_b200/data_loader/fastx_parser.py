"""``seq_parser(seq_fh, seq_type)`` with the reference's signature and record tuples
(``data_loader/fastx_parser.py:15-55``), driven by ``rd_scan_fastx`` instead of a per-line Python state
machine.  `seq_fh` is a text or binary file handle; `seq_type` is 'fastq' or anything else (= fasta),
like the reference."""
import ctypes

import numpy as np

from .. import _lib
from .fastx import RecordChunk, _p


def seq_parser(seq_fh, seq_type):
    data = seq_fh.read()
    if isinstance(data, str):
        data = data.encode("latin-1")
    fmt = "fastq" if seq_type == "fastq" else "fasta"
    buf = np.frombuffer(data, np.uint8)
    lib = _lib.load_library()
    pos = 0
    cap = 1 << 16
    while pos < buf.size:
        view = buf[pos:]
        hdr = np.empty(2 * cap, np.int64)
        plus = np.empty(2 * cap, np.int64) if fmt == "fastq" else None
        qual = np.empty(2 * cap, np.int64) if fmt == "fastq" else None
        seq = np.empty(view.size + 1, np.uint8)
        seq_off = np.empty(cap + 1, np.int64)
        consumed = ctypes.c_int64(0)
        n = lib.rd_scan_fastx(_p(view), view.size, _lib.FMT[fmt], 1, cap, _p(hdr), _p(plus), _p(qual), _p(seq),
                              view.size, _p(seq_off), ctypes.byref(consumed), 1)
        if n < 0:
            msg = lib.rd_fastx_last_error().decode("utf-8", "replace")
            raise IndexError(msg) if "blank line" in msg else ValueError(msg)
        if n == 0:
            return
        yield from RecordChunk(fmt, view, int(n), hdr, plus, qual, seq, seq_off).records()
        pos += consumed.value
