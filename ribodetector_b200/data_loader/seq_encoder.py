"""The reference's ``data_loader/seq_encoder.py`` surface (same names, argument meaning and errors) on
the B200 path.

* ``BASE_DICT`` / ``ZERO_LIST``            seq_encoder.py:11-18 — the table the CUDA encoders implement
* ``get_seq_format``                       seq_encoder.py:21-39
* ``load_reads`` / ``get_seq_chunks`` / ``get_pairedread_chunks``   seq_encoder.py:56-92 — same record
  tuples, parsed by ``rd_scan_fastx`` (C ABI) instead of the Python line state machine
* ``encode_read`` / ``encode_variable_len_read``   seq_encoder.py:126-145 — same values, produced ON THE
  GPU by ``rd_encode_onehot`` and returned as CUDA tensors (there is no CPU encoder in this package:
  without a visible B200 these raise).  The classify path never calls them — it consumes sequence
  bytes directly — they exist so code written against the reference keeps working and for parity tests.

The training-only helpers (``load_seqs``, ``get_read_rc_with_maxlen``, ``get_read_with_maxlen``,
``encode_seq_reads``; they need Bio.Seq) are out of scope.
"""
from itertools import islice

import numpy as np

from .fastx import FastxReader, get_seq_format  # noqa: F401

BASE_DICT = {'A': (1, 0, 0, 0), 'C': (0, 1, 0, 0), 'G': (0, 0, 1, 0), 'T': (0, 0, 0, 1), 'U': (0, 0, 0, 1)}
ZERO_LIST = (0, 0, 0, 0)

_ENCODER = None


def _encoder():
    """One library handle for the stand-alone encoder calls (device 0, shipped checkpoint)."""
    global _ENCODER
    if _ENCODER is None:
        from ..model import SeqModel
        from ..utils.weights import load_weights
        m = SeqModel()
        m.load_state_dict(load_weights())
        m.to("cuda")                      # raises without a GPU: no CPU fallback
        _ENCODER = m
    return _ENCODER


def _bytes(seq):
    return np.frombuffer(seq.encode("latin-1") if isinstance(seq, str) else bytes(seq), np.uint8).copy()


def encode_read(read):
    """one-hot rows of every base of `read` → CUDA float32 [len(read), 4]  (seq_encoder.py:126-127)."""
    b = _bytes(read)
    if b.size == 0:
        import torch
        return torch.zeros((0, 4), dtype=torch.float32, device=_encoder().device)
    rows, _ = _encoder().encode_onehot(b, np.array([0, b.size], np.int64), max(1, b.size), "ragged")
    return rows


def encode_variable_len_read(read, max_len=100):
    """first `max_len` bases, zero rows appended → CUDA float32 [max_len, 4]  (seq_encoder.py:130-145)."""
    b = _bytes(read)
    return _encoder().encode_onehot(b, np.array([0, b.size], np.int64), int(max_len), "padded")[0]


def load_reads(seq_file, label=None, max_len=100):
    """Whole file → list of record tuples (header, seq[, '+' line, qual]) like seq_encoder.py:56-65."""
    if label is not None:
        raise NotImplementedError("labelled loading is a training helper of the reference (out of scope)")
    out = []
    with FastxReader(seq_file) as reader:
        for chunk in reader:
            out.extend(chunk.records())
            chunk.release()
    return out


def _records(seq_file):
    with FastxReader(seq_file, max_records=1 << 18) as reader:      # closed at StopIteration and when abandoned
        for chunk in reader:
            yield from chunk.records()
            chunk.release()


def get_seq_chunks(seq_file, chunk_size=1048576):
    """Lists of up to `chunk_size` record tuples (seq_encoder.py:75-87)."""
    it = _records(seq_file)
    while True:
        chunk = list(islice(it, chunk_size))
        if not chunk:
            return
        yield chunk


def get_pairedread_chunks(r1_seq_file, r2_seq_file, chunk_size=1048576):
    """seq_encoder.py:90-92"""
    for r1_chunk, r2_chunk in zip(get_seq_chunks(r1_seq_file, chunk_size), get_seq_chunks(r2_seq_file, chunk_size)):
        yield r1_chunk, r2_chunk
