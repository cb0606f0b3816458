"""One-hot encoders, restated (test infrastructure; see oracle/__init__.py).

Reference: ``ribodetector/data_loader/seq_encoder.py``
  * table  ``BASE_DICT`` / ``ZERO_LIST``  :11-18   A,C,G,T → unit rows, U ≡ T, anything else
    (N, IUPAC, lower-case, '-') → all-zero row
  * ``encode_read``               :126-127  (caller slices ``seq[:max_len]``, detect.py:682)
  * ``encode_variable_len_read``  :130-145  first ``max_len`` bases, zero rows appended up to
    exactly ``max_len``
"""
import numpy as np

# byte → code; 0..3 = A,C,G,T/U ; 4 = zero row
CODE_LUT = np.full(256, 4, dtype=np.uint8)
for _b, _c in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3), (b"U", 3)):
    CODE_LUT[_b[0]] = _c

_ROWS = np.zeros((5, 4), dtype=np.float32)
_ROWS[0, 0] = _ROWS[1, 1] = _ROWS[2, 2] = _ROWS[3, 3] = 1.0


def as_bytes(read):
    if isinstance(read, str):
        read = read.encode("latin-1")
    return np.frombuffer(bytes(read), dtype=np.uint8)


def base_codes(read):
    """uint8 code per base (0..3, 4 = zero row)."""
    return CODE_LUT[as_bytes(read)]


def encode_read(read):
    """float32 [len(read), 4]  — seq_encoder.py:126-127."""
    return _ROWS[base_codes(read)]


def encode_variable_len_read(read, max_len=100):
    """float32 [max_len, 4] — truncate to the first max_len bases or right-pad with zero
    rows (seq_encoder.py:130-145)."""
    out = np.zeros((max_len, 4), dtype=np.float32)
    c = base_codes(read)[:max_len]
    out[: len(c)] = _ROWS[c]
    return out


def flatten_reads(reads):
    """list of str/bytes → (uint8 concatenation, int64 offsets[n+1]) — the layout the C ABI
    takes (include/rd_b200.h)."""
    bs = [r.encode("latin-1") if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), off
