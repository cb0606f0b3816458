"""Deterministic synthetic read generator: numpy PCG64, bases i.i.d. uniform
over ACGT with 0.1 % of positions replaced by 'N'.  There is no network, so every benchmark
and parity test runs on reads made here."""
import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
SEED_BASE = 20261017


def synth_reads_fixed(n, length, seed, n_frac=0.001):
    """n reads of exactly `length` bases → (uint8 [n*length], int64 offsets [n+1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = _ALPHA[rng.integers(0, 4, size=n * length, dtype=np.uint8)]
    if n_frac > 0 and seq.size:
        k = rng.binomial(seq.size, n_frac)
        seq[rng.integers(0, seq.size, size=k)] = ord("N")
    off = np.arange(n + 1, dtype=np.int64) * length
    return seq, off


_LUT4 = None


def _fill_block(args):
    out, seed, i, s, e, n_frac = args
    global _LUT4
    if _LUT4 is None:            # byte -> the four bases its 2-bit fields select, packed little-endian
        r = np.arange(256, dtype=np.uint32)
        _LUT4 = sum(_ALPHA[(r >> (2 * k)) & 3].astype(np.uint32) << (8 * k) for k in range(4)).astype(np.uint32)
    rng = np.random.Generator(np.random.PCG64([seed, i]))
    m = e - s
    raw = np.frombuffer(rng.bytes((m + 3) // 4), dtype=np.uint8)
    out[s:e] = _LUT4[raw].view(np.uint8)[:m]
    if n_frac > 0:
        k = rng.binomial(m, n_frac)
        out[s + rng.integers(0, m, size=k)] = ord("N")


def fill_bases(out, seed, n_frac=0.001, block=1 << 26, threads=8):
    """Fill a uint8 array with i.i.d. ACGT (+ n_frac 'N'), 64-MB block by block from (seed, block index): the
    generator for buffers of several GB (a quarter byte of randomness per base, blocks on a thread pool)."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = [(out, seed, i, s, min(out.size, s + block), n_frac) for i, s in enumerate(range(0, out.size, block))]
    with ThreadPoolExecutor(max(1, min(threads, len(jobs)))) as ex:
        list(ex.map(_fill_block, jobs))
    return out


def synth_reads(n, min_len, max_len, seed, n_frac=0.001):
    """n reads with length ~ U{min_len..max_len} → (uint8 bytes, int64 offsets [n+1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(min_len, max_len + 1, size=n, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    seq = _ALPHA[rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)]
    if n_frac > 0 and seq.size:
        k = rng.binomial(seq.size, n_frac)
        seq[rng.integers(0, seq.size, size=k)] = ord("N")
    return seq, off


def to_strings(seq, off):
    b = seq.tobytes()
    return [b[off[i]:off[i + 1]].decode("latin-1") for i in range(len(off) - 1)]


def fastq_text(n, length, seed, n_frac=0.001):
    """FASTQ text of n fixed-length reads: "@r%09d" / seq / "+" / 'I' x length, '\n' line ends → uint8 array."""
    seq, _ = synth_reads_fixed(n, length, seed, n_frac)
    w = 2 + 9 + 1 + length + 1 + 2 + length + 1
    rec = np.empty((n, w), np.uint8)
    rec[:, 0], rec[:, 1] = ord("@"), ord("r")
    idx = np.arange(n)
    for d in range(9):
        rec[:, 2 + 8 - d] = ord("0") + (idx // 10 ** d) % 10
    rec[:, 11] = ord("\n")
    rec[:, 12:12 + length] = seq.reshape(n, length)
    rec[:, 12 + length] = ord("\n")
    rec[:, 13 + length], rec[:, 14 + length] = ord("+"), ord("\n")
    rec[:, 15 + length:15 + 2 * length] = ord("I")
    rec[:, 15 + 2 * length] = ord("\n")
    return rec.reshape(-1)


def synth_weights(hidden_size, seed, gain=3.5):
    """A seeded random checkpoint of the reference architecture at another hidden size (state_dict keys of
    ``model/model.py:16-24``): U(-g/sqrt(H), g/sqrt(H)) like ``nn.LSTM``'s default init, times `gain` so that gates
    leave the linear range.  Used by oracle/gen_golden_arch.py and by the tests of the other hidden sizes — the
    reference ships one checkpoint (H = 128) only."""
    rng = np.random.Generator(np.random.PCG64([seed, hidden_size]))
    H = hidden_size
    a = gain / np.sqrt(H)
    shapes = (("rnn.weight_ih_l0", (4 * H, 4)), ("rnn.weight_hh_l0", (4 * H, H)),
              ("rnn.bias_ih_l0", (4 * H,)), ("rnn.bias_hh_l0", (4 * H,)),
              ("rnn.weight_ih_l0_reverse", (4 * H, 4)), ("rnn.weight_hh_l0_reverse", (4 * H, H)),
              ("rnn.bias_ih_l0_reverse", (4 * H,)), ("rnn.bias_hh_l0_reverse", (4 * H,)),
              ("out.weight", (2, 2 * H)), ("out.bias", (2,)))
    return {k: rng.uniform(-a, a, size=s).astype(np.float32) for k, s in shapes}
