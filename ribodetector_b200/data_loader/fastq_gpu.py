"""FASTQ or FASTA file(s) → label-partitioned output files with the record scan and the partition ON THE DEVICE
(``rd_fastq_submit`` / ``rd_fastq_collect``, csrc/rd_fastq_dev.cu).

The reference parses records one Python string at a time (``fastx_parser.py:15-47``), joins and routes
them in the main process (``detect.py:680,601-663``) and writes per batch (``detect.py:295-298``).
Here the host only moves bytes: a block of file text is read into page-locked memory, copied to the
GPU, indexed (K0), classified (K1-K3), partitioned by label (K4) and copied back as three contiguous
byte ranges that go to the output files with one ``write`` each.  A producer thread (read + submit) and
a consumer thread (collect + write) keep ``2 x n_devices`` blocks in flight, in file order.
"""
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .fastx import _host_array, get_seq_format, open_text, warn_if_truncated


class _Unit:
    """One (device, slot) pair with its page-locked input/output blocks."""

    def __init__(self, dev, slot, ends, block_bytes):
        self.dev, self.slot, self.ends, self.block_bytes = dev, slot, ends, block_bytes
        self.inp = self.out = None

    def ensure(self):                            # page-locking is slow: only units that get a block pay for it
        if self.inp is None:
            self.inp = [_host_array(self.block_bytes + 16, np.uint8, True) for _ in range(self.ends)]
            self.out = [_host_array(self.block_bytes + 32, np.uint8, True) for _ in range(self.ends)]
        return self


class FastqGpuStream:
    """``FastqGpuStream(models, inputs, ...).run(sinks)``.

    models   one SeqModel per device (blocks go round-robin over them)
    inputs   1 or 2 FASTQ paths (plain or .gz — gz is inflated on the host into the block buffer)
    sinks    dict with lists of binary file objects per end: 'non' (required), 'rrna', 'unc' (or None)
    """

    def __init__(self, models, inputs, max_len, mode="none", semantics="packed", precision=None,
                 block_bytes=None, threads=8):
        self.models = list(models)
        self.inputs = list(inputs)
        self.ends = len(self.inputs)
        if self.ends not in (1, 2):
            raise ValueError("one or two input files")
        fmts = [get_seq_format(p) for p in self.inputs]
        if len({f[:2] for f in fmts}) != 1:
            raise ValueError("the input files must be all FASTQ or all FASTA")
        self.fasta = fmts[0].startswith("fa")
        self.plain = [not f.endswith("gz") for f in fmts]
        self.max_len, self.mode, self.semantics, self.precision = int(max_len), mode, semantics, precision
        if block_bytes is None:                  # 256 MB blocks (~1.2 M 100-bp reads), less when the input is small
            est = max(os.path.getsize(p) * (1 if plain else 8) for p, plain in zip(self.inputs, self.plain))
            block_bytes = min(1 << 28, max(1 << 20, -(-(est + 4096) // (1 << 20)) * (1 << 20)))
        self.block_bytes = int(block_bytes)
        self.max_records = max(1, self.block_bytes // (32 if self.fasta else 16))
        self.threads = max(1, min(int(threads), 16))
        self.num_seqs = 0
        self.counts = np.zeros(3, np.int64)
        self.stage_seconds = {"read": 0.0, "submit": 0.0, "collect": 0.0, "write": 0.0}
        self.setup_seconds = 0.0                 # page-locking the block buffers (once per unit)

    # ---- reading -------------------------------------------------------------------------------------
    def _fill(self, e, buf, start):
        """Fill buf[start:block_bytes] from input e; returns the new fill level (sets self.eof[e])."""
        want = self.block_bytes - start
        if want <= 0 or self.eof[e]:
            return start
        mv = memoryview(buf)
        if not self.plain[e]:
            pos = start
            while pos < self.block_bytes:
                k = self.fh[e].readinto(mv[pos:self.block_bytes])
                if not k:
                    self.eof[e] = True
                    break
                pos += k
            return pos
        fd = self.fh[e].fileno()
        k = self.threads if want >= (1 << 24) else 1
        step = -(-want // k)

        def rd(i):
            lo, got = i * step, 0
            hi = min(want, lo + step)
            while lo + got < hi:
                r = os.preadv(fd, [mv[start + lo + got:start + hi]], self.file_pos[e] + lo + got)
                if r == 0:
                    break
                got += r
            return got, hi - lo

        total = 0
        for got, span in self.pool.map(rd, range(k)):
            total += got
            if got < span:
                self.eof[e] = True
                break
        self.file_pos[e] += total
        return start + total

    # ---- the pipeline ----------------------------------------------------------------------------------
    def run(self, sinks):
        ends = self.ends
        self.fh = [open_text(p, not plain, self.threads) for p, plain in zip(self.inputs, self.plain)]
        self.file_pos = [0] * ends
        self.eof = [False] * ends
        self.pool = ThreadPoolExecutor(self.threads)
        self.wpool = ThreadPoolExecutor(2)
        self.rpool = ThreadPoolExecutor(2)
        units = [_Unit(d, s, ends, self.block_bytes) for s in range(2) for d in range(len(self.models))]
        free_units = queue.Queue()
        for u in units:
            free_units.put(u)
        inflight = queue.Queue()
        errors = []
        busy = self.stage_seconds

        def consume():
            try:
                while True:
                    item = inflight.get()
                    if item is None:
                        return
                    u, n = item
                    t0 = time.perf_counter()
                    sizes, counts = self.models[u.dev].fastq_collect(u.slot)
                    t1 = time.perf_counter()
                    busy["collect"] += t1 - t0
                    def put(e):
                        s0, s1, s2 = (int(x) for x in sizes[e])
                        mv = memoryview(u.out[e])
                        if s0:
                            sinks["non"][e].write(mv[:s0])
                        if s1 and sinks.get("rrna"):
                            sinks["rrna"][e].write(mv[s0:s0 + s1])
                        if s2 and sinks.get("unc"):
                            sinks["unc"][e].write(mv[s0 + s1:s0 + s1 + s2])

                    if ends == 2:                # the two ends go to different files: write them side by side
                        list(self.wpool.map(put, range(2)))
                    else:
                        put(0)
                    busy["write"] += time.perf_counter() - t1
                    self.counts += counts
                    self.num_seqs += n
                    free_units.put(u)
            except BaseException as ex:          # noqa: BLE001 — surfaced on the producer thread
                errors.append(ex)
                free_units.put(None)

        t_out = threading.Thread(target=consume, daemon=True)
        t_out.start()
        tails = [np.zeros(0, np.uint8)] * ends
        try:
            while not errors:
                u = free_units.get()
                if u is None:
                    break
                t0 = time.perf_counter()
                u.ensure()
                self.setup_seconds += time.perf_counter() - t0
                t0 = time.perf_counter()
                def fill(e):
                    k = tails[e].size
                    u.inp[e][:k] = tails[e]
                    return self._fill(e, u.inp[e], k)

                # (two ends: both files are read / inflated side by side — zlib releases the GIL)
                fills = list(self.rpool.map(fill, range(ends))) if ends == 2 else [fill(0)]
                final = all(self.eof)
                t1 = time.perf_counter()
                busy["read"] += t1 - t0
                n, consumed, _ = self.models[u.dev].fastq_submit(
                    u.slot, u.inp, fills, final, self.max_records, self.max_len, u.out, mode=self.mode,
                    semantics=self.semantics, precision=self.precision, fasta=self.fasta)
                busy["submit"] += time.perf_counter() - t1
                tails = [u.inp[e][consumed[e]:fills[e]].copy() for e in range(ends)]
                if n:
                    inflight.put((u, n))
                else:
                    free_units.put(u)
                    if not final:             # every buffer is full or at end of file, and no record came out
                        if ends == 2 and any(self.eof):
                            raise RuntimeError("The two input files hold different numbers of reads.")
                        raise RuntimeError("a record does not fit the %d-byte block" % self.block_bytes)
                if final and (n == 0 or not any(t.size for t in tails)):
                    # (a capped block leaves whole records behind: they go round again; a truncated last record is dropped)
                    # one end holding a further COMPLETE record (four lines; the last one may lack its newline) means the
                    # files differ in length, as the host path (_pair_chunks) reports
                    def lines(t):
                        return int((t == 10).sum()) + int(t.size > 0 and t[-1] != 10)

                    def extra_record(t):              # FASTA: a header left over; FASTQ: four lines left over
                        return bool((t == 62).any()) if self.fasta else lines(t) >= 4
                    if ends == 2 and any(extra_record(t) for t in tails):
                        raise RuntimeError("The two input files hold different numbers of reads.")
                    if not self.fasta:
                        for t in tails:
                            warn_if_truncated(t, int(self.counts.sum()))
                    break
        finally:
            inflight.put(None)
            t_out.join()
            for m in self.models:                 # nothing may stay queued on a slot when buffers are dropped
                for s in range(2):
                    try:
                        m.fastq_collect(s)
                    except Exception:               # noqa: BLE001 — slot was not pending
                        pass
            for fh in self.fh:
                fh.close()
            self.pool.shutdown()
            self.wpool.shutdown()
            self.rpool.shutdown()
        if errors:
            raise errors[0]
        return self.counts
