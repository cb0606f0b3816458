"""``SeqModel`` — the reference's plugin object for the hot path, backed by librd_b200.so.

Mirrors ``ribodetector/model/model.py:10-37`` (and ``model_cpu.py:8-37`` when
``pack_seq=False``): same constructor keywords, ``load_state_dict`` takes the reference
checkpoint keys, ``.to(device)`` / ``.eval()`` / ``__call__`` behave as the batch loops expect
(``detect.py:93,115-119,185-188,286-287``), and ``config.json``'s ``arch.type = "SeqModel"``
resolves to this class through ``ConfigParser.init_obj`` (``parse_config.py:43-57``).

PyTorch is used only for device memory and streams; every computation is a CUDA kernel of
librd_b200.so called through the C ABI (``include/rd_b200.h``).  There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch
from torch.nn.utils.rnn import PackedSequence, pad_packed_sequence

from .. import _lib
from ..utils.weights import STATE_KEYS

_ARG_ORDER = (
    "rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0",
    "rnn.weight_ih_l0_reverse", "rnn.weight_hh_l0_reverse",
    "rnn.bias_ih_l0_reverse", "rnn.bias_hh_l0_reverse",
    "out.weight", "out.bias",
)


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return ctypes.c_void_p(t.ctypes.data)
    return ctypes.c_void_p(t.data_ptr())


class ReadBatch:
    """What the collate functions hand to the model in place of the reference's PackedSequence of one-hot
    rows (detect.py:685): the sequence bytes of a batch, already cut to `max_len`, plus offsets.
    ``.to(device)`` mirrors the reference loop's ``data.to(self.device, non_blocking=True)``."""

    def __init__(self, seq, off, max_len, pack_seq=True):
        self.seq, self.off, self.max_len, self.pack_seq = torch.as_tensor(seq), torch.as_tensor(off), int(max_len), bool(pack_seq)

    def to(self, device, non_blocking=False):
        return ReadBatch(self.seq.to(device, non_blocking=non_blocking), self.off.to(device, non_blocking=non_blocking),
                         self.max_len, self.pack_seq)

    def pin_memory(self):
        return ReadBatch(self.seq.pin_memory(), self.off.pin_memory(), self.max_len, self.pack_seq)

    def __len__(self):
        return self.off.numel() - 1


class SeqModel:
    """BiLSTM(4→H) + Linear(2H→2) classifier.  ``model(x)`` returns raw logits ``[B, 2]``.  H = 128 (the shipped
    checkpoint, ``config.json``) runs on the tensor-core kernels in the chosen ``precision``; any other multiple of 32 up to
    256 — ``SeqModel(**arch.args)`` in the reference takes any ``hidden_size``, ``model/model.py:11-29`` — runs on the
    fp32 CUDA-core kernel whatever the ``precision``."""

    def __init__(self, input_size=4, hidden_size=128, num_layers=1, num_classes=2,
                 batch_first=True, bidirectional=True, pack_seq=True, precision="tc_mixed"):
        if (input_size, num_layers, num_classes, batch_first, bidirectional) != (4, 1, 2, True, True):
            raise ValueError("SeqModel kernels implement input_size=4, num_layers=1, num_classes=2, "
                             "batch_first=True, bidirectional=True (the shipped architecture)")
        if hidden_size % 32 != 0 or not 32 <= hidden_size <= 256:
            raise ValueError("SeqModel kernels take hidden_size = a multiple of 32 between 32 and 256 (128, the shipped "
                             "size, runs on the tensor-core kernels; any other on the fp32 CUDA-core kernel)")
        if precision not in _lib.PREC:
            raise ValueError("precision must be one of %s" % sorted(_lib.PREC))
        self.hidden_size = hidden_size
        self.pack_seq = bool(pack_seq)
        self.precision = precision
        self.training = False
        self._weights = None
        self._handle = None
        self._device = None
        self._alphabet = None
        self._lib = _lib.load_library()          # raises if the extension is missing

    # ---- nn.Module-like surface used by the reference loop -----------------------------------
    def load_state_dict(self, state_dict, strict=True):
        sd = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith("module.") else k          # DataParallel checkpoints
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            sd[k] = np.ascontiguousarray(v, dtype=np.float32)
        missing = [k for k in STATE_KEYS if k not in sd]
        unexpected = [k for k in sd if k not in STATE_KEYS]
        if missing or (strict and unexpected):
            raise RuntimeError("Error(s) in loading state_dict for SeqModel: missing %s unexpected %s"
                               % (missing, unexpected))
        H = self.hidden_size
        shapes = {"rnn.weight_ih_l0": (4 * H, 4), "rnn.weight_hh_l0": (4 * H, H),
                  "rnn.bias_ih_l0": (4 * H,), "rnn.bias_hh_l0": (4 * H,), "out.weight": (2, 2 * H),
                  "out.bias": (2,)}
        for k in STATE_KEYS:
            want = shapes[k.replace("_reverse", "")]
            if sd[k].shape != want:
                raise RuntimeError("size mismatch for %s: %s vs %s" % (k, sd[k].shape, want))
        self._weights = {k: sd[k] for k in STATE_KEYS}
        if self._device is not None:
            self._create(self._device)
        return self

    def state_dict(self):
        return {k: torch.from_numpy(v.copy()) for k, v in (self._weights or {}).items()}

    def eval(self):
        self.training = False
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def to(self, device, non_blocking=False):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("ribodetector_b200.SeqModel runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("No visible CUDA devices! Set CUDA_VISIBLE_DEVICES.")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._device = torch.device("cuda", index)
        if self._weights is not None:
            self._create(self._device)
        return self

    def _create(self, device):
        self.close()
        h = ctypes.c_void_p()
        args = [_ptr(self._weights[k]) for k in _ARG_ORDER]
        rc = self._lib.rd_create(device.index, *args, self.hidden_size, ctypes.byref(h))
        _lib.check(self._lib, None, rc, "rd_create")
        self._handle = h

    def close(self):
        if self._handle is not None:
            self._lib.rd_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device(self):
        return self._device

    def _need(self):
        if self._handle is None:
            raise RuntimeError("SeqModel: call load_state_dict(...) and .to('cuda') first")
        return self._handle

    def kernel_launches(self):
        return int(self._lib.rd_kernel_launches(self._need()))

    def set_timing(self, enable=True):
        _lib.check(self._lib, self._need(), self._lib.rd_set_timing(self._need(), int(enable)), "rd_set_timing")

    def get_timing(self, reset=True):
        """→ {stage: (total_ms, launches)} for K1 plan, K2 lstm, K3 tail, pair (CUDA events)."""
        ms = (ctypes.c_double * 4)()
        cnt = (ctypes.c_int64 * 4)()
        rc = self._lib.rd_get_timing(self._need(), ctypes.cast(ms, ctypes.c_void_p),
                                     ctypes.cast(cnt, ctypes.c_void_p), int(reset))
        _lib.check(self._lib, self._handle, rc, "rd_get_timing")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(("plan", "lstm", "tail", "pair"))}

    # ---- device-resident entry points -----------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _dev_inputs(self, seq, off):
        seq = torch.as_tensor(seq)
        off = torch.as_tensor(off)
        if seq.dtype != torch.uint8 or off.dtype != torch.int64:
            raise ValueError("seq must be uint8 and offsets int64")
        return (seq.to(self._device, non_blocking=True).contiguous(),
                off.to(self._device, non_blocking=True).contiguous())

    def classify(self, seq, off, max_len, semantics=None, precision=None,
                 want_probs=False, want_labels=True, counts=None):
        """Device path: seq uint8[total], off int64[n+1] (moved to the device if needed).
        Returns (logits[n,2] f32, probs[n,2] f32 | None, labels[n] int8 | None); all device
        tensors, asynchronous on the current stream.  Replaces detect.py:284-288."""
        h = self._need()
        semantics = semantics or ("packed" if self.pack_seq else "padded")
        precision = precision or self.precision
        seq, off = self._dev_inputs(seq, off)
        n = off.numel() - 1
        with torch.cuda.device(self._device):
            logits = torch.empty((n, 2), dtype=torch.float32, device=self._device)
            probs = torch.empty((n, 2), dtype=torch.float32, device=self._device) if want_probs else None
            labels = torch.empty((n,), dtype=torch.int8, device=self._device) if want_labels else None
            rc = self._lib.rd_classify(h, _ptr(seq), _ptr(off), n, int(max_len), _lib.SEM[semantics],
                                       _lib.PREC[precision], _ptr(logits), _ptr(probs), _ptr(labels),
                                       _ptr(counts), self._stream())
        _lib.check(self._lib, h, rc, "rd_classify")
        return logits, probs, labels

    def pair_combine(self, logits1, logits2, mode="none", counts=None):
        """detect.py:616-663 on device logits → labels int8[n] in {-1,0,1}."""
        h = self._need()
        if mode not in _lib.PAIR:
            raise ValueError("ensure mode must be one of %s" % sorted(_lib.PAIR))
        l1 = logits1.to(self._device, torch.float32).contiguous()
        l2 = logits2.to(self._device, torch.float32).contiguous()
        if l1.shape != l2.shape or l1.dim() != 2 or l1.shape[1] != 2:
            raise ValueError("logits must both be [n, 2]")
        n = l1.shape[0]
        with torch.cuda.device(self._device):
            labels = torch.empty((n,), dtype=torch.int8, device=self._device)
            rc = self._lib.rd_pair_combine(h, _ptr(l1), _ptr(l2), n, _lib.PAIR[mode], _ptr(labels),
                                           _ptr(counts), self._stream())
        _lib.check(self._lib, h, rc, "rd_pair_combine")
        return labels

    def encode_onehot(self, seq, off, max_len, layout="padded"):
        """seq_encoder.py:126-145 on the device.  padded → [n, max_len, 4]; ragged → ([rows,4],
        row_off[n+1])."""
        h = self._need()
        seq, off = self._dev_inputs(seq, off)
        n = off.numel() - 1
        with torch.cuda.device(self._device):
            if layout == "padded":
                out = torch.empty((n, int(max_len), 4), dtype=torch.float32, device=self._device)
                row_off = None
            elif layout == "ragged":
                lens = torch.clamp(off[1:] - off[:-1], max=int(max_len))
                rows = int(lens.sum().item())
                out = torch.empty((rows, 4), dtype=torch.float32, device=self._device)
                row_off = torch.empty((n + 1,), dtype=torch.int64, device=self._device)
            else:
                raise ValueError("layout must be 'padded' or 'ragged'")
            rc = self._lib.rd_encode_onehot(h, _ptr(seq), _ptr(off), n, int(max_len), _lib.ONEHOT[layout],
                                            _ptr(out), _ptr(row_off), self._stream())
        _lib.check(self._lib, h, rc, "rd_encode_onehot")
        return out if row_off is None else (out, row_off)

    # ---- host-buffer entry points (the call a CLI / user makes) -------------------------------
    @staticmethod
    def _host(a, dtype):
        if isinstance(a, torch.Tensor):
            if a.is_cuda or a.dtype != dtype or not a.is_contiguous():
                raise ValueError("host tensors must be contiguous CPU tensors of dtype %s" % dtype)
            return a
        return torch.from_numpy(np.ascontiguousarray(a, dtype={torch.uint8: np.uint8, torch.int64: np.int64}[dtype]))

    def classify_host(self, seq, off, max_len, semantics=None, precision=None, want_logits=True,
                      want_probs=False, out=None):
        """HOST bytes in → HOST results out, synchronous: H2D, kernels and D2H are pipelined in
        chunks inside the library.  Returns dict(labels int8[n], counts int64[3], logits, probs)."""
        h = self._need()
        semantics = semantics or ("packed" if self.pack_seq else "padded")
        precision = precision or self.precision
        seq = self._host(seq, torch.uint8)
        off = self._host(off, torch.int64)
        n = off.numel() - 1
        out = out or {}
        labels = out.get("labels")
        if labels is None:
            labels = torch.empty((n,), dtype=torch.int8, pin_memory=True)
        logits = out.get("logits")
        if logits is None and want_logits:
            logits = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
        probs = out.get("probs")
        if probs is None and want_probs:
            probs = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
        counts = torch.zeros(3, dtype=torch.int64)
        rc = self._lib.rd_classify_host(h, _ptr(seq), _ptr(off), n, int(max_len), _lib.SEM[semantics],
                                        _lib.PREC[precision], _ptr(logits), _ptr(probs), _ptr(labels),
                                        _ptr(counts))
        _lib.check(self._lib, h, rc, "rd_classify_host")
        return {"labels": labels, "counts": counts, "logits": logits, "probs": probs}

    def classify_pairs_host(self, seq1, off1, seq2, off2, max_len, mode="none", semantics=None,
                            precision=None, want_logits=False, out=None):
        h = self._need()
        if mode not in _lib.PAIR:
            raise ValueError("ensure mode must be one of %s" % sorted(_lib.PAIR))
        semantics = semantics or ("packed" if self.pack_seq else "padded")
        precision = precision or self.precision
        seq1, seq2 = self._host(seq1, torch.uint8), self._host(seq2, torch.uint8)
        off1, off2 = self._host(off1, torch.int64), self._host(off2, torch.int64)
        n = off1.numel() - 1
        if off2.numel() - 1 != n:
            raise ValueError("R1 and R2 must hold the same number of reads")
        out = out or {}
        labels = out.get("labels")
        if labels is None:
            labels = torch.empty((n,), dtype=torch.int8, pin_memory=True)
        l1 = l2 = None
        if want_logits:
            l1 = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
            l2 = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
        counts = torch.zeros(3, dtype=torch.int64)
        rc = self._lib.rd_classify_pairs_host(h, _ptr(seq1), _ptr(off1), _ptr(seq2), _ptr(off2), n,
                                              int(max_len), _lib.SEM[semantics], _lib.PREC[precision],
                                              _lib.PAIR[mode], _ptr(l1), _ptr(l2), _ptr(labels), _ptr(counts))
        _lib.check(self._lib, h, rc, "rd_classify_pairs_host")
        return {"labels": labels, "counts": counts, "logits1": l1, "logits2": l2}

    # ---- FASTQ text on the device (K0 scan, K4 partition) ------------------------------------------
    def scan_fastq(self, text, final_chunk=True, max_records=None):
        """FASTQ text (uint8 tensor / bytes; moved to the device) → (d_text, rec int64[n, 8] device tensor of
        [begin, end) pairs for header / sequence / '+' / quality, n, consumed).  fastx_parser.py:15-47."""
        h = self._need()
        if isinstance(text, (bytes, bytearray, memoryview)):
            text = torch.frombuffer(bytearray(text), dtype=torch.uint8) if len(text) else torch.empty(0, dtype=torch.uint8)
        text = torch.as_tensor(text)
        if text.dtype != torch.uint8:
            raise ValueError("text must be uint8")
        d_text = text.to(self._device).contiguous()
        n_bytes = d_text.numel()
        cap = int(max_records) if max_records is not None else n_bytes // 8 + 1
        with torch.cuda.device(self._device):
            rec = torch.empty((max(cap, 1), 8), dtype=torch.int64, device=self._device)
            info = torch.empty(8, dtype=torch.int64, device=self._device)
            rc = self._lib.rd_scan_fastq_device(h, _ptr(d_text) if n_bytes else None, n_bytes, int(bool(final_chunk)), cap,
                                                _ptr(rec), _ptr(info), self._stream())
        _lib.check(self._lib, h, rc, "rd_scan_fastq_device")
        info = info.cpu().tolist()
        n = int(info[1])
        if info[4] >= 0 and info[4] // 4 < n:
            raise ValueError("FASTQ: %s in record %d" % ("blank line" if info[4] % 4 == 1 else "header without '@'", info[4] // 4))
        return d_text, rec[:n], n, int(info[2])

    def scan_fasta(self, text, final_chunk=True, max_records=None):
        """FASTA text → (d_buf, rec int64[n, 8], n, consumed): like scan_fastq; d_buf is the text followed by the region
        that holds the joined upper-cased sequences (rec[:, 2:4] index it).  fastx_parser.py:39-55."""
        h = self._need()
        if isinstance(text, (bytes, bytearray, memoryview)):
            text = torch.frombuffer(bytearray(text), dtype=torch.uint8) if len(text) else torch.empty(0, dtype=torch.uint8)
        text = torch.as_tensor(text)
        if text.dtype != torch.uint8:
            raise ValueError("text must be uint8")
        n_bytes = text.numel()
        seq_base = (n_bytes + 15) & ~15
        cap = int(max_records) if max_records is not None else n_bytes // 2 + 1
        with torch.cuda.device(self._device):
            d_buf = torch.empty(seq_base + n_bytes + 16, dtype=torch.uint8, device=self._device)
            d_buf[:n_bytes] = text.to(self._device)
            rec = torch.empty((max(cap, 1), 8), dtype=torch.int64, device=self._device)
            info = torch.empty(8, dtype=torch.int64, device=self._device)
            rc = self._lib.rd_scan_fasta_device(h, _ptr(d_buf), n_bytes, seq_base, int(bool(final_chunk)), cap,
                                                _ptr(rec), _ptr(info), self._stream())
        _lib.check(self._lib, h, rc, "rd_scan_fasta_device")
        info = info.cpu().tolist()
        if info[4] >= 0:
            raise ValueError("FASTA: line %d holds 4 MiB or more" % (info[4] // 4 + 1))
        n = int(info[1])
        return d_buf, rec[:n], n, int(info[2])

    def partition_fasta(self, d_buf, rec, labels):
        """partition_records for FASTA records ("header\\nSEQUENCE\\n")."""
        h = self._need()
        n = rec.shape[0]
        labels = torch.as_tensor(labels, dtype=torch.int8).to(self._device).contiguous()
        if labels.numel() != n:
            raise ValueError("labels/records mismatch")
        with torch.cuda.device(self._device):
            out = torch.empty(d_buf.numel() // 2 + 16, dtype=torch.uint8, device=self._device)
            sizes = torch.zeros(3, dtype=torch.int64, device=self._device)
            rc = self._lib.rd_partition_fasta_device(h, _ptr(d_buf) if n else None, _ptr(rec) if n else None, n,
                                                     _ptr(labels) if n else None, _ptr(out), _ptr(sizes), self._stream())
        _lib.check(self._lib, h, rc, "rd_partition_fasta_device")
        sizes = sizes.cpu()
        return out[:int(sizes.sum())], sizes

    def classify_records(self, d_text, rec, max_len, semantics=None, precision=None, counts=None):
        """rd_classify over the sequence lines of a record index → (logits[n,2], labels[n]) on the device."""
        h = self._need()
        semantics = semantics or ("packed" if self.pack_seq else "padded")
        precision = precision or self.precision
        n = rec.shape[0]
        with torch.cuda.device(self._device):
            logits = torch.empty((n, 2), dtype=torch.float32, device=self._device)
            labels = torch.empty((n,), dtype=torch.int8, device=self._device)
            rc = self._lib.rd_classify_records(h, _ptr(d_text), _ptr(rec), n, int(max_len), _lib.SEM[semantics],
                                               _lib.PREC[precision], _ptr(logits), None, _ptr(labels), _ptr(counts),
                                               self._stream())
        _lib.check(self._lib, h, rc, "rd_classify_records")
        return logits, labels

    def partition_records(self, d_text, rec, labels):
        """labels int8[n] in {0, 1, -1} → (out uint8 device tensor laid out [non-rRNA | rRNA | unclassified],
        sizes int64[3] on the host).  detect.py:680,601-663."""
        h = self._need()
        n = rec.shape[0]
        labels = torch.as_tensor(labels, dtype=torch.int8).to(self._device).contiguous()
        if labels.numel() != n:
            raise ValueError("labels/records mismatch")
        with torch.cuda.device(self._device):
            out = torch.empty(d_text.numel() + 1, dtype=torch.uint8, device=self._device)
            sizes = torch.zeros(3, dtype=torch.int64, device=self._device)
            rc = self._lib.rd_partition_records_device(h, _ptr(d_text) if n else None, _ptr(rec) if n else None, n,
                                                       _ptr(labels) if n else None, _ptr(out), _ptr(sizes), self._stream())
        _lib.check(self._lib, h, rc, "rd_partition_records_device")
        sizes = sizes.cpu()
        return out[:int(sizes.sum())], sizes

    def fastq_submit(self, slot, bufs, lens, final_chunk, max_records, max_len, outs, labels=None, mode="none",
                     semantics=None, precision=None, fasta=False):
        """Streaming form: host FASTQ (or, fasta=True, FASTA) block(s) in (numpy uint8, one per end), host outs (numpy
        uint8, capacity len + 2) filled by the time ``fastq_collect(slot)`` returns.
        → (n_records, consumed[ends], out_bytes[ends])."""
        h = self._need()
        ends = len(bufs)
        semantics = semantics or ("packed" if self.pack_seq else "padded")
        precision = precision or self.precision
        n = ctypes.c_int64(0)
        consumed = (ctypes.c_int64 * 2)()
        out_bytes = (ctypes.c_int64 * 2)()
        b2, l2, o2 = (_ptr(bufs[1]), int(lens[1]), _ptr(outs[1])) if ends == 2 else (None, 0, None)
        submit = self._lib.rd_fasta_submit if fasta else self._lib.rd_fastq_submit
        rc = submit(h, int(slot), ends, _ptr(bufs[0]), int(lens[0]), b2, l2, int(bool(final_chunk)),
                                       int(max_records), int(max_len), _lib.SEM[semantics], _lib.PREC[precision],
                                       _lib.PAIR[mode], _ptr(outs[0]), o2, _ptr(labels), ctypes.byref(n),
                                       ctypes.cast(consumed, ctypes.c_void_p), ctypes.cast(out_bytes, ctypes.c_void_p))
        if rc == _lib.RD_ERR_PARSE:
            raise ValueError(self._lib.rd_last_error(h).decode("utf-8", "replace"))
        _lib.check(self._lib, h, rc, "rd_fastq_submit")
        return int(n.value), [int(consumed[e]) for e in range(ends)], [int(out_bytes[e]) for e in range(ends)]

    def fastq_collect(self, slot):
        """→ (sizes int64[2, 3]: text bytes per label {non-rRNA, rRNA, unclassified} in out1 / out2, counts int64[3])."""
        sizes = (ctypes.c_int64 * 6)()
        counts = (ctypes.c_int64 * 3)()
        rc = self._lib.rd_fastq_collect(self._need(), int(slot), ctypes.cast(sizes, ctypes.c_void_p),
                                        ctypes.cast(counts, ctypes.c_void_p))
        if rc:
            raise _lib.RdError("rd_fastq_collect failed (code %d)" % rc)
        return np.array(list(sizes), np.int64).reshape(2, 3), np.array(list(counts), np.int64)

    # ---- drop-in __call__: the tensors the reference's collate functions produce -----------------

    def __call__(self, x):
        """x: a ReadBatch (this package's collate output), a PackedSequence of one-hot rows (detect.py:685) → packed semantics, or a padded
        one-hot tensor [B, T, 4] (detect.py:687 / detect_cpu.py:699-700) → padded semantics.
        One-hot rows are turned back into base bytes on the device (plumbing only)."""
        self._need()
        if isinstance(x, ReadBatch):
            return self.classify(x.seq, x.off, x.max_len, semantics="packed" if x.pack_seq else "padded",
                                 want_labels=False)[0]
        if isinstance(x, PackedSequence):
            padded, lens = pad_packed_sequence(x.to(self._device), batch_first=True)
            semantics = "packed"
        elif isinstance(x, torch.Tensor) and x.dim() == 3 and x.shape[2] == 4:
            padded = x.to(self._device)
            lens = torch.full((padded.shape[0],), padded.shape[1], dtype=torch.int64)
            semantics = "padded"
        else:
            raise TypeError("SeqModel expects a PackedSequence or a [B,T,4] tensor")
        B, T = padded.shape[0], padded.shape[1]
        if self._alphabet is None or self._alphabet.device != self._device:       # per instance: one model per GPU/thread
            self._alphabet = torch.tensor(list(b"ACGTN"), dtype=torch.uint8, device=self._device)
        code = torch.where(padded.sum(2) == 0, torch.full((), 4, device=self._device),
                           padded.argmax(2))
        bytes_bt = self._alphabet[code]
        lens_d = lens.to(self._device)
        mask = torch.arange(T, device=self._device)[None, :] < lens_d[:, None]
        seq = bytes_bt[mask].contiguous()
        off = torch.zeros(B + 1, dtype=torch.int64, device=self._device)
        off[1:] = torch.cumsum(lens_d, 0)
        logits, _, _ = self.classify(seq, off, max_len=T, semantics=semantics, want_labels=False)
        return logits

    forward = __call__

    def __str__(self):
        return "SeqModel(BiLSTM 4->%d, Linear %d->2) [librd_b200, precision=%s]\nTrainable parameters: %d" % (
            self.hidden_size, 2 * self.hidden_size, self.precision,
            2 * (4 * self.hidden_size * (4 + self.hidden_size + 2)) + 2 * 2 * self.hidden_size + 2)
