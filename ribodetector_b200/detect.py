"""``ribodetector`` — the reference's command line (``ribodetector/detect.py:763-809``) on the B200
hot path.  Same flags, same ``config.json`` model description, same output files and final count
lines; the batch loops of ``Predictor.run`` / ``run_with_chunks`` (``detect.py:121-523``) are one
streaming pipeline here.

FASTQ inputs (plain or gz; the default path, ``data_loader/fastq_gpu.py``): the host only moves bytes —

    producer thread : file block → page-locked buffer → rd_fastq_submit (H2D, K0 record scan; returns when the
                      record count and the cut position are known; K1-K3 classify and K4 label partition stay queued)
    consumer thread : rd_fastq_collect → one write() per label group and output file, input order kept
    2 x n_devices blocks of up to 256 MB are in flight, round-robin over the visible GPUs

FASTA inputs, ``--chunk_size`` runs and ``--host_ingest``:

    reader thread(s): file block → rd_scan_fastx (record index + sequence bytes, host C++)
    main thread     : rd_classify_host / rd_classify_pairs_host on every visible GPU (reads sharded
                      contiguously across devices, one handle per device)
    writer thread   : rd_partition_records → non-rRNA / rRNA / unclassified files, input order kept

Memory is bounded by the block / chunk size (``--chunk_size`` keeps its meaning: chunk = batch_size x
chunk_size reads, ``detect.py:370-371``), so there is no whole-file mode to run out of RAM.  ``-m`` only
feeds the reference's batch-size formula that the log line reports; ``-t`` sizes the host-side threads.
"""
import argparse
import math
import time
import os
import queue
import threading
from argparse import RawTextHelpFormatter
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import __version__
from .parse_config import ConfigParser

cd = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CHUNK_READS = 1 << 20


class colors:
    OKBLUE = '\033[94m'
    OKCYAN = '\033[96m'
    OKYELLOW = '\033[33m'
    FAIL = '\033[91m'
    ENDC = '\033[0m'
    BOLD = '\033[1m'


class Predictor:
    """Same life-cycle as the reference class: load_model() then detect()."""

    semantics = "packed"          # ribodetector_cpu overrides with "padded"
    gpu_flags = True

    def __init__(self, config, args):
        self.config = config
        self.args = args
        self.logger = config.get_logger('predict', 1, self.args.log)
        self.chunk_size = self.args.chunk_size

    # ---- detect.py:45-82 ---------------------------------------------------------------------------
    def get_state_dict(self):
        self.len = self.args.len
        if self.len < 40:
            self.logger.info('The accuracy will drop with reads shorter than 40.')
        model_file_ext = 'recall' if self.args.ensure == 'norrna' else 'mcc'
        self.state_file = os.path.join(cd, self.config['state_file'][model_file_ext])
        self.logger.info('Using high {} model'.format(model_file_ext.upper()))
        self.logger.info('Log file: {}'.format(self.args.log))

    # ---- detect.py:84-119 --------------------------------------------------------------------------
    def load_model(self):
        if getattr(self.args, "deviceid", None) is not None:
            os.environ["CUDA_VISIBLE_DEVICES"] = self.args.deviceid
        self.get_state_dict()
        import torch
        from .model import model as module_arch
        from .utils.weights import load_weights
        if not torch.cuda.is_available():
            self.logger.error('{}No visible CUDA devices!{} This build has no CPU path: the hot loop runs on '
                              'sm_100a kernels only'.format(colors.FAIL, colors.ENDC))
            raise RuntimeError("Set CUDA_VISIBLE_DEVICES to a B200.")
        n_dev = torch.cuda.device_count()
        want = self.config['n_gpu'] if getattr(self.args, "deviceid", None) is None else n_dev
        self.devices = list(range(max(1, min(n_dev, want))))
        state_dict = load_weights(self.state_file)
        self.models = []
        for d in self.devices:
            model = self.config.init_obj('arch', module_arch, precision=self.args.precision)
            model.load_state_dict(state_dict)
            self.models.append(model.to('cuda:%d' % d).eval())
        self.model = self.models[0]
        self.logger.info('Model using {} for read length {}{}{}{} loaded'.format(
            'cuda x%d' % len(self.devices), colors.BOLD, colors.OKCYAN, self.len, colors.ENDC))

    # ---- detect.py:525-584 -------------------------------------------------------------------------
    def detect(self):
        self.input = self.args.input
        self.output = self.args.output
        self.rrna = self.args.rrna
        num_inputs = len(self.input)
        num_rrna_outputs = None if self.rrna is None else len(self.rrna)
        if num_inputs != len(self.output) or num_inputs > 2 or num_inputs < 1:
            self.logger.error('{}The number of input and output sequence files is invalid!{}'.format(
                colors.FAIL, colors.ENDC))
            raise RuntimeError(
                "Input or output should have no more than two files and they should have the same number of files.")
        if num_rrna_outputs is not None and num_rrna_outputs != num_inputs:
            self.logger.error('{}The number of output rRNA sequence files is invalid!{}'.format(
                colors.FAIL, colors.ENDC))
            raise RuntimeError(
                "Ouput rRNA should have no more than two files and they should the same number with input files.")
        self.is_paired = num_inputs == 2
        memory = getattr(self.args, "memory", 32)
        batch_size_ = ((memory - 2) * 1024 * 1024) / ((2 if self.is_paired else 1) * self.len * 6.4)
        self.batch_size = 2 ** math.floor(math.log2(batch_size_))
        if self.gpu_flags:
            self.logger.info('Choose batch size: {}{}{}{} based on the given GPU RAM size {}GB and max read length {}'.format(
                colors.BOLD, colors.OKCYAN, self.batch_size, colors.ENDC, memory, self.len))
        self.chunk_reads = DEFAULT_CHUNK_READS if self.chunk_size is None else max(1, self.batch_size * self.chunk_size)
        self._label_buf = [None] * len(self.models)
        self.run()

    # ---- the streaming pipeline (detect.py:121-523) ---------------------------------------------------
    def _classify(self, chunks):
        """labels int8[n], counts int64[3] for one (pair of) RecordChunk(s), sharded across devices."""
        from .shard import shard_bounds
        n = chunks[0].n
        labels = np.empty(n, np.int8)
        counts = np.zeros(3, np.int64)
        world = len(self.models) if n >= 4096 else 1

        def work(r):
            b, e = shard_bounds(n, r, world)
            if e <= b:
                return
            m = self.models[r]
            if self._label_buf[r] is None or self._label_buf[r].numel() < e - b:
                import torch
                self._label_buf[r] = torch.empty(max(e - b, self.chunk_reads), dtype=torch.int8, pin_memory=True)
            out = {"labels": self._label_buf[r][:e - b]}
            if self.is_paired:
                res = m.classify_pairs_host(chunks[0].seq, chunks[0].seq_off[b:e + 1], chunks[1].seq,
                                            chunks[1].seq_off[b:e + 1], self.len, mode=self.args.ensure,
                                            semantics=self.semantics, out=out)
            else:
                res = m.classify_host(chunks[0].seq, chunks[0].seq_off[b:e + 1], self.len, semantics=self.semantics,
                                      want_logits=False, out=out)
            labels[b:e] = res["labels"].numpy()
            return res["counts"].numpy()

        if world == 1:
            counts += work(0)
        else:
            with ThreadPoolExecutor(world) as ex:
                for c in ex.map(work, range(world)):
                    if c is not None:
                        counts += c
        return labels, counts

    def _pair_chunks(self, r1, r2):
        """Re-slice two independent chunk streams into chunks with equal record counts."""
        c1 = c2 = None
        o1 = o2 = 0
        while True:
            if c1 is None or o1 == c1.n:
                c1, o1 = next(r1, None), 0
            if c2 is None or o2 == c2.n:
                c2, o2 = next(r2, None), 0
            if c1 is None or c2 is None:
                if (c1 is None) != (c2 is None):
                    raise RuntimeError("The two input files hold different numbers of reads.")
                return
            k = min(c1.n - o1, c2.n - o2)
            yield c1.view(o1, o1 + k), c2.view(o2, o2 + k)
            o1 += k
            o2 += k

    def run(self):
        from .data_loader import FastxReader, open_for_write, partition_records
        ends = 2 if self.is_paired else 1
        threads = max(1, min(int(self.args.threads), os.cpu_count() or 1))
        want_unc = self.is_paired and self.args.ensure == 'both'
        if self.rrna is not None:
            self.logger.info('Writing output rRNA sequences into file: {}{}{}'.format(
                colors.OKBLUE, ", ".join(self.rrna), colors.ENDC))
        self.logger.info('Writing output non-rRNA sequences into file: {}{}{}'.format(
            colors.OKBLUE, ", ".join(self.output), colors.ENDC))
        fh_non = [open_for_write(f, threads) for f in self.output]
        fh_rrna = [open_for_write(f, threads) for f in self.rrna] if self.rrna is not None else None
        fh_unc = None
        if want_unc:
            unc = [f + '.unclassified.gz' for f in self.output]
            fh_unc = [open_for_write(f, threads) for f in unc]
            self.logger.info('Writing unclassified sequences into file: {}{}{}'.format(
                colors.OKYELLOW, ", ".join(unc), colors.ENDC))

        if self._device_ingest():
            # the file text goes to the GPU as it is: record scan (K0) and label partition (K4) run there too
            from .data_loader.fastq_gpu import FastqGpuStream
            stream = FastqGpuStream(self.models, self.input, self.len, mode=self.args.ensure, semantics=self.semantics,
                                    precision=self.args.precision, threads=threads)
            try:
                total = stream.run({"non": fh_non, "rrna": fh_rrna, "unc": fh_unc})
            finally:
                for fh in fh_non + (fh_rrna or []) + (fh_unc or []):
                    fh.close()
            self.stage_seconds = stream.stage_seconds
            self.logger.debug('stage busy seconds: %s (page-locking buffers: %.2f s)', stream.stage_seconds, stream.setup_seconds)
            self.setup_seconds = stream.setup_seconds
            self._report(stream.num_seqs, total, want_unc)
            return

        readers = [FastxReader(f, max_records=self.chunk_reads, threads=max(1, threads // ends), pinned=True)
                   for f in self.input]
        q_in = queue.Queue(maxsize=3)
        q_out = queue.Queue(maxsize=3)
        errors = []

        busy = {"read": 0.0, "classify": 0.0, "write": 0.0}
        scratch = [{}, {}]

        def produce():
            try:
                stream = self._pair_chunks(iter(readers[0]), iter(readers[1])) if self.is_paired else ((c,) for c in readers[0])
                t0 = time.perf_counter()
                for chunks in stream:
                    busy["read"] += time.perf_counter() - t0
                    q_in.put(chunks)
                    t0 = time.perf_counter()
            except BaseException as e:          # noqa: BLE001 — surfaced on the main thread
                errors.append(e)
            finally:
                q_in.put(None)

        def consume():
            try:
                while True:
                    item = q_out.get()
                    if item is None:
                        return
                    chunks, labels = item
                    t0 = time.perf_counter()
                    for e in range(ends):
                        outs, _ = partition_records(chunks[e], labels, (True, fh_rrna is not None, want_unc), threads,
                                                    scratch=scratch[e])
                        if outs[0] is not None:
                            fh_non[e].write(memoryview(outs[0]))
                        if outs[1] is not None:
                            fh_rrna[e].write(memoryview(outs[1]))
                        if outs[2] is not None:
                            fh_unc[e].write(memoryview(outs[2]))
                        chunks[e].release()
                    busy["write"] += time.perf_counter() - t0
            except BaseException as e:          # noqa: BLE001
                errors.append(e)
                while q_out.get() is not None:
                    pass

        t_in = threading.Thread(target=produce, daemon=True)
        t_out = threading.Thread(target=consume, daemon=True)
        t_in.start()
        t_out.start()
        num_seqs = 0
        total = np.zeros(3, np.int64)
        try:
            while True:
                chunks = q_in.get()
                if chunks is None or errors:
                    break
                t0 = time.perf_counter()
                labels, counts = self._classify(chunks)
                busy["classify"] += time.perf_counter() - t0
                num_seqs += chunks[0].n
                total += counts
                q_out.put((chunks, labels))
        finally:
            q_out.put(None)
            t_out.join()
            for fh in fh_non + (fh_rrna or []) + (fh_unc or []):
                fh.close()
            for r in readers:
                r.close()
        if errors:
            raise errors[0]

        busy["read"] -= sum(r.wait_seconds for r in readers)      # time blocked on buffer back-pressure is not work
        self.stage_seconds = busy
        self.logger.debug('stage busy seconds: %s', busy)
        self._report(num_seqs, total, want_unc)

    def _device_ingest(self):
        """FASTQ or FASTA inputs (plain or gz; all of one kind) take the device ingest path unless --host_ingest asks for
        the host scanner."""
        from .data_loader import get_seq_format
        if getattr(self.args, "host_ingest", False) or self.chunk_size is not None:
            return False
        kinds = {get_seq_format(f)[:2] for f in self.input}
        return len(kinds) == 1

    def _report(self, num_seqs, total, want_unc):
        self.num_seqs, self.num_nonrrna, self.num_rrna, self.num_unknown = num_seqs, int(total[0]), int(total[1]), int(total[2])
        self.logger.info('Processed {}{}{}{} sequences in total'.format(colors.BOLD, colors.OKCYAN, num_seqs, colors.ENDC))
        self.logger.info('Detected {}{}{}{} non-rRNA sequences'.format(colors.BOLD, colors.OKCYAN, self.num_nonrrna, colors.ENDC))
        self.logger.info('Detected {}{}{}{} rRNA sequences'.format(colors.BOLD, colors.OKCYAN, self.num_rrna, colors.ENDC))
        if want_unc:
            self.logger.info('Discarded {}{}{}{} unclassified sequences'.format(
                colors.BOLD, colors.OKCYAN, self.num_unknown, colors.ENDC))

    run_with_chunks = run

    # ---- detect.py:586-663, for code that drives the model batch by batch like the reference loop ----------
    @staticmethod
    def generate_chunks(reads, n):
        for i in range(0, len(reads), n):
            yield reads[i:i + n]

    @staticmethod
    def generate_paired_read_chunks(reads, n):
        r1, r2 = reads
        for i in range(0, len(r1), n):
            yield r1[i:i + n], r2[i:i + n]

    @staticmethod
    def separate_reads(reads, labels):
        from collections import defaultdict
        reads_dict = defaultdict(list)
        for read, label in zip(reads, labels):
            reads_dict[int(label)].append(read)
        return reads_dict

    def separate_paired_reads(self, r1_reads, r1_outs, r2_reads, r2_outs):
        """Pair rule of `-e` on the two ends' logits (rd_pair_combine on the device), then routing."""
        from collections import defaultdict
        labels = self.model.pair_combine(r1_outs, r2_outs, self.args.ensure).cpu().tolist()
        r1_dict, r2_dict = defaultdict(list), defaultdict(list)
        for r1, r2, label in zip(r1_reads, r2_reads, labels):
            r1_dict[label].append(r1)
            r2_dict[label].append(r2)
        return r1_dict, r2_dict


# ---- the reference's feed contract (detect.py:666-726), kept for code written against it ------------------
def unlabeled_read_collate_fn(batch, max_len=100, pack_seq=True):
    """batch of record tuples → (list of record texts, ReadBatch).  Same signature as the reference;
    the encoded data is the byte form the CUDA path consumes instead of a PackedSequence of one-hot rows."""
    from .model import ReadBatch
    read_list = ['\n'.join(read) for read in batch]
    seqs = [read[1][:max_len].encode("latin-1") for read in batch]
    off = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return read_list, ReadBatch(np.frombuffer(b"".join(seqs), np.uint8).copy(), off, max_len, pack_seq)


def unlabeled_paired_read_collate_fn(batch, max_len=100, pack_seq=True):
    r1_list, r1_data = unlabeled_read_collate_fn([p[0] for p in batch], max_len, pack_seq)
    r2_list, r2_data = unlabeled_read_collate_fn([p[1] for p in batch], max_len, pack_seq)
    return r1_list, r1_data, r2_list, r2_data


def build_parser(gpu=True):
    args = argparse.ArgumentParser(description='rRNA sequence detector', formatter_class=RawTextHelpFormatter)
    args.add_argument('-c', '--config', default=None, type=str, help='Path of config file')
    if gpu:
        args.add_argument('-d', '--deviceid', default=None, type=str,
                          help='Indices of GPUs to enable. Quotated comma-separated device ID numbers. (default: all)')
    args.add_argument('-l', '--len', type=int, required=True,
                      help='Sequencing read length. Note: the accuracy reduces for reads shorter than 40.')
    args.add_argument('-i', '--input', default=None, type=str, nargs='*', required=True,
                      help='Path of input sequence files (fasta and fastq), the second file will be considered as second end if two files given.')
    args.add_argument('-o', '--output', default=None, type=str, nargs='*', required=True,
                      help='Path of the output sequence files after rRNAs removal (same number of files as input). \n(Note: 2 times slower to write gz files)')
    args.add_argument('-r', '--rrna', default=None, type=str, nargs='*',
                      help='Path of the output sequence file of detected rRNAs (same number of files as input)')
    args.add_argument('-e', '--ensure', default="none", type=str, choices=['rrna', 'norrna', 'both', 'none'],
                      help='''Ensure which classificaion has high confidence for paired end reads.
norrna: output only high confident non-rRNAs, the rest are clasified as rRNAs;
rrna: vice versa, only high confident rRNAs are classified as rRNA and the rest output as non-rRNAs;
both: both non-rRNA and rRNA prediction with high confidence;
none: give label based on the mean probability of read pair.
      (Only applicable for paired end reads, discard the read pair when their predicitons are discordant)''')
    args.add_argument('-t', '--threads', default=10 if gpu else 20, type=int,
                      help='Number of threads to use. (default: {})'.format(10 if gpu else 20))
    if gpu:
        args.add_argument('-m', '--memory', default=32, type=int, help='Amount (GB) of GPU RAM. (default: 12)')
    args.add_argument('--chunk_size', default=None, type=int,
                      help='Use this parameter when having low memory. Parsing the file in chunks.\n'
                           'chunk = batch_size x chunk_size reads; without it chunks of 1 Mi reads are streamed.')
    args.add_argument('--log', default=None, type=str, help='Log file name')
    args.add_argument('--precision', default='tc_mixed', choices=['tc_mixed', 'tc_exact', 'tc_auto', 'tc_fast', 'tc_mixed_raw', 'fp32'],
                      help='(extension) arithmetic of the recurrent contraction: tc_mixed = tcgen05 fp16 pass + 8-bit correction pass,\nlow-margin reads re-run in tc_exact (default; labels = tc_exact); tc_exact = 3-pass fp16 split, fp32-grade logits;\ntc_auto = tc_fast + tc_exact on low-margin reads; tc_fast = one fp16 pass; fp32 = CUDA cores')
    args.add_argument('--host_ingest', action='store_true',
                      help='(extension) scan FASTQ records and partition the output on the host instead of the GPU\n'
                           '(FASTA inputs and --chunk_size runs always do)')
    args.add_argument('-v', '--version', action='version', version='%(prog)s {version}'.format(version=__version__))
    return args


def main(argv=None):
    args = build_parser(gpu=True).parse_args(argv)
    config_file = os.path.join(cd, 'config.json') if args.config is None else args.config
    config = ConfigParser.from_json(config_file)
    seq_pred = Predictor(config, args)
    seq_pred.load_model()
    seq_pred.detect()
    return seq_pred


if __name__ == '__main__':
    main()
