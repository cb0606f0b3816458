/*
 * rd_b200.h — C ABI of the B200-native RiboDetector hot path (librd_b200.so).
 *
 * The reference (hzi-bifo/RiboDetector, pure Python) has no FFI; its plugin seam for this path
 * is the object that `ConfigParser.init_obj('arch', module)` builds
 * (ribodetector/parse_config.py:43-57, called at ribodetector/detect.py:93) plus the collate
 * function that feeds it (detect.py:666-726).  Each entry point below cites the reference
 * interface it replaces.  The Python side (ribodetector_b200/model/model.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer; the library owns the handle
 *     (device weight images, LUTs, scratch).  Nothing allocated here is returned to the caller.
 *   - every function returns RD_OK (0) or an RD_ERR_* code and never throws; the message is
 *     available from rd_last_error().
 *   - reads are passed as a byte concatenation `seq` plus `off[n+1]` (int64, off[0] may be
 *     non-zero; read i is seq[off[i] .. off[i+1])).  Bytes are the sequence line as it
 *     appears in the FASTQ/FASTA record; the encoding table is the reference's BASE_DICT.
 *   - "d_" parameters are device pointers on the handle's device; functions taking a `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) are asynchronous on it.
 *   - a handle is bound to one device; calls on one handle must not overlap (thread-compatible,
 *     not re-entrant), mirroring the reference's single model object per process.
 */
#ifndef RD_B200_H
#define RD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rd_handle rd_handle;

#define RD_ABI_VERSION 1

/* return codes */
#define RD_OK               0
#define RD_ERR_INVALID      1   /* bad argument (maps to ValueError in the Python wrapper)      */
#define RD_ERR_CUDA         2   /* CUDA runtime failure / no device (RuntimeError)             */
#define RD_ERR_EMPTY_READ   3   /* zero-length read under packed semantics: torch's            */
                                /* pack_sequence raises on it too (detect.py:685) → RuntimeError*/
#define RD_ERR_NOMEM        4
#define RD_ERR_UNSUPPORTED  5   /* shape outside what the kernels were built for               */
#define RD_ERR_PARSE        6   /* malformed FASTQ/FASTA text (ValueError)                      */

/* which "last valid step" convention:
 *   PACKED  = `ribodetector`     : detect.py:681-685 + model/model.py:32-37,114-119
 *   PADDED  = `ribodetector_cpu` : detect_cpu.py:699-703 + model/model_cpu.py:29-37,57-62     */
#define RD_SEM_PACKED 0
#define RD_SEM_PADDED 1

/* arithmetic of the forward-direction recurrent contraction
 *   FP32     : fp32 FFMA on CUDA cores, accurate exp/tanh — the on-device fp32 reference
 *   TC_EXACT : tcgen05 kind::f16 (cta_group::2), fp16 hi/lo split of W_hh and h_t (3 MMA passes,
 *              fp32 accumulate in TMEM), ex2/rcp activations accurate to a few ulp:
 *              |dlogit| <= 2e-4 * max(1, max_len/100) vs the reference's fp32 model (measured 1.5e-5)
 *   TC_FAST  : tcgen05 kind::f16, single pass, tanh.approx activations: |dlogit| <= 5e-2 * max(1, max_len/100)^3
 *              (measured over 2^20 reads per length: 2.8e-2 at 100 bp, 5.5e-2 at 150 bp, 0.48 at 300 bp) */
#define RD_PREC_FP32     0
#define RD_PREC_TC_EXACT 1
#define RD_PREC_TC_FAST  2
/*   TC_AUTO  : two passes — TC_FAST over every read, then TC_EXACT over the reads whose fast margin
 *              |l1 - l0| is below 0.25 * max(1, max_len/100) (5x the fast mode's logit error bound, ~1 % of
 *              reads): LABELS equal TC_EXACT's, logits are exact-grade inside the band and fast-grade
 *              (|dlogit| <= 5e-2 * max(1, max_len/100)^3) outside it                                    */
#define RD_PREC_TC_AUTO  3
/*   TC_MIXED_RAW : the TC_EXACT structure with the two correction passes in tcgen05 kind::f8f6f4 (e5m2, K = 32 per
 *              MMA): fp16 main pass + one 8-bit pass over [W_lo | W_hi] . [h_hi ; h_lo] = 17 MMAs per chunk
 *              instead of 25, same few-ulp ex2/rcp activations.  With s = max(1, max_len/100):
 *              |dlogit| <= 3e-3 * s^3, |dprob| <= 1e-3 * s^3 (SURVEY.md 8c's tolerance at 100 bp; measured over 2^20
 *              reads per length: 1.7e-3 / 7.6e-4 at 100 bp, 7.0e-2 / 6.1e-3 at 300 bp, where the 3-pass mode itself
 *              moves 1.6e-3 away from the fp32 CUDA-core kernel — profiles/r2_prec_err_big.txt)
 *   TC_MIXED : two passes like TC_AUTO — TC_MIXED_RAW over every read, then TC_EXACT over the reads whose margin is
 *              below RD_BAND_MIXED * s^2 (12x / 2.6x the largest margin error measured for the raw kernel at
 *              100 / 300 bp; 0.15 % / 1.5 % of random reads): LABELS equal TC_EXACT's, logits exact-grade inside the
 *              band and to the TC_MIXED_RAW tolerance outside it.  The default of the command line and of bench.py.
 *              (Pairs: the per-read guarantee carries over to the rrna / norrna / both rules, which combine per-end
 *              LABELS; RD_PAIR_NONE takes the argmax of the SUM of the two ends' logits, detect.py:655-661, so its label
 *              equals TC_EXACT's for every pair whose summed margin is outside twice the logit tolerance.) */
#define RD_PREC_TC_MIXED 4
#define RD_PREC_TC_MIXED_RAW 5
#define RD_BAND_FAST  0.25f
#define RD_BAND_MIXED 0.04f
#define RD_PREC_LAST     RD_PREC_TC_MIXED_RAW

/* paired-end combination, detect.py:616-663 (`-e/--ensure`) */
#define RD_PAIR_NONE   0   /* argmax(logits_r1 + logits_r2)            detect.py:655-661 */
#define RD_PAIR_RRNA   1   /* 1 iff both ends 1                        detect.py:620-630 */
#define RD_PAIR_NORRNA 2   /* 0 iff both ends 0                        detect.py:631-641 */
#define RD_PAIR_BOTH   3   /* concordant label else -1 (unclassified)  detect.py:642-654 */

/* one-hot output layouts of rd_encode_onehot */
#define RD_ONEHOT_RAGGED 0 /* [sum_i min(len_i,max_len), 4] fp32, read after read in input order:
                              seq_encoder.py:126-127 encode_read(seq[:max_len]) (detect.py:682) */
#define RD_ONEHOT_PADDED 1 /* [n, max_len, 4] fp32, zero rows appended:
                              seq_encoder.py:130-145 encode_variable_len_read                   */

#define RD_MAX_LEN 4096    /* largest supported -l / max_len */

/* sequence file formats of the host-side record scanner / writer */
#define RD_FMT_FASTQ 0
#define RD_FMT_FASTA 1

int rd_abi_version(void);

/* Replaces SeqModel.__init__ + load_state_dict + .to('cuda').eval()
 * (model/model.py:11-29, detect.py:93,115-119).  All weight pointers are HOST fp32 arrays in
 * the reference state_dict layout: w_ih [4H,4], w_hh [4H,H], b_ih [4H], b_hh [4H] (gate row
 * order i,f,g,o) for the forward ("_l0") and reverse ("_l0_reverse") directions, w_out [2,2H],
 * b_out [2].  hidden = 128 (the shipped checkpoint) runs on the tensor-core kernels; any other
 * multiple of 32 between 32 and 256 is accepted (the reference's SeqModel takes any hidden_size)
 * and runs every RD_PREC_* on the fp32 CUDA-core kernel; anything else returns
 * RD_ERR_UNSUPPORTED.  Uploads the weights, builds the gate-input table, the tensor-core weight
 * images (H = 128) and the reverse-direction logit LUT on `device`. */
int rd_create(int device,
              const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
              const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r,
              const float* w_out, const float* b_out,
              int hidden, rd_handle** out);

void rd_destroy(rd_handle* h);

/* message of the last failure on this handle (h == NULL: last failure of rd_create). */
const char* rd_last_error(const rd_handle* h);

/* Pre-size the scratch for batches of up to n reads of up to max_len steps (optional; the
 * classify calls grow it on demand, which synchronises the device). */
int rd_reserve(rd_handle* h, int64_t n, int max_len);

/* Replaces encode_read / encode_variable_len_read (seq_encoder.py:126-145) as called from the
 * collate functions (detect.py:681-682, detect_cpu.py:699-700).  d_out is fp32, layout above;
 * d_row_off (RAGGED only, may be NULL) receives int64[n+1] row offsets into d_out. */
int rd_encode_onehot(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n,
                     int max_len, int layout, float* d_out, int64_t* d_row_off, void* stream);

/* Replaces collate + `model(data)` + argmax for one batch of single reads
 * (detect.py:284-288: unlabeled_read_collate_fn → SeqModel.forward1 → torch.argmax).
 * Outputs (each may be NULL): d_logits [n,2] fp32 raw logits in input order (model.py:36-37),
 * d_probs [n,2] fp32 softmax of the logits, d_labels [n] int8 argmax (ties → 0),
 * d_counts int64[3] += {#label0 (non-rRNA), #label1 (rRNA), 0}. */
int rd_classify(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n,
                int max_len, int semantics, int precision,
                float* d_logits, float* d_probs, int8_t* d_labels, int64_t* d_counts,
                void* stream);

/* Replaces Predictor.separate_paired_reads' label rule (detect.py:616-663) for n pairs.
 * d_labels [n] int8 in {-1,0,1}; d_counts int64[3] += {non-rRNA, rRNA, unclassified} pairs. */
int rd_pair_combine(rd_handle* h, const float* d_logits1, const float* d_logits2, int64_t n,
                    int mode, int8_t* d_labels, int64_t* d_counts, void* stream);

/* Host-buffer forms: what the reference's batch loops do per file/chunk (detect.py:284-298 and
 * :183-206).  Synchronous.  seq/off/outputs are HOST pointers (pinned or pageable); the library
 * pipelines H2D → kernels → D2H in chunks on its own streams.  counts is int64[3], overwritten.
 * Outputs may be NULL except labels. */
int rd_classify_host(rd_handle* h, const uint8_t* seq, const int64_t* off, int64_t n,
                     int max_len, int semantics, int precision,
                     float* logits, float* probs, int8_t* labels, int64_t* counts);

int rd_classify_pairs_host(rd_handle* h,
                           const uint8_t* seq1, const int64_t* off1,
                           const uint8_t* seq2, const int64_t* off2, int64_t n,
                           int max_len, int semantics, int precision, int mode,
                           float* logits1, float* logits2, int8_t* labels, int64_t* counts);

/* Diagnostics: number of kernel launches issued through this handle so far, and the reverse-
 * direction logit LUT (host copy, fp32 [kmax+1,5,2]) for tests. */
int64_t rd_kernel_launches(const rd_handle* h);
int rd_reverse_lut(rd_handle* h, int kmax, float* out);

/* Per-stage device timing for bench.py's roofline: when enabled, every classify / pair call
 * brackets its stages with CUDA events on the launching stream.  rd_get_timing synchronises the
 * device and returns accumulated milliseconds and launch counts per stage
 * (0 = K1 plan/encode, 1 = K2 LSTM, 2 = K3 tail, 3 = pair combine); reset != 0 clears them. */
int rd_set_timing(rd_handle* h, int enable);
int rd_get_timing(rd_handle* h, double* ms4, int64_t* count4, int reset);

/* ---- host side of the path's edges (no GPU involved) ------------------------------------------------
 * Replaces seq_parser (data_loader/fastx_parser.py:15-55) for one buffer of UNCOMPRESSED text: scans
 * complete records of buf[0..len) and returns their number (>= 0) or -RD_ERR_* (message from
 * rd_fastx_last_error).  Reference semantics: FASTQ lines rstrip()ped, not upper-cased, a truncated
 * final record dropped, blank lines an error; FASTA lines strip()ped, joined, upper-cased.
 *   final_chunk  != 0 when buf ends at end of file (otherwise an unfinished record is left for the
 *                next call: *consumed = offset of the first byte not consumed)
 *   hdr          int64[2*max_records]  [begin, end) of each header line in buf (incl. '@' / '>')
 *   plus, qual   int64[2*max_records]  FASTQ only: the '+' line and the quality line
 *   seq_out      the sequences, concatenated, exactly as the reference's record[1] (this is the
 *                `seq` argument of rd_classify*); seq_off int64[max_records+1] their offsets
 * Stops early when max_records or seq_cap is reached.  `threads` host threads index a FASTQ buffer
 * in parallel (records are exactly four lines, so newline counts per segment locate them). */
int64_t rd_scan_fastx(const uint8_t* buf, int64_t len, int format, int final_chunk, int64_t max_records,
                      int64_t* hdr, int64_t* plus, int64_t* qual, uint8_t* seq_out, int64_t seq_cap,
                      int64_t* seq_off, int64_t* consumed, int threads);

/* Replaces '\n'.join(record) + separate_reads / separate_paired_reads routing + the writes
 * (detect.py:680,601-614,295-298): appends "header\nseq\n[plus\nqual\n]" of every record to the
 * stream of its label (0 -> out_non, 1 -> out_rrna, -1 -> out_unc), preserving input order.
 * sizes3 receives the byte counts; output pointers may be NULL (then only sizes are computed —
 * call once with NULLs to size the buffers, once more to fill them). */
int rd_partition_records(const uint8_t* buf, int format, int64_t n, const int64_t* hdr, const int64_t* plus,
                         const int64_t* qual, const uint8_t* seq, const int64_t* seq_off, const int8_t* labels,
                         uint8_t* out_non, uint8_t* out_rrna, uint8_t* out_unc, int64_t* sizes3, int threads);

const char* rd_fastx_last_error(void);

/* Replaces `gzip.open` (seq_encoder.py:43-53) for BGZF-framed .gz text (bgzip, bcl2fastq / BCL Convert output): every
 * member carries its compressed size, so the complete members of in[0..in_len) are located without inflating and
 * inflated on `threads` host threads into out (CRC-32 checked), as many as fit in out_cap.  Returns the bytes written
 * (0: no complete member fits yet), *in_used = input bytes consumed; -RD_ERR_UNSUPPORTED if in does not start with a
 * BGZF member (use a serial gzip reader), -RD_ERR_PARSE on a corrupt member. */
int64_t rd_bgzf_inflate(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_cap, int64_t* in_used, int threads);

/* The serial form for any other .gz (members cannot be located without inflating): a zlib stream driven directly on the
 * caller's buffers, concatenated members and zero padding handled, CRC-32 checked.  rd_gz_inflate returns the bytes
 * written into out (stops when in or out is exhausted), *in_used = input consumed, *mid_member = 1 while inside a member
 * (end of file there = truncated file); -RD_ERR_PARSE on a corrupt stream (message from rd_fastx_last_error). */
typedef struct rd_gz rd_gz;
rd_gz* rd_gz_open(void);
int64_t rd_gz_inflate(rd_gz* g, const uint8_t* in, int64_t in_len, int64_t* in_used, uint8_t* out, int64_t out_cap,
                      int* mid_member);
void rd_gz_close(rd_gz* g);

/* ---- the same edges on the device, for uncompressed FASTQ text resident in HBM --------------------------
 * K0.  Replaces seq_parser's FASTQ branch (fastx_parser.py:15-47) with rd_scan_fastx's semantics: d_buf[0..len)
 * (16-byte aligned) is scanned in one pass; d_rec receives int64[8] per record = [begin, end) of the header,
 * sequence, '+' and quality lines after rstrip(); d_info (device int64[8]) receives [0] newlines seen,
 * [1] records n (complete ones, <= max_records), [2] consumed = first byte after record n-1, [4] -1 or
 * 4 * (first malformed record) + kind (1 blank line, 2 header without '@').  Asynchronous on `stream`. */
int rd_scan_fastq_device(rd_handle* h, const uint8_t* d_buf, int64_t len, int final_chunk, int64_t max_records,
                         int64_t* d_rec, int64_t* d_info, void* stream);

/* rd_classify over the sequence lines of a record index: read i = d_buf[d_rec[8i+2] .. d_rec[8i+3]). */
int rd_classify_records(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n, int max_len,
                        int semantics, int precision, float* d_logits, float* d_probs, int8_t* d_labels,
                        int64_t* d_counts, void* stream);

/* K4.  Replaces '\n'.join(record) + separate_reads / separate_paired_reads routing (detect.py:680,601-663):
 * d_out receives the text of the label-0 records, then the label-1 records, then the label -1 records, each
 * group in input order; d_sizes3 (device int64[3]) their byte counts.  d_out needs len + 1 bytes. */
int rd_partition_records_device(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n,
                                const int8_t* d_labels, uint8_t* d_out, int64_t* d_sizes3, void* stream);

/* Streaming form over HOST buffers (what Predictor.run does per chunk, detect.py:284-298 / :183-206): block of
 * FASTQ text in (one per end), partitioned record text out.  Two slots (0, 1) double-buffer the pipeline
 * H2D -> K0 -> K1..K3 -> K4 -> D2H on the handle's streams.
 *   rd_fastq_submit  copies buf1 (and buf2: the mates, ends = 2) to the device, scans them and returns, once n
 *                    is known, *n_records = complete records taken (the same number from each end),
 *                    consumed2[e] = bytes of buf_e they cover (cut the next block there), out_bytes2[e] = bytes
 *                    that will land in out_e; classify, partition and the copies back to out1/out2 (host, capacity
 *                    len_e + 1) and labels (host int8[n], may be NULL) are left running.
 *   rd_fastq_collect waits for that slot; sizes6 = {non-rRNA, rRNA, unclassified} text bytes in out1 then out2
 *                    (laid out in that order), counts3 = records per label.
 * Calls on one handle come from one thread, except that rd_fastq_collect may run on a second thread. */
int rd_fastq_submit(rd_handle* h, int slot, int ends, const uint8_t* buf1, int64_t len1, const uint8_t* buf2,
                    int64_t len2, int final_chunk, int64_t max_records, int max_len, int semantics, int precision,
                    int mode, uint8_t* out1, uint8_t* out2, int8_t* labels, int64_t* n_records, int64_t* consumed2,
                    int64_t* out_bytes2);
int rd_fastq_collect(rd_handle* h, int slot, int64_t* sizes6, int64_t* counts3);

/* ---- the same for FASTA text -----------------------------------------------------------------------------
 * K0 for seq_parser's FASTA branch (fastx_parser.py:39-55) with rd_scan_fastx's semantics: lines strip()ped, blank
 * ones skipped, a line starting with '>' opens a record, the other lines of a record are joined and UPPER-CASED; a
 * record is complete when the next header (or, final_chunk != 0, the end of the text) is seen; at the end of the file a
 * last header without sequence is dropped; sequence lines before the very first header stay attached to it.
 * The joined sequences are written to d_buf[seq_base ..) — seq_base >= len, 16-byte aligned, d_buf holding at least
 * seq_base + len bytes — so that a FASTA record looks like a FASTQ one to rd_classify_records:
 *   d_rec int64[8] per record = [begin, end) of the header line (after strip) in d_buf, [begin, end) of the joined
 *   sequence in d_buf (inside the region at seq_base), the offset of the record's first line, three unused slots.
 * d_info (device int64[8]): [0] newlines seen, [1] records n, [2] consumed = offset of the first line not consumed,
 * [3] bytes in the sequence region, [4] -1 or 4 * (line index) + 2 for a line of 4 MiB or more. */
int rd_scan_fasta_device(rd_handle* h, uint8_t* d_buf, int64_t len, int64_t seq_base, int final_chunk,
                         int64_t max_records, int64_t* d_rec, int64_t* d_info, void* stream);

/* K4 for FASTA records: "header\nSEQUENCE\n" of every record, grouped [non-rRNA | rRNA | unclassified], input order
 * inside a group (detect.py:680,601-663 with the 2-tuple records of fastx_parser.py:53-55). */
int rd_partition_fasta_device(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n,
                              const int8_t* d_labels, uint8_t* d_out, int64_t* d_sizes3, void* stream);

/* rd_fastq_submit for blocks of FASTA text (same slots, collected with rd_fastq_collect; out_e needs len_e + 2). */
int rd_fasta_submit(rd_handle* h, int slot, int ends, const uint8_t* buf1, int64_t len1, const uint8_t* buf2,
                    int64_t len2, int final_chunk, int64_t max_records, int max_len, int semantics, int precision,
                    int mode, uint8_t* out1, uint8_t* out2, int8_t* labels, int64_t* n_records, int64_t* consumed2,
                    int64_t* out_bytes2);

#ifdef __cplusplus
}
#endif
#endif /* RD_B200_H */
