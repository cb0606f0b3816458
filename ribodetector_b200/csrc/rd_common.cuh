// rd_common.cuh — shared definitions of librd_b200 (handle, layouts, launch bookkeeping).
//
// HBM data layout of one classify batch (all scratch is owned by the handle):
//   d_seq   uint8  [total]            caller's sequence bytes (FASTQ sequence lines, concatenated)
//   d_off   int64  [n+1]              caller's read offsets into d_seq
//   plan    uint32 [n]                per read: nfwd (13 b) | krev (13 b) << 13 | crev (3 b) << 26
//                                     | invalid << 29      (see rd_plan.cu)
//   perm    int32  [tiles*128]        slot → read index, slots sorted by nfwd DESCENDING
//                                     (length-bucketed tiles, SURVEY §5 "long-context"); -1 = pad
//   splan   uint32 [tiles*128]        plan[] gathered into slot order (0 for pad slots)
//   codes   uint8  [tiles][L][128]    ONLY for precision fp32 (CUDA-core kernel): base codes 0..3 = A,C,G,T/U,
//                                     4 = zero row, transposed so the 128 reads of a tile at step t are one
//                                     128-B line.  The tensor-core kernels read the caller's sequence bytes
//                                     directly (one byte per read per step through L1) and convert in registers.
//   logits  fp32   [n][2]             caller's output, input order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/rd_b200.h"

#define RD_H        128            // hidden size
#define RD_G4       512            // 4 gates x hidden
#define RD_TILE     128            // reads per tile (= TMEM lanes = UMMA M)

#define PLAN_NFWD(p)   ((p) & 0x1FFFu)
#define PLAN_KREV(p)   (((p) >> 13) & 0x1FFFu)
#define PLAN_CREV(p)   (((p) >> 26) & 0x7u)
#define PLAN_INVALID(p) (((p) >> 29) & 0x1u)

struct rd_tc_state;                // tensor-core weight images (rd_lstm_tc.cu)
struct rd_fq_state;                // FASTQ-on-device scratch and streaming slots (rd_fastq_dev.cu)

// BASE_DICT of the reference (seq_encoder.py:11-18): A C G T U -> 0 1 2 3 3; everything else (N, IUPAC,
// lower case, '-') -> 4 = the zero row
__device__ __forceinline__ uint32_t rd_base_code(uint32_t b) {
    uint32_t c = 4u;
    c = (b == 'A') ? 0u : c;
    c = (b == 'C') ? 1u : c;
    c = (b == 'G') ? 2u : c;
    c = (b == 'T' || b == 'U') ? 3u : c;
    return c;
}

struct rd_handle {
    int device = 0;
    int sm_count = 0;
    int hidden = RD_H;             // hidden size of the loaded model; the tensor-core kernels need RD_H, any other
                                   // multiple of 32 up to 256 runs every precision on the fp32 kernel (rd_lstm_fp32.cu)
    bool fp32_attr_set = false, lut_attr_set = false;
    int fp32_ctas_per_sm = 1;
    std::string err;
    int64_t launches = 0;

    // weights (device, fp32)
    float* d_tab_f = nullptr;      // [5][512] fwd gate-input table: W_ih^T rows + b_ih + b_hh; row 4 = bias only
    float* d_whh_g4 = nullptr;     // [H (k)][H (u)][4]: W_hh^T with the four gates of a unit adjacent (rd_lstm_fp32.cu)
    float* d_tab_r = nullptr;      // [5][512] reverse direction table
    float* d_whh_r_t = nullptr;    // [128][512]
    float* d_wout = nullptr;       // [2][256]
    float* d_bout = nullptr;       // [2]
    float* d_revlut = nullptr;     // [RD_MAX_LEN][5][2] reverse-half logit contributions; rows [0, lut_rows) are built
    double* d_lutstate = nullptr;  // [2][128] fp64 (h, c) of the reverse chain after lut_rows zero steps
    int lut_rows = 0;
    rd_tc_state* tc = nullptr;
    rd_fq_state* fq = nullptr;

    // scratch
    int64_t cap_n = 0;             // reads
    int64_t cap_slots = 0;         // tiles*128
    uint32_t* d_plan = nullptr;
    uint32_t* d_splan = nullptr;
    int32_t* d_perm = nullptr;
    uint32_t* d_splan2 = nullptr;  // TC_AUTO: the slots of the reads re-run in exact mode (same order, compacted)
    int32_t* d_perm2 = nullptr;
    int64_t* d_band = nullptr;     // [slots/256 + 2]: per-block band counts -> exclusive offsets; last = total
    int64_t cap_band = 0;
    int32_t* d_hist = nullptr;     // [RD_MAX_LEN+2] histogram → bucket starts
    int32_t* d_cursor = nullptr;   // [RD_MAX_LEN+2]
    int32_t* d_ctrl = nullptr;     // [8]: 0 = status flags, 1 = work counter, 2 = max nfwd
    int64_t* d_blocksum = nullptr; // scan scratch for ragged one-hot
    int64_t cap_blocksum = 0;

    // per-stage timing (rd_set_timing)
    bool timing = false;
    struct TimedSpan { cudaEvent_t a, b; int which; };
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
    double t_ms[4] = {0, 0, 0, 0};
    int64_t t_cnt[4] = {0, 0, 0, 0};

    // host pipeline (rd_classify_host)
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    static const int NSTAGE = 2;
    uint8_t* d_stage_seq[2][NSTAGE] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int64_t* d_stage_off[2][NSTAGE] = {{nullptr, nullptr}, {nullptr, nullptr}};
    float* d_stage_logits[2][NSTAGE] = {{nullptr, nullptr}, {nullptr, nullptr}};
    float* d_stage_probs[NSTAGE] = {nullptr, nullptr};
    int8_t* d_stage_labels[NSTAGE] = {nullptr, nullptr};
    int64_t* d_stage_counts = nullptr;
    int64_t cap_stage_bytes = 0, cap_stage_n = 0;
    cudaEvent_t ev_in[NSTAGE] = {nullptr, nullptr}, ev_cmp[NSTAGE] = {nullptr, nullptr},
                ev_out[NSTAGE] = {nullptr, nullptr};
};

#define RD_CUDA(h, call)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                \
            return RD_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// kernels' host launchers (each returns RD_OK / RD_ERR_*; all async on `st`)
int rd_launch_plan(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n, int max_len,
                   int semantics, int64_t* n_tiles_out, cudaStream_t st, int ostride = 1);
int rd_launch_onehot(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n, int max_len,
                     int layout, float* d_out, int64_t* d_row_off, cudaStream_t st);
int rd_launch_lstm_fp32(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n_tiles, int max_len,
                           float* d_logits, cudaStream_t st, int ostride = 1);
int rd_launch_lstm_tc(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n_tiles, int max_len,
                      int precision, float* d_logits, cudaStream_t st, const uint32_t* d_splan = nullptr,
                      const int32_t* d_perm = nullptr, const int64_t* d_n_reads = nullptr, int ostride = 1);
// read i of a batch is seq[off[i*ostride] .. off[i*ostride + 1]): ostride = 1 for the caller's off[n+1], 8 for the
// record index of rd_scan_fastq_device (off = rec + 2: the sequence line of every record)
int rd_classify_device(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n, int max_len, int semantics,
                       int precision, float* d_logits, float* d_probs, int8_t* d_labels, int64_t* d_counts,
                       cudaStream_t st, int ostride = 1);      // K1 -> K2 -> K3 on `st`, scratch grown on demand
int rd_launch_band_select(rd_handle* h, const float* d_logits, int64_t n_tiles, float tau, cudaStream_t st,
                          const float* d_mate_logits = nullptr);      // mate != NULL: band on the pair's summed margin
// RD_PAIR_NONE under a two-pass precision: re-run in TC_EXACT both ends of the pairs whose SUMMED first-pass margin is
// inside the band (detect.py:655-661 decides on the sum), so that the pair label equals TC_EXACT's
int rd_pair_none_refine(rd_handle* h, const uint8_t* const d_seq[2], const int64_t* const d_off[2], int64_t n, int max_len,
                        int semantics, int precision, float* const d_logits[2], cudaStream_t st, int ostride = 1);
int rd_launch_tail(rd_handle* h, const float* d_logits, int64_t n, float* d_probs, int8_t* d_labels,
                   int64_t* d_counts, cudaStream_t st);
int rd_launch_pair(rd_handle* h, const float* d_l1, const float* d_l2, int64_t n, int mode,
                   int8_t* d_labels, int64_t* d_counts, cudaStream_t st);
int rd_build_reverse_lut(rd_handle* h, int rows, cudaStream_t st);      // extends the table to `rows` rows if shorter
int rd_tc_create(rd_handle* h, const float* w_hh, const float* w_ih, const float* b_ih, const float* b_hh);
void rd_tc_destroy(rd_handle* h);
void rd_fq_destroy(rd_handle* h);
