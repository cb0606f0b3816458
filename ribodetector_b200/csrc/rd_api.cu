// rd_api.cu — the C ABI of librd_b200.so (include/rd_b200.h): handle life-cycle, scratch
// management, the device-pointer entry points and the host-buffer pipeline.
#include <new>
#include <vector>
#include <algorithm>
#include <cstring>
#include "rd_common.cuh"

static thread_local std::string g_create_err;     // rd_create has no handle to carry its message

static int fail(rd_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

extern "C" int rd_abi_version(void) { return RD_ABI_VERSION; }

extern "C" const char* rd_last_error(const rd_handle* h) {
    return h ? h->err.c_str() : g_create_err.c_str();
}

extern "C" int64_t rd_kernel_launches(const rd_handle* h) { return h ? h->launches : 0; }

static void free_scratch(rd_handle* h) {
    cudaFree(h->d_plan); cudaFree(h->d_splan); cudaFree(h->d_perm);
    cudaFree(h->d_splan2); cudaFree(h->d_perm2); cudaFree(h->d_band);
    h->d_plan = h->d_splan = nullptr; h->d_perm = nullptr;
    h->d_splan2 = nullptr; h->d_perm2 = nullptr; h->d_band = nullptr;
    h->cap_n = h->cap_slots = h->cap_band = 0;
}

static int ensure_band(rd_handle* h) {          // TC_AUTO scratch, sized like the slot tables
    if (h->cap_band >= h->cap_slots && h->d_perm2) return RD_OK;
    RD_CUDA(h, cudaDeviceSynchronize());
    cudaFree(h->d_splan2); cudaFree(h->d_perm2); cudaFree(h->d_band);
    h->d_splan2 = nullptr; h->d_perm2 = nullptr; h->d_band = nullptr;
    RD_CUDA(h, cudaMalloc(&h->d_splan2, sizeof(uint32_t) * (h->cap_slots + 2 * RD_TILE)));
    RD_CUDA(h, cudaMalloc(&h->d_perm2, sizeof(int32_t) * (h->cap_slots + 2 * RD_TILE)));
    RD_CUDA(h, cudaMalloc(&h->d_band, sizeof(int64_t) * (h->cap_slots / 256 + 4)));
    h->cap_band = h->cap_slots;
    return RD_OK;
}

static int ensure_scratch(rd_handle* h, int64_t n) {
    int64_t tiles = (n + RD_TILE - 1) / RD_TILE;
    int64_t slots = (tiles + 1) * RD_TILE;                 // + one pad tile: the exact kernel works on tile pairs
    if (n <= h->cap_n && slots <= h->cap_slots) return RD_OK;
    RD_CUDA(h, cudaDeviceSynchronize());
    int64_t nn = std::max(n, h->cap_n), ss = std::max(slots, h->cap_slots);
    free_scratch(h);
    RD_CUDA(h, cudaMalloc(&h->d_plan, sizeof(uint32_t) * std::max<int64_t>(nn, 1)));
    RD_CUDA(h, cudaMalloc(&h->d_splan, sizeof(uint32_t) * std::max<int64_t>(ss, 1)));
    RD_CUDA(h, cudaMalloc(&h->d_perm, sizeof(int32_t) * std::max<int64_t>(ss, 1)));
    h->cap_n = nn; h->cap_slots = ss;
    return RD_OK;
}

extern "C" int rd_reserve(rd_handle* h, int64_t n, int max_len) {
    if (!h) return RD_ERR_INVALID;
    if (n < 0 || max_len < 1 || max_len > RD_MAX_LEN) return fail(h, RD_ERR_INVALID, "rd_reserve: bad n/max_len");
    RD_CUDA(h, cudaSetDevice(h->device));
    return ensure_scratch(h, n);
}

extern "C" int rd_create(int device,
                         const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                         const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r,
                         const float* w_out, const float* b_out, int hidden, rd_handle** out) {
    if (!out) return fail(nullptr, RD_ERR_INVALID, "rd_create: out is NULL");
    *out = nullptr;
    if (hidden < 32 || hidden > 256 || hidden % 32 != 0)
        return fail(nullptr, RD_ERR_UNSUPPORTED, "rd_create: hidden_size must be a multiple of 32 between 32 and 256 "
                                                 "(128 runs on the tensor-core kernels, the others on the fp32 CUDA-core kernel)");
    if (!w_ih_f || !w_hh_f || !b_ih_f || !b_hh_f || !w_ih_r || !w_hh_r || !b_ih_r || !b_hh_r || !w_out || !b_out)
        return fail(nullptr, RD_ERR_INVALID, "rd_create: NULL weight pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RD_ERR_CUDA, std::string("rd_create: no visible CUDA device (") +
                                              cudaGetErrorString(e) + ")");
    if (device < 0 || device >= ndev) return fail(nullptr, RD_ERR_INVALID, "rd_create: bad device index");
    rd_handle* h = new (std::nothrow) rd_handle();
    if (!h) return fail(nullptr, RD_ERR_NOMEM, "rd_create: out of host memory");
    h->device = device;
    h->hidden = hidden;
    const int H = hidden, G4 = 4 * hidden;
    auto bail = [&](int code) { g_create_err = h->err; rd_destroy(h); return code; };
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(_e); return bail(RD_ERR_CUDA); } } while (0)
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;

    // gate-input tables: row c<4 = W_ih[:,c] + b_ih + b_hh ; row 4 = b_ih + b_hh   (x_t is one-hot or zero)
    std::vector<float> tab_f(5 * G4), tab_r(5 * G4), whh_r_t(H * G4);
    for (int code = 0; code < 5; ++code)
        for (int j = 0; j < G4; ++j) {
            float bf = b_ih_f[j] + b_hh_f[j], br = b_ih_r[j] + b_hh_r[j];
            tab_f[code * G4 + j] = code < 4 ? w_ih_f[j * 4 + code] + bf : bf;
            tab_r[code * G4 + j] = code < 4 ? w_ih_r[j * 4 + code] + br : br;
        }
    for (int j = 0; j < G4; ++j)
        for (int k = 0; k < H; ++k) {
            whh_r_t[k * G4 + j] = w_hh_r[j * H + k];
        }
    CK(cudaMalloc(&h->d_tab_f, sizeof(float) * 5 * G4));
    CK(cudaMalloc(&h->d_tab_r, sizeof(float) * 5 * G4));
    CK(cudaMalloc(&h->d_whh_r_t, sizeof(float) * H * G4));
    {                                              // the fp32 kernel's image: the four gates of a unit in one 16-byte word
        std::vector<float> g4((size_t)(H + 1) * G4, 0.0f);      // + one zero row: the kernel prefetches one k ahead
        for (int k = 0; k < H; ++k)
            for (int u = 0; u < H; ++u)
                for (int g = 0; g < 4; ++g) g4[((size_t)k * H + u) * 4 + g] = w_hh_f[(size_t)(g * H + u) * H + k];
        CK(cudaMalloc(&h->d_whh_g4, sizeof(float) * (H + 1) * G4));
        CK(cudaMemcpy(h->d_whh_g4, g4.data(), sizeof(float) * (H + 1) * G4, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&h->d_wout, sizeof(float) * 2 * 2 * H));
    CK(cudaMalloc(&h->d_bout, sizeof(float) * 2));
    CK(cudaMalloc(&h->d_revlut, sizeof(float) * RD_MAX_LEN * 5 * 2));
    CK(cudaMalloc(&h->d_lutstate, sizeof(double) * 2 * H));
    CK(cudaMalloc(&h->d_hist, sizeof(int32_t) * (RD_MAX_LEN + 2)));
    CK(cudaMalloc(&h->d_cursor, sizeof(int32_t) * (RD_MAX_LEN + 2)));
    CK(cudaMalloc(&h->d_ctrl, sizeof(int32_t) * 8));
    CK(cudaMemcpy(h->d_tab_f, tab_f.data(), sizeof(float) * 5 * G4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_tab_r, tab_r.data(), sizeof(float) * 5 * G4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_whh_r_t, whh_r_t.data(), sizeof(float) * H * G4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_wout, w_out, sizeof(float) * 2 * 2 * H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_bout, b_out, sizeof(float) * 2, cudaMemcpyHostToDevice));
    CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int s = 0; s < rd_handle::NSTAGE; ++s) {
        CK(cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_cmp[s], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming));
    }
    CK(cudaMalloc(&h->d_stage_counts, sizeof(int64_t) * 4));
    int rc = rd_build_reverse_lut(h, 512, 0);
    if (rc != RD_OK) return bail(rc);
    if (hidden == RD_H) {                          // tensor-core weight images (the kernels are laid out for H = 128)
        rc = rd_tc_create(h, w_hh_f, w_ih_f, b_ih_f, b_hh_f);
        if (rc != RD_OK) return bail(rc);
    }
    CK(cudaDeviceSynchronize());
#undef CK
    *out = h;
    return RD_OK;
}

extern "C" void rd_destroy(rd_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    rd_tc_destroy(h);
    rd_fq_destroy(h);
    free_scratch(h);
    cudaFree(h->d_tab_f); cudaFree(h->d_tab_r); cudaFree(h->d_whh_r_t); cudaFree(h->d_whh_g4);
    cudaFree(h->d_wout); cudaFree(h->d_bout); cudaFree(h->d_revlut); cudaFree(h->d_lutstate);
    cudaFree(h->d_hist); cudaFree(h->d_cursor); cudaFree(h->d_ctrl); cudaFree(h->d_blocksum);
    for (int e = 0; e < 2; ++e)
        for (int s = 0; s < rd_handle::NSTAGE; ++s) {
            cudaFree(h->d_stage_seq[e][s]); cudaFree(h->d_stage_off[e][s]); cudaFree(h->d_stage_logits[e][s]);
        }
    for (int s = 0; s < rd_handle::NSTAGE; ++s) {
        cudaFree(h->d_stage_probs[s]); cudaFree(h->d_stage_labels[s]);
        if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
        if (h->ev_cmp[s]) cudaEventDestroy(h->ev_cmp[s]);
        if (h->ev_out[s]) cudaEventDestroy(h->ev_out[s]);
    }
    cudaFree(h->d_stage_counts);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_cmp) cudaStreamDestroy(h->s_cmp);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
}

extern "C" int rd_reverse_lut(rd_handle* h, int kmax, float* out) {
    if (!h || !out || kmax < 0 || kmax >= RD_MAX_LEN) return fail(h, RD_ERR_INVALID, "rd_reverse_lut: bad arguments");
    RD_CUDA(h, cudaSetDevice(h->device));
    int rc = rd_build_reverse_lut(h, kmax + 1, 0);
    if (rc) return rc;
    RD_CUDA(h, cudaMemcpy(out, h->d_revlut, sizeof(float) * (kmax + 1) * 10, cudaMemcpyDeviceToHost));
    return RD_OK;
}

static int check_common(rd_handle* h, int64_t n, int max_len, const char* who) {
    if (!h) return RD_ERR_INVALID;
    if (n < 0) return fail(h, RD_ERR_INVALID, std::string(who) + ": n < 0");
    if (n > (int64_t)1 << 30) return fail(h, RD_ERR_UNSUPPORTED, std::string(who) + ": n > 2^30 per call");
    if (max_len < 1 || max_len > RD_MAX_LEN)
        return fail(h, RD_ERR_INVALID, std::string(who) + ": max_len outside [1, RD_MAX_LEN]");
    return RD_OK;
}

extern "C" int rd_encode_onehot(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n,
                                int max_len, int layout, float* d_out, int64_t* d_row_off, void* stream) {
    int rc = check_common(h, n, max_len, "rd_encode_onehot");
    if (rc) return rc;
    if (layout != RD_ONEHOT_RAGGED && layout != RD_ONEHOT_PADDED)
        return fail(h, RD_ERR_INVALID, "rd_encode_onehot: unknown layout");
    if (n && (!d_off || !d_out)) return fail(h, RD_ERR_INVALID, "rd_encode_onehot: NULL buffer");
    RD_CUDA(h, cudaSetDevice(h->device));
    return rd_launch_onehot(h, d_seq, d_off, n, max_len, layout, d_out, d_row_off, (cudaStream_t)stream);
}

static cudaEvent_t take_event(rd_handle* h) {
    if (!h->ev_pool.empty()) { cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct StageTimer {      // brackets one stage with events on the launching stream when timing is on
    rd_handle* h; cudaStream_t st; cudaEvent_t a = nullptr; int which;
    StageTimer(rd_handle* h_, int which_, cudaStream_t st_) : h(h_), st(st_), which(which_) {
        if (h->timing) { a = take_event(h); cudaEventRecord(a, st); }
    }
    ~StageTimer() {
        if (a) { cudaEvent_t b = take_event(h); cudaEventRecord(b, st); h->spans.push_back({a, b, which}); }
    }
};

extern "C" int rd_set_timing(rd_handle* h, int enable) {
    if (!h) return RD_ERR_INVALID;
    h->timing = enable != 0;
    return RD_OK;
}

extern "C" int rd_get_timing(rd_handle* h, double* ms4, int64_t* count4, int reset) {
    if (!h) return RD_ERR_INVALID;
    RD_CUDA(h, cudaSetDevice(h->device));
    RD_CUDA(h, cudaDeviceSynchronize());
    for (auto& s : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { h->t_ms[s.which] += ms; h->t_cnt[s.which] += 1; }
        h->ev_pool.push_back(s.a); h->ev_pool.push_back(s.b);
    }
    h->spans.clear();
    for (int i = 0; i < 4; ++i) {
        if (ms4) ms4[i] = h->t_ms[i];
        if (count4) count4[i] = h->t_cnt[i];
        if (reset) { h->t_ms[i] = 0; h->t_cnt[i] = 0; }
    }
    return RD_OK;
}

static float band_tau(int precision, int max_len) {
    // the largest first-pass margin error over 2^20 reads per length grows like the square of the length for both
    // first-pass kernels (profiles/r2_prec_err_big*.txt: fast 0.055 / 0.11 / 0.96, mixed 0.0033 / 0.009 / 0.14 at 100 / 150 / 300 bp)
    const float len_scale = max_len > 100 ? (float)max_len / 100.0f : 1.0f;
    return (precision == RD_PREC_TC_AUTO ? RD_BAND_FAST : RD_BAND_MIXED) * len_scale * len_scale;
}

int rd_classify_device(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n, int max_len,
                           int semantics, int precision, float* d_logits, float* d_probs,
                           int8_t* d_labels, int64_t* d_counts, cudaStream_t st, int ostride) {
    // fp32, and every precision of a handle with a hidden size other than 128, run the fp32 CUDA-core kernel
    const bool fp32 = h->hidden != RD_H || precision == RD_PREC_FP32;
    int rc = ensure_scratch(h, n);
    if (rc) return rc;
    if (semantics == RD_SEM_PADDED && max_len > h->lut_rows) {      // krev < max_len: extend the reverse-direction table
        rc = rd_build_reverse_lut(h, max_len, st);
        if (rc) return rc;
    }
    int64_t tiles = 0;
    {
        StageTimer tm(h, 0, st);
        rc = rd_launch_plan(h, d_seq, d_off, n, max_len, semantics, &tiles, st, ostride);
    }
    if (rc) return rc;
    {
        StageTimer tm(h, 1, st);
        if (fp32) rc = rd_launch_lstm_fp32(h, d_seq, d_off, tiles, max_len, d_logits, st, ostride);
        else if (precision == RD_PREC_TC_AUTO || precision == RD_PREC_TC_MIXED) {
            // two passes: the cheaper kernel over everything, then the exact kernel over the reads whose margin from the
            // first pass is inside a band several times the first pass's error bound (their slots are compacted on the
            // device and their count stays there, so nothing synchronises with the host): LABELS equal TC_EXACT's
            const bool fast = precision == RD_PREC_TC_AUTO;
            const float tau = band_tau(precision, max_len);
            rc = ensure_band(h);
            if (!rc) rc = rd_launch_lstm_tc(h, d_seq, d_off, tiles, max_len, fast ? RD_PREC_TC_FAST : RD_PREC_TC_MIXED_RAW, d_logits, st,
                                            nullptr, nullptr, nullptr, ostride);
            if (!rc) rc = rd_launch_band_select(h, d_logits, tiles, tau, st);
            const int64_t nb = (tiles * RD_TILE + 255) / 256;
            if (!rc) rc = rd_launch_lstm_tc(h, d_seq, d_off, tiles, max_len, RD_PREC_TC_EXACT, d_logits, st, h->d_splan2,
                                            h->d_perm2, h->d_band + nb, ostride);
        } else rc = rd_launch_lstm_tc(h, d_seq, d_off, tiles, max_len, precision, d_logits, st, nullptr, nullptr, nullptr, ostride);
    }
    if (rc) return rc;
    if (d_probs || d_labels || d_counts) {
        StageTimer tm(h, 2, st);
        rc = rd_launch_tail(h, d_logits, n, d_probs, d_labels, d_counts, st);
    }
    return rc;
}

int rd_pair_none_refine(rd_handle* h, const uint8_t* const d_seq[2], const int64_t* const d_off[2], int64_t n, int max_len,
                        int semantics, int precision, float* const d_logits[2], cudaStream_t st, int ostride) {
    if (n == 0 || h->hidden != RD_H || (precision != RD_PREC_TC_AUTO && precision != RD_PREC_TC_MIXED)) return RD_OK;
    const float tau = 2.0f * band_tau(precision, max_len);          // two first-pass margins add up
    int rc = ensure_band(h);
    for (int e = 0; e < 2 && !rc; ++e) {
        int64_t tiles = 0;
        rc = rd_launch_plan(h, d_seq[e], d_off[e], n, max_len, semantics, &tiles, st, ostride);   // (the ends share the scratch)
        if (!rc) rc = rd_launch_band_select(h, d_logits[e], tiles, tau, st, d_logits[1 - e]);
        const int64_t nb = (tiles * RD_TILE + 255) / 256;
        if (!rc) rc = rd_launch_lstm_tc(h, d_seq[e], d_off[e], tiles, max_len, RD_PREC_TC_EXACT, d_logits[e], st, h->d_splan2,
                                        h->d_perm2, h->d_band + nb, ostride);
    }
    return rc;
}

extern "C" int rd_classify(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n,
                           int max_len, int semantics, int precision,
                           float* d_logits, float* d_probs, int8_t* d_labels, int64_t* d_counts,
                           void* stream) {
    int rc = check_common(h, n, max_len, "rd_classify");
    if (rc) return rc;
    if (semantics != RD_SEM_PACKED && semantics != RD_SEM_PADDED)
        return fail(h, RD_ERR_INVALID, "rd_classify: unknown semantics");
    if (precision < RD_PREC_FP32 || precision > RD_PREC_LAST)
        return fail(h, RD_ERR_INVALID, "rd_classify: unknown precision");
    if (n == 0) return RD_OK;
    if (!d_off || !d_logits) return fail(h, RD_ERR_INVALID, "rd_classify: d_off and d_logits are required");
    RD_CUDA(h, cudaSetDevice(h->device));
    return rd_classify_device(h, d_seq, d_off, n, max_len, semantics, precision, d_logits, d_probs,
                           d_labels, d_counts, (cudaStream_t)stream);
}

extern "C" int rd_pair_combine(rd_handle* h, const float* d_logits1, const float* d_logits2, int64_t n,
                               int mode, int8_t* d_labels, int64_t* d_counts, void* stream) {
    if (!h) return RD_ERR_INVALID;
    if (n < 0 || mode < RD_PAIR_NONE || mode > RD_PAIR_BOTH)
        return fail(h, RD_ERR_INVALID, "rd_pair_combine: bad n or mode");
    if (n == 0) return RD_OK;
    if (!d_logits1 || !d_logits2) return fail(h, RD_ERR_INVALID, "rd_pair_combine: NULL logits");
    RD_CUDA(h, cudaSetDevice(h->device));
    StageTimer tm(h, 3, (cudaStream_t)stream);
    return rd_launch_pair(h, d_logits1, d_logits2, n, mode, d_labels, d_counts, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// host-buffer pipeline: chunk c → stage c % NSTAGE;  H2D (s_in) → kernels (s_cmp) → D2H (s_out)
static const int64_t CHUNK_READS = (int64_t)1 << 21;      // 2 Mi reads per chunk
static const int64_t FIRST_CHUNK_READS = (int64_t)1 << 18; // the first chunk is small: its H2D is the only copy no kernel hides

static int ensure_stage(rd_handle* h, int64_t n, int64_t bytes, int ends, bool want_probs) {
    if (n <= h->cap_stage_n && bytes <= h->cap_stage_bytes &&
        (ends == 1 || h->d_stage_seq[1][0]) && (!want_probs || h->d_stage_probs[0])) return RD_OK;
    RD_CUDA(h, cudaDeviceSynchronize());
    int64_t nn = std::max(n, h->cap_stage_n), bb = std::max(bytes, h->cap_stage_bytes);
    for (int e = 0; e < 2; ++e)
        for (int s = 0; s < rd_handle::NSTAGE; ++s) {
            cudaFree(h->d_stage_seq[e][s]); cudaFree(h->d_stage_off[e][s]); cudaFree(h->d_stage_logits[e][s]);
            h->d_stage_seq[e][s] = nullptr; h->d_stage_off[e][s] = nullptr; h->d_stage_logits[e][s] = nullptr;
        }
    for (int s = 0; s < rd_handle::NSTAGE; ++s) {
        cudaFree(h->d_stage_probs[s]); cudaFree(h->d_stage_labels[s]);
        h->d_stage_probs[s] = nullptr; h->d_stage_labels[s] = nullptr;
    }
    for (int e = 0; e < ends; ++e)
        for (int s = 0; s < rd_handle::NSTAGE; ++s) {
            RD_CUDA(h, cudaMalloc(&h->d_stage_seq[e][s], std::max<int64_t>(bb, 1)));
            RD_CUDA(h, cudaMalloc(&h->d_stage_off[e][s], sizeof(int64_t) * (nn + 1)));
            RD_CUDA(h, cudaMalloc(&h->d_stage_logits[e][s], sizeof(float) * 2 * std::max<int64_t>(nn, 1)));
        }
    for (int s = 0; s < rd_handle::NSTAGE; ++s) {
        RD_CUDA(h, cudaMalloc(&h->d_stage_probs[s], sizeof(float) * 2 * std::max<int64_t>(nn, 1)));
        RD_CUDA(h, cudaMalloc(&h->d_stage_labels[s], std::max<int64_t>(nn, 1)));
    }
    h->cap_stage_n = nn; h->cap_stage_bytes = bb;
    return RD_OK;
}

// rebase offsets on the device: off[i] -= base   (tiny elementwise kernel)
__global__ void rebase_kernel(int64_t* off, int64_t n1, int64_t base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) off[i] -= base;
}

static int classify_host_run(rd_handle* h, int ends,
                             const uint8_t* const seq[2], const int64_t* const off[2], int64_t n,
                             int max_len, int semantics, int precision, int mode,
                             float* const logits[2], float* probs, int8_t* labels, int64_t* counts) {
    int rc = check_common(h, n, max_len, "rd_classify_host");
    if (rc) return rc;
    if (semantics != RD_SEM_PACKED && semantics != RD_SEM_PADDED)
        return fail(h, RD_ERR_INVALID, "rd_classify_host: unknown semantics");
    if (precision < RD_PREC_FP32 || precision > RD_PREC_LAST)
        return fail(h, RD_ERR_INVALID, "rd_classify_host: unknown precision");
    if (ends == 2 && (mode < RD_PAIR_NONE || mode > RD_PAIR_BOTH))
        return fail(h, RD_ERR_INVALID, "rd_classify_pairs_host: unknown mode");
    if (counts) counts[0] = counts[1] = counts[2] = 0;
    if (n == 0) return RD_OK;
    for (int e = 0; e < ends; ++e)
        if (!off[e] || !seq[e]) return fail(h, RD_ERR_INVALID, "rd_classify_host: NULL input buffer");
    if (!labels) return fail(h, RD_ERR_INVALID, "rd_classify_host: labels is required");
    RD_CUDA(h, cudaSetDevice(h->device));

    // (a chunk is ~2 Mi READS whatever the number of ends: a pair chunk holds half as many units, so that a 2 Mi-pair
    //  call is still cut into pieces whose copies hide behind the kernels of the piece before)
    const int64_t chunk_cap = CHUNK_READS / ends;
    const int64_t chunk = std::min<int64_t>(chunk_cap, n);
    std::vector<int64_t> cut;                                  // chunk c = reads [cut[c], cut[c+1])
    cut.push_back(0);
    if (n > chunk_cap) cut.push_back(FIRST_CHUNK_READS / ends);
    while (cut.back() < n) cut.push_back(std::min(n, cut.back() + chunk));
    int64_t max_bytes = 0;
    for (int e = 0; e < ends; ++e)
        for (size_t c = 0; c + 1 < cut.size(); ++c)
            max_bytes = std::max(max_bytes, off[e][cut[c + 1]] - off[e][cut[c]]);
    rc = ensure_stage(h, chunk, max_bytes, ends, probs != nullptr);
    if (rc) return rc;
    rc = ensure_scratch(h, chunk);
    if (rc) return rc;
    RD_CUDA(h, cudaMemsetAsync(h->d_stage_counts, 0, sizeof(int64_t) * 4, h->s_cmp));

    const int64_t nchunks = (int64_t)cut.size() - 1;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int st = (int)(c % rd_handle::NSTAGE);
        const int64_t s = cut[(size_t)c], t = cut[(size_t)c + 1], m = t - s;
        // reference behaviour: a zero-length read cannot be packed (torch pack_sequence raises).  Checked chunk by
        // chunk so that the scan of chunk c runs while the GPU works on chunk c-1.
        if (semantics == RD_SEM_PACKED)
            for (int e = 0; e < ends; ++e) {
                const int64_t* o = off[e];
                int bad = 0;
                for (int64_t i = s; i < t; ++i) bad |= (o[i + 1] <= o[i]);       // branch-free: vectorises
                if (bad) {
                    int64_t i = s;
                    while (o[i + 1] > o[i]) ++i;
                    cudaDeviceSynchronize();
                    return fail(h, RD_ERR_EMPTY_READ, "zero-length read at index " + std::to_string(i) +
                                                          " cannot be classified under packed semantics");
                }
            }
        // stage buffers are free once the D2H of chunk c-NSTAGE has been issued and finished
        if (c >= rd_handle::NSTAGE) RD_CUDA(h, cudaStreamWaitEvent(h->s_in, h->ev_out[st], 0));
        for (int e = 0; e < ends; ++e) {
            const int64_t b0 = off[e][s], nb = off[e][t] - b0;
            RD_CUDA(h, cudaMemcpyAsync(h->d_stage_seq[e][st], seq[e] + b0, nb, cudaMemcpyHostToDevice, h->s_in));
            RD_CUDA(h, cudaMemcpyAsync(h->d_stage_off[e][st], off[e] + s, sizeof(int64_t) * (m + 1),
                                       cudaMemcpyHostToDevice, h->s_in));
            if (b0 != 0) {
                rebase_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, h->s_in>>>(h->d_stage_off[e][st], m + 1, b0);
                h->launches += 1;
            }
        }
        RD_CUDA(h, cudaEventRecord(h->ev_in[st], h->s_in));
        RD_CUDA(h, cudaStreamWaitEvent(h->s_cmp, h->ev_in[st], 0));
        if (ends == 1) {
            rc = rd_classify_device(h, h->d_stage_seq[0][st], h->d_stage_off[0][st], m, max_len, semantics, precision,
                                 h->d_stage_logits[0][st], probs ? h->d_stage_probs[st] : nullptr,
                                 h->d_stage_labels[st], h->d_stage_counts, h->s_cmp);
            if (rc) return rc;
        } else {
            for (int e = 0; e < 2; ++e) {
                rc = rd_classify_device(h, h->d_stage_seq[e][st], h->d_stage_off[e][st], m, max_len, semantics,
                                     precision, h->d_stage_logits[e][st], nullptr, nullptr, nullptr, h->s_cmp);
                if (rc) return rc;
            }
            if (mode == RD_PAIR_NONE) {
                const uint8_t* sq[2] = {h->d_stage_seq[0][st], h->d_stage_seq[1][st]};
                const int64_t* of[2] = {h->d_stage_off[0][st], h->d_stage_off[1][st]};
                float* lg[2] = {h->d_stage_logits[0][st], h->d_stage_logits[1][st]};
                rc = rd_pair_none_refine(h, sq, of, m, max_len, semantics, precision, lg, h->s_cmp);
                if (rc) return rc;
            }
            rc = rd_launch_pair(h, h->d_stage_logits[0][st], h->d_stage_logits[1][st], m, mode,
                                h->d_stage_labels[st], h->d_stage_counts, h->s_cmp);
            if (rc) return rc;
        }
        RD_CUDA(h, cudaEventRecord(h->ev_cmp[st], h->s_cmp));
        RD_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_cmp[st], 0));
        RD_CUDA(h, cudaMemcpyAsync(labels + s, h->d_stage_labels[st], m, cudaMemcpyDeviceToHost, h->s_out));
        for (int e = 0; e < ends; ++e)
            if (logits[e])
                RD_CUDA(h, cudaMemcpyAsync(logits[e] + 2 * s, h->d_stage_logits[e][st], sizeof(float) * 2 * m,
                                           cudaMemcpyDeviceToHost, h->s_out));
        if (probs)
            RD_CUDA(h, cudaMemcpyAsync(probs + 2 * s, h->d_stage_probs[st], sizeof(float) * 2 * m,
                                       cudaMemcpyDeviceToHost, h->s_out));
        RD_CUDA(h, cudaEventRecord(h->ev_out[st], h->s_out));
    }
    RD_CUDA(h, cudaStreamSynchronize(h->s_cmp));
    int64_t cnt[4];
    RD_CUDA(h, cudaMemcpy(cnt, h->d_stage_counts, sizeof(int64_t) * 3, cudaMemcpyDeviceToHost));
    RD_CUDA(h, cudaStreamSynchronize(h->s_out));
    if (counts) { counts[0] = cnt[0]; counts[1] = cnt[1]; counts[2] = cnt[2]; }
    return RD_OK;
}

// A call that fails half way has copies in flight from and into the CALLER's buffers: drain the device before the error
// is returned, so that the caller may free or reuse them at once.
static int classify_host_impl(rd_handle* h, int ends,
                              const uint8_t* const seq[2], const int64_t* const off[2], int64_t n,
                              int max_len, int semantics, int precision, int mode,
                              float* const logits[2], float* probs, int8_t* labels, int64_t* counts) {
    const int rc = classify_host_run(h, ends, seq, off, n, max_len, semantics, precision, mode, logits, probs, labels, counts);
    if (rc != RD_OK && h && rc != RD_ERR_INVALID) {
        const std::string msg = h->err;            // (keep the first error's message)
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        h->err = msg;
    }
    return rc;
}

extern "C" int rd_classify_host(rd_handle* h, const uint8_t* seq, const int64_t* off, int64_t n,
                                int max_len, int semantics, int precision,
                                float* logits, float* probs, int8_t* labels, int64_t* counts) {
    const uint8_t* seqs[2] = {seq, nullptr};
    const int64_t* offs[2] = {off, nullptr};
    float* lg[2] = {logits, nullptr};
    return classify_host_impl(h, 1, seqs, offs, n, max_len, semantics, precision, RD_PAIR_NONE, lg, probs,
                              labels, counts);
}

extern "C" int rd_classify_pairs_host(rd_handle* h, const uint8_t* seq1, const int64_t* off1,
                                      const uint8_t* seq2, const int64_t* off2, int64_t n,
                                      int max_len, int semantics, int precision, int mode,
                                      float* logits1, float* logits2, int8_t* labels, int64_t* counts) {
    const uint8_t* seqs[2] = {seq1, seq2};
    const int64_t* offs[2] = {off1, off2};
    float* lg[2] = {logits1, logits2};
    return classify_host_impl(h, 2, seqs, offs, n, max_len, semantics, precision, mode, lg, nullptr,
                              labels, counts);
}
