// rd_tail.cu — K3: softmax / argmax / pair combination / label counts, and the one-time
// reverse-direction logit LUT.
//
// Replaces: torch.argmax(output, dim=1)                 detect.py:288,481      (ties → 0)
//           Predictor.separate_paired_reads label rule   detect.py:616-663
//           the count accumulation                       detect.py:193-194,292-293
// HBM-bound: 8 B in, 9 B out per read.
#include "rd_common.cuh"

__device__ __forceinline__ void count_labels(int label, bool valid, int64_t* counts) {
    // warp-aggregated: one atomic per warp per class
    unsigned m0 = __ballot_sync(0xffffffffu, valid && label == 0);
    unsigned m1 = __ballot_sync(0xffffffffu, valid && label == 1);
    unsigned mu = __ballot_sync(0xffffffffu, valid && label < 0);
    if ((threadIdx.x & 31) == 0 && counts) {
        if (m0) atomicAdd(reinterpret_cast<unsigned long long*>(counts + 0), (unsigned long long)__popc(m0));
        if (m1) atomicAdd(reinterpret_cast<unsigned long long*>(counts + 1), (unsigned long long)__popc(m1));
        if (mu) atomicAdd(reinterpret_cast<unsigned long long*>(counts + 2), (unsigned long long)__popc(mu));
    }
}

__global__ void __launch_bounds__(256)
tail_kernel(const float2* __restrict__ logits, int64_t n, float2* __restrict__ probs,
            int8_t* __restrict__ labels, int64_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    int label = 0;
    if (valid) {
        float2 l = logits[i];
        label = l.y > l.x ? 1 : 0;                  // first max on ties, like torch.argmax
        if (probs) {
            float m = fmaxf(l.x, l.y);
            float e0 = expf(l.x - m), e1 = expf(l.y - m);
            float inv = 1.0f / (e0 + e1);
            probs[i] = make_float2(e0 * inv, e1 * inv);
        }
        if (labels) labels[i] = (int8_t)label;
    }
    count_labels(label, valid, counts);
}

__global__ void __launch_bounds__(256)
pair_kernel(const float2* __restrict__ l1, const float2* __restrict__ l2, int64_t n, int mode,
            int8_t* __restrict__ labels, int64_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    int label = 0;
    if (valid) {
        float2 a = l1[i], b = l2[i];
        int la = a.y > a.x ? 1 : 0, lb = b.y > b.x ? 1 : 0;
        if (mode == RD_PAIR_RRNA) label = (la & lb);
        else if (mode == RD_PAIR_NORRNA) label = (la | lb);
        else if (mode == RD_PAIR_BOTH) label = (la == lb) ? la : -1;
        else {
            float s0 = a.x + b.x, s1 = a.y + b.y;   // fp32 add of logits, detect.py:657
            label = s1 > s0 ? 1 : 0;
        }
        if (labels) labels[i] = (int8_t)label;
    }
    count_labels(label, valid, counts);
}

// Reverse direction.  The classifier reads the BiLSTM output only at the last valid step, where
// the reverse LSTM has consumed k zero-input steps (k = 0 under packed semantics) plus one
// base: its hidden state is a function of (k, code) only.  Computed once in fp64 and folded
// through W_out[:, H:2H] into logit offsets lut[k][code][2].   
__global__ void __launch_bounds__(1024, 1)
reverse_lut_kernel(const float* __restrict__ whh_r_t,  // [H][4H]
                   const float* __restrict__ tab_r,    // [5][4H]
                   const float* __restrict__ wout,     // [2][2H]
                   int H, int k0, int k1, double* __restrict__ state,   // rows [k0, k1); state = (h, c) after k0 zero steps
                   float* __restrict__ lut) {
    extern __shared__ double lut_sm[];                // h[H] | c[H] | z[5][4H] | hrev[5][H]
    const int G4 = 4 * H;
    double* h = lut_sm;
    double* c = h + H;
    double* z = c + H;
    double* hrev = z + 5 * G4;
    const int j = threadIdx.x;                       // gate row (blockDim.x = 4H)
    if (j < H) { h[j] = k0 ? state[j] : 0.0; c[j] = k0 ? state[H + j] : 0.0; }
    __syncthreads();
    for (int k = k0; k < k1; ++k) {
        double dot = 0.0;
        for (int m = 0; m < H; ++m) dot += (double)whh_r_t[m * G4 + j] * h[m];
#pragma unroll
        for (int code = 0; code < 5; ++code) z[code * G4 + j] = dot + (double)tab_r[code * G4 + j];
        __syncthreads();
        double c_next = 0.0, h_next = 0.0;
        if (j < H) {
#pragma unroll
            for (int code = 0; code < 5; ++code) {
                const double* zc = z + code * G4;
                double ig = 1.0 / (1.0 + exp(-zc[j]));
                double fg = 1.0 / (1.0 + exp(-zc[H + j]));
                double gg = tanh(zc[2 * H + j]);
                double og = 1.0 / (1.0 + exp(-zc[3 * H + j]));
                double cc = fg * c[j] + ig * gg;
                double hh = og * tanh(cc);
                hrev[code * H + j] = hh;
                if (code == 4) { c_next = cc; h_next = hh; }
            }
        }
        __syncthreads();
        if (j < 10) {                                 // 5 codes x 2 classes
            int code = j >> 1, cls = j & 1;
            double s = 0.0;
            for (int u = 0; u < H; ++u) s += (double)wout[cls * 2 * H + H + u] * hrev[code * H + u];
            lut[((int64_t)k * 5 + code) * 2 + cls] = (float)s;
        }
        if (j < H) { h[j] = h_next; c[j] = c_next; }
        __syncthreads();
    }
    if (j < H) { state[j] = h[j]; state[H + j] = c[j]; }       // the chain resumes here when the table is extended
}

// ---- TC_AUTO: order-preserving compaction of the slots whose fast-pass margin is inside the band -------------
// (slots are sorted by step count; keeping their order keeps the compacted tiles length-bucketed)
// (mate != NULL: the band is on the SUMMED margin of a read and its mate, the quantity RD_PAIR_NONE decides on)
__device__ __forceinline__ bool in_band(const int32_t* perm, const float2* logits, const float2* mate, int64_t s, float tau) {
    const int32_t rd = perm[s];
    if (rd < 0) return false;
    const float2 l = logits[rd];
    float m = l.y - l.x;
    if (mate) { const float2 o = mate[rd]; m += o.y - o.x; }
    return fabsf(m) < tau;                                // NaN (invalid read) compares false
}

__global__ void __launch_bounds__(256)
band_count_kernel(const int32_t* __restrict__ perm, const float2* __restrict__ logits, const float2* __restrict__ mate,
                  int64_t slots, float tau, int64_t* __restrict__ band) {
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool f = s < slots && in_band(perm, logits, mate, s, tau);
    const int c = __syncthreads_count(f);
    if (threadIdx.x == 0) band[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
band_scan_kernel(int64_t* __restrict__ band, int64_t nb, int32_t* __restrict__ perm2,
                 uint32_t* __restrict__ splan2) {   // exclusive scan in place; band[nb] = total; pads the compacted tables
    __shared__ int64_t part[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nb ? band[i] : 0;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const int64_t a = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
            __syncthreads();
            part[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < nb) band[i] = carry + part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) band[nb] = carry;
    __syncthreads();
    // "no read" entries behind the compacted slots: the rest of the last tile plus one pad tile for the CTA pair
    const int64_t total = carry;
    if (threadIdx.x < 2 * RD_TILE) { perm2[total + threadIdx.x] = -1; splan2[total + threadIdx.x] = 0u; }
}

__global__ void __launch_bounds__(256)
band_write_kernel(const int32_t* __restrict__ perm, const uint32_t* __restrict__ splan, const float2* __restrict__ logits,
                  const float2* __restrict__ mate, int64_t slots, float tau, const int64_t* __restrict__ band,
                  int32_t* __restrict__ perm2, uint32_t* __restrict__ splan2) {
    __shared__ int warp_cnt[8];
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool f = s < slots && in_band(perm, logits, mate, s, tau);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[w] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int i = 0; i < w; ++i) before += warp_cnt[i];
    if (f) {
        const int64_t dst = band[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
        perm2[dst] = perm[s];
        splan2[dst] = splan[s];
    }
}

int rd_launch_band_select(rd_handle* h, const float* d_logits, int64_t n_tiles, float tau, cudaStream_t st,
                          const float* d_mate_logits) {
    const int64_t slots = n_tiles * RD_TILE;
    const int64_t nb = (slots + 255) / 256;
    const float2* lg = reinterpret_cast<const float2*>(d_logits);
    const float2* mt = reinterpret_cast<const float2*>(d_mate_logits);
    band_count_kernel<<<(unsigned)nb, 256, 0, st>>>(h->d_perm, lg, mt, slots, tau, h->d_band);
    band_scan_kernel<<<1, 1024, 0, st>>>(h->d_band, nb, h->d_perm2, h->d_splan2);
    band_write_kernel<<<(unsigned)nb, 256, 0, st>>>(h->d_perm, h->d_splan, lg, mt, slots, tau, h->d_band, h->d_perm2, h->d_splan2);
    h->launches += 3;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

int rd_launch_tail(rd_handle* h, const float* d_logits, int64_t n, float* d_probs, int8_t* d_labels,
                   int64_t* d_counts, cudaStream_t st) {
    if (n == 0) return RD_OK;
    tail_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float2*>(d_logits), n, reinterpret_cast<float2*>(d_probs), d_labels,
        d_counts);
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

int rd_launch_pair(rd_handle* h, const float* d_l1, const float* d_l2, int64_t n, int mode,
                   int8_t* d_labels, int64_t* d_counts, cudaStream_t st) {
    if (n == 0) return RD_OK;
    pair_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float2*>(d_l1), reinterpret_cast<const float2*>(d_l2), n, mode,
        d_labels, d_counts);
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

// The table is a serial chain of fp64 LSTM steps on one CTA (8.5 us per row): rd_create builds the rows every packed call
// (row 0) and every padded call up to -l 512 needs; a padded call with a longer -l extends it from the saved state.
int rd_build_reverse_lut(rd_handle* h, int rows, cudaStream_t st) {
    if (rows > RD_MAX_LEN) rows = RD_MAX_LEN;
    if (rows <= h->lut_rows) return RD_OK;
    const int H = h->hidden;
    const size_t smem = sizeof(double) * (size_t)(27 * H);           // h, c, z[5][4H], hrev[5][H]
    if (!h->lut_attr_set) {
        // (the attribute belongs to the function on this device, not to the handle: always the largest size, H = 256)
        RD_CUDA(h, cudaFuncSetAttribute(reverse_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 27 * 256)));
        h->lut_attr_set = true;
    }
    reverse_lut_kernel<<<1, 4 * H, smem, st>>>(h->d_whh_r_t, h->d_tab_r, h->d_wout, H, h->lut_rows, rows, h->d_lutstate, h->d_revlut);
    h->launches += 1;
    h->lut_rows = rows;
    RD_CUDA(h, cudaGetLastError());
    RD_CUDA(h, cudaStreamSynchronize(st));      // rare (first use of a longer -l): later calls may come on other streams
    return RD_OK;
}
