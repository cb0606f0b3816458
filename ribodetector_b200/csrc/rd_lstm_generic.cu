// rd_lstm_generic.cu — K2 for hidden sizes other than 128 (any multiple of 32 up to 256): forward-direction LSTM on
// CUDA cores in fp32 with the FC tail fused.  The reference's SeqModel(**arch.args) takes any hidden_size
// (model/model.py:11-29); the tensor-core kernels (rd_lstm_tc.cu) and the tuned fp32 kernel (rd_lstm_simt.cu) are laid
// out for the shipped H = 128, so every precision of a handle with another H runs here.
//
// A CTA of 4H threads owns a group of 16 consecutive slots (the plan sorts slots by step count, so a group's reads have
// equal or similar lengths) and reads the caller's sequence bytes in place.  Per step:
//   phase 1: thread j = gate row j: z[r][j] = tab[code_r][j] + sum_k W_hh^T[k][j] * h[k][r] for the 16 reads — one
//            coalesced weight load (L1/L2; the [H][4H] image is at most 1 MB) and four 16-byte broadcast loads of h
//            per 16 FMAs; h is kept [k][read] in shared memory for that
//   phase 2: thread j = (unit j mod H, reads 4(j/H)..4(j/H)+3): accurate expf/tanhf gates, c in registers, h back to
//            shared memory — all 4H threads take part, so small H is not bound by a quarter-full cell phase.
// Bound: the FP32 pipe (8·H² FMA per read-step) against L2 weight traffic (16·H² B per group-step).
//
// Replaces `self.rnn(x, None)` + `last_items` + `self.out` (model/model.py:33-36) for the forward direction; the reverse
// direction enters through the logit LUT (rd_tail.cu), exactly as in the other kernels.
#include <algorithm>
#include "rd_common.cuh"

namespace {

constexpr int GEN_READS = 16;           // reads per work group
constexpr int GEN_Q = GEN_READS / 4;    // reads per thread in the cell phase

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(1024, 1)
lstm_generic_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off, int ostride,
                    const uint32_t* __restrict__ splan, const int32_t* __restrict__ perm, int L, int64_t n_slots, int H,
                    const float* __restrict__ whh_t,    // [H][4H]
                    const float* __restrict__ tab,      // [5][4H]
                    const float* __restrict__ wout,     // [2][2H]
                    const float* __restrict__ bout, const float* __restrict__ revlut, float* __restrict__ logits) {
    extern __shared__ __align__(16) float sm[];
    const int G4 = 4 * H;
    float* h_s = sm;                                  // [H][GEN_READS]
    float* z_s = sm + GEN_READS * H;                  // [GEN_READS][4H]
    __shared__ int nf_s[GEN_READS], len_s[GEN_READS], rd_s[GEN_READS];
    __shared__ int64_t beg_s[GEN_READS];
    __shared__ uint32_t plan_s[GEN_READS], code_s[GEN_READS];
    const int j = threadIdx.x;
    const int u = j % H, q = j / H;                   // cell phase: unit u of reads GEN_Q*q .. GEN_Q*q + GEN_Q-1
    const int64_t n_groups = (n_slots + GEN_READS - 1) / GEN_READS;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        __syncthreads();
        if (j < GEN_READS) {
            const int64_t slot = grp * GEN_READS + j;
            const uint32_t p = slot < n_slots ? splan[slot] : 0u;
            const int32_t rd = slot < n_slots ? perm[slot] : -1;
            plan_s[j] = p; rd_s[j] = rd;
            nf_s[j] = rd >= 0 ? (int)PLAN_NFWD(p) : 0;
            const int64_t b = rd >= 0 ? off[(int64_t)rd * ostride] : 0;
            const int64_t l = rd >= 0 ? off[(int64_t)rd * ostride + 1] - b : 0;
            beg_s[j] = b; len_s[j] = (int)(l < (int64_t)L ? l : (int64_t)L);
        }
        for (int i = j; i < GEN_READS * H; i += blockDim.x) h_s[i] = 0.f;
        __syncthreads();
        const int T = nf_s[0];                        // slots are sorted by step count, descending
        float c[GEN_Q];
        int nf[GEN_Q];
#pragma unroll
        for (int r = 0; r < GEN_Q; ++r) { c[r] = 0.f; nf[r] = nf_s[GEN_Q * q + r]; }
        for (int t = 0; t < T; ++t) {
            if (j < GEN_READS) code_s[j] = t < len_s[j] ? rd_base_code(seq[beg_s[j] + t]) : 4u;
            __syncthreads();
            float acc[GEN_READS];
#pragma unroll
            for (int r = 0; r < GEN_READS; ++r) acc[r] = tab[code_s[r] * G4 + j];
#pragma unroll 4
            for (int k = 0; k < H; ++k) {
                const float w = __ldg(whh_t + (int64_t)k * G4 + j);
                const float4* hk = reinterpret_cast<const float4*>(h_s + k * GEN_READS);
#pragma unroll
                for (int r4 = 0; r4 < GEN_READS / 4; ++r4) {
                    const float4 hv = hk[r4];
                    acc[4 * r4 + 0] = fmaf(w, hv.x, acc[4 * r4 + 0]);
                    acc[4 * r4 + 1] = fmaf(w, hv.y, acc[4 * r4 + 1]);
                    acc[4 * r4 + 2] = fmaf(w, hv.z, acc[4 * r4 + 2]);
                    acc[4 * r4 + 3] = fmaf(w, hv.w, acc[4 * r4 + 3]);
                }
            }
#pragma unroll
            for (int r = 0; r < GEN_READS; ++r) z_s[r * G4 + j] = acc[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < GEN_Q; ++r) {
                if (t < nf[r]) {
                    const int rr = GEN_Q * q + r;
                    const float* z = z_s + rr * G4;
                    const float ig = sigmoid_acc(z[u]), fg = sigmoid_acc(z[H + u]);
                    const float gg = tanhf(z[2 * H + u]), og = sigmoid_acc(z[3 * H + u]);
                    c[r] = fmaf(fg, c[r], ig * gg);
                    h_s[u * GEN_READS + rr] = og * tanhf(c[r]);
                }
            }
            __syncthreads();
        }
        // FC tail: logits = W_out[:, :H] . h_fwd + revlut[krev][crev] + b_out   (model.py:36); one warp per read
        const int nwarp = blockDim.x >> 5, lane = j & 31;
        for (int r = j >> 5; r < GEN_READS; r += nwarp) {
            if (rd_s[r] < 0) continue;
            float a0 = 0.f, a1 = 0.f;
            for (int v = lane; v < H; v += 32) {
                const float hv = h_s[v * GEN_READS + r];
                a0 = fmaf(wout[v], hv, a0);
                a1 = fmaf(wout[2 * H + v], hv, a1);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d); }
            if (lane == 0) {
                const uint32_t p = plan_s[r];
                const float* lut = revlut + ((int64_t)PLAN_KREV(p) * 5 + PLAN_CREV(p)) * 2;
                float l0 = a0 + lut[0] + bout[0], l1 = a1 + lut[1] + bout[1];
                if (PLAN_INVALID(p)) { l0 = __int_as_float(0x7fc00000); l1 = l0; }
                *reinterpret_cast<float2*>(logits + (int64_t)rd_s[r] * 2) = make_float2(l0, l1);
            }
        }
    }
}

}  // namespace

int rd_launch_lstm_generic(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n_tiles, int max_len,
                           float* d_logits, cudaStream_t st, int ostride) {
    if (n_tiles == 0) return RD_OK;
    const int H = h->hidden;
    const size_t smem = sizeof(float) * GEN_READS * 5 * H;
    if (!h->generic_attr_set) {
        RD_CUDA(h, cudaFuncSetAttribute(lstm_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        RD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_generic_kernel, 4 * H, smem));
        h->generic_ctas_per_sm = per_sm > 0 ? per_sm : 1;
        h->generic_attr_set = true;
    }
    const int64_t n_slots = n_tiles * RD_TILE;
    const int64_t groups = (n_slots + GEN_READS - 1) / GEN_READS;
    const int grid = (int)std::min<int64_t>(groups, (int64_t)h->sm_count * h->generic_ctas_per_sm);
    lstm_generic_kernel<<<grid, 4 * H, smem, st>>>(d_seq, d_off, ostride, h->d_splan, h->d_perm, max_len, n_slots, H, h->d_whh_t,
                                                    h->d_tab_f, h->d_wout, h->d_bout, h->d_revlut, d_logits);
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}
