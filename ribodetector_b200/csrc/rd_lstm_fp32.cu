// rd_lstm_fp32.cu — K2 on CUDA cores in fp32 (RD_PREC_FP32, and EVERY precision of a handle whose hidden size is not
// the shipped 128): forward-direction LSTM with the FC tail fused, for any hidden size that is a multiple of 32 up to
// 256.  The reference's SeqModel(**arch.args) takes any hidden_size (model/model.py:11-29); the tensor-core kernels
// (rd_lstm_tc.cu) are laid out for H = 128.  This is also the on-device fp32 reference the tensor-core modes are
// checked against at sizes the CPU oracle cannot reach (fp32 FFMA, expf-based gates, no reduced-precision step).
//
// A CTA of H x S threads (S = 512 / H read subgroups) owns 16 S consecutive slots (the plan sorts slots by step count,
// so a group's reads have equal or similar lengths) and reads the caller's sequence bytes in place.  Thread (u, s) owns
// hidden unit u — all four gates — of the 16 reads of subgroup s: 64 fp32 accumulators and 16 cell states in registers.
// Per step and k: ONE 16-byte weight load (the image [k][u][i,f,g,o], coalesced, L1/L2 — prefetched one k ahead) and four
// 16-byte broadcast loads of h (kept [k][read] in shared memory, double-buffered over the steps) feed 64 FMAs; the gates
// and the cell update follow in the same thread with no exchange, and the 16 new h values go back to shared memory as
// four conflict-free 16-byte stores.  One barrier per step, over the subgroup's H threads only.
// Bound: the FP32 pipe (4·H² FMA per read-step).
//
// Replaces `self.rnn(x, None)` + `last_items` + `self.out` (model/model.py:33-36) for the forward direction; the reverse
// direction enters through the logit LUT (rd_tail.cu), exactly as in the other kernels.
#include <algorithm>
#include "rd_common.cuh"

namespace {

constexpr int GEN_C = 16;               // reads per thread
constexpr int GEN_THREADS = 512;        // at most (S = GEN_THREADS / H subgroups of H threads)
constexpr size_t GEN_SMEM_MAX = 100 * 1024;   // >= the dynamic shared memory of every supported H
constexpr int GEN_PAD = 4;              // h row = 16 S + 4 floats: 16-byte stores of consecutive units hit distinct banks

// Gates: expf (2 ulp) and an approximate-reciprocal division (2 ulp), branch-free.  tanh as 1 - 2 / (e^2x + 1): absolute
// error <= 2e-7 everywhere (the relative error grows for |x| -> 0, where the value itself vanishes) — the same size as
// the rounding of the H-term fp32 dot products that feed it; e^2x = inf gives 1, e^2x = 0 gives -1.
// (Measured and dropped: the FMAs issued as packed fma.rn.f32x2 — 27 % fewer issue slots in the k loop, bit-identical
// results, no change in time (43.4 vs 43.1 TFLOP/s at H = 128): the loop is not bound by issue slots alone; and the gates
// straight from ex2.approx / rcp.approx — 5 % faster, but the kernel's error against the
// fp64 oracle grows from 1.15e-5 to 1.8e-5 at 100 bp, and this kernel is the arbiter of the tensor-core modes.)
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) { return 1.0f - __fdividef(2.0f, expf(2.0f * x) + 1.0f); }

__global__ void __launch_bounds__(GEN_THREADS, 1)
lstm_fp32_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off, int ostride,
                    const uint32_t* __restrict__ splan, const int32_t* __restrict__ perm, int L, int64_t n_slots, int H, int S,
                    const float4* __restrict__ whh_g4,  // [H + 1 (k; the last row is padding)][H (u)] {i, f, g, o}
                    const float* __restrict__ tab,      // [5][4H]
                    const float* __restrict__ wout,     // [2][2H]
                    const float* __restrict__ bout, const float* __restrict__ revlut, float* __restrict__ logits) {
    extern __shared__ __align__(16) float sm[];
    const int RPC = GEN_C * S;                        // reads per CTA
    const int HROW = RPC + GEN_PAD;
    float* h_s = sm;                                  // [2][H][HROW]
    float* tab_s = sm + 2 * H * HROW;                 // [5][4H]
    int* nf_s = reinterpret_cast<int*>(tab_s + 20 * H);          // [RPC]
    int* len_s = nf_s + RPC;                          // [RPC]
    int* rd_s = len_s + RPC;                          // [RPC]
    uint32_t* plan_s = reinterpret_cast<uint32_t*>(rd_s + RPC);  // [RPC]
    uint32_t* code_s = plan_s + RPC;                  // [2][RPC]
    int64_t* beg_s = reinterpret_cast<int64_t*>(code_s + 2 * RPC);   // [RPC]  (8-byte aligned: every array above holds a multiple of 16 words)
    const int tid = threadIdx.x;
    const int u = tid % H, s = tid / H;               // s is warp-uniform (H is a multiple of 32)
    const int r0 = GEN_C * s;
    for (int i = tid; i < 20 * H; i += blockDim.x) tab_s[i] = tab[i];
    const int64_t n_groups = (n_slots + RPC - 1) / RPC;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        __syncthreads();
        for (int j = tid; j < RPC; j += blockDim.x) {
            const int64_t slot = grp * RPC + j;
            const uint32_t p = slot < n_slots ? splan[slot] : 0u;
            const int32_t rd = slot < n_slots ? perm[slot] : -1;
            plan_s[j] = p; rd_s[j] = rd;
            const int nf = rd >= 0 ? (int)PLAN_NFWD(p) : 0;
            nf_s[j] = nf;
            const int64_t b = rd >= 0 ? off[(int64_t)rd * ostride] : 0;
            const int64_t l = rd >= 0 ? off[(int64_t)rd * ostride + 1] - b : 0;
            const int len = (int)(l < (int64_t)L ? l : (int64_t)L);
            beg_s[j] = b; len_s[j] = len;
            code_s[j] = (nf > 0 && len > 0) ? rd_base_code(seq[b]) : 4u;
        }
        for (int i = tid; i < H * HROW; i += blockDim.x) h_s[i] = 0.f;           // buffer 0 = h_{-1}
        __syncthreads();
        // From here to the FC tail a subgroup runs on its own: its 16 reads' h rows are written and read by its H threads
        // only, so the per-step barrier covers just those (a named barrier; one warp when H = 32) and the subgroups of a CTA
        // drift apart — one's gate arithmetic overlaps another's FMA loop, and a subgroup of shorter reads finishes early.
        const int T = nf_s[r0];                       // slots are sorted by step count, descending
        float c[GEN_C];
#pragma unroll
        for (int r = 0; r < GEN_C; ++r) c[r] = 0.f;
        for (int t = 0; t < T; ++t) {
            const int cur = t & 1;
            const float* hc = h_s + cur * H * HROW + r0;
            float* hn = h_s + (cur ^ 1) * H * HROW;
            // next step's base codes (read after this step's barrier)
            if (u < GEN_C && t + 1 < T) {
                const int j = r0 + u;
                code_s[(cur ^ 1) * RPC + j] = t + 1 < len_s[j] ? rd_base_code(seq[beg_s[j] + t + 1]) : 4u;
            }
            float acc[4][GEN_C];
#pragma unroll
            for (int r = 0; r < GEN_C; ++r) {
                const float* tr = tab_s + code_s[cur * RPC + r0 + r] * 4 * H + u;
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[g][r] = tr[g * H];
            }
            const float4* wp = whh_g4 + u;                    // row k of the image; row H is padding for the last prefetch
            const float* hp = hc;
            float4 w = __ldg(wp);
#pragma unroll 2
            for (int k = 0; k < H; ++k) {
                wp += H;
                const float4 wn = __ldg(wp);                  // one k ahead
                const float4* hk = reinterpret_cast<const float4*>(hp);
                hp += HROW;
#pragma unroll
                for (int r4 = 0; r4 < GEN_C / 4; ++r4) {
                    const float4 hv = hk[r4];
                    const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[0][4 * r4 + q] = fmaf(w.x, hh[q], acc[0][4 * r4 + q]);
                        acc[1][4 * r4 + q] = fmaf(w.y, hh[q], acc[1][4 * r4 + q]);
                        acc[2][4 * r4 + q] = fmaf(w.z, hh[q], acc[2][4 * r4 + q]);
                        acc[3][4 * r4 + q] = fmaf(w.w, hh[q], acc[3][4 * r4 + q]);
                    }
                }
                w = wn;
            }
#pragma unroll
            for (int r4 = 0; r4 < GEN_C / 4; ++r4) {
                const float4 ho4 = *reinterpret_cast<const float4*>(hc + u * HROW + 4 * r4);   // a finished read keeps its state
                const float ho[4] = {ho4.x, ho4.y, ho4.z, ho4.w};
                float hv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = 4 * r4 + q;
                    const bool act = t < nf_s[r0 + r];
                    const float ig = sigmoid_acc(acc[0][r]), fg = sigmoid_acc(acc[1][r]);
                    const float gg = tanh_acc(acc[2][r]), og = sigmoid_acc(acc[3][r]);
                    const float cn = fmaf(fg, c[r], ig * gg);
                    const float hn1 = og * tanh_acc(cn);
                    c[r] = act ? cn : c[r];
                    hv[q] = act ? hn1 : ho[q];
                }
                *reinterpret_cast<float4*>(hn + u * HROW + r0 + 4 * r4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            }
            if (H == 32) __syncwarp();
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "r"(H) : "memory");
        }
        __syncthreads();
        // FC tail: logits = W_out[:, :H] . h_fwd + revlut[krev][crev] + b_out   (model.py:36); one warp per read
        const int nwarp = blockDim.x >> 5, lane = tid & 31;
        for (int r = tid >> 5; r < RPC; r += nwarp) {
            if (rd_s[r] < 0) continue;
            const float* hf = h_s + (nf_s[r & ~(GEN_C - 1)] & 1) * H * HROW;      // the buffer the read's subgroup wrote last
            float a0 = 0.f, a1 = 0.f;
            for (int v = lane; v < H; v += 32) {
                const float hv = hf[v * HROW + r];
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

int rd_launch_lstm_fp32(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n_tiles, int max_len,
                           float* d_logits, cudaStream_t st, int ostride) {
    if (n_tiles == 0) return RD_OK;
    const int H = h->hidden;
    const int S = std::max(1, GEN_THREADS / H);
    const int RPC = GEN_C * S;
    const size_t smem = sizeof(float) * (2 * (size_t)H * (RPC + GEN_PAD) + 20 * (size_t)H) + sizeof(int) * 6 * (size_t)RPC +
                        sizeof(int64_t) * (size_t)RPC;
    if (smem > GEN_SMEM_MAX) { h->err = "rd_lstm_fp32: shared memory budget exceeded"; return RD_ERR_UNSUPPORTED; }
    if (!h->fp32_attr_set) {
        // the attribute belongs to the function on this device, not to the handle: handles of different hidden sizes share
        // it, so it is always set to the largest size any H needs (H = 256: 95 KB)
        RD_CUDA(h, cudaFuncSetAttribute(lstm_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_MAX));
        int per_sm = 1;
        RD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_fp32_kernel, H * S, smem));
        h->fp32_ctas_per_sm = per_sm > 0 ? per_sm : 1;
        h->fp32_attr_set = true;
    }
    const int64_t n_slots = n_tiles * RD_TILE;
    const int64_t groups = (n_slots + RPC - 1) / RPC;
    const int grid = (int)std::min<int64_t>(groups, (int64_t)h->sm_count * h->fp32_ctas_per_sm);
    lstm_fp32_kernel<<<grid, H * S, smem, st>>>(d_seq, d_off, ostride, h->d_splan, h->d_perm, max_len, n_slots, H, S,
                                                   reinterpret_cast<const float4*>(h->d_whh_g4), h->d_tab_f, h->d_wout, h->d_bout,
                                                   h->d_revlut, d_logits);
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}
