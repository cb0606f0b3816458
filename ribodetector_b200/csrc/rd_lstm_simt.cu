// rd_lstm_simt.cu — K2 (RD_PREC_FP32): forward-direction LSTM on CUDA cores in fp32, with the
// FC tail fused.  This is the on-device fp32 reference the tensor-core kernels are checked
// against at sizes the CPU oracle cannot reach; it is also a complete product path.
//
// Replaces `self.rnn(x, None)` + `last_items` + `self.out` (model/model.py:33-36) for the
// forward direction; the reverse direction enters through the logit LUT (rd_tail.cu).
//
// Work item = half tile (64 reads).  256 threads: warp w owns reads 8w..8w+7, lane l owns hidden
// units 4l..4l+3 (all four gates) → 8x16 fp32 accumulators per thread.  Per step:
//   z[64,512] = h[64,128] . W_hh^T[128,512]     (h broadcast from smem, W_hh^T streamed through L1/L2)
//   gates (+ input table row select), c/h update in registers, h written back to smem.
// Bound: FP32 FFMA pipe (65 536 FMA per read-step) — see DESIGN.md.
#include "rd_common.cuh"

#define SIMT_READS 64
#define SIMT_THREADS 256
#define SIMT_SMEM_HS (2 * SIMT_READS * RD_H * 4)
#define SIMT_SMEM_TAB (5 * RD_G4 * 4)
#define SIMT_SMEM_BYTES (SIMT_SMEM_HS + SIMT_SMEM_TAB + 2 * SIMT_READS + SIMT_READS * 4)

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(SIMT_THREADS, 1)
lstm_simt_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ splan,
                 const int32_t* __restrict__ perm, int L, int n_work,
                 const float4* __restrict__ whh_t4,   // [128][512] as float4
                 const float* __restrict__ tab,       // [5][512]
                 const float* __restrict__ wout,      // [2][256]
                 const float* __restrict__ bout,      // [2]
                 const float* __restrict__ revlut,    // [RD_MAX_LEN][5][2]
                 float* __restrict__ logits, int32_t* __restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    float (*hs)[SIMT_READS][RD_H] = reinterpret_cast<float (*)[SIMT_READS][RD_H]>(smem_dyn);   // [2] 64 KB
    float (*tab_s)[RD_G4] = reinterpret_cast<float (*)[RD_G4]>(smem_dyn + SIMT_SMEM_HS);      // [5] 10 KB
    uint8_t (*code_s)[SIMT_READS] =
        reinterpret_cast<uint8_t (*)[SIMT_READS]>(smem_dyn + SIMT_SMEM_HS + SIMT_SMEM_TAB);     // [2]
    int* nf_s = reinterpret_cast<int*>(smem_dyn + SIMT_SMEM_HS + SIMT_SMEM_TAB + 2 * SIMT_READS);
    __shared__ int s_work;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 5 * RD_G4; i += SIMT_THREADS) (&tab_s[0][0])[i] = tab[i];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_work = atomicAdd(work_counter, 1);
        __syncthreads();
        const int work = s_work;
        if (work >= n_work) break;
        const int64_t tile = work >> 1;
        const int half = work & 1;
        const int64_t slot0 = tile * RD_TILE + half * SIMT_READS;
        const uint8_t* cbase = codes + tile * (int64_t)L * RD_TILE + half * SIMT_READS;

        if (tid < SIMT_READS) nf_s[tid] = (int)PLAN_NFWD(splan[slot0 + tid]);
        for (int i = tid; i < SIMT_READS * RD_H; i += SIMT_THREADS) (&hs[0][0][0])[i] = 0.f;
        __syncthreads();
        const int T = nf_s[0];                         // slots sorted descending
        if (tid < SIMT_READS && T > 0) code_s[0][tid] = cbase[tid];

        float c[8][4], hreg[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[r][j] = 0.f; hreg[r][j] = 0.f; }
        int nf[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) nf[r] = nf_s[warp * 8 + r];
        __syncthreads();

        for (int t = 0; t < T; ++t) {
            const int cur = t & 1, nxt = cur ^ 1;
            // prefetch next step's codes
            if (tid < SIMT_READS && t + 1 < T) code_s[nxt][tid] = cbase[(int64_t)(t + 1) * RD_TILE + tid];

            float acc[8][4][4];                        // [read][gate][unit]
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[r][g][j] = 0.f;

#pragma unroll 1
            for (int k4 = 0; k4 < RD_H / 4; ++k4) {
                float4 hv[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    hv[r] = *reinterpret_cast<const float4*>(&hs[cur][warp * 8 + r][k4 * 4]);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float4 w[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        w[g] = __ldg(&whh_t4[(k4 * 4 + kk) * (RD_G4 / 4) + g * (RD_H / 4) + lane]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const float hk = kk == 0 ? hv[r].x : kk == 1 ? hv[r].y : kk == 2 ? hv[r].z : hv[r].w;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            acc[r][g][0] = fmaf(hk, w[g].x, acc[r][g][0]);
                            acc[r][g][1] = fmaf(hk, w[g].y, acc[r][g][1]);
                            acc[r][g][2] = fmaf(hk, w[g].z, acc[r][g][2]);
                            acc[r][g][3] = fmaf(hk, w[g].w, acc[r][g][3]);
                        }
                    }
                }
            }

#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int rd = warp * 8 + r;
                if (t < nf[r]) {
                    const int code = code_s[cur][rd];
                    const float4 ti = *reinterpret_cast<const float4*>(&tab_s[code][0 * RD_H + lane * 4]);
                    const float4 tf = *reinterpret_cast<const float4*>(&tab_s[code][1 * RD_H + lane * 4]);
                    const float4 tg = *reinterpret_cast<const float4*>(&tab_s[code][2 * RD_H + lane * 4]);
                    const float4 to = *reinterpret_cast<const float4*>(&tab_s[code][3 * RD_H + lane * 4]);
                    const float zi[4] = {ti.x, ti.y, ti.z, ti.w};
                    const float zf[4] = {tf.x, tf.y, tf.z, tf.w};
                    const float zg[4] = {tg.x, tg.y, tg.z, tg.w};
                    const float zo[4] = {to.x, to.y, to.z, to.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float ig = sigmoid_acc(acc[r][0][j] + zi[j]);
                        const float fg = sigmoid_acc(acc[r][1][j] + zf[j]);
                        const float gg = tanhf(acc[r][2][j] + zg[j]);
                        const float og = sigmoid_acc(acc[r][3][j] + zo[j]);
                        c[r][j] = fmaf(fg, c[r][j], ig * gg);
                        hreg[r][j] = og * tanhf(c[r][j]);
                    }
                }
                *reinterpret_cast<float4*>(&hs[nxt][rd][lane * 4]) =
                    make_float4(hreg[r][0], hreg[r][1], hreg[r][2], hreg[r][3]);
            }
            __syncthreads();
        }

        // fused FC tail: logits = W_out[:, :H] . h_fwd + revlut[krev][crev] + b_out  (model.py:36)
        const float4 w0 = *reinterpret_cast<const float4*>(&wout[lane * 4]);
        const float4 w1 = *reinterpret_cast<const float4*>(&wout[2 * RD_H + lane * 4]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float p0 = w0.x * hreg[r][0] + w0.y * hreg[r][1] + w0.z * hreg[r][2] + w0.w * hreg[r][3];
            float p1 = w1.x * hreg[r][0] + w1.y * hreg[r][1] + w1.z * hreg[r][2] + w1.w * hreg[r][3];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                p0 += __shfl_xor_sync(0xffffffffu, p0, d);
                p1 += __shfl_xor_sync(0xffffffffu, p1, d);
            }
            if (lane == 0) {
                const int64_t slot = slot0 + warp * 8 + r;
                const int32_t rd = perm[slot];
                if (rd >= 0) {
                    const uint32_t p = splan[slot];
                    const float* lut = revlut + ((int64_t)PLAN_KREV(p) * 5 + PLAN_CREV(p)) * 2;
                    float l0 = p0 + lut[0] + bout[0];
                    float l1 = p1 + lut[1] + bout[1];
                    if (PLAN_INVALID(p)) { l0 = __int_as_float(0x7fc00000); l1 = l0; }
                    logits[(int64_t)rd * 2 + 0] = l0;
                    logits[(int64_t)rd * 2 + 1] = l1;
                }
            }
        }
    }
}

int rd_launch_lstm_simt(rd_handle* h, int64_t n_tiles, int L, float* d_logits, cudaStream_t st) {
    if (n_tiles == 0) return RD_OK;
    int n_work = (int)(n_tiles * 2);
    int grid = h->sm_count < n_work ? h->sm_count : n_work;
    if (!h->simt_attr_set) {      // per handle: function attributes are per device
        RD_CUDA(h, cudaFuncSetAttribute(lstm_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SIMT_SMEM_BYTES));
        h->simt_attr_set = true;
    }
    lstm_simt_kernel<<<grid, SIMT_THREADS, SIMT_SMEM_BYTES, st>>>(
        h->d_codes, h->d_splan, h->d_perm, L, n_work, reinterpret_cast<const float4*>(h->d_whh_t),
        h->d_tab_f, h->d_wout, h->d_bout, h->d_revlut, d_logits, h->d_ctrl + 1);
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}
