// tc_probe.cu — one-instruction known-answer test of the tcgen05 operand conventions that
// rd_lstm_tc.cu relies on (run on the B200 box; prints max |D - ref| per variant):
//   B operand  : shared memory, K-major, SWIZZLE_NONE canonical layout
//                byte(n,k) = (k/8)*LBO + (n/8)*SBO + (n%8)*16 + (k%8)*2
//   A operand  : variant SS = same canonical layout in shared memory;
//                variant TS = tensor memory, lane m, column k/2, fp16 pair packed little-endian
//   D          : tensor memory fp32, lane m, column n  (tcgen05.ld.32x32b)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 64, K = 32;         // two K=16 instructions

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
    return d;                                   // layout_type = 0 (SWIZZLE_NONE)
}

__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                               // D = f32
    d |= 0u << 7;                               // A = f16
    d |= 0u << 10;                              // B = f16
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int variant) {
    __shared__ __align__(128) __half sA[M * K];
    __shared__ __align__(128) __half sB[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // canonical no-swizzle K-major images
    for (int i = tid; i < M * K; i += 128) {
        int m = i / K, k = i % K;
        int off = (k / 8) * (M * 8) + (m / 8) * 64 + (m % 8) * 8 + (k % 8);   // in halfs
        sA[off] = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        int n = i / K, k = i % K;
        int off = (k / 8) * (N * 8) + (n / 8) * 64 + (n % 8) * 8 + (k % 8);
        sB[off] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");          // generic smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t dcol = tmem;            // D: columns [0, 64)
    const uint32_t acol = tmem + 64;       // A (TS variant): columns [64, 64 + K/2)
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

    if (variant == 1) {
        // each thread = one row m: pack its K halfs into K/2 columns
        uint32_t r[K / 2];
        for (int c = 0; c < K / 2; ++c) {
            __half2 h2 = __halves2half2(A[tid * K + 2 * c], A[tid * K + 2 * c + 1]);
            r[c] = *reinterpret_cast<uint32_t*>(&h2);
        }
        for (int c = 0; c < K / 2; c += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(acol + lane_off + c),
                         "r"(r[c]), "r"(r[c + 1]), "r"(r[c + 2]), "r"(r[c + 3]));
        asm volatile("tcgen05.wait::st.sync.aligned;");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }

    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N);
        for (int kc = 0; kc < K / 16; ++kc) {
            uint64_t bdesc = make_desc(smem_u32(sB) + kc * 2 * (N * 16), N * 16, 128);
            uint32_t acc = kc > 0;
            if (variant == 0) {
                uint64_t adesc = make_desc(smem_u32(sA) + kc * 2 * (M * 16), M * 16, 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol),
                             "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(dcol),
                             "r"(acol + kc * 8), "l"(bdesc), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // everyone waits for the MMA
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(dcol + lane_off + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    srand(1);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.0f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 250.0f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, sizeof(__half) * M * K)); CK(cudaMalloc(&dB, sizeof(__half) * N * K)); CK(cudaMalloc(&dD, sizeof(float) * M * N));
    CK(cudaMemcpy(dA, hA.data(), sizeof(__half) * M * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(__half) * N * K, cudaMemcpyHostToDevice));
    int bad = 0;
    for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemset(dD, 0xFF, sizeof(float) * M * N));
        probe_kernel<<<1, 128>>>(dA, dB, dD, variant);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
        double mx = 0; int nbad = 0;
        for (int i = 0; i < M * N; ++i) { double d = fabs((double)out[i] - ref[i]); if (!(d <= 1e-3)) ++nbad; if (d > mx || d != d) mx = d; }
        printf("variant %d (%s): max|D-ref| = %.3e, mismatches = %d / %d  D[0][0..3] = %f %f %f %f  ref = %f %f %f %f\n", variant,
               variant ? "A in TMEM (TS)" : "A in SMEM (SS)", mx, nbad, M * N, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
        bad += nbad;
    }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad ? 1 : 0;
}
