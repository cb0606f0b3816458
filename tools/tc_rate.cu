// tc_rate.cu — micro-benchmarks behind the K2 design numbers in DESIGN.md (run on the B200 box):
//   1. cycles per tcgen05.mma (kind::f16, M=128/256, K=16) as a function of N, A from TMEM or SMEM
//   2. MUFU throughput per SM: tanh.approx.f32, ex2+rcp, tanh.approx.f16x2
//   3. tcgen05.ld 32x32b.x32 throughput per SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_rate tc_rate.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---- 1. MMA rate ------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int a_in_tmem, int iters, long long* out, int nd = 2) {
    extern __shared__ __align__(1024) unsigned char smem[];     // B: [2 k-groups][256 rows][16 B] = 8 KB, A: [2][128][16] = 4 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < 12288 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;   // all ones
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    // zero the A region in TMEM (cols 448..463)
    {
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tmem + lane_off + 448), "r"(0x3C003C00u));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tmem + lane_off + 452), "r"(0x3C003C00u));
        asm volatile("tcgen05.wait::st.sync.aligned;");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (warp == 0 && rank == 0) {
        uint32_t el; asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(el));
        const bool elected = el != 0;
        const uint32_t idesc = make_idesc(128 * CG, N);
        const uint64_t bdesc = make_desc(smem_u32(smem), 256 * 16, 128);
        const uint64_t adesc = make_desc(smem_u32(smem) + 8192, 128 * 16, 128);
        long long t0 = clock64();
        int dsel = 0;
        for (int i = 0; i < iters; ++i) {
            dsel = dsel + 1 == nd ? 0 : dsel + 1;
            const uint32_t d = tmem + (uint32_t)(dsel * N);   // rotate over nd independent accumulators
            if (!elected) continue;
            if (a_in_tmem) {
                if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(tmem + 448), "l"(bdesc), "r"(idesc), "r"(1u));
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(tmem + 448), "l"(bdesc), "r"(idesc), "r"(1u));
            } else {
                if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u));
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u));
            }
        }
        if (!elected) {} else if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (elected) out[blockIdx.x / CG] = t1 - t0;
    } else if (CG == 2 && tid == 0) {
        mbar_wait(smem_u32(&bar), 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;");
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

// ---- 2. MUFU rates ----------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 1) mufu_kernel(int iters, float* sink, long long* cyc) {
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = 0.001f * (threadIdx.x + 1) + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
            else if (MODE == 1) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j])); }
            else if (MODE == 2) { uint32_t u = __float_as_uint(x[j]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); x[j] = __uint_as_float(u); }
            else if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
            else if (MODE == 5) { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4427f * x[j])); asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(x[j]) : "f"(1.0f + e)); }
            else if (MODE == 6) { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4427f * x[j])); x[j] = __fdividef(1.0f, 1.0f + e); }
            else if (MODE == 7) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j])); x[j] += 1.5f; }
            else if (MODE == 4) { uint32_t u = __float_as_uint(x[j]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); x[j] = __uint_as_float(u); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 2b. FMA / ALU pipe rates ------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(512, 1) fma_kernel(int iters, float* sink, long long* cyc) {
    float x[8], a = 1.0001f + threadIdx.x * 1e-7f, b = 0.5f;
    for (int j = 0; j < 8; ++j) x[j] = 0.001f * (threadIdx.x + 1) + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(a), "f"(b));
            else if (MODE == 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(a));
            else { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(a), "f"(b)); asm volatile("min.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(a)); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 3. tcgen05.ld rate -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) ldtm_kernel(int iters, int nwarps, uint32_t* sink, long long* cyc) {
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tmem + (uint32_t)((i & 7) * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            acc ^= v[0] ^ v[31];
        }
    }
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_s));
}

int main() {
    long long* d_cyc; float* d_sink;
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024)); CK(cudaMalloc(&d_sink, sizeof(float) * 256 * 1024));
    long long h[8];
    const int iters = 2000;
    CK(cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    CK(cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    printf("== cycles per tcgen05.mma kind::f16 K=16 (back-to-back, one issuing thread, 1 CTA or 1 pair) ==\n");
    int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    for (int cg = 1; cg <= 2; ++cg)
        for (int a_tm = 1; a_tm >= 0; --a_tm)
            for (int N : Ns) {
                if (cg == 1) mma_rate_kernel<1><<<1, 128, 16384>>>(N, a_tm, iters, d_cyc, N <= 192 ? 2 : 1);
                else {
                    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 16384;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<2>, N, a_tm, iters, d_cyc, 2));
                }
                CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
                printf("cta_group::%d M=%d A-in-%s N=%3d : %7.1f cycles/MMA  (%.0f MAC/cycle/SM)\n", cg, 128 * cg, a_tm ? "TMEM" : "SMEM", N,
                       (double)h[0] / iters, 128.0 * N * 16 / ((double)h[0] / iters));
            }
    printf("== same, rotating over nd independent accumulators (A in TMEM, cta_group::1) ==\n");
    for (int N : {32, 64, 128})
        for (int nd : {1, 2, 3, 4, 6}) {
            if (nd * N > 448) continue;
            mma_rate_kernel<1><<<1, 128, 16384>>>(N, 1, iters, d_cyc, nd);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
            printf("N=%3d nd=%d : %7.1f cycles/MMA  (%.0f MAC/cycle/SM)\n", N, nd, (double)h[0] / iters, 128.0 * N * 16 / ((double)h[0] / iters));
        }
    printf("== MUFU throughput, 8 warps/SM, 8 independent chains per thread ==\n");
    const char* names[] = {"tanh.approx.f32", "ex2+rcp (f32)", "tanh.approx.f16x2", "ex2.approx.f32", "ex2.approx.f16x2", "sigmoid=rcp(1+ex2)", "sigmoid via fdividef", "rcp.approx only"};
    for (int mode = 0; mode < 8; ++mode) {
        const int it = 4000;
        if (mode == 0) mufu_kernel<0><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 1) mufu_kernel<1><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 2) mufu_kernel<2><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 3) mufu_kernel<3><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 4) mufu_kernel<4><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 5) mufu_kernel<5><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 6) mufu_kernel<6><<<148, 256>>>(it, d_sink, d_cyc);
        if (mode == 7) mufu_kernel<7><<<148, 256>>>(it, d_sink, d_cyc);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
        double ops = (double)it * 8 * 256 * ((mode == 1 || mode == 5 || mode == 6) ? 2 : 1);
        printf("%-20s : %.2f MUFU instr-lanes/cycle/SM (%s)\n", names[mode], ops / (double)h[0], mode == 2 || mode == 4 ? "x2 results" : "x1");
    }
    printf("== FMA-pipe / ALU-pipe throughput (8 independent chains per thread) ==\n");
    for (int nw : {4, 8, 16}) {
        const int it = 4000;
        const char* nm[] = {"ffma", "fmnmx (alu)", "ffma+fmnmx"};
        for (int mode = 0; mode < 3; ++mode) {
            if (mode == 0) fma_kernel<0><<<148, nw * 32>>>(it, d_sink, d_cyc);
            if (mode == 1) fma_kernel<1><<<148, nw * 32>>>(it, d_sink, d_cyc);
            if (mode == 2) fma_kernel<2><<<148, nw * 32>>>(it, d_sink, d_cyc);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
            printf("%2d warps/SM %-12s: %.1f lanes/cycle/SM\n", nw, nm[mode], (double)it * 8 * nw * 32 * (mode == 2 ? 2 : 1) / (double)h[0]);
        }
    }
    printf("== tcgen05.ld.32x32b.x32 (4 KB per warp instruction) ==\n");
    for (int nw : {1, 4, 8}) {
        ldtm_kernel<<<1, 256>>>(4000, nw, (uint32_t*)d_sink, d_cyc);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
        printf("%d warps: %.1f cycles per ld per warp, %.0f B/cycle/SM\n", nw, (double)h[0] / 4000, 4096.0 * nw * 4000 / (double)h[0]);
    }
    return 0;
}
