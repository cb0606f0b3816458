// rd_lstm_tc.cu — K2 on the 5th-generation tensor cores (RD_PREC_TC_FAST / TC_EXACT / TC_MIXED_RAW).
//
// Replaces `self.rnn(x, None)` + `last_items` + `self.out` (model/model.py:33-36) for the forward
// direction; the reverse direction enters through the logit LUT (rd_tail.cu).
//
// One persistent CTA per SM owns one tile of 128 reads at a time (TMEM lane = read).  Per step t
//     Z[128, 512] = [h_{t-1} | onehot(x_t), 1] . [W_hh ; W_ih ; b]^T          (tcgen05.mma, kind::f16)
// with  A = h_{t-1} resident in TENSOR MEMORY, written there by the activation warps with tcgen05.st (fp16;
//           EXACT: fp16 hi + fp16 lo residual; MIXED: fp16 hi + e5m2 hi + e5m2 lo residual), plus the 16-wide
//           one-hot/bias chunk in shared memory,
//       B = the recurrent weights, staged ONCE per CTA into shared memory by the TMA bulk-copy
//           engine (cp.async.bulk) in the K-major SWIZZLE_NONE canonical layout,
//       D = fp32 accumulators in tensor memory, produced in 4 column chunks of 128 (= 32 hidden units x 4 gates)
//           through a ring of 2 chunk buffers (N = 128 is the smallest N at which a tcgen05.mma costs its ~70-cycle
//           floor, tools/tc_rate.cu).  Columns are laid out so that one tcgen05.ld.32x32b.x16 hands a thread i,f,g,o
//           of 4 hidden units of its read (a "half-chunk"; see col_to_row).
// Sixteen activation warps (4 TMEM lane quarters x 4 unit groups) drain the chunks half-chunk by half-chunk in a
// software pipeline — the tcgen05.ld of the next half-chunk is in flight while the cells of the current one run:
// ex2/rcp (or tanh) on the XU pipe, cell update in fp32 registers, h_t back into the other A buffer; each thread reads
// its read's next base straight from the caller's sequence bytes.  One elected lane of a warp-uniform warp issues the
// MMAs.  The recurrence is pipelined as a wavefront: chunk 0 of step t+1 accumulates K-chunk kc as soon as the
// activation warps have published the 16 hidden units of K-chunk kc of step t (h_ready[kc]; the publish of a
// half-chunk is deferred behind the next one's cells so that tcgen05.wait::st never stalls), so the tensor pipe
// trails the activation pipe by one half-chunk instead of one step.
//
// FAST  : cta_group::1, one fp16 pass (9 MMAs of 128x128x16 per chunk), tanh.approx activations.
// EXACT : cta_group::2 — a CTA pair shares the weights (each SM holds half of the N columns of
//         W_hi and W_lo: 136 KB) and runs two tiles (M = 256) in lock step; three fp16 passes
//         W_lo.h_hi + W_hi.h_lo + W_hi.h_hi (25 MMAs per chunk, small products first) with fp32
//         accumulation; merged-denominator ex2/rcp activations accurate to a few ulp.
// MIXED : the EXACT structure with the two correction passes in kind::f8f6f4 (e5m2, K = 32 per MMA): the fp16 main
//         pass W_hi.h_hi plus ONE 8-bit pass over K = 256, [W_lo8 | W_hi8] . [h_hi8 ; h_lo8], i.e. 17 MMAs per chunk
//         instead of 25.  The corrections are ~2^-12 of the products, so 3 significant bits keep them to ~2^-15:
//         |dlogit| <= 3e-3, |dprob| <= 1e-3 at 100 bp (1.7e-3 / 7.6e-4 measured over 2^20 reads).
//         h is carried as h * 2^-6 (free: the scale rides in the cell's last FMA) and W_hh as W * 2^6, which
//         puts h_hi (as the top byte of its fp16), the fp16 residuals and the weights in e5m2's normal range.
//         (e4m3 would halve the error but needs a 2^15 scale between the passes — scale-input-d, verified in
//         tools/tc_probe8.cu — which rules out issuing main products before the last correction: DESIGN.md.)
// See DESIGN.md for the layout tables and the roofline arithmetic.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "rd_common.cuh"

#ifdef RD_TC_PROFILE
__device__ unsigned long long g_tc_prof[16];
#define PROF_DECL(x) long long x = 0
#define PROF_T0() const long long _p0 = clock64()
#define PROF_ADD(x) x += clock64() - _p0
#else
#define PROF_DECL(x)
#define PROF_T0()
#define PROF_ADD(x)
#endif

#ifndef RD_TC_EXACT_EARLY_MAIN
#define RD_TC_EXACT_EARLY_MAIN 1
#endif
#ifndef RD_TC_DEFER_PUBLISH
#define RD_TC_DEFER_PUBLISH 1
#endif
// measurement-only switches (tools/ab_bench.sh; results are garbage): run the kernel without its tensor-core work or
// without its cell arithmetic, to see what each half costs in time, clock and power on its own
#ifndef RD_TC_SKIP_MMA
#define RD_TC_SKIP_MMA 0
#endif
#ifndef RD_TC_SKIP_CELLS
#define RD_TC_SKIP_CELLS 0
#endif
#define RD_TC_G_FAST 4          // unit-group warps per TMEM lane quarter (the activation loop is written for 4)
#define RD_TC_G_EXACT 4

namespace {

// activation warps: 4 TMEM lane quarters x G unit-group lanes; warp (q, g) owns the 8-unit groups j = g (mod G)
__host__ __device__ constexpr int epi_warps(int G) { return 4 * G; }
__host__ __device__ constexpr int tc_threads(int G) { return 4 * G * 32 + 32; }     // + the MMA / allocator warp
constexpr int KCHUNKS = 8;                            // K-chunks of h (16 hidden units each) per step
constexpr int MMA_N = 128;                            // D columns per MMA chunk = 32 hidden units x 4 gates
constexpr int MMA_CHUNKS = RD_G4 / MMA_N;             // 4 per step
constexpr int NBUF = 2;                               // D chunk ring (2 x 128 columns)
constexpr int KG_H = 16;                              // 8-wide k-groups of h
constexpr int KG_X = 2;                               // k-groups of the one-hot/bias chunk
constexpr int X_BYTES = 2 * RD_TILE * 16;             // one x buffer: [2 k-groups][128 rows][16 B]

enum { M_FAST = 0, M_EXACT = 1, M_MIXED = 2 };

template <int MODE>
struct Cfg {
    static constexpr bool EXACT = MODE != M_FAST;                 // two weight images, CTA pair (EXACT and MIXED)
    static constexpr int CG = EXACT ? 2 : 1;                      // CTAs per MMA (cta_group)
    static constexpr int NL = RD_G4 / CG;                         // weight rows resident per CTA
    static constexpr int NB = MMA_N / CG;                         // weight rows per CTA per MMA chunk
    static constexpr int HI_BYTES = (KG_H + KG_X) * NL * 16;      // FAST 147456, EXACT 73728
    static constexpr int LO_BYTES = EXACT ? KG_H * NL * 16 : 0;   // EXACT 65536 (fp16, K = 128); MIXED 65536 (e5m2, K = 256)
    static constexpr int LBO = NL * 16;                           // bytes between k-groups
    // tensor memory: A buffer s at column s*ACOLS = [h_hi 64 | h_lo 64 (EXACT)] or [h_hi 64 | h_hi8 32 | h_lo8 32 (MIXED)];
    // D ring at DCOL0
    static constexpr int ACOLS = EXACT ? 128 : 64;
    static constexpr int DCOL0 = 256;
    static constexpr int M = 128 * CG;
    // shared memory carve-up (bytes)
    static constexpr int OFF_LO = HI_BYTES;
    static constexpr int OFF_X = HI_BYTES + LO_BYTES;             // 2 x X_BYTES one-hot/bias A operand
    static constexpr int OFF_WOUT = OFF_X + 2 * X_BYTES;          // float [2][128]
    static constexpr int OFF_PART = OFF_WOUT + 2 * RD_H * 4;      // float2 [4][128]
    static constexpr int OFF_BAR = OFF_PART + 4 * RD_TILE * 8;    // mbarriers
    static constexpr int N_BAR = 2 + KCHUNKS + 2 * NBUF;
    static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
    static constexpr int SMEM_BYTES = OFF_TMEM + 16;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // barrier with remote arrivals
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive on the barrier at the same offset in CTA `cta` of the cluster.  RELAXED on purpose: a
// release at cluster scope costs MEMBAR.ALL.GPU + CCTL.IVALL per arrive (measured: ~2500 cycles per
// LSTM step).  What the arrive publishes never travels through memory the waiter reads: it is this
// SM's own tensor memory (completed by tcgen05.wait::st, ordered by tcgen05.fence::before_thread_sync)
// and this SM's own shared memory (made visible to the async proxy by fence.proxy.async), both
// consumed later by this SM's own tensor core when the leader issues the cta_group::2 MMA.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the same, pinned behind the computation of `dep` (an unused operand: keeps ptxas from hoisting the wait above the
// arithmetic that is meant to cover the store latency)
__device__ __forceinline__ void tc_wait_st_after(uint32_t dep) {
    asm volatile("tcgen05.wait::st.sync.aligned; // %0" ::"r"(dep) : "memory");
}


__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE (layout verified by tools/tc_probe.cu)
//   byte(n, k) = (k/8)*LBO + (n/8)*SBO + (n%8)*16 + (k%8)*2 ; fields are in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = f16, both K-major, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// kind::f8f6f4: A and B format e5m2 (1) at bits 7 and 10; one MMA is K = 32 (two 16-byte k-groups; tools/tc_probe8.cu)
__host__ __device__ constexpr uint32_t make_idesc8(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void mma_ts8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                         "r"(bar), "h"((uint16_t)3) : "memory");
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- activations ----------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_mufu(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_mufu(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_mufu(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// One LSTM cell update from the four gate pre-activations of one hidden unit.  The weight images
// carry a per-gate scale (rd_tc_create) so that the values read from tensor memory are directly
// the arguments the activation hardware wants:
//   FAST : zi,zf,zo = z/2, zg = z          sigmoid(z) = 0.5 tanh(z/2) + 0.5     (MUFU.TANH, 5 per unit)
//          (option RD_TC_FAST_FMA_FORGET: zf = -z log2e and the forget gate on the FMA pipe, see sigmoid_fma)
//   EXACT: zi,zf,zo = -z log2e, zg = -2 z log2e, so A = 2^zi = e^-z etc. and
//            c' = f c + i g = (c (1+A)(1+B) + (1-B)(1+F)) / ((1+F)(1+A)(1+B))       A = e^-zi, B = e^-2zg, F = e^-zf
//            h  = o tanh(c') = (1-D) / ((1+O)(1+D))                                  O = e^-zo, D = e^-2c'
//          i.e. 5 MUFU.EX2 + 2 MUFU.RCP per unit (instead of 5 + 5), each within a few ulp.  Exponent
//          arguments are clamped at e^28 (sigmoid/tanh are saturated to fp32 precision long before)
//          so the triple product stays finite.
constexpr float EXACT_SCALE_IFO = -1.4426950408889634f;     // -log2(e)
constexpr float EXACT_SCALE_G = -2.8853900817779268f;       // -2 log2(e)
constexpr float EXACT_CLAMP = 40.395461f;                   // 28 log2(e)
// Kept as an option (off): measured on the box, evaluating the forget gate on the FMA pipe does NOT pay —
// cycles per tile-step 6400 -> 6360 while the extra 16 instructions per unit pull the kernel onto the power
// cap (SM clock 1965 -> 1905 MHz, 58.1 -> 56.7 M reads/s).  The FMA
// rate is not the limit (tools/tc_rate.cu: FFMA and FMNMX each sustain ~125 lanes/clk/SM); the step time is
// set by the per-chunk dependency structure, see DESIGN.md.
#ifndef RD_TC_FAST_FMA_FORGET
#define RD_TC_FAST_FMA_FORGET 0
#endif
constexpr bool FAST_FMA_FORGET = RD_TC_FAST_FMA_FORGET != 0;
__device__ __forceinline__ float sigmoid_fma(float x) {       // 1 / (1 + 2^x), no MUFU
    x = fminf(fmaxf(x, -60.0f), 60.0f);
    const float t = x + 12582912.0f;                             // 1.5 * 2^23: round(x) lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);                      // [-0.5, 0.5]
    const float p = fmaf(f, fmaf(f, fmaf(f, 0.05500893f, 0.24221097f), 0.6932829f), 1.0f);   // 2^f, rel err 1.0e-4
    const float e = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));            // * 2^round(x)
    const float d = 1.0f + e;
    float y = __int_as_float(0x7EF311C7 - __float_as_int(d));   // 1/d to 5 %
    y = fmaf(y, fmaf(-d, y, 1.0f), y);                          // Newton: 2.5e-3
    y = fmaf(y, fmaf(-d, y, 1.0f), y);                          // 6.5e-6
    return y;
}
// 1/d on the FMA pipe: bit-trick seed (5 %), one cubic step (1.3e-4), one Newton step (1e-7); d in [1, 1e25].
// Option (off): with it the exact cell needs 6 MUFU instead of 7, but the kernel is power-capped — measured
// 29.2 M reads/s at 1 586 MHz against 29.6-30.1 M at 1 635-1 650 MHz without it.
#ifndef RD_TC_EXACT_FMA_RCP
#define RD_TC_EXACT_FMA_RCP 0
#endif
constexpr bool EXACT_FMA_RCP = RD_TC_EXACT_FMA_RCP != 0;
__device__ __forceinline__ float rcp_fma(float d) {
    float y = __int_as_float(0x7EF311C7 - __float_as_int(d));
    float r = fmaf(-d, y, 1.0f);
    y = fmaf(y, fmaf(r, r, r), y);
    r = fmaf(-d, y, 1.0f);
    return fmaf(y, r, y);
}
// MIXED carries h as h * MIXED_HS (the weight image as W / MIXED_HS); the residual h - fp16(h) is scaled by MIXED_LS
// before it is rounded to e5m2 (the W_hi8 image by 1 / MIXED_LS)
constexpr float MIXED_HS = 0.015625f;                       // 2^-6
constexpr float MIXED_LS = 16384.0f;                        // 2^14
template <int MODE>
__device__ __forceinline__ void lstm_cell(float zi, float zf, float zg, float zo, float c_old, float& c_new, float& h_new) {
    if constexpr (MODE != M_FAST) {
        const float A = ex2_mufu(fminf(zi, EXACT_CLAMP));
        const float B = ex2_mufu(fminf(zg, EXACT_CLAMP));
        const float F = ex2_mufu(fminf(zf, EXACT_CLAMP));
        const float B1 = 1.0f + B;
        const float P = fmaf(A, B1, B1);                          // (1+A)(1+B)
        const float Q = 1.0f + F;
        const float num = fmaf(c_old, P, fmaf(-B, Q, Q));         // c (1+A)(1+B) + (1-B)(1+F)
        c_new = num * (RD_TC_EXACT_FMA_RCP >= 2 ? rcp_fma(Q * P) : rcp_mufu(Q * P));
        const float O = ex2_mufu(zo);                             // unclamped: O = inf gives den = inf, 1/den = 0, h = 0 (D is finite)
        const float D = ex2_mufu(fminf(EXACT_SCALE_G * c_new, EXACT_CLAMP));
        const float D1 = 1.0f + D;
        const float den = fmaf(O, D1, D1);                        // (1+O)(1+D)
        const float hnum = MODE == M_MIXED ? fmaf(-MIXED_HS, D, MIXED_HS) : 1.0f - D;
        h_new = hnum * (EXACT_FMA_RCP ? rcp_fma(den) : rcp_mufu(den));
    } else {
        const float ig = fmaf(tanh_mufu(zi), 0.5f, 0.5f);
        const float fg = FAST_FMA_FORGET ? sigmoid_fma(zf) : fmaf(tanh_mufu(zf), 0.5f, 0.5f);
        const float gg = tanh_mufu(zg);
        const float og = fmaf(tanh_mufu(zo), 0.5f, 0.5f);
        c_new = fmaf(fg, c_old, ig * gg);
        h_new = og * tanh_mufu(c_new);
    }
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// two e5m2 values in 16 bits, `lo` in the low byte (= the lower k index)
__device__ __forceinline__ uint32_t e5m2x2_from_h2(uint32_t h2) {
    uint16_t r;
    asm("cvt.rn.satfinite.e5m2x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
    return r;
}
__device__ __forceinline__ uint32_t e5m2x2_from_f32(float lo, float hi) {
    uint16_t r;
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}

// one-hot/bias chunk of a read at one step, k = [onehot4 (W_ih hi rows) | onehot4 (W_ih lo rows) | 1 | 1 | 0 x6]:
// k-group 0 (this 16-byte row) changes per step, k-group 1 = {1, 1, 0...} is written once per kernel.
__device__ __forceinline__ void st_x_row(uint32_t saddr, uint32_t code) {
    const uint32_t a = (code == 0u ? 0x00003C00u : 0u) | (code == 1u ? 0x3C000000u : 0u);
    const uint32_t b = (code == 2u ? 0x00003C00u : 0u) | (code == 3u ? 0x3C000000u : 0u);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
template <int MODE, int G>
__global__ void __launch_bounds__(tc_threads(G), 1)
lstm_tc_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off, int ostride,   // the caller's reads
               const uint32_t* __restrict__ splan, const int32_t* __restrict__ perm, int L, int n_tiles_arg,
               const int64_t* __restrict__ n_reads_dev,   // non-NULL: the slot count lives on the device (TC_AUTO)
               const uint8_t* __restrict__ img_hi,     // [CG][HI_BYTES] weight images (rd_tc_create)
               const uint8_t* __restrict__ img_lo,     // [CG][LO_BYTES]
               const float* __restrict__ wout,         // [2][256]
               const float* __restrict__ bout,         // [2]
               const float* __restrict__ revlut,       // [RD_MAX_LEN][5][2]
               float* __restrict__ logits) {
    using C = Cfg<MODE>;
    constexpr bool EXACT = C::EXACT;                  // split precision on a CTA pair (M_EXACT and M_MIXED)
    constexpr int CG = C::CG;
    constexpr int EPI_WARPS = epi_warps(G), EPI_THREADS = EPI_WARPS * 32, TC_THREADS = tc_threads(G);
    static_assert(G == 4, "the activation warps are laid out as 4 lane quarters x 4 unit groups");
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_hi = s_base, s_lo = s_base + C::OFF_LO, s_x = s_base + C::OFF_X;
    float* wout_s = reinterpret_cast<float*>(smem + C::OFF_WOUT);
    float2* part_s = reinterpret_cast<float2*>(smem + C::OFF_PART);
    const uint32_t bar_w = s_base + C::OFF_BAR;               // weights landed
    const uint32_t bar_tile = bar_w + 8;                      // tile set up (x_0 written)      [leader]
    const uint32_t bar_h = bar_w + 16;                        // h_ready[8]                     [leader]
    const uint32_t bar_full = bar_h + 8 * KCHUNKS;            // acc_full[NBUF]                 [each CTA]
    const uint32_t bar_empty = bar_full + 8 * NBUF;           // acc_empty[NBUF]                [leader]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = n_reads_dev ? (int)((*n_reads_dev + RD_TILE - 1) / RD_TILE) : n_tiles_arg;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;      // a unit = one CTA (pair)

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_tile, EPI_WARPS * CG);
        for (int i = 0; i < KCHUNKS; ++i) mbar_init(bar_h + 8 * i, EPI_WARPS * CG);   // every activation warp publishes 4 units of each K-chunk
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, EPI_WARPS * CG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // stage this CTA's weight image once: TMA bulk copies, completion on bar_w
        mbar_expect_tx(bar_w, C::HI_BYTES + C::LO_BYTES);
        const uint8_t* src_hi = img_hi + (size_t)rank * C::HI_BYTES;
        for (int o = 0; o < C::HI_BYTES; o += 8192) bulk_g2s(s_hi + o, src_hi + o, 8192, bar_w);
        if constexpr (EXACT) {
            const uint8_t* src_lo = img_lo + (size_t)rank * C::LO_BYTES;
            for (int o = 0; o < C::LO_BYTES; o += 8192) bulk_g2s(s_lo + o, src_lo + o, 8192, bar_w);
        }
    }
    for (int i = tid; i < 2 * RD_H; i += TC_THREADS)      // fwd half of W_out (MIXED: the cell hands out h * MIXED_HS)
        wout_s[i] = wout[(i >> 7) * 2 * RD_H + (i & 127)] * (MODE == M_MIXED ? 1.0f / MIXED_HS : 1.0f);
    if (tid < 2 * RD_TILE) {      // constant k-group 1 of both x buffers: k8 = k9 = 1.0 (bias hi / lo rows), rest 0
        const uint32_t a = s_x + (uint32_t)((tid >> 7) * X_BYTES + RD_TILE * 16 + (tid & 127) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(a), "r"(0x3C003C00u), "r"(0u) : "memory");
    }
    fence_async_smem();
    if (warp == EPI_WARPS) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync();        // peer's barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == EPI_WARPS) {
        // ============ MMA issuer (leader CTA; warp-uniform control flow, one elected lane issues) ============
        if (rank == 0) {
            const bool elected_lane = elect_one();
            const bool elected = elected_lane && !RD_TC_SKIP_MMA;        // (RD_TC_SKIP_MMA: only the commits are issued)
            mbar_wait(bar_w, 0);
            // (the peer's half of the weights has landed before its activation warps first arrive on
            //  bar_tile: they wait on their own bar_w first)
            constexpr uint32_t idesc = make_idesc(C::M, MMA_N);
            uint32_t hcnt = 0, it = 0;
            PROF_DECL(pw_empty); PROF_DECL(pw_h); PROF_DECL(pw_tile); PROF_DECL(pw_chunk); PROF_DECL(p_chunk0);
#ifdef RD_TC_PROFILE
            const long long p_begin = clock64();
#endif
            for (int tile = unit * CG; tile < n_tiles; tile += n_units * CG, ++it) {
                const int T = (int)PLAN_NFWD(splan[(int64_t)tile * RD_TILE]);
                { PROF_T0();
                if constexpr (CG == 2) mbar_wait_cluster(bar_tile, it & 1); else mbar_wait(bar_tile, it & 1);
                PROF_ADD(pw_tile); }
                tc_fence_after();
                for (int t = 0; t < T; ++t) {
                    const uint32_t rd = (uint32_t)((t + 1) & 1);                   // operand buffer written during step t-1
                    const uint32_t abuf = tmem + rd * C::ACOLS;
                    const uint64_t xdesc = make_desc(s_x + rd * X_BYTES, RD_TILE * 16, 128);
#pragma unroll
                    for (int mc = 0; mc < MMA_CHUNKS; ++mc) {
                        const int buf = mc & 1;
                        const uint32_t eparity = ((mc >> 1) & 1) ^ 1;
                        { PROF_T0();
                        if constexpr (CG == 2) mbar_wait_cluster(bar_empty + 8 * buf, eparity);
                        else mbar_wait(bar_empty + 8 * buf, eparity);
                        PROF_ADD(pw_empty); }
                        if (mc == 0 && t > 0) {
                            PROF_T0();
                            if constexpr (CG == 2) mbar_wait_cluster(bar_h, hcnt & 1); else mbar_wait(bar_h, hcnt & 1);
                            PROF_ADD(pw_h);
                        }
                        tc_fence_after();
#ifdef RD_TC_PROFILE
                        if (mc == 2) p_chunk0 = clock64();
#endif
                        const uint32_t d = tmem + (uint32_t)(C::DCOL0 + buf * MMA_N);
                        const uint32_t boff = (uint32_t)(mc * C::NB * 16);         // chunk's rows inside a k-group
                        const uint64_t xb = make_desc(s_hi + KG_H * C::LBO + boff, C::LBO, 128);
                        if (EXACT && t > 0) {
                            // Accumulate the 16 small correction products first (|W_lo.h_hi|, |W_hi.h_lo| ~ 2^-11 |z|),
                            // then the input chunk and the 8 large products: the tensor core's fp32 accumulator then
                            // rounds at large magnitude 9 times instead of 25 (measured: 1.6x lower logit error at
                            // 100 bp, 2.4x at 300 bp).  In the wavefront chunk (mc == 0) everything that depends only on
                            // K-chunks 0..5 of h_t — their corrections, the input chunk, their large products — is
                            // issued while the activation warps still work on the step's last chunk (which produces
                            // K-chunks 6, 7); only 4 corrections + 2 large products follow the last h_ready, instead
                            // of 2 + 1 + 8.  (RD_TC_EXACT_EARLY_MAIN=0 restores "all corrections, then all large".)
                            constexpr int KSPLIT = (RD_TC_EXACT_EARLY_MAIN != 0) ? 6 : KCHUNKS;
                            const int ksplit = mc == 0 ? KSPLIT : KCHUNKS;
#pragma unroll
                            for (int kc = 0; kc < KCHUNKS; ++kc) {
                                if (kc == ksplit && elected) {                // (mc == 0 only) early input chunk + large products
                                    mma_ss<CG>(d, xdesc, xb, idesc, 1u);
#pragma unroll
                                    for (int k2 = 0; k2 < KSPLIT; ++k2)
                                        mma_ts<CG>(d, abuf + 8 * k2, make_desc(s_hi + 2 * k2 * C::LBO + boff, C::LBO, 128), idesc, 1u);
                                }
                                if (mc == 0 && kc > 0) {
                                    PROF_T0();
                                    mbar_wait_cluster(bar_h + 8 * kc, hcnt & 1);
                                    tc_fence_after();
                                    PROF_ADD(pw_h);
                                }
                                if constexpr (MODE == M_MIXED) {
                                    // 8-bit corrections, one K = 32 MMA per 32 hidden units (= K-chunks kc-1, kc) and term:
                                    // W_lo8 . h_hi8 (image k-groups 0..7, A columns 64..95), W_hi8 . h_lo8 (8..15, 96..127)
                                    if (elected && (kc & 1)) {
                                        const int q8 = kc >> 1;
                                        constexpr uint32_t idesc8 = make_idesc8(C::M, MMA_N);
                                        mma_ts8<CG>(d, abuf + 64 + 8 * q8, make_desc(s_lo + 2 * q8 * C::LBO + boff, C::LBO, 128), idesc8, q8 > 0 ? 1u : 0u);
                                        mma_ts8<CG>(d, abuf + 96 + 8 * q8, make_desc(s_lo + (8 + 2 * q8) * C::LBO + boff, C::LBO, 128), idesc8, 1u);
                                    }
                                } else if (elected) {
                                    const uint64_t bhi = make_desc(s_hi + 2 * kc * C::LBO + boff, C::LBO, 128);
                                    const uint64_t blo = make_desc(s_lo + 2 * kc * C::LBO + boff, C::LBO, 128);
                                    mma_ts<CG>(d, abuf + 8 * kc, blo, idesc, kc > 0 ? 1u : 0u);      // W_lo . h_hi
                                    mma_ts<CG>(d, abuf + 64 + 8 * kc, bhi, idesc, 1u);               // W_hi . h_lo
                                }
                            }
                            if (elected) {
                                if (ksplit == KCHUNKS) mma_ss<CG>(d, xdesc, xb, idesc, 1u);
#pragma unroll
                                for (int kc = 0; kc < KCHUNKS; ++kc)
                                    if (kc >= ksplit || ksplit == KCHUNKS)
                                        mma_ts<CG>(d, abuf + 8 * kc, make_desc(s_hi + 2 * kc * C::LBO + boff, C::LBO, 128), idesc, 1u);
                            }
                        } else {
                            // input projection + biases (A from shared memory): overwrites D
                            if (elected) mma_ss<CG>(d, xdesc, xb, idesc, 0u);
                            if (t > 0) {
#pragma unroll
                                for (int kc = 0; kc < KCHUNKS; ++kc) {
                                    if (mc == 0 && kc > 0) {
                                        PROF_T0();
                                        mbar_wait(bar_h + 8 * kc, hcnt & 1);
                                        tc_fence_after();
                                        PROF_ADD(pw_h);
                                    }
                                    if (elected) mma_ts<CG>(d, abuf + 8 * kc, make_desc(s_hi + 2 * kc * C::LBO + boff, C::LBO, 128), idesc, 1u);
                                }
                            }
                        }
                        if (elected_lane) mma_commit<CG>(bar_full + 8 * buf);
                        __syncwarp();
#ifdef RD_TC_PROFILE
                        if (mc == 2 && t > 0) {      // time from the chunk's first issue to its completion (perturbs the pipeline)
                            mbar_wait(bar_full + 8 * buf, (mc >> 1) & 1);
                            pw_chunk += clock64() - p_chunk0;
                        }
#endif
                    }
                    if (t > 0) ++hcnt;
                }
            }
#ifdef RD_TC_PROFILE
            if (blockIdx.x == 0 && elected) {
                g_tc_prof[0] = clock64() - p_begin; g_tc_prof[1] = pw_empty; g_tc_prof[2] = pw_h; g_tc_prof[3] = pw_tile; g_tc_prof[15] = pw_chunk;
            }
#endif
        }
    } else {
        // =============================== activation warps ===============================
        const int q = warp & 3, par = warp >> 2;                         // lane quarter, unit-group lane g
        const int row = q * 32 + lane;                                   // read slot inside the tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t x_row = s_x + (uint32_t)(row * 16);               // this read's k-group-0 row in x buffer 0
        mbar_wait(bar_w, 0);                                             // this CTA's weights have landed
        uint32_t it = 0;
        PROF_DECL(pe_full); PROF_DECL(pe_ld); PROF_DECL(pe_st); PROF_DECL(pe_full0);
#ifdef RD_TC_PROFILE
        const long long pe_begin = clock64();
#endif
        for (int tile0 = unit * CG; tile0 < n_tiles; tile0 += n_units * CG, ++it) {
            const int tile = tile0 + (int)rank;
            const int64_t slot = (int64_t)tile * RD_TILE + row;
            const int T = (int)PLAN_NFWD(splan[(int64_t)tile0 * RD_TILE]);   // steps of the unit's longest read
            const bool have_tile = tile < n_tiles;
            const uint32_t myplan = have_tile ? splan[slot] : 0u;
            const int nf = (int)PLAN_NFWD(myplan);
            // this thread's read: its bytes are consumed one per step straight from the caller's buffer
            // (consecutive steps hit the same 128-B line in L1; HBM sees every base once)
            const int32_t my_rd = have_tile ? perm[slot] : -1;
            const int64_t my_b = my_rd >= 0 ? off[(int64_t)my_rd * ostride] : 0;
            const int64_t my_l64 = my_rd >= 0 ? off[(int64_t)my_rd * ostride + 1] - my_b : 0;
            const int my_len = (int)(my_l64 < (int64_t)L ? my_l64 : (int64_t)L);
            const uint8_t* cptr = seq + my_b;
            auto code_at = [&](int t) -> uint32_t { return t < my_len ? rd_base_code(__ldg(cptr + t)) : 4u; };
            float c[MMA_CHUNKS][2][4];                                   // cell states: [MMA chunk][half][unit]
#pragma unroll
            for (int g = 0; g < MMA_CHUNKS; ++g)
#pragma unroll
                for (int u = 0; u < 4; ++u) { c[g][0][u] = 0.f; c[g][1][u] = 0.f; }
            float p0 = 0.f, p1 = 0.f;
            uint32_t code_next = 4u, code_next2 = 4u;
            if (par == 0) {
                const uint32_t code0 = code_at(0);
                code_next = code_at(1);
                st_x_row(x_row + X_BYTES, code0);                        // x_0 -> operand buffer 1
                fence_async_smem();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2 && rank != 0) mbar_arrive_remote(bar_tile, 0); else mbar_arrive(bar_tile); }

            for (int t = 0; t < T; ++t) {
                if (par == 0) code_next2 = code_at(t + 2);
                const bool last = t == nf - 1;
                const bool any_last = __any_sync(0xffffffffu, last);       // warp-uniform: the FC is skipped on most steps
                const bool more = t + 1 < T;
                const uint32_t wr = (uint32_t)(t & 1);                                   // h_t, x_{t+1} go to buffer t&1
                const uint32_t awr = tmem + wr * C::ACOLS + lane_off;
                // One half-chunk = this thread's i,f,g,o of 4 hidden units (16 accumulator columns).  Across the four
                // unit-group warps of a lane quarter a half-chunk is 16 hidden units = one K-chunk of the next step.
                auto half_step = [&](const uint32_t (&v)[16], const int mc, const int half) {
                    const int u0 = 32 * mc + 16 * half + 4 * par;        // first of the 4 hidden units
                    float hv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float cn;
                        if (RD_TC_SKIP_CELLS) { cn = c[mc][half][u]; hv[u] = __uint_as_float(v[u] ^ v[4 + u] ^ v[8 + u] ^ v[12 + u]) * 1e-30f; }
                        else lstm_cell<MODE>(__uint_as_float(v[u]), __uint_as_float(v[4 + u]), __uint_as_float(v[8 + u]),
                                             __uint_as_float(v[12 + u]), c[mc][half][u], cn, hv[u]);
                        c[mc][half][u] = cn;     // (a read that has finished keeps stepping on zero rows: nothing reads its state)
                    }
                    if (any_last && last) {        // fused FC: this thread's units of W_out[:, :H] . h_fwd   (model.py:36)
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            p0 = fmaf(wout_s[u0 + u], hv[u], p0);
                            p1 = fmaf(wout_s[RD_H + u0 + u], hv[u], p1);
                        }
                    }
                    if (more) {
                        const uint32_t hcol = awr + (uint32_t)(u0 >> 1);
                        __half2 ha = __floats2half2_rn(hv[0], hv[1]), hb = __floats2half2_rn(hv[2], hv[3]);
                        const uint32_t hia = *reinterpret_cast<uint32_t*>(&ha), hib = *reinterpret_cast<uint32_t*>(&hb);
                        // K-chunk kc of h_t is published (h_ready[kc]) once every store of it has landed in tensor memory.
                        // The publish of the PREVIOUS half-chunk sits here, behind this half-chunk's cells, so that
                        // tcgen05.wait::st finds its stores long complete instead of stalling the warp ~100 cycles eight
                        // times a step; only the step's last half-chunk is published right after its own stores.
                        auto publish = [&](const int kc, const uint32_t dep) {
                            PROF_T0();
                            if (RD_TC_DEFER_PUBLISH) tc_wait_st_after(dep); else tc_wait_st();
                            tc_fence_before();
                            __syncwarp();
                            const uint32_t hb_bar = bar_h + 8 * (uint32_t)kc;
                            if (lane == 0) { if (CG == 2 && rank != 0) mbar_arrive_remote(hb_bar, 0); else mbar_arrive(hb_bar); }
                            PROF_ADD(pe_st);
                        };
                        const int kc = 2 * mc + half;
                        if (RD_TC_DEFER_PUBLISH && kc > 0) publish(kc - 1, hia);
                        if constexpr (MODE == M_MIXED) {
                            // fp16(h'), its top byte as e5m2 (h_hi8), and the residual h' - fp16(h') as e5m2 (h_lo8)
                            const float2 ba = __half22float2(ha), bb = __half22float2(hb);
                            const uint32_t hi8 = e5m2x2_from_h2(hia) | (e5m2x2_from_h2(hib) << 16);
                            const uint32_t lo8 = e5m2x2_from_f32((hv[0] - ba.x) * MIXED_LS, (hv[1] - ba.y) * MIXED_LS) |
                                                 (e5m2x2_from_f32((hv[2] - bb.x) * MIXED_LS, (hv[3] - bb.y) * MIXED_LS) << 16);
                            tmem_st2(hcol, hia, hib);
                            tmem_st1(awr + (uint32_t)(64 + (u0 >> 2)), hi8);
                            tmem_st1(awr + (uint32_t)(96 + (u0 >> 2)), lo8);
                        } else if constexpr (EXACT) {
                            const float2 ba = __half22float2(ha), bb = __half22float2(hb);
                            tmem_st2(hcol, hia, hib);
                            tmem_st2(hcol + 64, pack_h2(hv[0] - ba.x, hv[1] - ba.y), pack_h2(hv[2] - bb.x, hv[3] - bb.y));
                        } else {
                            tmem_st2(hcol, hia, hib);
                        }
                        if (mc == 0 && half == 0 && par == 0) {           // x_{t+1} rides with K-chunk 0
                            st_x_row(x_row + wr * X_BYTES, code_next);
                            fence_async_smem();
                        }
                        if (!RD_TC_DEFER_PUBLISH || kc == KCHUNKS - 1) publish(kc, hia);
                    }
                };
                // Software pipeline over the step's 8 half-chunks: the tcgen05.ld of the next half-chunk is in flight
                // while the cells of the current one run on the XU/FMA pipes (two 16-register buffers).
                uint32_t va[16], vb[16];
                const uint32_t dthr = tmem + (uint32_t)(C::DCOL0 + par * 32) + lane_off;   // this thread's columns of D buffer 0
                { PROF_T0();
                mbar_wait(bar_full, 0);
                PROF_ADD(pe_full);
#ifdef RD_TC_PROFILE
                pe_full0 += clock64() - _p0;
#endif
                }
                tc_fence_after();
                tmem_ld16(dthr, va);
#pragma unroll
                for (int mc = 0; mc < MMA_CHUNKS; ++mc) {
                    const int buf = mc & 1;
                    { PROF_T0(); tc_wait_ld(); PROF_ADD(pe_ld); }
                    tmem_ld16(dthr + (uint32_t)(buf * MMA_N + 16), vb);
                    half_step(va, mc, 0);
                    { PROF_T0(); tc_wait_ld(); PROF_ADD(pe_ld); }
                    tc_fence_before();                                   // this warp has drained its part of the D buffer
                    __syncwarp();
                    if (lane == 0) { if (CG == 2 && rank != 0) mbar_arrive_remote(bar_empty + 8 * buf, 0); else mbar_arrive(bar_empty + 8 * buf); }
                    if (mc + 1 < MMA_CHUNKS) {
                        { PROF_T0();
                        mbar_wait(bar_full + 8 * ((mc + 1) & 1), ((mc + 1) >> 1) & 1);
                        PROF_ADD(pe_full); }
                        tc_fence_after();
                        tmem_ld16(dthr + (uint32_t)(((mc + 1) & 1) * MMA_N), va);
                    }
                    half_step(vb, mc, 1);
                }
                code_next = code_next2;
            }

            // tile tail: combine the two unit-group parities, add the reverse-direction LUT and bias
            part_s[par * RD_TILE + row] = make_float2(p0, p1);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            if (par == 0 && have_tile) {
                const int32_t rd = my_rd;
                if (rd >= 0) {
                    float2 a = part_s[row];
#pragma unroll
                    for (int g = 1; g < G; ++g) { const float2 b = part_s[g * RD_TILE + row]; a.x += b.x; a.y += b.y; }
                    const float* lut = revlut + ((int64_t)PLAN_KREV(myplan) * 5 + PLAN_CREV(myplan)) * 2;
                    float l0 = a.x + lut[0] + bout[0];
                    float l1 = a.y + lut[1] + bout[1];
                    if (PLAN_INVALID(myplan)) { l0 = __int_as_float(0x7fc00000); l1 = l0; }
                    *reinterpret_cast<float2*>(logits + (int64_t)rd * 2) = make_float2(l0, l1);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        }
#ifdef RD_TC_PROFILE
        if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5)) {
            unsigned long long* o = g_tc_prof + (warp == 5 ? 10 : 4);
            o[0] = clock64() - pe_begin; o[1] = pe_full; o[2] = pe_ld; o[3] = pe_st; o[4] = pe_full0;
        }
#endif
    }

    // teardown: nobody leaves while the pair still references this CTA's memories
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync();
    if (warp == EPI_WARPS) {
        __syncwarp();
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---- host side: weight images ----------------------------------------------------------------------
// D column n  <->  hidden unit / gate:  n = 128*mc + 32*par + 16*half + 4*gate + u4  (MMA chunk mc, activation-warp
// unit group par, half-chunk, gate in (i,f,g,o)) holds  unit = 32*mc + 16*half + 4*par + u4: one tcgen05.ld.x16 hands a
// thread i,f,g,o of 4 units, and the four unit groups of a half-chunk together own one 16-unit K-chunk of h.
inline int col_to_row(int n) {
    const int mc = n / 128, r = n % 128, par = r / 32, half = (r % 32) / 16, gate = (r % 16) / 4, u4 = r % 4;
    return gate * RD_H + 32 * mc + 16 * half + 4 * par + u4;
}

// image[rank][kg][n_local][8] halfs; MMA chunk cc takes rows cc*NB .. cc*NB+NB-1 of each rank,
// which are D columns cc*128 + rank*NB + i  (cta_group::2: each CTA supplies half of the N columns).
// lo (bytes): M_EXACT  [rank][kg 0..15][n_local][8] halfs  = fp16(W - W_hi)
//             M_MIXED  [rank][kg 0..15][n_local][16] e5m2: k-groups 0..7 = e5m2(W' - fp16(W')) for k = 16 kg + i,
//                      k-groups 8..15 = e5m2(fp16(W') / MIXED_LS), with W' = W / MIXED_HS (see lstm_cell)
void build_images(const float* w_hh, const float* w_ih, const float* b_ih, const float* b_hh, int mode,
                  std::vector<__half>& hi, std::vector<uint8_t>& lo) {
    const bool exact = mode != M_FAST;
    const int cg = exact ? 2 : 1;
    const int NL = RD_G4 / cg, NB = MMA_N / cg;
    hi.assign((size_t)cg * (KG_H + KG_X) * NL * 8, __float2half(0.f));
    lo.assign(exact ? (size_t)cg * KG_H * NL * 16 : 0, 0);
    __half* lo16 = reinterpret_cast<__half*>(lo.data());
    for (int rank = 0; rank < cg; ++rank)
        for (int nl = 0; nl < NL; ++nl) {
            const int cc = nl / NB, i = nl % NB;
            const int n = cc * MMA_N + rank * NB + i;
            const int row = col_to_row(n);
            // per-gate scale folded into the weights (see lstm_cell): gate = row / H in (i, f, g, o)
            const int gate = row / RD_H;
            const bool is_g = gate == 2;
            const double sc = exact ? (is_g ? (double)EXACT_SCALE_G : (double)EXACT_SCALE_IFO)
                                    : (is_g ? 1.0 : (gate == 1 && FAST_FMA_FORGET) ? (double)EXACT_SCALE_IFO : 0.5);
            const double hs = mode == M_MIXED ? 1.0 / (double)MIXED_HS : 1.0;       // h is carried as h * MIXED_HS
            for (int k = 0; k < RD_H; ++k) {
                const float w = (float)(sc * hs * (double)w_hh[row * RD_H + k]);
                const __half whi = __float2half_rn(w);
                const size_t at = (((size_t)rank * (KG_H + KG_X) + k / 8) * NL + nl) * 8 + k % 8;
                hi[at] = whi;
                if (mode == M_EXACT) {
                    const size_t al = (((size_t)rank * KG_H + k / 8) * NL + nl) * 8 + k % 8;
                    lo16[al] = __float2half_rn(w - __half2float(whi));
                } else if (mode == M_MIXED) {
                    const size_t a_lo = (((size_t)rank * KG_H + k / 16) * NL + nl) * 16 + k % 16;
                    const size_t a_hi = (((size_t)rank * KG_H + 8 + k / 16) * NL + nl) * 16 + k % 16;
                    lo[a_lo] = (uint8_t)__nv_cvt_float_to_fp8(w - __half2float(whi), __NV_SATFINITE, __NV_E5M2);
                    lo[a_hi] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(whi) / MIXED_LS, __NV_SATFINITE, __NV_E5M2);
                }
            }
            // x chunk: k = 0..3 W_ih hi, 4..7 W_ih lo, 8 bias hi, 9 bias lo
            const float bias = (float)(sc * ((double)b_ih[row] + (double)b_hh[row]));
            for (int cdx = 0; cdx < 4; ++cdx) {
                const float w = (float)(sc * (double)w_ih[row * 4 + cdx]);
                const __half whi = __float2half_rn(w);
                const size_t a0 = (((size_t)rank * (KG_H + KG_X) + KG_H) * NL + nl) * 8;
                hi[a0 + cdx] = whi;
                hi[a0 + 4 + cdx] = __float2half_rn(w - __half2float(whi));
            }
            const __half bhi = __float2half_rn(bias);
            const size_t a1 = (((size_t)rank * (KG_H + KG_X) + KG_H + 1) * NL + nl) * 8;
            hi[a1 + 0] = bhi;
            hi[a1 + 1] = __float2half_rn(bias - __half2float(bhi));
        }
}

}  // namespace

constexpr int G_FAST = RD_TC_G_FAST, G_EXACT = RD_TC_G_EXACT;

struct rd_tc_state {
    uint8_t* d_img_fast = nullptr;      // [1][147456]
    uint8_t* d_img_hi = nullptr;        // [2][73728]   M_EXACT
    uint8_t* d_img_lo = nullptr;        // [2][65536]
    uint8_t* d_img_mhi = nullptr;       // [2][73728]   M_MIXED
    uint8_t* d_img_mlo = nullptr;       // [2][65536]
    bool attr_fast = false, attr_exact = false, attr_mixed = false;
};

static int upload(rd_handle* h, uint8_t** dst, const void* src, size_t bytes) {
    RD_CUDA(h, cudaMalloc(dst, bytes));
    RD_CUDA(h, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return RD_OK;
}

int rd_tc_create(rd_handle* h, const float* w_hh, const float* w_ih, const float* b_ih, const float* b_hh) {
    rd_tc_state* s = new (std::nothrow) rd_tc_state();
    if (!s) { h->err = "rd_tc_create: out of host memory"; return RD_ERR_NOMEM; }
    h->tc = s;
    std::vector<__half> hi;
    std::vector<uint8_t> lo;
    int rc;
    build_images(w_hh, w_ih, b_ih, b_hh, M_FAST, hi, lo);
    if ((rc = upload(h, &s->d_img_fast, hi.data(), hi.size() * sizeof(__half)))) return rc;
    build_images(w_hh, w_ih, b_ih, b_hh, M_EXACT, hi, lo);
    if ((rc = upload(h, &s->d_img_hi, hi.data(), hi.size() * sizeof(__half)))) return rc;
    if ((rc = upload(h, &s->d_img_lo, lo.data(), lo.size()))) return rc;
    build_images(w_hh, w_ih, b_ih, b_hh, M_MIXED, hi, lo);
    if ((rc = upload(h, &s->d_img_mhi, hi.data(), hi.size() * sizeof(__half)))) return rc;
    if ((rc = upload(h, &s->d_img_mlo, lo.data(), lo.size()))) return rc;
    return RD_OK;
}

#ifdef RD_TC_PROFILE
extern "C" int rd_debug_prof(unsigned long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 2;
}
#endif

void rd_tc_destroy(rd_handle* h) {
    if (!h->tc) return;
    cudaFree(h->tc->d_img_fast); cudaFree(h->tc->d_img_hi); cudaFree(h->tc->d_img_lo);
    cudaFree(h->tc->d_img_mhi); cudaFree(h->tc->d_img_mlo);
    delete h->tc;
    h->tc = nullptr;
}

template <int MODE>
static int launch_pair_kernel(rd_handle* h, bool* attr_set, const uint8_t* d_seq, const int64_t* d_off, int ostride,
                              const uint32_t* splan, const int32_t* perm, int L, int nt, const int64_t* d_n_reads,
                              const uint8_t* ihi, const uint8_t* ilo, float* d_logits, cudaStream_t st) {
    using C = Cfg<MODE>;
    if (!*attr_set) {
        RD_CUDA(h, cudaFuncSetAttribute(lstm_tc_kernel<MODE, G_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        *attr_set = true;
    }
    const int pairs = (nt + 1) / 2, max_pairs = h->sm_count / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
    cfg.blockDim = dim3(tc_threads(G_EXACT));
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const float* wout = h->d_wout; const float* bout = h->d_bout; const float* lut = h->d_revlut;
    RD_CUDA(h, cudaLaunchKernelEx(&cfg, lstm_tc_kernel<MODE, G_EXACT>, d_seq, d_off, ostride, splan, perm, L, nt, d_n_reads,
                                  ihi, ilo, wout, bout, lut, d_logits));
    return RD_OK;
}

int rd_launch_lstm_tc(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n_tiles, int L, int precision,
                      float* d_logits, cudaStream_t st, const uint32_t* d_splan, const int32_t* d_perm,
                      const int64_t* d_n_reads, int ostride) {
    if (n_tiles == 0) return RD_OK;            // (with d_n_reads: an upper bound used to size the grid)
    if (!d_splan) d_splan = h->d_splan;
    if (!d_perm) d_perm = h->d_perm;
    rd_tc_state* s = h->tc;
    if (!s) { h->err = "tensor-core state missing"; return RD_ERR_UNSUPPORTED; }
    if (precision == RD_PREC_TC_FAST) {
        using C = Cfg<M_FAST>;
        if (!s->attr_fast) {
            RD_CUDA(h, cudaFuncSetAttribute(lstm_tc_kernel<M_FAST, G_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
            s->attr_fast = true;
        }
        int grid = (int)(n_tiles < h->sm_count ? n_tiles : h->sm_count);
        lstm_tc_kernel<M_FAST, G_FAST><<<grid, tc_threads(G_FAST), C::SMEM_BYTES, st>>>(
            d_seq, d_off, ostride, d_splan, d_perm, L, (int)n_tiles, d_n_reads, s->d_img_fast, nullptr, h->d_wout, h->d_bout,
            h->d_revlut, d_logits);
    } else if (precision == RD_PREC_TC_MIXED || precision == RD_PREC_TC_MIXED_RAW) {
        int rc = launch_pair_kernel<M_MIXED>(h, &s->attr_mixed, d_seq, d_off, ostride, d_splan, d_perm, L, (int)n_tiles, d_n_reads,
                                             s->d_img_mhi, s->d_img_mlo, d_logits, st);
        if (rc) return rc;
    } else {
        int rc = launch_pair_kernel<M_EXACT>(h, &s->attr_exact, d_seq, d_off, ostride, d_splan, d_perm, L, (int)n_tiles, d_n_reads,
                                             s->d_img_hi, s->d_img_lo, d_logits, st);
        if (rc) return rc;
    }
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}
