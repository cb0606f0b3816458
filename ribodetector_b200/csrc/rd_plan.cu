// rd_plan.cu — K1: sequence bytes → per-read step plan, length-bucketed tiles and transposed
// base codes; plus the fp32 one-hot materialisation (the reference's encoder output).
//
// Replaces, on the device:
//   BASE_DICT / encode_read / encode_variable_len_read   seq_encoder.py:11-18,126-145
//   the truncation `seq[:max_len]`                        detect.py:682,714,717
//   pack_sequence's length sort (enforce_sorted=False)    detect.py:685  → counting sort by steps
//   last_items / last_out_items index math                model.py:114-119, model_cpu.py:57-62
//                                                         → (nfwd, krev, crev) per read
// All kernels here are HBM-bound byte/integer work: coalesced 128-B lines, no tensor cores.
#include "rd_common.cuh"

__device__ __forceinline__ uint32_t base_code(uint8_t b) { return rd_base_code(b); }

// ---------------------------------------------------------------------------------------------
// plan: one thread per read
__global__ void __launch_bounds__(256)
plan_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off, int ostride, int64_t n, int L,
            int semantics, uint32_t* __restrict__ plan, int32_t* __restrict__ hist,
            int32_t* __restrict__ ctrl) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t b = off[i * ostride];                 // read i = seq[off[i*ostride] .. off[i*ostride + 1])
    int64_t len64 = off[i * ostride + 1] - b;
    int len = (int)(len64 < (int64_t)L ? len64 : (int64_t)L);
    if (len < 0) len = 0;
    uint32_t nfwd, krev, crev, invalid = 0;
    if (semantics == RD_SEM_PACKED) {
        nfwd = (uint32_t)len;
        krev = 0;
        if (len == 0) {
            crev = 4; invalid = 1;
            atomicOr(&ctrl[0], 1);
        } else {
            crev = base_code(seq[b + len - 1]);
        }
    } else {
        int p = len - 1;
        crev = 4;
        while (p >= 0) {
            crev = base_code(seq[b + p]);
            if (crev != 4u) break;
            --p;
        }
        if (p < 0) { p = L - 1; crev = 4; }
        nfwd = (uint32_t)(p + 1);
        krev = (uint32_t)(L - 1 - p);
    }
    plan[i] = nfwd | (krev << 13) | (crev << 26) | (invalid << 29);
    // warp-aggregated histogram update: fixed-length inputs put every read in one bucket
    unsigned peers = __match_any_sync(__activemask(), nfwd);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[nfwd], __popc(peers));
}

// bucket starts for keys in DESCENDING order: bstart[key] = sum_{k' > key} hist[k'].
// One block; RD_MAX_LEN+1 keys.
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(int32_t* __restrict__ hist, int nkeys) {
    __shared__ int32_t part[1024];
    int tid = threadIdx.x;
    int per = (nkeys + 1023) / 1024;
    int hi = nkeys - 1 - tid * per;            // this thread owns keys hi, hi-1, ..., hi-per+1
    int32_t s = 0;
    for (int j = 0; j < per; ++j) {
        int k = hi - j;
        if (k >= 0) s += hist[k];
    }
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {        // inclusive Hillis-Steele scan
        int32_t v = (tid >= d) ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int32_t run = part[tid] - s;                 // exclusive prefix
    for (int j = 0; j < per; ++j) {
        int k = hi - j;
        if (k >= 0) {
            int32_t c = hist[k];
            hist[k] = run;
            run += c;
        }
    }
}

// scatter reads into slots: per-CTA shared histogram → one global reservation per (CTA, key)
__global__ void __launch_bounds__(256)
scatter_kernel(const uint32_t* __restrict__ plan, int64_t n, const int32_t* __restrict__ bstart,
               int32_t* __restrict__ cursor, int32_t* __restrict__ perm,
               uint32_t* __restrict__ splan, int kmin, int kspan) {
    extern __shared__ int32_t sh[];              // [kspan] counts, then [kspan] bases
    int32_t* cnt = sh;
    int32_t* base = sh + kspan;
    for (int k = threadIdx.x; k < kspan; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = 0;
    int key = 0, rank = 0;
    if (i < n) {
        p = plan[i];
        key = (int)PLAN_NFWD(p) - kmin;
        rank = atomicAdd(&cnt[key], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kspan; k += blockDim.x) {
        int32_t c = cnt[k];
        base[k] = c ? atomicAdd(&cursor[k + kmin], c) : 0;
    }
    __syncthreads();
    if (i < n) {
        int32_t slot = bstart[key + kmin] + base[key] + rank;
        perm[slot] = (int32_t)i;
        splan[slot] = p;
    }
}

// ---------------------------------------------------------------------------------------------
// one-hot materialisation (parity artefact of the reference encoders)
// A write-only stream of 16 B per base.  Every store instruction of a warp covers 32 CONSECUTIVE rows (512 contiguous
// bytes = 4 full 128-B lines); a thread owns rows lane, lane + 32, ... of its warp's 256-row span, so 8 independent
// streaming stores are in flight per thread.  One integer division per thread; the other rows follow by stepping.
constexpr int OH_ROWS = 8;          // rows per thread
__device__ __forceinline__ float4 onehot_row(uint32_t c) {
    return make_float4(c == 0u ? 1.f : 0.f, c == 1u ? 1.f : 0.f, c == 2u ? 1.f : 0.f, c == 3u ? 1.f : 0.f);
}

__global__ void __launch_bounds__(256)
onehot_padded_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off, int64_t n,
                     int L, float4* __restrict__ out) {
    const int64_t total = n * (int64_t)L;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t r = warp * (32 * OH_ROWS) + lane;
    if (r >= total) return;
    int64_t i;
    int t;
    if (total < ((int64_t)1 << 32)) {
        const uint32_t q = (uint32_t)r / (uint32_t)L;
        i = q; t = (int)((uint32_t)r - q * (uint32_t)L);
    } else {
        i = r / L; t = (int)(r - i * L);
    }
    int64_t b = off[i];
    int64_t len = off[i + 1] - b;
#pragma unroll
    for (int k = 0; k < OH_ROWS; ++k) {
        if (r < total) {
            if (t >= L) {
                do { t -= L; ++i; } while (t >= L);
                b = off[i]; len = off[i + 1] - b;
            }
            const uint32_t c = (int64_t)t < len ? base_code(seq[b + t]) : 4u;
            __stcs(out + r, onehot_row(c));
        }
        r += 32; t += 32;
    }
}

__global__ void __launch_bounds__(256)
rowlen_blocksum_kernel(const int64_t* __restrict__ off, int64_t n, int L,
                       int64_t* __restrict__ blocksum) {
    __shared__ int64_t red[256];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t v = 0;
    if (i < n) { v = off[i + 1] - off[i]; v = v < L ? v : L; }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) blocksum[blockIdx.x] = red[0];
}

// single block, serial-in-chunks exclusive scan of the block sums (n/256 entries)
__global__ void __launch_bounds__(1024)
blocksum_scan_kernel(int64_t* __restrict__ blocksum, int64_t nb) {
    __shared__ int64_t part[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < nb ? blocksum[i] : 0;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            int64_t a = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
            __syncthreads();
            part[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < nb) blocksum[i] = carry + part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
rowoff_kernel(const int64_t* __restrict__ off, int64_t n, int L, const int64_t* __restrict__ blocksum,
              int64_t* __restrict__ row_off) {
    __shared__ int64_t sc[256];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t v = 0;
    if (i < n) { v = off[i + 1] - off[i]; v = v < L ? v : L; }
    sc[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        int64_t a = threadIdx.x >= d ? sc[threadIdx.x - d] : 0;
        __syncthreads();
        sc[threadIdx.x] += a;
        __syncthreads();
    }
    int64_t excl = blocksum[blockIdx.x] + sc[threadIdx.x] - v;
    if (i < n) row_off[i] = excl;
    if (i == n - 1) row_off[n] = excl + v;
}

__global__ void __launch_bounds__(256)
onehot_ragged_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off,
                     const int64_t* __restrict__ row_off, int64_t n, int L,
                     float4* __restrict__ out) {
    // The output rows of all reads form one contiguous range [0, row_off[n]).  Like the padded writer, a warp owns a span
    // of 256 consecutive rows and every store instruction covers 32 consecutive rows (512 B) whatever the read lengths
    // (a warp per read left the last store of every 100-bp read with 4 of 32 lanes).  The read holding a span's first
    // row is found by a warp-wide 32-ary search in row_off (5 rounds for 2^22 reads), once per warp: a warp walks a
    // run of consecutive spans, its lanes stepping from read to read.
    const int lane = threadIdx.x & 31;
    const int64_t total = row_off[n];
    const int64_t n_spans = (total + 32 * OH_ROWS - 1) / (32 * OH_ROWS);
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per_warp = (n_spans + n_warps - 1) / n_warps;          // a warp owns CONSECUTIVE spans: one search
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t span0 = warp * per_warp, span1 = span0 + per_warp < n_spans ? span0 + per_warp : n_spans;
    (void)L;                                          // rows beyond max_len do not exist in row_off
    if (span0 >= span1) return;
    const int64_t base = span0 * (32 * OH_ROWS);
    int64_t lo = 0, hi = n;                           // row_off[lo] <= base < row_off[hi]
    while (hi - lo > 1) {
        const int64_t width = hi - lo;
        const int64_t p = lo + (((int64_t)(lane + 1) * width) >> 5);
        const unsigned m = __ballot_sync(0xffffffffu, row_off[p] <= base);
        const int c = __popc(m);                      // the predicate is monotone in the probe index
        const int64_t nlo = c ? lo + (((int64_t)c * width) >> 5) : lo;
        const int64_t nhi = c < 32 ? lo + (((int64_t)(c + 1) * width) >> 5) : hi;
        lo = nlo; hi = nhi;
    }
    int64_t i = lo, r = base + lane;
    int64_t ro = row_off[i], ro_next = row_off[i + 1], b = off[i];
    // the next read's offsets are loaded one read ahead, so stepping into it costs no dependent load
    int64_t ro_next2 = row_off[i + 2 <= n ? i + 2 : n], b_next = off[i + 1 <= n ? i + 1 : n];
    for (int64_t span = span0; span < span1; ++span) {
#pragma unroll
        for (int k = 0; k < OH_ROWS; ++k) {
            if (r < total) {
                while (r >= ro_next) {
                    ++i; ro = ro_next; ro_next = ro_next2; b = b_next;
                    ro_next2 = row_off[i + 2 <= n ? i + 2 : n];
                    b_next = off[i + 1 <= n ? i + 1 : n];
                }
                __stcs(out + r, onehot_row(base_code(seq[b + (r - ro)])));
            }
            r += 32;
        }
    }
}

// ---------------------------------------------------------------------------------------------
int rd_launch_plan(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n, int L,
                   int semantics, int64_t* n_tiles_out, cudaStream_t st, int ostride) {
    int64_t tiles = (n + RD_TILE - 1) / RD_TILE;
    *n_tiles_out = tiles;
    if (n == 0) return RD_OK;
    int nkeys = RD_MAX_LEN + 1;
    RD_CUDA(h, cudaMemsetAsync(h->d_hist, 0, sizeof(int32_t) * (RD_MAX_LEN + 2), st));
    RD_CUDA(h, cudaMemsetAsync(h->d_cursor, 0, sizeof(int32_t) * (RD_MAX_LEN + 2), st));
    RD_CUDA(h, cudaMemsetAsync(h->d_ctrl, 0, sizeof(int32_t) * 8, st));
    // the scatter fills slots [0, n) exactly (the buckets partition them); only the pad slots of the last
    // tile (and the pad tile the CTA-pair kernel may touch) need "no read" entries
    const int64_t pad = (tiles + 1) * RD_TILE - n;
    RD_CUDA(h, cudaMemsetAsync(h->d_perm + n, 0xFF, sizeof(int32_t) * pad, st));
    RD_CUDA(h, cudaMemsetAsync(h->d_splan + n, 0, sizeof(uint32_t) * pad, st));
    unsigned nb = (unsigned)((n + 255) / 256);
    plan_kernel<<<nb, 256, 0, st>>>(d_seq, d_off, ostride, n, L, semantics, h->d_plan, h->d_hist, h->d_ctrl);
    bucket_scan_kernel<<<1, 1024, 0, st>>>(h->d_hist, nkeys);
    // keys lie in [0, L]; shared histogram spans L+1 keys
    int kspan = L + 1;
    scatter_kernel<<<nb, 256, sizeof(int32_t) * 2 * kspan, st>>>(h->d_plan, n, h->d_hist, h->d_cursor,
                                                                 h->d_perm, h->d_splan, 0, kspan);
    h->launches += 3;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

int rd_launch_onehot(rd_handle* h, const uint8_t* d_seq, const int64_t* d_off, int64_t n, int L,
                     int layout, float* d_out, int64_t* d_row_off, cudaStream_t st) {
    if (n == 0) return RD_OK;
    if (layout == RD_ONEHOT_PADDED) {
        int64_t total = n * (int64_t)L;
        onehot_padded_kernel<<<(unsigned)((total + 256 * OH_ROWS - 1) / (256 * OH_ROWS)), 256, 0, st>>>(
            d_seq, d_off, n, L, reinterpret_cast<float4*>(d_out));
        h->launches += 1;
    } else {
        int64_t nb = (n + 255) / 256;
        if (nb > h->cap_blocksum) {
            if (h->d_blocksum) cudaFree(h->d_blocksum);
            h->d_blocksum = nullptr;
            RD_CUDA(h, cudaMalloc(&h->d_blocksum, sizeof(int64_t) * nb));
            h->cap_blocksum = nb;
        }
        int64_t* ro = d_row_off;
        int64_t* tmp = nullptr;
        if (!ro) {
            RD_CUDA(h, cudaMalloc(&tmp, sizeof(int64_t) * (n + 1)));
            ro = tmp;
        }
        rowlen_blocksum_kernel<<<(unsigned)nb, 256, 0, st>>>(d_off, n, L, h->d_blocksum);
        blocksum_scan_kernel<<<1, 1024, 0, st>>>(h->d_blocksum, nb);
        rowoff_kernel<<<(unsigned)nb, 256, 0, st>>>(d_off, n, L, h->d_blocksum, ro);
        onehot_ragged_kernel<<<(unsigned)(h->sm_count * 8), 256, 0, st>>>(
            d_seq, d_off, ro, n, L, reinterpret_cast<float4*>(d_out));
        h->launches += 4;
        if (tmp) {
            RD_CUDA(h, cudaStreamSynchronize(st));
            cudaFree(tmp);
        }
    }
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}
