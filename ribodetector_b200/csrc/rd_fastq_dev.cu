// rd_fastq_dev.cu — the edges of the path on the device: FASTQ (and FASTA) text resident in
// HBM → record index (K0), labels → label-partitioned record text (K4), and the host-buffer streaming form
// that chains  H2D → K0 → K1..K3 → K4 → D2H  over two slots.
//
// Replaces, for uncompressed FASTQ text,
//   seq_parser                       ribodetector/data_loader/fastx_parser.py:15-47   (K0)
//   '\n'.join(record)                ribodetector/detect.py:680,711-712               (K4)
//   separate_reads / separate_paired_reads routing + fh.write   detect.py:601-663,295-298 (K4)
// with the semantics rd_scan_fastx / rd_partition_records keep on the host (rd_fastx.cu): lines
// rstrip()ped, not upper-cased, a truncated final record dropped, blank lines an error.
//
// All kernels here are HBM-bound byte/integer work:
//   K0  nl_index_kernel   one pass over the text: 64 B per thread (4 x 16-B loads), newline ranks from a
//                         single-pass chained scan (decoupled look-back over 16-KB tiles) → line_end[]
//       record_kernel     one thread per record: the four [begin, end) line ranges after rstrip → rec[n][8]
//   K4  part_sum / part_scan / part_copy   per-label exclusive offsets of the record texts, then one warp
//                         per record writes "hdr\nseq\nplus\nqual\n" into [non-rRNA | rRNA | unclassified]
// Algorithmic HBM bytes per record of B text bytes: K0 = B + 32 (line ends) + 64 (index), K4 = 2 B + 64 + 1.
#include <algorithm>
#include <new>
#include "rd_common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_THREADS * 64;                 // 16 KB of text per CTA
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;

__device__ __forceinline__ bool py_space(uint32_t c) {        // what str.rstrip() removes for ASCII text
    return c == ' ' || (c >= 0x09u && c <= 0x0du) || (c >= 0x1cu && c <= 0x1fu);
}

// ---- K0a: newline positions, in order, in one pass -------------------------------------------------------
// line_end[i] = position of the i-th '\n', with two hints for record_kernel in the top bits (the bytes next to
// a newline are in L1 when it is found, so record_kernel need not touch the text again in the common case):
constexpr unsigned long long LE_SPACE = 1ull << 63;     // the byte before the newline is <= 0x20: the line may need rstrip()
constexpr unsigned long long LE_AT = 1ull << 62;        // the byte after the newline is '@'
constexpr unsigned long long LE_POS = LE_AT - 1;

// 4-bit mask of the bytes of w equal to '\n' (exact SWAR zero-byte test, then the four 0x80 flags gathered by a multiply)
__device__ __forceinline__ uint32_t nl_nibble(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    const uint32_t t = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);       // 0x80 in every zero byte of x
    return ((t >> 7) * 0x01020408u) >> 24;
}

__global__ void __launch_bounds__(SCAN_THREADS)
nl_index_kernel(const uint8_t* __restrict__ buf, int64_t len, int64_t ntiles, unsigned long long* desc, int* ticket,
                unsigned long long* __restrict__ line_end, int64_t cap, int64_t* __restrict__ info) {
    __shared__ int64_t s_tile, s_base;
    __shared__ int s_warp[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1);               // tiles are taken in launch order: look-back never waits on a CTA that has not started
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t p0 = tile * SCAN_TILE + (int64_t)tid * 64;
    uint32_t mlo = 0u, mhi = 0u;                               // bit b of (mhi:mlo): byte p0 + b is '\n'
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t pos = p0 + 16 * j;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (pos + 16 <= len) {
            v = __ldg(reinterpret_cast<const uint4*>(buf + pos));
        } else if (pos < len) {                                // the text's last, partial 16 bytes
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int b = 0; b < 16; ++b)
                if (pos + b < len) w[b >> 2] |= (uint32_t)buf[pos + b] << (8 * (b & 3));
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        const uint32_t m16 = nl_nibble(v.x) | (nl_nibble(v.y) << 4) | (nl_nibble(v.z) << 8) | (nl_nibble(v.w) << 12);
        if (j < 2) mlo |= m16 << (16 * j); else mhi |= m16 << (16 * (j - 2));
    }
    const int cnt = __popc(mlo) + __popc(mhi);
    // exclusive rank of this thread's first newline inside the tile
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) wbase += c;
        tile_total += c;
    }
    // chained scan across tiles: publish this tile's aggregate, sum the predecessors' (32 at a time)
    if (warp == 0) {
        if (lane == 0) atomicExch(&desc[tile], (tile == 0 ? ST_INC : ST_AGG) | (unsigned long long)tile_total);
        int64_t excl = 0;
        if (tile > 0) {
            int64_t look = tile - 1;
            while (true) {
                const int64_t idx = look - lane;
                unsigned long long d = ST_INC;                 // before tile 0: inclusive prefix 0
                if (idx >= 0) {
                    const volatile unsigned long long* p = desc + idx;
                    do { d = *p; } while ((d & ST_MASK) == 0ull);
                }
                const unsigned inc = __ballot_sync(0xffffffffu, (d & ST_MASK) == ST_INC);
                const int first = inc ? __ffs(inc) - 1 : 32;
                int64_t v = lane <= first ? (int64_t)(d & ~ST_MASK) : 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
                excl += v;
                if (inc) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(&desc[tile], ST_INC | (unsigned long long)(excl + tile_total));
        }
        if (lane == 0) {
            s_base = excl;
            if (tile == ntiles - 1) info[0] = excl + tile_total;     // newlines in the whole text
        }
    }
    __syncthreads();
    int64_t rank = s_base + wbase + (incl - cnt);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t mm = half ? mhi : mlo;
        while (mm) {
            const int64_t pos = p0 + 32 * half + (__ffs(mm) - 1);
            mm &= mm - 1;
            if (rank < cap) {
                unsigned long long v = (unsigned long long)pos;
                if (pos == 0 || buf[pos - 1] <= 0x20) v |= LE_SPACE;             // (L1 hits: this CTA just loaded the lines)
                if (pos + 1 < len && buf[pos + 1] == '@') v |= LE_AT;
                line_end[rank] = v;
            }
            ++rank;
        }
    }
}

// ---- K0b: records.  info: [0] newlines (in), [1] records n, [2] consumed bytes, [4] first bad record * 4 + kind
__global__ void __launch_bounds__(256)
record_kernel(const uint8_t* __restrict__ buf, int64_t len, int final_chunk, int64_t max_records,
              const unsigned long long* __restrict__ line_end, int64_t* __restrict__ rec, int64_t* __restrict__ info) {
    const int64_t n_nl = info[0];
    const bool open_tail = final_chunk && len > 0 && buf[len - 1] != '\n';      // last line without a newline
    const int64_t n_lines = n_nl + (open_tail ? 1 : 0);
    int64_t n = n_lines / 4;                        // a truncated final record is dropped / left for the next block
    if (n > max_records) n = max_records;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) {
        info[1] = n;
        info[2] = n == 0 ? 0 : (4 * n - 1 < n_nl ? (int64_t)(line_end[4 * n - 1] & LE_POS) + 1 : len);
    }
    if (r >= n) return;
    // the five newlines around this record: the one closing the previous record, then its own four
    unsigned long long le[5];
    le[0] = r == 0 ? 0ull : line_end[4 * r - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) le[k + 1] = 4 * r + k < n_nl ? line_end[4 * r + k] : ((unsigned long long)len | LE_SPACE);
    int64_t v[8];
    int kind = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t b = (r == 0 && k == 0) ? 0 : (int64_t)(le[k] & LE_POS) + 1;
        int64_t x = (int64_t)(le[k + 1] & LE_POS);
        if (le[k + 1] & LE_SPACE)
            while (x > b && py_space(buf[x - 1])) --x;                           // line.rstrip()
        if (x == b && kind == 0) kind = 1;                                       // blank line: the reference raises IndexError
        v[2 * k] = b;
        v[2 * k + 1] = x;
    }
    const bool at = r == 0 ? buf[0] == '@' : (le[0] & LE_AT) != 0ull;
    if (kind == 0 && !at) kind = 2;
    if (kind) atomicMin(reinterpret_cast<unsigned long long*>(info + 4), (unsigned long long)(r * 4 + kind));
    longlong2* o = reinterpret_cast<longlong2*>(rec + 8 * r);
    o[0] = make_longlong2(v[0], v[1]);
    o[1] = make_longlong2(v[2], v[3]);
    o[2] = make_longlong2(v[4], v[5]);
    o[3] = make_longlong2(v[6], v[7]);
}

// ---- K0 for FASTA text ---------------------------------------------------------------------------------------
// seq_parser's FASTA branch (fastx_parser.py:39-55): lines strip()ped, blank ones skipped, a line starting with '>' opens
// a record, the other lines of a record are joined and upper-cased; a record is complete when the next header (or the
// end of the file) is seen; at the end of the file a header without sequence is dropped; sequence lines before the very
// first header stay attached to it.  On the device: newline index (nl_index_kernel, as for FASTQ) → one thread per line
// (fa_line_kernel: strip, classify) → a two-component scan over the lines (headers so far = record id, sequence bytes so
// far = where the line's bases go) → fa_emit_kernel writes the record index, fa_copy_kernel the joined upper-cased
// sequences into the region BEHIND the text in the same buffer, so that a FASTA record looks like a FASTQ one to the
// classify and partition kernels:  rec[r] = { hdr [b,e) , seq [b,e) (in the region at seq_base) , raw line begin, -, -, - }.
constexpr unsigned long long FA_HDR = 1ull << 63, FA_SEQ = 1ull << 62;      // line descriptor: type | len << 40 | begin
constexpr int FA_LEN_BITS = 22;                                             // a stripped line holds < 4 Mi bytes

// lines the chunk is scanned over: every complete line that fits the line index, plus an unterminated last line at the
// end of the file; *final_eff = the scanned lines really end the file (a chunk with more lines than the index holds is
// handled like a chunk cut short: its remaining text comes round again)
__device__ __forceinline__ int64_t fa_lines(const uint8_t* buf, int64_t len, int final_chunk, int64_t cap, const int64_t* info,
                                            int64_t* n_nl_out, bool* final_eff) {
    const bool all = info[0] <= cap;
    const int64_t n_nl = all ? info[0] : cap;
    *final_eff = final_chunk && all;
    *n_nl_out = n_nl;
    return n_nl + ((*final_eff && len > 0 && buf[len - 1] != '\n') ? 1 : 0);
}

__global__ void __launch_bounds__(256)
fa_line_kernel(const uint8_t* __restrict__ buf, int64_t len, int final_chunk, const unsigned long long* __restrict__ line_end,
               unsigned long long* __restrict__ line_desc, int64_t cap, int64_t* __restrict__ info) {
    int64_t n_nl; bool fin;
    const int64_t n_lines = fa_lines(buf, len, final_chunk, cap, info, &n_nl, &fin);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;                 // grid-stride: the grid is fixed
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += stride) {
        int64_t b = i == 0 ? 0 : (int64_t)(line_end[i - 1] & LE_POS) + 1;
        int64_t x = i < n_nl ? (int64_t)(line_end[i] & LE_POS) : len;
        while (x > b && py_space(buf[x - 1])) --x;                          // line.strip()
        while (b < x && py_space(buf[b])) ++b;
        unsigned long long d = (unsigned long long)b;
        if (x > b) {
            if (x - b >= ((int64_t)1 << FA_LEN_BITS)) atomicMin(reinterpret_cast<unsigned long long*>(info + 4), (unsigned long long)(i * 4 + 2));
            d |= (unsigned long long)(x - b) << 40;
            d |= buf[b] == '>' ? FA_HDR : FA_SEQ;
        }
        line_desc[i] = d;
    }
}

__device__ __forceinline__ int fa_len(unsigned long long d) { return (int)((d >> 40) & ((1u << FA_LEN_BITS) - 1)); }

// per 256-line block: {headers, sequence bytes}
__global__ void __launch_bounds__(256)
fa_sum_kernel(const uint8_t* __restrict__ buf, int64_t len, int final_chunk, int64_t cap, const int64_t* __restrict__ info,
              const unsigned long long* __restrict__ line_desc, int64_t* __restrict__ blocksum) {
    __shared__ int64_t s[2][8];
    int64_t n_nl; bool fin;
    const int64_t n_lines = fa_lines(buf, len, final_chunk, cap, info, &n_nl, &fin);
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long d = i < n_lines ? line_desc[i] : 0ull;
    int64_t v[2] = {(d & FA_HDR) ? 1 : 0, (d & FA_SEQ) ? fa_len(d) : 0};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], k);
        if ((threadIdx.x & 31) == 0) s[c][threadIdx.x >> 5] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        int64_t t = 0;
        for (int w = 0; w < 8; ++w) t += s[threadIdx.x][w];
        blocksum[(int64_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
}

// one CTA: exclusive scan of the block sums; then the chunk's record count and `consumed`
//   info: [0] newlines (in), [1] records n, [2] consumed, [3] sequence bytes of the n records' region, [4] error,
//         [5] headers in the chunk, [6] lines
__global__ void __launch_bounds__(1024)
fa_scan_kernel(const uint8_t* __restrict__ buf, int64_t len, int final_chunk_in, int64_t cap, int64_t* __restrict__ blocksum,
               int64_t nblk, int64_t max_records, int64_t* __restrict__ info) {
    __shared__ int64_t part[1024];
    int64_t n_nl; bool fin;
    const int64_t n_lines = fa_lines(buf, len, final_chunk_in, cap, info, &n_nl, &fin);
    const int final_chunk = fin ? 1 : 0;
    __shared__ int64_t total[2];
    const int tid = threadIdx.x;
    const int64_t per = (nblk + 1023) / 1024;
    const int64_t lo = tid * per, hi = lo + per < nblk ? lo + per : nblk;
    for (int c = 0; c < 2; ++c) {
        int64_t sum = 0;
        for (int64_t i = lo; i < hi; ++i) sum += blocksum[i * 2 + c];
        part[tid] = sum;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const int64_t o = tid >= d ? part[tid - d] : 0;
            __syncthreads();
            part[tid] += o;
            __syncthreads();
        }
        int64_t run = part[tid] - sum;
        if (tid == 1023) total[c] = part[1023];
        __syncthreads();
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t v = blocksum[i * 2 + c];
            blocksum[i * 2 + c] = run;
            run += v;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int64_t headers = total[0];
        // the record opened by the chunk's last header is complete only at the end of the file
        int64_t n = final_chunk ? headers : (headers > 0 ? headers - 1 : 0);
        if (final_chunk && headers == 0 && total[1] > 0) n = 1;            // header-less sequence: one record, '' header
        if (n > max_records) n = max_records;
        info[1] = n;                                                        // (fa_emit_kernel drops a trailing empty record)
        info[3] = total[1];
        info[5] = headers;
        info[6] = n_lines;
        // every header's record taken (only at the end of the file): all of the text is consumed; otherwise fa_emit_kernel
        // stores the line start of header n, the first record left for the next chunk
        info[2] = (n >= headers && (n > 0 || fin)) ? len : 0;
        info[7] = n_nl;
    }
}

// per line: record id and sequence offset from the scans; headers write the index, sequence lines are copied by
// fa_copy_kernel.  seq_base = offset of the sequence region inside buf.
__global__ void __launch_bounds__(256)
fa_emit_kernel(const uint8_t* __restrict__ buf, int64_t len, int final_chunk_in, int64_t cap,
               const unsigned long long* __restrict__ line_desc, const unsigned long long* __restrict__ line_end,
               const int64_t* __restrict__ blocksum, int64_t seq_base, int64_t* __restrict__ rec, int64_t rec_cap,
               int64_t* __restrict__ seq_at, int64_t* __restrict__ info) {
    __shared__ int64_t s_w[2][8];
    int64_t n_nl; bool fin;
    const int64_t n_lines = fa_lines(buf, len, final_chunk_in, cap, info, &n_nl, &fin);
    const int final_chunk = fin ? 1 : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = (int64_t)blockIdx.x * 256 + tid;
    const unsigned long long d = i < n_lines ? line_desc[i] : 0ull;
    const int64_t mine[2] = {(d & FA_HDR) ? 1 : 0, (d & FA_SEQ) ? fa_len(d) : 0};
    int64_t incl[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int64_t v = mine[c];
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, v, k);
            if (lane >= k) v += o;
        }
        incl[c] = v;
        if (lane == 31) s_w[c][warp] = v;
    }
    __syncthreads();
    int64_t ex[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int64_t before = blocksum[(int64_t)blockIdx.x * 2 + c];
        for (int w = 0; w < warp; ++w) before += s_w[c][w];
        ex[c] = before + incl[c] - mine[c];                                 // headers / sequence bytes BEFORE this line
    }
    if (i >= n_lines) return;
    const int64_t n = info[1], headers = info[5], total_seq = info[3];
    if (i < n_lines) seq_at[i] = ex[1];                                     // where this line's bases go (fa_copy_kernel)
    if (d & FA_HDR) {
        const int64_t r = ex[0];                                            // this header opens record r
        const int64_t b = (int64_t)(d & ((1ull << 40) - 1)), e = b + fa_len(d);
        const int64_t raw = i == 0 ? 0 : (int64_t)(line_end[i - 1] & LE_POS) + 1;
        if (r < rec_cap) {
            rec[8 * r + 0] = b; rec[8 * r + 1] = e;
            rec[8 * r + 2] = seq_base + (r == 0 ? 0 : ex[1]);              // lines before the first header stay with it
            rec[8 * r + 4] = raw;
            rec[8 * r + 5] = 0; rec[8 * r + 6] = 0; rec[8 * r + 7] = 0;
            if (r == headers - 1) rec[8 * r + 3] = seq_base + total_seq;    // the chunk's last record ends with the region
        }
        if (r >= 1 && r - 1 < rec_cap) rec[8 * (r - 1) + 3] = seq_base + ex[1];
        if (r == n) info[2] = raw;                                          // the first record not taken starts here
        // at the end of the file a last record without sequence is dropped (fastx_parser.py:54)
        if (final_chunk && r == headers - 1 && r == n - 1 && ex[1] == total_seq && !(r == 0 && total_seq > 0)) info[1] = n - 1;
    }
    if (headers == 0 && i == 0 && n == 1) {                                 // header-less sequence at the end of the file
        rec[0] = 0; rec[1] = 0; rec[2] = seq_base; rec[3] = seq_base + total_seq; rec[4] = 0; rec[5] = rec[6] = rec[7] = 0;
    }
}

// half a warp per sequence line: bases upper-cased into the sequence region
__global__ void __launch_bounds__(256)
fa_copy_kernel(uint8_t* __restrict__ buf, const unsigned long long* __restrict__ line_desc, const int64_t* __restrict__ seq_at,
               const int64_t* __restrict__ info, int64_t seq_base, int64_t seq_limit) {
    const int hl = threadIdx.x & 15;
    const int64_t n_lines = info[6], stride = ((int64_t)gridDim.x * blockDim.x) >> 4;      // grid-stride: the grid is fixed
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; i < n_lines; i += stride) {
        const unsigned long long d = line_desc[i];
        if (!(d & FA_SEQ)) continue;
        const int64_t b = (int64_t)(d & ((1ull << 40) - 1));
        const int l = fa_len(d);
        const int64_t dst = seq_at[i];
        if (dst + l > seq_limit) continue;                                  // (cannot happen: the region is as large as the text)
        uint8_t* o = buf + seq_base + dst;
        for (int j = hl; j < l; j += 16) {
            const uint8_t c = buf[b + j];
            o[j] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;          // .upper()
        }
    }
}

// ---- K4: label-partitioned record text ------------------------------------------------------------------
__device__ __forceinline__ int label_class(int8_t l) { return l == 0 ? 0 : (l == 1 ? 1 : 2); }

// (nlines = 4: FASTQ "hdr\nseq\nplus\nqual\n"; nlines = 2: FASTA "hdr\nseq\n", the last two index slots hold other data)
__device__ __forceinline__ int rec_text_len(const int64_t* __restrict__ rec, int64_t r, int nlines) {
    const longlong2* p = reinterpret_cast<const longlong2*>(rec + 8 * r);
    const longlong2 a = p[0], b = p[1];
    if (nlines == 2) return (int)((a.y - a.x) + (b.y - b.x)) + 2;
    const longlong2 c = p[2], d = p[3];
    return (int)((a.y - a.x) + (b.y - b.x) + (c.y - c.x) + (d.y - d.x)) + 4;
}

// per 256-record block: text bytes of each class
__global__ void __launch_bounds__(256)
part_sum_kernel(const int64_t* __restrict__ rec, const int8_t* __restrict__ labels, int64_t n, int64_t* __restrict__ blocksum,
                int nlines) {
    __shared__ int64_t s[3][8];
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    int64_t v[3] = {0, 0, 0};
    if (r < n) {
        const int c = label_class(labels[r]);
        const int l = rec_text_len(rec, r, nlines);
        v[0] = c == 0 ? l : 0; v[1] = c == 1 ? l : 0; v[2] = c == 2 ? l : 0;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], d);
        if ((threadIdx.x & 31) == 0) s[c][threadIdx.x >> 5] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int64_t t = 0;
        for (int w = 0; w < 8; ++w) t += s[threadIdx.x][w];
        blocksum[(int64_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

// one CTA: exclusive scan of the block sums per class, class bases (non-rRNA | rRNA | unclassified), sizes3
__global__ void __launch_bounds__(1024)
part_scan_kernel(int64_t* __restrict__ blocksum, int64_t nblk, int64_t* __restrict__ sizes3) {
    __shared__ int64_t part[1024];
    __shared__ int64_t total[3];
    const int tid = threadIdx.x;
    const int64_t per = (nblk + 1023) / 1024;                  // consecutive entries owned by one thread
    const int64_t lo = tid * per, hi = lo + per < nblk ? lo + per : nblk;
    for (int c = 0; c < 3; ++c) {
        int64_t s = 0;
        for (int64_t i = lo; i < hi; ++i) s += blocksum[i * 3 + c];
        part[tid] = s;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {                   // inclusive Hillis-Steele scan of the per-thread sums
            const int64_t o = tid >= d ? part[tid - d] : 0;
            __syncthreads();
            part[tid] += o;
            __syncthreads();
        }
        int64_t run = part[tid] - s;
        if (tid == 1023) total[c] = part[1023];
        __syncthreads();
        // class c starts after the classes before it
        run += c == 0 ? 0 : (c == 1 ? total[0] : total[0] + total[1]);
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t v = blocksum[i * 3 + c];
            blocksum[i * 3 + c] = run;
            run += v;
        }
        __syncthreads();
    }
    if (tid < 3) sizes3[tid] = total[tid];
}

// 256 records per CTA: in-block exclusive offsets per class, then half a warp per record writes its text
__global__ void __launch_bounds__(256)
part_copy_kernel(const uint8_t* __restrict__ buf, const int64_t* __restrict__ rec, const int8_t* __restrict__ labels,
                 int64_t n, const int64_t* __restrict__ blockbase, uint8_t* __restrict__ out, int nlines) {
    __shared__ longlong2 s_rec[256][4];          // the block's slice of the record index: read from HBM once, coalesced
    __shared__ int64_t s_dst[256];
    __shared__ int s_wsum[3][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 256;
    {
        const longlong2* g = reinterpret_cast<const longlong2*>(rec + 8 * r0);
        longlong2* sflat = &s_rec[0][0];
        const int64_t avail = (n - r0 < 256 ? n - r0 : 256) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = j * 256 + tid;
            if (idx < avail) sflat[idx] = g[idx];
        }
    }
    __syncthreads();
    const int64_t r = r0 + tid;
    int cls = 0, len = 0;
    if (r < n) {
        cls = label_class(labels[r]);
        const longlong2 a = s_rec[tid][0], b = s_rec[tid][1], c = s_rec[tid][2], d = s_rec[tid][3];
        len = nlines == 2 ? (int)((a.y - a.x) + (b.y - b.x)) + 2
                          : (int)((a.y - a.x) + (b.y - b.x) + (c.y - c.x) + (d.y - d.x)) + 4;
    }
    int incl[3], mine[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mine[c] = (r < n && cls == c) ? len : 0;
        int v = mine[c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        incl[c] = v;
        if (lane == 31) s_wsum[c][warp] = v;
    }
    __syncthreads();
    if (r < n) {
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_wsum[cls][w];
        s_dst[tid] = blockbase[(int64_t)blockIdx.x * 3 + cls] + before + (incl[cls] - mine[cls]);
    }
    __syncthreads();
    if (nlines == 4) {
        // FASTQ.  A record nothing was stripped from is ONE contiguous range of the input, and so is a RUN of such records
        // that follow each other in the text and carry the same label: their texts (closing newlines included) are adjacent
        // in the input and land adjacent in the output.  With a few per cent of the reads labelled rRNA a warp's 32 records
        // are two or three runs of several KB: each is copied by the whole warp as 16-byte words aligned on the OUTPUT, the
        // input words realigned with funnel shifts (two aligned 16-byte loads per store; the second one hits L1).
        const int slot = warp * 32 + lane;
        const bool valid = r0 + slot < n;
        longlong2 a = make_longlong2(0, 0), d = a;
        bool plain = false;
        int mycls = -1;
        if (valid) {
            a = s_rec[slot][0];
            const longlong2 b2 = s_rec[slot][1], c2 = s_rec[slot][2];
            d = s_rec[slot][3];
            plain = a.y + 1 == b2.x && b2.y + 1 == c2.x && c2.y + 1 == d.x;
            mycls = label_class(labels[r0 + slot]);
        }
        const long long pend = __shfl_up_sync(0xffffffffu, d.y, 1);         // previous record: end of its quality line,
        const int pcls = __shfl_up_sync(0xffffffffu, mycls, 1);            // its class, and whether it is plain
        const int pplain = __shfl_up_sync(0xffffffffu, (int)plain, 1);
        const bool joined = lane > 0 && valid && plain && pplain && pcls == mycls && pend + 1 == a.x;
        uint32_t heads = __ballot_sync(0xffffffffu, valid && !joined);
        while (heads) {
            const int first = __ffs(heads) - 1;
            heads &= heads - 1;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            const int stop = heads ? __ffs(heads) - 1 : 32 - __clz(vmask);   // one past the run's last record
            const long long src0 = __shfl_sync(0xffffffffu, a.x, first);
            const long long end = __shfl_sync(0xffffffffu, d.y, stop - 1);
            const int is_plain = __shfl_sync(0xffffffffu, (int)plain, first);
            uint8_t* o = out + s_dst[warp * 32 + first];
            if (is_plain) {
                const uint8_t* src = buf + src0;
                const int nbytes = (int)(end - src0);                        // up to, not including, the last closing '\n'
                int head = (int)((16u - (uint32_t)(uintptr_t)o) & 15u);
                if (head > nbytes) head = nbytes;
                if (lane < head) o[lane] = src[lane];
                const int nvec = (nbytes - head) >> 4;
                const uint8_t* s2 = src + head;
                const uint32_t sh = (uint32_t)(uintptr_t)s2 & 15u, ws = sh >> 2, bs = 8u * (sh & 3u);
                const uint4* sa = reinterpret_cast<const uint4*>(s2 - sh);
                uint4* ov = reinterpret_cast<uint4*>(o + head);
                // four independent 512-byte rows of the run in flight per warp (all loads first, then the stores)
                for (int v0 = lane; v0 < nvec; v0 += 128) {
                    uint4 A[4], B[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int v = v0 + 32 * q;
                        A[q] = v < nvec ? __ldg(sa + v) : make_uint4(0u, 0u, 0u, 0u);
                        B[q] = (sh && v < nvec) ? __ldg(sa + v + 1) : make_uint4(0u, 0u, 0u, 0u);   // (the aligned word holding the vector's last bytes)
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int v = v0 + 32 * q;
                        uint32_t w0, w1, w2, w3, w4;
                        if (ws == 0) { w0 = A[q].x; w1 = A[q].y; w2 = A[q].z; w3 = A[q].w; w4 = B[q].x; }
                        else if (ws == 1) { w0 = A[q].y; w1 = A[q].z; w2 = A[q].w; w3 = B[q].x; w4 = B[q].y; }
                        else if (ws == 2) { w0 = A[q].z; w1 = A[q].w; w2 = B[q].x; w3 = B[q].y; w4 = B[q].z; }
                        else { w0 = A[q].w; w1 = B[q].x; w2 = B[q].y; w3 = B[q].z; w4 = B[q].w; }
                        if (v < nvec)
                            __stcs(ov + v, make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs),
                                                      __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs)));
                    }
                }
                const int done = head + 16 * nvec;
                if (lane < nbytes - done) o[done + lane] = src[done + lane];
                if (lane == 0) o[nbytes] = (uint8_t)'\n';
            } else {
                // a record with stripped line ends (CRLF, trailing blanks): byte by byte through its four line ranges
                const int s1 = warp * 32 + first;
                const longlong2 ra = s_rec[s1][0], rb = s_rec[s1][1], rc = s_rec[s1][2], rd = s_rec[s1][3];
                const int t0 = (int)(ra.y - ra.x) + 1, t1 = t0 + (int)(rb.y - rb.x) + 1;
                const int t2 = t1 + (int)(rc.y - rc.x) + 1, t3 = t2 + (int)(rd.y - rd.x) + 1;
                for (int j = lane; j < t3; j += 32) {     // output byte j: which line it belongs to, or the '\n' closing one
                    int64_t sp; int e2;
                    if (j < t0) { sp = ra.x + j; e2 = t0; }
                    else if (j < t1) { sp = rb.x + (j - t0); e2 = t1; }
                    else if (j < t2) { sp = rc.x + (j - t1); e2 = t2; }
                    else { sp = rd.x + (j - t2); e2 = t3; }
                    o[j] = j == e2 - 1 ? (uint8_t)'\n' : buf[sp];
                }
            }
        }
        return;
    }
    // FASTA (two lines per record; the sequence lives in the joined region behind the text): half a warp per record
    const int sub = lane >> 4, hl = lane & 15;
    for (int i = 0; i < 16; ++i) {
        if (r0 + warp * 32 + 2 * i >= n) break;
        const int slot = warp * 32 + 2 * i + sub;
        if (r0 + slot >= n) continue;
        const longlong2 a = s_rec[slot][0], b = s_rec[slot][1];
        const int t0 = (int)(a.y - a.x) + 1, t1 = t0 + (int)(b.y - b.x) + 1;
        uint8_t* o = out + s_dst[slot];
        for (int j = hl; j < t1; j += 16) {
            int64_t src; int end;
            if (j < t0) { src = a.x + j; end = t0; }
            else { src = b.x + (j - t0); end = t1; }
            o[j] = j == end - 1 ? (uint8_t)'\n' : buf[src];
        }
    }
}

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------------------
struct rd_fq_state {
    static const int NSLOT = 2;
    // scan scratch (used on s_in only, in order)
    unsigned long long* d_desc = nullptr; int64_t cap_desc = 0;
    int* d_ticket = nullptr;
    unsigned long long* d_line_end[2] = {nullptr, nullptr}; int64_t cap_lines[2] = {0, 0};
    unsigned long long* d_line_desc[2] = {nullptr, nullptr}; int64_t cap_line_desc[2] = {0, 0};    // FASTA: per-line type/length/begin
    int64_t* d_seq_at[2] = {nullptr, nullptr}; int64_t cap_seq_at[2] = {0, 0};                     // FASTA: per-line sequence offset
    int64_t* d_fa_blocksum = nullptr; int64_t cap_fa_blk = 0;
    // partition scratch (s_cmp only)
    int64_t* d_blocksum = nullptr; int64_t cap_blk = 0;
    // streaming slots
    uint8_t* d_buf[NSLOT][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; int64_t cap_buf[NSLOT][2] = {{0, 0}, {0, 0}};
    uint8_t* d_out[NSLOT][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; int64_t cap_out[NSLOT][2] = {{0, 0}, {0, 0}};
    int64_t* d_rec[NSLOT][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; int64_t cap_rec[NSLOT][2] = {{0, 0}, {0, 0}};
    int8_t* d_labels[NSLOT] = {nullptr, nullptr}; int64_t cap_labels[NSLOT] = {0, 0};
    float* d_logits[2] = {nullptr, nullptr}; int64_t cap_logits[2] = {0, 0};
    int64_t* d_info = nullptr;          // [2 ends][8]
    int64_t* h_info = nullptr;          // pinned mirror + one spare line
    int64_t* d_res[NSLOT] = {nullptr, nullptr};   // [9]: sizes3 end 0, sizes3 end 1, counts3
    int64_t* h_res[NSLOT] = {nullptr, nullptr};   // pinned
    cudaEvent_t ev_cmp[NSLOT] = {nullptr, nullptr}, ev_out[NSLOT] = {nullptr, nullptr};
    bool pending[NSLOT] = {false, false};
};

static int fq_state(rd_handle* h, rd_fq_state** out) {
    if (!h->fq) {
        rd_fq_state* s = new (std::nothrow) rd_fq_state();
        if (!s) { h->err = "rd_fastq: out of host memory"; return RD_ERR_NOMEM; }
        h->fq = s;
        RD_CUDA(h, cudaMalloc(&s->d_ticket, sizeof(int) * 4));
        RD_CUDA(h, cudaMalloc(&s->d_info, sizeof(int64_t) * 16));
        RD_CUDA(h, cudaMallocHost(&s->h_info, sizeof(int64_t) * 24));
        for (int i = 0; i < rd_fq_state::NSLOT; ++i) {
            RD_CUDA(h, cudaMalloc(&s->d_res[i], sizeof(int64_t) * 9));
            RD_CUDA(h, cudaMallocHost(&s->h_res[i], sizeof(int64_t) * 9));
            RD_CUDA(h, cudaEventCreateWithFlags(&s->ev_cmp[i], cudaEventDisableTiming));
            RD_CUDA(h, cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
        }
    }
    *out = h->fq;
    return RD_OK;
}

void rd_fq_destroy(rd_handle* h) {
    rd_fq_state* s = h->fq;
    if (!s) return;
    cudaFree(s->d_desc); cudaFree(s->d_ticket); cudaFree(s->d_blocksum); cudaFree(s->d_info);
    cudaFreeHost(s->h_info);
    for (int e = 0; e < 2; ++e) { cudaFree(s->d_line_end[e]); cudaFree(s->d_logits[e]); cudaFree(s->d_line_desc[e]); cudaFree(s->d_seq_at[e]); }
    cudaFree(s->d_fa_blocksum);
    for (int i = 0; i < rd_fq_state::NSLOT; ++i) {
        for (int e = 0; e < 2; ++e) { cudaFree(s->d_buf[i][e]); cudaFree(s->d_out[i][e]); cudaFree(s->d_rec[i][e]); }
        cudaFree(s->d_labels[i]); cudaFree(s->d_res[i]); cudaFreeHost(s->h_res[i]);
        if (s->ev_cmp[i]) cudaEventDestroy(s->ev_cmp[i]);
        if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
    }
    delete s;
    h->fq = nullptr;
}

template <typename T>
static int grow(rd_handle* h, T** p, int64_t* cap, int64_t want) {       // device buffer of at least `want` elements
    if (want <= *cap && *p) return RD_OK;
    RD_CUDA(h, cudaDeviceSynchronize());
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    const int64_t n = std::max<int64_t>(want + want / 8, 16);
    RD_CUDA(h, cudaMalloc(p, sizeof(T) * (size_t)n));
    *cap = n;
    return RD_OK;
}

static int launch_scan(rd_handle* h, rd_fq_state* s, int e, const uint8_t* d_buf, int64_t len, int final_chunk,
                       int64_t max_records, int64_t* d_rec, int64_t* d_info, cudaStream_t st) {
    const int64_t ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    int rc = grow(h, &s->d_desc, &s->cap_desc, ntiles);
    if (!rc) rc = grow(h, &s->d_line_end[e], &s->cap_lines[e], 4 * max_records);
    if (rc) return rc;
    RD_CUDA(h, cudaMemsetAsync(d_info, 0, sizeof(int64_t) * 4, st));
    RD_CUDA(h, cudaMemsetAsync(d_info + 4, 0xFF, sizeof(int64_t), st));            // "no bad record"
    if (ntiles > 0) {
        RD_CUDA(h, cudaMemsetAsync(s->d_desc, 0, sizeof(unsigned long long) * ntiles, st));
        RD_CUDA(h, cudaMemsetAsync(s->d_ticket, 0, sizeof(int), st));
        nl_index_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(d_buf, len, ntiles, s->d_desc, s->d_ticket,
                                                                     s->d_line_end[e], 4 * max_records, d_info);
        h->launches += 1;
    }
    const int64_t upper = std::max<int64_t>(1, std::min<int64_t>(max_records, len / 8 + 1));   // a record is >= 8 bytes
    record_kernel<<<(unsigned)((upper + 255) / 256), 256, 0, st>>>(d_buf, len, final_chunk, max_records, s->d_line_end[e],
                                                                   d_rec, d_info);
    h->launches += 1;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

static int launch_partition(rd_handle* h, rd_fq_state* s, const uint8_t* d_buf, const int64_t* d_rec, int64_t n,
                            const int8_t* d_labels, uint8_t* d_out, int64_t* d_sizes3, cudaStream_t st, int nlines = 4) {
    if (n == 0) {
        RD_CUDA(h, cudaMemsetAsync(d_sizes3, 0, sizeof(int64_t) * 3, st));
        return RD_OK;
    }
    const int64_t nblk = (n + 255) / 256;
    int rc = grow(h, &s->d_blocksum, &s->cap_blk, nblk * 3);
    if (rc) return rc;
    part_sum_kernel<<<(unsigned)nblk, 256, 0, st>>>(d_rec, d_labels, n, s->d_blocksum, nlines);
    part_scan_kernel<<<1, 1024, 0, st>>>(s->d_blocksum, nblk, d_sizes3);
    part_copy_kernel<<<(unsigned)nblk, 256, 0, st>>>(d_buf, d_rec, d_labels, n, s->d_blocksum, d_out, nlines);
    h->launches += 3;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

// FASTA: K0 over d_buf[0..len); the joined upper-cased sequences go to d_buf[seq_base ..) (seq_base >= len, capacity len)
static int launch_scan_fasta(rd_handle* h, rd_fq_state* s, int e, uint8_t* d_buf, int64_t len, int64_t seq_base, int final_chunk,
                             int64_t max_records, int64_t* d_rec, int64_t* d_info, cudaStream_t st) {
    const int64_t ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    const int64_t cap = std::max<int64_t>(4, std::min<int64_t>(4 * max_records, len / 2 + 2));      // lines indexed
    const int64_t nblk = (cap + 255) / 256;
    int rc = grow(h, &s->d_desc, &s->cap_desc, ntiles);
    if (!rc) rc = grow(h, &s->d_line_end[e], &s->cap_lines[e], cap);
    if (!rc) rc = grow(h, &s->d_line_desc[e], &s->cap_line_desc[e], cap);
    if (!rc) rc = grow(h, &s->d_seq_at[e], &s->cap_seq_at[e], cap);
    if (!rc) rc = grow(h, &s->d_fa_blocksum, &s->cap_fa_blk, nblk * 2);
    if (rc) return rc;
    RD_CUDA(h, cudaMemsetAsync(d_info, 0, sizeof(int64_t) * 8, st));
    RD_CUDA(h, cudaMemsetAsync(d_info + 4, 0xFF, sizeof(int64_t), st));            // "no bad line"
    if (ntiles > 0) {
        RD_CUDA(h, cudaMemsetAsync(s->d_desc, 0, sizeof(unsigned long long) * ntiles, st));
        RD_CUDA(h, cudaMemsetAsync(s->d_ticket, 0, sizeof(int), st));
        nl_index_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(d_buf, len, ntiles, s->d_desc, s->d_ticket,
                                                                     s->d_line_end[e], cap, d_info);
        h->launches += 1;
    }
    fa_line_kernel<<<(unsigned)std::min<int64_t>(nblk, (int64_t)h->sm_count * 8), 256, 0, st>>>(d_buf, len, final_chunk, s->d_line_end[e], s->d_line_desc[e], cap, d_info);
    fa_sum_kernel<<<(unsigned)nblk, 256, 0, st>>>(d_buf, len, final_chunk, cap, d_info, s->d_line_desc[e], s->d_fa_blocksum);
    fa_scan_kernel<<<1, 1024, 0, st>>>(d_buf, len, final_chunk, cap, s->d_fa_blocksum, nblk, max_records, d_info);
    fa_emit_kernel<<<(unsigned)nblk, 256, 0, st>>>(d_buf, len, final_chunk, cap, s->d_line_desc[e], s->d_line_end[e],
                                                   s->d_fa_blocksum, seq_base, d_rec, max_records, s->d_seq_at[e], d_info);
    fa_copy_kernel<<<(unsigned)(h->sm_count * 8), 256, 0, st>>>(d_buf, s->d_line_desc[e], s->d_seq_at[e], d_info, seq_base, len);
    h->launches += 5;
    RD_CUDA(h, cudaGetLastError());
    return RD_OK;
}

static int fq_fail(rd_handle* h, int code, const std::string& msg) { h->err = msg; return code; }

static int parse_error_fasta(rd_handle* h, int64_t key) {
    h->err = "FASTA: line " + std::to_string(key / 4 + 1) + " of the block holds 4 MiB or more (use --host_ingest)";
    return RD_ERR_PARSE;
}

static int parse_error(rd_handle* h, int64_t key) {
    const int64_t r = key / 4;
    h->err = (key & 3) == 1 ? "FASTQ: blank line in record " + std::to_string(r) + " (the reference parser raises IndexError here)"
                            : "FASTQ: record " + std::to_string(r) + " does not start with '@'";
    return RD_ERR_PARSE;
}

// ---- C ABI ---------------------------------------------------------------------------------------------------
extern "C" int rd_scan_fastq_device(rd_handle* h, const uint8_t* d_buf, int64_t len, int final_chunk, int64_t max_records,
                                    int64_t* d_rec, int64_t* d_info, void* stream) {
    if (!h) return RD_ERR_INVALID;
    if (len < 0 || max_records < 0 || max_records > ((int64_t)1 << 30) || !d_info || (max_records && !d_rec) || (len && !d_buf))
        return fq_fail(h, RD_ERR_INVALID, "rd_scan_fastq_device: bad arguments");
    if (((uintptr_t)d_buf & 15) != 0) return fq_fail(h, RD_ERR_INVALID, "rd_scan_fastq_device: d_buf must be 16-byte aligned");
    RD_CUDA(h, cudaSetDevice(h->device));
    rd_fq_state* s = nullptr;
    int rc = fq_state(h, &s);
    if (rc) return rc;
    return launch_scan(h, s, 0, d_buf, len, final_chunk, max_records, d_rec, d_info, (cudaStream_t)stream);
}

extern "C" int rd_scan_fasta_device(rd_handle* h, uint8_t* d_buf, int64_t len, int64_t seq_base, int final_chunk,
                                    int64_t max_records, int64_t* d_rec, int64_t* d_info, void* stream) {
    if (!h) return RD_ERR_INVALID;
    if (len < 0 || max_records < 0 || max_records > ((int64_t)1 << 30) || !d_info || (max_records && !d_rec) || (len && !d_buf) ||
        seq_base < len)
        return fq_fail(h, RD_ERR_INVALID, "rd_scan_fasta_device: bad arguments");
    if (((uintptr_t)d_buf & 15) != 0) return fq_fail(h, RD_ERR_INVALID, "rd_scan_fasta_device: d_buf must be 16-byte aligned");
    RD_CUDA(h, cudaSetDevice(h->device));
    rd_fq_state* s = nullptr;
    int rc = fq_state(h, &s);
    if (rc) return rc;
    return launch_scan_fasta(h, s, 0, d_buf, len, seq_base, final_chunk, max_records, d_rec, d_info, (cudaStream_t)stream);
}

extern "C" int rd_classify_records(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n, int max_len,
                                   int semantics, int precision, float* d_logits, float* d_probs, int8_t* d_labels,
                                   int64_t* d_counts, void* stream) {
    if (!h) return RD_ERR_INVALID;
    if (n < 0 || n > ((int64_t)1 << 30) || max_len < 1 || max_len > RD_MAX_LEN ||
        (semantics != RD_SEM_PACKED && semantics != RD_SEM_PADDED) || precision < RD_PREC_FP32 || precision > RD_PREC_LAST)
        return fq_fail(h, RD_ERR_INVALID, "rd_classify_records: bad arguments");
    if (n == 0) return RD_OK;
    if (!d_buf || !d_rec || !d_logits) return fq_fail(h, RD_ERR_INVALID, "rd_classify_records: NULL buffer");
    RD_CUDA(h, cudaSetDevice(h->device));
    return rd_classify_device(h, d_buf, d_rec + 2, n, max_len, semantics, precision, d_logits, d_probs, d_labels, d_counts,
                              (cudaStream_t)stream, 8);
}

static int partition_device(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n, const int8_t* d_labels,
                            uint8_t* d_out, int64_t* d_sizes3, void* stream, int nlines) {
    if (!h) return RD_ERR_INVALID;
    if (n < 0 || !d_sizes3 || (n && (!d_buf || !d_rec || !d_labels || !d_out)))
        return fq_fail(h, RD_ERR_INVALID, "rd_partition_records_device: bad arguments");
    RD_CUDA(h, cudaSetDevice(h->device));
    rd_fq_state* s = nullptr;
    int rc = fq_state(h, &s);
    if (rc) return rc;
    return launch_partition(h, s, d_buf, d_rec, n, d_labels, d_out, d_sizes3, (cudaStream_t)stream, nlines);
}

extern "C" int rd_partition_records_device(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n,
                                           const int8_t* d_labels, uint8_t* d_out, int64_t* d_sizes3, void* stream) {
    return partition_device(h, d_buf, d_rec, n, d_labels, d_out, d_sizes3, stream, 4);
}

extern "C" int rd_partition_fasta_device(rd_handle* h, const uint8_t* d_buf, const int64_t* d_rec, int64_t n,
                                         const int8_t* d_labels, uint8_t* d_out, int64_t* d_sizes3, void* stream) {
    return partition_device(h, d_buf, d_rec, n, d_labels, d_out, d_sizes3, stream, 2);
}

static int fastx_submit(rd_handle* h, int format, int slot, int ends, const uint8_t* buf1, int64_t len1, const uint8_t* buf2,
                        int64_t len2, int final_chunk, int64_t max_records, int max_len, int semantics, int precision,
                        int mode, uint8_t* out1, uint8_t* out2, int8_t* labels, int64_t* n_records,
                        int64_t* consumed2, int64_t* out_bytes2) {
    const bool fasta = format == RD_FMT_FASTA;
    if (!h) return RD_ERR_INVALID;
    if (slot < 0 || slot >= rd_fq_state::NSLOT || (ends != 1 && ends != 2) || len1 < 0 || (ends == 2 && len2 < 0) ||
        max_records < 1 || max_records > ((int64_t)1 << 30) || max_len < 1 || max_len > RD_MAX_LEN || !n_records ||
        !consumed2 || !out_bytes2 || (len1 && !buf1) || (ends == 2 && len2 && !buf2) || !out1 || (ends == 2 && !out2) ||
        (semantics != RD_SEM_PACKED && semantics != RD_SEM_PADDED) || precision < RD_PREC_FP32 || precision > RD_PREC_LAST ||
        (ends == 2 && (mode < RD_PAIR_NONE || mode > RD_PAIR_BOTH)))
        return fq_fail(h, RD_ERR_INVALID, "rd_fastq_submit: bad arguments");
    RD_CUDA(h, cudaSetDevice(h->device));
    rd_fq_state* s = nullptr;
    int rc = fq_state(h, &s);
    if (rc) return rc;
    if (s->pending[slot]) return fq_fail(h, RD_ERR_INVALID, "rd_fastq_submit: slot still holds an uncollected block");
    const uint8_t* bufs[2] = {buf1, buf2};
    const int64_t lens[2] = {len1, ends == 2 ? len2 : 0};
    uint8_t* outs[2] = {out1, out2};
    *n_records = 0;
    consumed2[0] = consumed2[1] = 0;
    out_bytes2[0] = out_bytes2[1] = 0;
    int64_t seq_base[2] = {0, 0};                     // FASTA: the joined sequences live behind the text, in the same buffer
    for (int e = 0; e < ends; ++e) {
        seq_base[e] = (lens[e] + 15) & ~(int64_t)15;
        rc = grow(h, &s->d_buf[slot][e], &s->cap_buf[slot][e], fasta ? seq_base[e] + lens[e] + 16 : lens[e] + 16);
        if (!rc) rc = grow(h, &s->d_out[slot][e], &s->cap_out[slot][e], lens[e] + 16);
        if (!rc) rc = grow(h, &s->d_rec[slot][e], &s->cap_rec[slot][e], 8 * max_records);
        if (!rc) rc = grow(h, &s->d_logits[e], &s->cap_logits[e], 2 * max_records);
        if (rc) return rc;
    }
    rc = grow(h, &s->d_labels[slot], &s->cap_labels[slot], max_records);
    if (rc) return rc;
    // H2D + K0 on the copy-in stream; the host needs n and `consumed` before it can cut the next block
    for (int e = 0; e < ends; ++e) {
        if (lens[e])
            RD_CUDA(h, cudaMemcpyAsync(s->d_buf[slot][e], bufs[e], (size_t)lens[e], cudaMemcpyHostToDevice, h->s_in));
        rc = fasta ? launch_scan_fasta(h, s, e, s->d_buf[slot][e], lens[e], seq_base[e], final_chunk, max_records,
                                       s->d_rec[slot][e], s->d_info + 8 * e, h->s_in)
                   : launch_scan(h, s, e, s->d_buf[slot][e], lens[e], final_chunk, max_records, s->d_rec[slot][e], s->d_info + 8 * e, h->s_in);
        if (rc) return rc;
    }
    RD_CUDA(h, cudaMemcpyAsync(s->h_info, s->d_info, sizeof(int64_t) * 16, cudaMemcpyDeviceToHost, h->s_in));
    RD_CUDA(h, cudaStreamSynchronize(h->s_in));
    int64_t n = s->h_info[1];
    if (ends == 2) n = std::min(n, s->h_info[8 + 1]);
    for (int e = 0; e < ends; ++e) {
        const int64_t bad = s->h_info[8 * e + 4];
        if (bad >= 0 && fasta) return parse_error_fasta(h, bad);
        if (bad >= 0 && bad / 4 < n) return parse_error(h, bad);
        consumed2[e] = s->h_info[8 * e + 2];
    }
    if (ends == 2)
        for (int e = 0; e < 2; ++e)
            if (s->h_info[8 * e + 1] > n) {           // this end holds more records than its mate: give the extra ones back
                if (n == 0) { consumed2[e] = 0; continue; }
                if (fasta) {                          // record n of this end starts at the raw line begin kept in its index entry
                    RD_CUDA(h, cudaMemcpyAsync(s->h_info + 16, s->d_rec[slot][e] + (8 * n + 4), sizeof(int64_t), cudaMemcpyDeviceToHost, h->s_in));
                    RD_CUDA(h, cudaStreamSynchronize(h->s_in));
                    consumed2[e] = s->h_info[16];
                    continue;
                }
                RD_CUDA(h, cudaMemcpyAsync(s->h_info + 16, s->d_line_end[e] + (4 * n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, h->s_in));
                RD_CUDA(h, cudaStreamSynchronize(h->s_in));
                consumed2[e] = (int64_t)((unsigned long long)s->h_info[16] & ((1ull << 62) - 1)) + 1;
            }
    *n_records = n;
    if (n == 0) return RD_OK;
    rc = rd_reserve(h, n, max_len);
    if (rc) return rc;
    // K1..K3 and K4 on the compute stream (everything it reads was completed above)
    RD_CUDA(h, cudaMemsetAsync(s->d_res[slot], 0, sizeof(int64_t) * 9, h->s_cmp));
    if (ends == 1) {
        rc = rd_classify_device(h, s->d_buf[slot][0], s->d_rec[slot][0] + 2, n, max_len, semantics, precision, s->d_logits[0],
                                nullptr, s->d_labels[slot], s->d_res[slot] + 6, h->s_cmp, 8);
        if (rc) return rc;
    } else {
        for (int e = 0; e < 2; ++e) {
            rc = rd_classify_device(h, s->d_buf[slot][e], s->d_rec[slot][e] + 2, n, max_len, semantics, precision,
                                    s->d_logits[e], nullptr, nullptr, nullptr, h->s_cmp, 8);
            if (rc) return rc;
        }
        if (mode == RD_PAIR_NONE) {
            const uint8_t* sq[2] = {s->d_buf[slot][0], s->d_buf[slot][1]};
            const int64_t* of[2] = {s->d_rec[slot][0] + 2, s->d_rec[slot][1] + 2};
            float* lg[2] = {s->d_logits[0], s->d_logits[1]};
            rc = rd_pair_none_refine(h, sq, of, n, max_len, semantics, precision, lg, h->s_cmp, 8);
            if (rc) return rc;
        }
        rc = rd_launch_pair(h, s->d_logits[0], s->d_logits[1], n, mode, s->d_labels[slot], s->d_res[slot] + 6, h->s_cmp);
        if (rc) return rc;
    }
    for (int e = 0; e < ends; ++e) {
        rc = launch_partition(h, s, s->d_buf[slot][e], s->d_rec[slot][e], n, s->d_labels[slot], s->d_out[slot][e],
                              s->d_res[slot] + 3 * e, h->s_cmp, fasta ? 2 : 4);
        if (rc) return rc;
    }
    RD_CUDA(h, cudaEventRecord(s->ev_cmp[slot], h->s_cmp));
    RD_CUDA(h, cudaStreamWaitEvent(h->s_out, s->ev_cmp[slot], 0));
    for (int e = 0; e < ends; ++e) {
        // the text of n records is at most the bytes they took in the block (+1: an open last line gains its '\n')
        // (FASTA: "hdr\nSEQ\n" is never longer than the lines it came from, except for a header-less record: + 2)
        out_bytes2[e] = std::min<int64_t>(consumed2[e] + (fasta ? 2 : 1), lens[e] + (fasta ? 2 : 1));
        RD_CUDA(h, cudaMemcpyAsync(outs[e], s->d_out[slot][e], (size_t)out_bytes2[e], cudaMemcpyDeviceToHost, h->s_out));
    }
    if (labels) RD_CUDA(h, cudaMemcpyAsync(labels, s->d_labels[slot], (size_t)n, cudaMemcpyDeviceToHost, h->s_out));
    RD_CUDA(h, cudaMemcpyAsync(s->h_res[slot], s->d_res[slot], sizeof(int64_t) * 9, cudaMemcpyDeviceToHost, h->s_out));
    RD_CUDA(h, cudaEventRecord(s->ev_out[slot], h->s_out));
    s->pending[slot] = true;
    return RD_OK;
}

extern "C" int rd_fastq_submit(rd_handle* h, int slot, int ends, const uint8_t* buf1, int64_t len1, const uint8_t* buf2,
                               int64_t len2, int final_chunk, int64_t max_records, int max_len, int semantics, int precision,
                               int mode, uint8_t* out1, uint8_t* out2, int8_t* labels, int64_t* n_records,
                               int64_t* consumed2, int64_t* out_bytes2) {
    return fastx_submit(h, RD_FMT_FASTQ, slot, ends, buf1, len1, buf2, len2, final_chunk, max_records, max_len, semantics, precision,
                        mode, out1, out2, labels, n_records, consumed2, out_bytes2);
}

extern "C" int rd_fasta_submit(rd_handle* h, int slot, int ends, const uint8_t* buf1, int64_t len1, const uint8_t* buf2,
                               int64_t len2, int final_chunk, int64_t max_records, int max_len, int semantics, int precision,
                               int mode, uint8_t* out1, uint8_t* out2, int8_t* labels, int64_t* n_records,
                               int64_t* consumed2, int64_t* out_bytes2) {
    return fastx_submit(h, RD_FMT_FASTA, slot, ends, buf1, len1, buf2, len2, final_chunk, max_records, max_len, semantics, precision,
                        mode, out1, out2, labels, n_records, consumed2, out_bytes2);
}

extern "C" int rd_fastq_collect(rd_handle* h, int slot, int64_t* sizes6, int64_t* counts3) {
    if (!h) return RD_ERR_INVALID;
    rd_fq_state* s = h->fq;
    if (!s || slot < 0 || slot >= rd_fq_state::NSLOT || !s->pending[slot])
        return RD_ERR_INVALID;                                 // (no message: may run beside rd_fastq_submit on another thread)
    cudaError_t e = cudaEventSynchronize(s->ev_out[slot]);
    s->pending[slot] = false;
    if (e != cudaSuccess) return RD_ERR_CUDA;
    for (int i = 0; i < 6; ++i) if (sizes6) sizes6[i] = s->h_res[slot][i];
    for (int i = 0; i < 3; ++i) if (counts3) counts3[i] = s->h_res[slot][6 + i];
    return RD_OK;
}
