// rd_fastx.cu — host side of the path's edges: FASTQ/FASTA record scanning and
// label-partitioned record writing.  Pure host C++ (no kernels; lives in librd_b200.so so that one
// library carries the whole C ABI).
//
// Replaces  seq_parser                    ribodetector/data_loader/fastx_parser.py:15-55
//           the record text '\n'.join(r)  ribodetector/detect.py:680,711-712
//           separate_reads routing        ribodetector/detect.py:601-614 (labels come from K3)
// The reference parses with Python str methods, one line at a time (~1e5 reads/s per process); this
// is one memchr-driven pass over the byte buffer.  Semantics kept: FASTQ lines are rstrip()ped and
// NOT upper-cased, FASTA lines are strip()ped, joined and upper-cased, a truncated final FASTQ record
// is dropped, blank lines inside a FASTQ file are an error (the reference raises IndexError).
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../../include/rd_b200.h"

namespace {

thread_local std::string g_fx_err;

inline bool py_space(unsigned char c) {        // what str.strip() removes for ASCII text
    return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f);
}

}  // namespace

extern "C" const char* rd_fastx_last_error(void) { return g_fx_err.c_str(); }

// ---- FASTQ: strictly 4 lines per record (the reference's state machine), so once the newline positions
// are known every record is independent: count newlines per segment in parallel, prefix-sum, index.
static int64_t scan_fastq(const uint8_t* buf, int64_t len, int final_chunk, int64_t max_records, int64_t* hdr,
                          int64_t* plus, int64_t* qual, uint8_t* seq_out, int64_t seq_cap, int64_t* seq_off,
                          int64_t* consumed, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (len < (1 << 20)) threads = 1;
    std::vector<int64_t> cnt((size_t)threads + 1, 0);
    auto seg = [&](int t) { return len * t / threads; };
    auto par = [&](auto&& fn) {
        if (threads == 1) { fn(0); return; }
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(fn, t);
        for (auto& th : pool) th.join();
    };
    par([&](int t) {
        int64_t c = 0;
        const uint8_t* p = buf + seg(t);
        const uint8_t* e = buf + seg(t + 1);
        while (p < e) {
            const uint8_t* q = static_cast<const uint8_t*>(memchr(p, '\n', (size_t)(e - p)));
            if (!q) break;
            ++c; p = q + 1;
        }
        cnt[(size_t)t + 1] = c;
    });
    for (int t = 0; t < threads; ++t) cnt[(size_t)t + 1] += cnt[(size_t)t];
    const int64_t n_nl = cnt[(size_t)threads];
    int64_t n_lines = n_nl;
    const bool open_tail = final_chunk && len > 0 && buf[len - 1] != '\n';
    if (open_tail) ++n_lines;                                   // last line without a newline
    int64_t n = n_lines / 4;                                    // a truncated final record is dropped / left for later
    if (n > max_records) n = max_records;
    seq_off[0] = 0;
    if (n == 0) { *consumed = 0; return 0; }
    // line_end[i] = position of the newline ending line i (len for an open last line); only 4n needed
    std::vector<int64_t> line_end((size_t)(4 * n));
    par([&](int t) {
        int64_t i = cnt[(size_t)t];
        const uint8_t* p = buf + seg(t);
        const uint8_t* e = buf + seg(t + 1);
        while (p < e && i < 4 * n) {
            const uint8_t* q = static_cast<const uint8_t*>(memchr(p, '\n', (size_t)(e - p)));
            if (!q) break;
            line_end[(size_t)i++] = q - buf;
            p = q + 1;
        }
    });
    if (4 * n > n_nl) line_end[(size_t)(4 * n - 1)] = len;     // the open last line closes the last record
    std::vector<int64_t> bad((size_t)threads, -1);
    std::vector<int> bad_kind((size_t)threads, 0);
    par([&](int t) {
        const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
        for (int64_t r = lo; r < hi; ++r) {
            int64_t b[4], e[4];
            for (int k = 0; k < 4; ++k) {
                const int64_t li = 4 * r + k;
                b[k] = li == 0 ? 0 : line_end[(size_t)(li - 1)] + 1;
                int64_t x = line_end[(size_t)li];
                while (x > b[k] && py_space(buf[x - 1])) --x;   // line.rstrip()
                e[k] = x;
                if (x == b[k] && bad[(size_t)t] < 0) { bad[(size_t)t] = r; bad_kind[(size_t)t] = 1; }
            }
            if (e[0] > b[0] && buf[b[0]] != '@' && bad[(size_t)t] < 0) { bad[(size_t)t] = r; bad_kind[(size_t)t] = 2; }
            hdr[2 * r] = b[0]; hdr[2 * r + 1] = e[0];
            plus[2 * r] = b[2]; plus[2 * r + 1] = e[2];
            qual[2 * r] = b[3]; qual[2 * r + 1] = e[3];
            seq_off[r + 1] = e[1] - b[1];                       // length for now
        }
    });
    for (int t = 0; t < threads; ++t)
        if (bad[(size_t)t] >= 0) {
            g_fx_err = bad_kind[(size_t)t] == 1
                           ? "FASTQ: blank line in record " + std::to_string(bad[(size_t)t]) + " (the reference parser raises IndexError here)"
                           : "FASTQ: record " + std::to_string(bad[(size_t)t]) + " does not start with '@'";
            return -RD_ERR_PARSE;
        }
    int64_t fit = n;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t nxt = seq_off[r] + seq_off[r + 1];
        if (nxt > seq_cap) { fit = r; break; }
        seq_off[r + 1] = nxt;
    }
    n = fit;
    par([&](int t) {
        const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
        for (int64_t r = lo; r < hi; ++r) {
            const int64_t sb = (4 * r + 1 == 0 ? 0 : line_end[(size_t)(4 * r)] + 1);
            memcpy(seq_out + seq_off[r], buf + sb, (size_t)(seq_off[r + 1] - seq_off[r]));
        }
    });
    if (n == 0) { *consumed = 0; return 0; }
    const int64_t last = line_end[(size_t)(4 * n - 1)];
    *consumed = last >= len ? len : last + 1;
    return n;
}

extern "C" int64_t rd_scan_fastx(const uint8_t* buf, int64_t len, int format, int final_chunk, int64_t max_records,
                                 int64_t* hdr, int64_t* plus, int64_t* qual, uint8_t* seq_out, int64_t seq_cap,
                                 int64_t* seq_off, int64_t* consumed, int threads) {
    g_fx_err.clear();
    if (!buf || len < 0 || max_records < 0 || !hdr || !seq_out || !seq_off || !consumed ||
        (format != RD_FMT_FASTQ && format != RD_FMT_FASTA) || (format == RD_FMT_FASTQ && (!plus || !qual))) {
        g_fx_err = "rd_scan_fastx: bad arguments";
        return -RD_ERR_INVALID;
    }
    int64_t n = 0, sfill = 0;
    seq_off[0] = 0;
    *consumed = 0;
    if (format == RD_FMT_FASTQ)
        return scan_fastq(buf, len, final_chunk, max_records, hdr, plus, qual, seq_out, seq_cap, seq_off, consumed, threads);
    // FASTA: a record is complete when the next header line (or EOF) is seen.  Every call starts at a
    // record boundary: an unfinished record is rolled back and rescanned with the next chunk.
    bool have_header = false;        // a '>' line opened the current record
    bool open = false;               // current record has a header or (file start only) header-less sequence
    int64_t rec_start = 0, hb = 0, he = 0, srec = 0, p = 0;
    auto rollback = [&]() { sfill = srec; *consumed = open ? rec_start : p; seq_off[n] = sfill; };
    while (true) {
        if (p >= len) {
            if (!final_chunk) { rollback(); return n; }
            if (sfill > srec && n < max_records) {                 // `if seq != ''` at EOF (fastx_parser.py:54-55)
                hdr[2 * n] = hb; hdr[2 * n + 1] = he;
                seq_off[++n] = sfill;
                *consumed = len;
            } else if (sfill > srec) {
                rollback();                                         // no room: hand it out on the next call
            } else {
                sfill = srec; *consumed = len;                      // trailing header without sequence is dropped
            }
            seq_off[n] = sfill;
            return n;
        }
        const uint8_t* nl = static_cast<const uint8_t*>(memchr(buf + p, '\n', (size_t)(len - p)));
        if (!nl && !final_chunk) { rollback(); return n; }          // partial last line
        const int64_t le = nl ? nl - buf : len;
        int64_t l = p, r = le;
        while (r > l && py_space(buf[r - 1])) --r;
        while (l < r && py_space(buf[l])) ++l;                      // line.strip()
        if (r > l) {
            if (buf[l] == '>') {
                if (have_header) {                                  // closes the previous record, even an empty one
                    hdr[2 * n] = hb; hdr[2 * n + 1] = he;
                    seq_off[++n] = sfill;
                    srec = sfill;
                    open = false;
                    if (n == max_records) { *consumed = p; return n; }
                }
                // (sequence lines before the very first header stay attached to it, like the reference)
                if (!open) rec_start = p;
                open = true; have_header = true; hb = l; he = r;
            } else {
                if (!open) { open = true; rec_start = p; hb = he = l; }
                if (sfill + (r - l) > seq_cap) {
                    if (n == 0) { g_fx_err = "FASTA: a single record exceeds the sequence buffer"; return -RD_ERR_NOMEM; }
                    rollback();
                    return n;
                }
                for (int64_t i = l; i < r; ++i) {
                    const uint8_t c = buf[i];
                    seq_out[sfill++] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;   // .upper()
                }
            }
        }
        p = nl ? le + 1 : len;
    }
}

// One output stream: header \n seq \n [plus \n qual \n] for every record whose label == want.
static int64_t emit_class(const uint8_t* buf, int format, int64_t lo, int64_t hi, const int64_t* hdr, const int64_t* plus,
                          const int64_t* qual, const uint8_t* seq, const int64_t* seq_off, const int8_t* labels, int want,
                          uint8_t* out) {
    int64_t w = 0;
    for (int64_t i = lo; i < hi; ++i) {
        if (labels[i] != want) continue;
        const int64_t hl = hdr[2 * i + 1] - hdr[2 * i], sl = seq_off[i + 1] - seq_off[i];
        if (out) { memcpy(out + w, buf + hdr[2 * i], (size_t)hl); out[w + hl] = '\n'; }
        w += hl + 1;
        if (out) { memcpy(out + w, seq + seq_off[i], (size_t)sl); out[w + sl] = '\n'; }
        w += sl + 1;
        if (format == RD_FMT_FASTQ) {
            const int64_t pl = plus[2 * i + 1] - plus[2 * i], ql = qual[2 * i + 1] - qual[2 * i];
            if (out) { memcpy(out + w, buf + plus[2 * i], (size_t)pl); out[w + pl] = '\n'; }
            w += pl + 1;
            if (out) { memcpy(out + w, buf + qual[2 * i], (size_t)ql); out[w + ql] = '\n'; }
            w += ql + 1;
        }
    }
    return w;
}

extern "C" int rd_partition_records(const uint8_t* buf, int format, int64_t n, const int64_t* hdr, const int64_t* plus,
                                    const int64_t* qual, const uint8_t* seq, const int64_t* seq_off, const int8_t* labels,
                                    uint8_t* out_non, uint8_t* out_rrna, uint8_t* out_unc, int64_t* sizes3, int threads) {
    g_fx_err.clear();
    if (!buf || n < 0 || !hdr || !seq || !seq_off || !labels || !sizes3 ||
        (format != RD_FMT_FASTQ && format != RD_FMT_FASTA) || (format == RD_FMT_FASTQ && (!plus || !qual))) {
        g_fx_err = "rd_partition_records: bad arguments";
        return RD_ERR_INVALID;
    }
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n < 4096) threads = 1;
    uint8_t* outs[3] = {out_non, out_rrna, out_unc};
    const int wants[3] = {0, 1, -1};
    // pass 1: bytes per (thread range, class); pass 2: fill at the prefix offsets
    std::vector<int64_t> part((size_t)threads * 3, 0);
    auto span = [&](int t, int64_t& lo, int64_t& hi) { lo = n * t / threads; hi = n * (t + 1) / threads; };
    auto run = [&](bool fill) {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) {
            auto work = [&, t]() {
                int64_t lo, hi;
                span(t, lo, hi);
                for (int c = 0; c < 3; ++c) {
                    if (fill) {
                        if (!outs[c]) continue;
                        int64_t base = 0;
                        for (int u = 0; u < t; ++u) base += part[(size_t)u * 3 + c];
                        emit_class(buf, format, lo, hi, hdr, plus, qual, seq, seq_off, labels, wants[c], outs[c] + base);
                    } else {
                        part[(size_t)t * 3 + c] = emit_class(buf, format, lo, hi, hdr, plus, qual, seq, seq_off, labels, wants[c], nullptr);
                    }
                }
            };
            if (threads == 1) work(); else pool.emplace_back(work);
        }
        for (auto& th : pool) th.join();
    };
    run(false);
    for (int c = 0; c < 3; ++c) {
        sizes3[c] = 0;
        for (int t = 0; t < threads; ++t) sizes3[c] += part[(size_t)t * 3 + c];
    }
    if (out_non || out_rrna || out_unc) run(true);
    return RD_OK;
}

// ---- parallel inflate of BGZF text (bgzip, bcl2fastq / BCL Convert FASTQ.gz): every gzip member carries its own
// compressed size in a 'BC' extra subfield, so member boundaries are known without inflating and the members
// (<= 64 KB of text each) can be inflated side by side.  The reference reads .gz through Python's gzip module on
// one thread (seq_encoder.py:43-53 `gzip.open`), which is what bounds it on compressed inputs.
#include <zlib.h>

namespace {
struct BgzfMember { int64_t in_off; int32_t in_len, hdr_len; int64_t out_off; uint32_t isize; };

// -> total member size, or 0 if [p, p+avail) does not start with a complete BGZF member header, or -1 if it is not BGZF
inline int64_t bgzf_member_size(const uint8_t* p, int64_t avail, int32_t* hdr_len) {
    if (avail < 18) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return -1;
    const int xlen = p[10] | (p[11] << 8);
    if (avail < 12 + xlen) return 0;
    int64_t bsize = -1;
    for (int i = 12; i + 4 <= 12 + xlen;) {
        const int slen = p[i + 2] | (p[i + 3] << 8);
        if (p[i] == 'B' && p[i + 1] == 'C' && slen == 2 && i + 6 <= 12 + xlen) bsize = (p[i + 4] | (p[i + 5] << 8)) + 1;
        i += 4 + slen;
    }
    if (bsize < 0 || (p[3] & ~4)) return -1;          // no BC subfield, or name/comment/hcrc fields we do not expect in BGZF
    *hdr_len = 12 + xlen;
    return bsize;
}
}  // namespace

extern "C" int64_t rd_bgzf_inflate(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_cap, int64_t* in_used,
                                   int threads) {
    g_fx_err.clear();
    if (!in || in_len < 0 || !out || out_cap < 0 || !in_used) { g_fx_err = "rd_bgzf_inflate: bad arguments"; return -RD_ERR_INVALID; }
    *in_used = 0;
    std::vector<BgzfMember> mem;
    int64_t ip = 0, op = 0;
    while (ip < in_len) {
        int32_t hl = 0;
        const int64_t sz = bgzf_member_size(in + ip, in_len - ip, &hl);
        if (sz < 0) {
            if (mem.empty()) { g_fx_err = "not a BGZF stream (gzip member without a 'BC' size field)"; return -RD_ERR_UNSUPPORTED; }
            break;                                       // (the caller sees in_used < in_len and an unreadable next member)
        }
        if (sz == 0 || ip + sz > in_len || sz < hl + 8) break;       // incomplete member: wait for more input
        const uint8_t* t = in + ip + sz - 4;
        const uint32_t isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        if (op + (int64_t)isize > out_cap) break;                   // output full
        mem.push_back({ip, (int32_t)sz, hl, op, isize});
        ip += sz;
        op += isize;
    }
    if (mem.empty()) return 0;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if ((int64_t)mem.size() < 4 * threads) threads = 1;
    std::vector<int> bad((size_t)threads, 0);
    auto work = [&](int t) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { bad[(size_t)t] = 1; return; }
        const size_t lo = mem.size() * (size_t)t / (size_t)threads, hi = mem.size() * (size_t)(t + 1) / (size_t)threads;
        for (size_t i = lo; i < hi; ++i) {
            const BgzfMember& m = mem[i];
            zs.next_in = const_cast<Bytef*>(in + m.in_off + m.hdr_len);
            zs.avail_in = (uInt)(m.in_len - m.hdr_len - 8);
            zs.next_out = out + m.out_off;
            zs.avail_out = m.isize;
            const int rc = inflate(&zs, Z_FINISH);
            const uint8_t* c = in + m.in_off + m.in_len - 8;
            const uint32_t want = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24);
            if (rc != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(0L, out + m.out_off, m.isize) != want) { bad[(size_t)t] = 1; break; }
            inflateReset(&zs);
        }
        inflateEnd(&zs);
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < threads; ++t)
        if (bad[(size_t)t]) { g_fx_err = "BGZF: corrupt member (inflate or CRC-32 failed)"; return -RD_ERR_PARSE; }
    *in_used = ip;
    return op;
}

// ---- serial gzip stream (any .gz that is not BGZF-framed): zlib driven directly on the caller's buffers — no
// intermediate bytes objects, CRC-32 checked inside zlib, concatenated members handled.  Replaces `gzip.open`
// (seq_encoder.py:43-53) where the members cannot be located without inflating.
struct rd_gz { z_stream zs; bool in_member; };

extern "C" rd_gz* rd_gz_open(void) {
    rd_gz* g = new (std::nothrow) rd_gz();
    if (!g) return nullptr;
    memset(&g->zs, 0, sizeof(g->zs));
    if (inflateInit2(&g->zs, 15 + 16) != Z_OK) { delete g; return nullptr; }
    g->in_member = false;
    return g;
}

extern "C" void rd_gz_close(rd_gz* g) {
    if (!g) return;
    inflateEnd(&g->zs);
    delete g;
}

// Inflates from in[0..in_len) into out[0..out_cap) until either is exhausted.  Returns the bytes written (>= 0) or
// -RD_ERR_PARSE; *in_used = input consumed; *mid_member = 1 if the stream stands inside a member (so end of file here
// means a truncated file).  Zero padding between / after members is skipped like Python's gzip module does.
extern "C" int64_t rd_gz_inflate(rd_gz* g, const uint8_t* in, int64_t in_len, int64_t* in_used, uint8_t* out,
                                 int64_t out_cap, int* mid_member) {
    g_fx_err.clear();
    if (!g || in_len < 0 || out_cap < 0 || !in_used || (in_len && !in) || (out_cap && !out)) {
        g_fx_err = "rd_gz_inflate: bad arguments";
        return -RD_ERR_INVALID;
    }
    int64_t ip = 0, op = 0;
    while (ip < in_len && op < out_cap) {
        if (!g->in_member) {
            while (ip < in_len && in[ip] == 0) ++ip;                  // padding between members
            if (ip == in_len) break;
            g->in_member = true;
        }
        const int64_t ni = std::min<int64_t>(in_len - ip, 1 << 30), no = std::min<int64_t>(out_cap - op, 1 << 30);
        g->zs.next_in = const_cast<Bytef*>(in + ip);
        g->zs.avail_in = (uInt)ni;
        g->zs.next_out = out + op;
        g->zs.avail_out = (uInt)no;
        const int rc = inflate(&g->zs, Z_NO_FLUSH);
        ip += ni - (int64_t)g->zs.avail_in;
        op += no - (int64_t)g->zs.avail_out;
        if (rc == Z_STREAM_END) {
            g->in_member = false;
            if (inflateReset(&g->zs) != Z_OK) { g_fx_err = "gzip: inflateReset failed"; return -RD_ERR_PARSE; }
        } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
            g_fx_err = std::string("gzip: corrupt stream (") + (g->zs.msg ? g->zs.msg : "inflate failed") + ")";
            return -RD_ERR_PARSE;
        } else if (rc == Z_BUF_ERROR && g->zs.avail_in != 0 && g->zs.avail_out != 0) {
            g_fx_err = "gzip: inflate made no progress";
            return -RD_ERR_PARSE;
        }
    }
    *in_used = ip;
    if (mid_member) *mid_member = g->in_member ? 1 : 0;
    return op;
}
