// herro_oracle.cpp — CPU restatement of the HERRO features → (collate) → consensus path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under herro_b200/ may include, link or dlopen this
// file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs can check the CUDA path against an independent, literal CPU
// statement of the reference algorithm.
//
// PARITY STATUS: *parity unpinned* for windowing / features / consensus — the reference
// (lbcb-sci/herro @ 9cd0296) is a Rust crate that cannot be compiled in this image (no
// cargo/rustc, crates not vendored) and it ships no live test or golden vector for these
// modules.  What IS pinned: the 2-bit codec against all 12 known-answer
// tests of src/haec_io.rs:191-299, the token tables of src/features.rs:24-42 /
// src/inference.rs:23-31 / src/consensus.rs:18-19, and the hand-derived window cases of
// SURVEY.md App. E (tests/test_oracle_*.py).  Every function below cites the reference
// lines it follows; the structure (first-pass [L,1+max(n,30)] matrix, HashMap counters,
// re-stacked [L',31] matrix) is kept deliberately literal — the CUDA path is structured
// differently, so agreement between the two is meaningful.
//
// Build: see oracle/Makefile  (g++ -O2 -shared -fPIC).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace ho {

[[noreturn]] static void panic(const std::string& msg) { throw std::runtime_error(msg); }

// ---------------------------------------------------------------------------------------
// haec_io.rs — 2-bit codec (src/haec_io.rs:7-17,121-173)
// ---------------------------------------------------------------------------------------
static uint64_t base_encoding(uint8_t b) {  // src/haec_io.rs:7-15
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 255;  // H12: OR-ed in unmasked by encode()
    }
}
static const uint8_t BASE_DECODING[4] = {'A', 'C', 'G', 'T'};  // src/haec_io.rs:17

struct HAECSeq {
    std::vector<uint64_t> data;
    size_t length = 0;
};

// src/haec_io.rs:121-136
static HAECSeq encode(const uint8_t* seq, size_t n) {
    HAECSeq s;
    s.length = n;
    s.data.reserve((n + 31) / 32);
    uint64_t block = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t c = base_encoding(seq[i]);
        block |= c << ((i << 1) & 63);
        if (((i + 1) & 31) == 0 || i == n - 1) {
            s.data.push_back(block);
            block = 0;
        }
    }
    return s;
}

// src/haec_io.rs:138-173
static void decode(const HAECSeq& s, size_t start, size_t end, bool is_reversed, uint8_t* buffer) {
    if (end > s.length) panic("Out of bounds for 2-bit sequence decoding.");
    if (start >= end) return;
    uint64_t rc_mask = is_reversed ? 3 : 0;
    for (size_t i0 = start; i0 < end; i0++) {
        size_t idx = i0 - start;
        size_t i = i0;
        if (is_reversed) i = end - idx - 1;
        uint64_t code = ((s.data[i >> 5] >> ((i << 1) & 63)) & 3) ^ rc_mask;
        buffer[idx] = BASE_DECODING[code];
    }
}

struct HAECRecord {  // src/haec_io.rs:19-24
    std::string id;
    bool has_description = false;
    std::string description;
    HAECSeq seq;
    std::vector<uint8_t> qual;
};

// ---------------------------------------------------------------------------------------
// overlaps.rs — Overlap / Alignment (src/overlaps.rs:44-101)
// ---------------------------------------------------------------------------------------
enum Strand : uint8_t { Forward = 0, Reverse = 1 };

struct Overlap {
    uint32_t qid, qlen, qstart, qend;
    Strand strand;
    uint32_t tid, tlen, tstart, tend;
    uint32_t return_other_id(uint32_t id) const { return qid == id ? tid : qid; }  // :84-90
};

struct Alignment {
    Overlap overlap;
    std::string cigar;
};

// ---------------------------------------------------------------------------------------
// aligners.rs — CigarIter (src/aligners.rs:252-293)
// ---------------------------------------------------------------------------------------
enum OpKind : uint8_t { Match = 0, Mismatch = 1, Insertion = 2, Deletion = 3 };
struct CigarOp {
    OpKind kind;
    uint32_t len;
};
struct CigarItem {
    CigarOp op;
    size_t start, end;  // byte range relative to the slice handed to the iterator
};

struct CigarIter {
    const uint8_t* data;
    size_t n;
    size_t pos = 0;
    CigarIter(const uint8_t* d, size_t n_) : data(d), n(n_) {}
    bool next(CigarItem& out) {
        if (pos >= n) return false;
        size_t start = pos;
        uint32_t len = 0;
        while (true) {
            if (pos >= n) panic("index out of bounds in CigarIter");  // Rust slice index panic
            uint8_t c = data[pos];
            if (c < '0' || c > '9') break;
            len = len * 10 + (uint32_t)(c - '0');
            pos++;
        }
        if (!(len > 0)) panic("Length has to be longer than 0");
        OpKind k;
        switch (data[pos]) {
            case 'M': k = Match; break;
            case 'I': k = Insertion; break;
            case 'D': k = Deletion; break;
            default: panic(std::string("Unexpected cigar operation ") + (char)data[pos]);
        }
        pos++;
        out.op = {k, len};
        out.start = start;
        out.end = pos;
        return true;
    }
};

// ---------------------------------------------------------------------------------------
// windowing.rs — OverlapWindow + extract_windows (src/windowing.rs:6-16,44-293)
// ---------------------------------------------------------------------------------------
struct OverlapWindow {
    const Overlap* overlap;
    uint32_t ovl_index;  // (ours) which alignment of the target this came from
    uint32_t tstart, qstart, qend;
    size_t cigar_start_idx;
    uint32_t cigar_start_offset;
    size_t cigar_end_idx;
    uint32_t cigar_end_offset;
};
using Windows = std::vector<std::vector<OverlapWindow>>;

// src/windowing.rs:275-293
static CigarOp get_last_cigar_op(const std::string& cigar) {
    uint8_t op = (uint8_t)cigar[cigar.size() - 1];
    uint32_t len = 0, p10 = 1;
    for (size_t k = cigar.size() - 1; k-- > 0;) {
        uint8_t c = (uint8_t)cigar[k];
        if (c < '0' || c > '9') break;
        len += (uint32_t)(c - '0') * p10;
        p10 *= 10;
    }
    switch (op) {
        case 'M': return {Match, len};
        case 'I': return {Insertion, len};
        case 'D': return {Deletion, len};
        default: panic("Invalid cigar op");
    }
}

template <class T>
struct Opt {
    bool some = false;
    T v{};
    void set(T x) { some = true; v = x; }
    T unwrap() const {
        if (!some) panic("called `Option::unwrap()` on a `None` value");
        return v;
    }
};

// src/windowing.rs:44-273
static void extract_windows(Windows& windows, const Overlap* overlap, uint32_t ovl_index,
                            const std::string& cigar, uint32_t tshift, uint32_t qshift,
                            bool is_target, uint32_t window_size) {
    if ((is_target && (overlap->tend - overlap->tstart) < window_size) ||
        ((overlap->qend - overlap->qstart) < window_size))
        return;  // :53-57

    uint32_t first_window, last_window, tstart, tpos, qpos = 0;
    uint32_t zeroth_window_thresh = (uint32_t)(0.1f * (float)window_size);  // :65
    uint32_t nth_window_thresh =
        is_target ? overlap->tlen - zeroth_window_thresh : overlap->qlen - zeroth_window_thresh;

    if (is_target) {  // :74-88
        first_window = overlap->tstart < zeroth_window_thresh
                           ? 0
                           : (overlap->tstart + window_size - 1) / window_size;
        last_window = overlap->tend > nth_window_thresh ? (overlap->tend - 1) / window_size + 1
                                                        : overlap->tend / window_size;
        tstart = overlap->tstart;
        tpos = overlap->tstart;
    } else {  // :89-104
        first_window = overlap->qstart < zeroth_window_thresh
                           ? 0
                           : (overlap->qstart + window_size - 1) / window_size;
        last_window = overlap->qend > nth_window_thresh ? (overlap->qend - 1) / window_size + 1
                                                        : overlap->qend / window_size;
        tstart = overlap->qstart;
        tpos = overlap->qstart;
    }

    // :106  (u32 subtraction: a negative difference would overflow-panic in debug, wrap in
    // release and then not be < 1; with valid PAF input last >= first - see DESIGN.md)
    if ((int64_t)last_window - (int64_t)first_window < 1) return;

    Opt<uint32_t> t_window_start, q_window_start, cigar_start_offset;
    Opt<size_t> cigar_start_idx;

    tpos += tshift;
    qpos += qshift;

    if (tpos % window_size == 0 || tstart < zeroth_window_thresh) {  // :120-125
        t_window_start.set(tpos);
        q_window_start.set(qpos);
        cigar_start_idx.set(0);
        cigar_start_offset.set(0);
    }

    // materialise ops so that peek() is trivial
    std::vector<CigarItem> items;
    {
        CigarIter it((const uint8_t*)cigar.data(), cigar.size());
        CigarItem ci;
        while (it.next(ci)) items.push_back(ci);
    }

    auto push = [&](size_t w, uint32_t ts, uint32_t qs, uint32_t qe, size_t csi, uint32_t cso,
                    size_t cei, uint32_t ceo) {
        if (w >= windows.size()) panic("index out of bounds: windows");
        windows[w].push_back(OverlapWindow{overlap, ovl_index, ts, qs, qe, csi, cso, cei, ceo});
    };

    for (size_t k = 0; k < items.size(); k++) {
        const CigarOp op = items[k].op;
        const size_t rs = items[k].start, re = items[k].end;
        uint32_t tnew, qnew;
        bool is_m = (op.kind == Match || op.kind == Mismatch);
        if (is_m) {
            tnew = tpos + op.len;
            qnew = qpos + op.len;
        } else if (op.kind == Deletion) {
            tnew = tpos + op.len;
            qnew = qpos;
        } else {  // Insertion :132-135
            qpos += op.len;
            continue;
        }

        uint32_t current_w = tpos / window_size;
        uint32_t new_w = tnew / window_size;
        uint32_t diff_w = new_w - current_w;

        if (diff_w == 0) {  // :142-147
            tpos = tnew;
            qpos = qnew;
            continue;
        }

        for (uint32_t i = 1; i < diff_w; i++) {  // :150-195
            uint32_t offset = (current_w + i) * window_size - tpos;
            uint32_t q_start_new = is_m ? qpos + offset : qpos;

            if (cigar_start_idx.some) {
                push((size_t)(current_w + i) - 1, t_window_start.unwrap(), q_window_start.unwrap(),
                     q_start_new, cigar_start_idx.unwrap(), cigar_start_offset.unwrap(), re, offset);
                t_window_start.set(tpos + offset);
                q_window_start.set(is_m ? qpos + offset : qpos);
                cigar_start_idx.set(rs);
                cigar_start_offset.set(offset);
            } else {
                t_window_start.set(tpos + offset);
                q_window_start.set(is_m ? qpos + offset : qpos);
                cigar_start_idx.set(rs);
                cigar_start_offset.set(offset);
            }
        }

        // :197-254
        uint32_t offset = new_w * window_size - tpos;
        uint32_t qend = is_m ? qpos + offset : qpos;

        size_t cigar_end_idx, next_cigar_start_idx;
        uint32_t cigar_end_offset, next_cigar_start_offset;
        if (tnew == new_w * window_size) {
            if (k + 1 < items.size() && items[k + 1].op.kind == Insertion) {
                uint32_t l = items[k + 1].op.len;
                qend += l;
                cigar_end_idx = items[k + 1].end;
                cigar_end_offset = l;
            } else {
                cigar_end_idx = re;
                cigar_end_offset = op.len;
            }
            next_cigar_start_idx = cigar_end_idx;
            next_cigar_start_offset = 0;
        } else {
            cigar_end_idx = re;
            cigar_end_offset = offset;
            next_cigar_start_idx = rs;
            next_cigar_start_offset = cigar_end_offset;
        }

        if (cigar_start_idx.some) {
            push((size_t)new_w - 1, t_window_start.unwrap(), q_window_start.unwrap(), qend,
                 cigar_start_idx.unwrap(), cigar_start_offset.unwrap(), cigar_end_idx,
                 cigar_end_offset);
            t_window_start.set(tpos + offset);
            q_window_start.set(qend);
            cigar_start_idx.set(next_cigar_start_idx);
            cigar_start_offset.set(next_cigar_start_offset);
        } else {
            t_window_start.set(tpos + offset);
            q_window_start.set(qend);
            cigar_start_idx.set(next_cigar_start_idx);
            cigar_start_offset.set(next_cigar_start_offset);
        }

        tpos = tnew;
        qpos = qnew;
    }

    // :260-272
    if (tpos > nth_window_thresh && tpos % window_size != 0) {
        push((size_t)last_window - 1, t_window_start.unwrap(), q_window_start.unwrap(), qpos,
             cigar_start_idx.unwrap(), cigar_start_offset.unwrap(), cigar.size(),
             get_last_cigar_op(cigar).len);
    }
}

// ---------------------------------------------------------------------------------------
// features.rs
// ---------------------------------------------------------------------------------------
static const size_t TOP_K_SORT = 30;  // src/features.rs:22

static uint8_t base_lower(uint8_t b) {  // src/features.rs:24-32
    switch (b) {
        case 'A': return 'a';
        case 'C': return 'c';
        case 'G': return 'g';
        case 'T': return 't';
        default: return 255;
    }
}
static uint8_t base_forward(uint8_t b) {  // src/features.rs:34-42
    switch (b) {
        case '#': case '*': return '*';
        case 'A': case 'a': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        default: return 255;
    }
}

struct SupportedPos {  // src/features.rs:896-900
    uint16_t pos;
    uint8_t ins;
    bool operator==(const SupportedPos& o) const { return pos == o.pos && ins == o.ins; }
};

struct Mat {  // row-major [rows, cols] u8, like ndarray Array2<u8> in standard layout
    size_t rows = 0, cols = 0;
    std::vector<uint8_t> d;
    Mat() {}
    Mat(size_t r, size_t c, uint8_t fill) : rows(r), cols(c), d(r * c, fill) {}
    uint8_t& at(size_t r, size_t c) { return d[r * cols + c]; }
    uint8_t at(size_t r, size_t c) const { return d[r * cols + c]; }
};

// Effective (clipped) op length — src/features.rs:82-90,182-188,591-614 (App. A.3)
static inline uint32_t sub_or_panic(uint32_t a, uint32_t b, const char* what) {
    if (a < b) panic(std::string("attempt to subtract with overflow: ") + what);
    return a - b;
}

// src/features.rs:44-95
static std::vector<uint16_t> get_max_ins_for_window(const std::vector<OverlapWindow>& overlaps,
                                                   const std::vector<const std::string*>& cigar_of_ovl,
                                                   size_t tstart, size_t window_length) {
    std::vector<uint16_t> max_ins(window_length, 0);
    for (const auto& ow : overlaps) {
        size_t tpos = (size_t)ow.tstart - tstart;
        const std::string& cigar = *cigar_of_ovl[ow.ovl_index];
        size_t slice_len = ow.cigar_end_idx - ow.cigar_start_idx;
        CigarIter it((const uint8_t*)cigar.data() + ow.cigar_start_idx, slice_len);
        CigarItem ci;
        while (it.next(ci)) {
            size_t l = ci.op.len;
            if (ci.op.kind == Insertion) {
                if (!(tpos <= max_ins.size())) panic("Length is bigger than the tseq (max_ins)");
                if (tpos == 0) panic("attempt to subtract with overflow: max_ins[tpos - 1]");
                if (tpos - 1 >= max_ins.size()) panic("index out of bounds: max_ins");
                max_ins[tpos - 1] = std::max(max_ins[tpos - 1], (uint16_t)l);
                continue;
            }
            if (ci.start == 0 && ci.end == slice_len) {
                tpos += sub_or_panic(ow.cigar_end_offset, ow.cigar_start_offset, "end_off-start_off");
            } else if (ci.start == 0) {
                tpos += l - ow.cigar_start_offset;
            } else if (ci.end == slice_len) {
                tpos += ow.cigar_end_offset;
            } else {
                tpos += l;
            }
        }
    }
    return max_ins;
}

// src/features.rs:97-108
static std::pair<uint32_t, uint32_t> get_query_region(const OverlapWindow& w, uint32_t tid) {
    uint32_t qstart, qend;
    if (w.overlap->tid == tid) {
        qstart = w.overlap->qstart;
        qend = w.overlap->qend;
    } else {
        qstart = w.overlap->tstart;
        qend = w.overlap->tend;
    }
    if (w.overlap->strand == Forward) return {qstart + w.qstart, qstart + w.qend};
    return {qend - w.qend, qend - w.qstart};
}

// src/features.rs:110-237.  `bases`/`quals` are column `col` of the [L, ncols] matrices.
static void get_features_for_ol_window(Mat& bases, Mat& quals, size_t col, const OverlapWindow& window,
                                       const std::string& cigar_full, const HAECRecord& query,
                                       size_t offset, uint32_t tid, const std::vector<uint16_t>& max_ins,
                                       std::vector<uint8_t>& qbuffer) {
    uint32_t qstart, qend;
    if (window.overlap->tid == tid) {
        qstart = window.overlap->qstart;
        qend = window.overlap->qend;
    } else {
        qstart = window.overlap->tstart;
        qend = window.overlap->tend;
    }
    size_t qlen = (size_t)(window.qend - window.qstart);
    std::vector<uint8_t> qb(qlen), qq(qlen);  // the (base, qual) stream of query_iter
    if (window.overlap->strand == Forward) {
        size_t rs = (size_t)qstart + window.qstart, re = (size_t)qstart + window.qend;
        decode(query.seq, rs, re, false, qbuffer.data());
        if (re > query.qual.size()) panic("qual slice out of range");
        for (size_t i = 0; i < qlen; i++) {
            qb[i] = qbuffer[i];
            qq[i] = query.qual[rs + i];
        }
    } else {
        size_t rs = (size_t)qend - window.qend, re = (size_t)qend - window.qstart;
        decode(query.seq, rs, re, true, qbuffer.data());
        if (re > query.qual.size()) panic("qual slice out of range");
        for (size_t i = 0; i < qlen; i++) {
            qb[i] = base_lower(qbuffer[i]);
            qq[i] = query.qual[re - 1 - i];  // quals.iter().rev()
        }
    }
    size_t qi = 0;
    auto next_query = [&](uint8_t& b, uint8_t& q) {
        if (qi >= qlen) panic("Base and its quality should be present.");
        b = qb[qi];
        q = qq[qi];
        qi++;
    };

    size_t slice_len = window.cigar_end_idx - window.cigar_start_idx;
    CigarIter it((const uint8_t*)cigar_full.data() + window.cigar_start_idx, slice_len);

    uint8_t gap = window.overlap->strand == Forward ? '*' : '#';
    size_t L = bases.rows;
    for (size_t r = 0; r < L; r++) bases.at(r, col) = gap;  // :163

    size_t tpos = offset;
    size_t idx = offset;
    for (size_t i = 0; i < offset; i++) idx += max_ins[i];  // :166
    if (idx > 0) {
        if (idx > L) panic("slice out of range (.. idx)");
        for (size_t r = 0; r < idx; r++) bases.at(r, col) = '.';
    }

    CigarItem ci;
    while (it.next(ci)) {
        size_t l = ci.op.len;
        if (ci.start == 0 && ci.end == slice_len) {
            l = sub_or_panic(window.cigar_end_offset, window.cigar_start_offset, "end_off-start_off");
        } else if (ci.start == 0) {
            if (l < window.cigar_start_offset) panic("attempt to subtract with overflow: l -= start_off");
            l -= window.cigar_start_offset;
        } else if (ci.end == slice_len) {
            l = window.cigar_end_offset;
        }

        switch (ci.op.kind) {
            case Match:
            case Mismatch:
                for (size_t i = 0; i < l; i++) {
                    uint8_t b, q;
                    next_query(b, q);
                    if (idx >= L) panic("index out of bounds: bases[idx]");
                    bases.at(idx, col) = b;
                    quals.at(idx, col) = q;
                    if (tpos + i >= max_ins.size()) panic("index out of bounds: max_ins[tpos+i]");
                    idx += 1 + max_ins[tpos + i];
                }
                tpos += l;
                break;
            case Deletion:
                for (size_t i = 0; i < l; i++) {
                    if (tpos + i >= max_ins.size()) panic("index out of bounds: max_ins[tpos+i]");
                    idx += 1 + max_ins[tpos + i];
                }
                tpos += l;
                break;
            case Insertion: {
                if (tpos == 0) panic("attempt to subtract with overflow: max_ins[tpos-1]");
                size_t mi = max_ins[tpos - 1];
                if (idx < mi) panic("attempt to subtract with overflow: idx -= max_ins");
                idx -= mi;
                for (size_t i = 0; i < l; i++) {
                    uint8_t b, q;
                    next_query(b, q);
                    if (idx + i >= L) panic("index out of bounds: bases[idx+i]");
                    bases.at(idx + i, col) = b;
                    quals.at(idx + i, col) = q;
                }
                idx += mi;
                break;
            }
        }
    }

    if (idx < L) {
        for (size_t r = idx; r < L; r++) bases.at(r, col) = '.';  // :233-236
    }
}

// src/features.rs:239-266
static void write_target_for_window(size_t tstart, const HAECRecord& target,
                                    const std::vector<uint16_t>& max_ins, Mat& bases, Mat& quals,
                                    size_t window_length, const std::vector<uint8_t>& tbuffer) {
    for (size_t r = 0; r < bases.rows; r++) bases.at(r, 0) = '*';
    size_t tpos = 0;
    for (size_t i = 0; i < window_length; i++) {
        bases.at(tpos, 0) = tbuffer[tstart + i];
        quals.at(tpos, 0) = target.qual[tstart + i];
        tpos += 1 + max_ins[i];
    }
}

// src/features.rs:315-324
static bool overlap_window_filter(const uint8_t* cigar, size_t n) {
    CigarIter it(cigar, n);
    CigarItem ci;
    bool long_indel = false;
    while (it.next(ci)) {
        if ((ci.op.kind == Insertion || ci.op.kind == Deletion) && ci.op.len > 50) {
            long_indel = true;
            break;  // Iterator::any short-circuits
        }
    }
    return !long_indel;
}

// src/features.rs:585-679
static float calculate_accuracy(const OverlapWindow& window, const std::string& cigar,
                                const uint8_t* tseq, size_t tlen, const uint8_t* qseq, size_t qlen) {
    size_t tpos = 0, qpos = 0;
    size_t m = 0, s = 0, i = 0, d = 0;
    size_t slice_len = window.cigar_end_idx - window.cigar_start_idx;
    CigarIter it((const uint8_t*)cigar.data() + window.cigar_start_idx, slice_len);
    CigarItem ci;
    while (it.next(ci)) {
        size_t len;
        if (ci.start == 0 && ci.end == slice_len) {
            if (!(window.cigar_end_offset > window.cigar_start_offset)) panic("assert end_off > start_off");
            len = window.cigar_end_offset - window.cigar_start_offset;
        } else if (ci.start == 0) {
            if (!(ci.op.len > window.cigar_start_offset)) panic("assert op_len > start_off");
            len = ci.op.len - window.cigar_start_offset;
        } else if (ci.end == slice_len) {
            len = window.cigar_end_offset;
        } else {
            len = ci.op.len;
        }
        if (!(len > 0)) panic("Operation length cannot be 0");
        if (ci.op.kind != Insertion && !(tpos + len <= tlen)) panic("Length is bigger than the tseq");
        if (ci.op.kind != Deletion && !(qpos + len <= qlen)) panic("Length is bigger than the qseq");
        switch (ci.op.kind) {
            case Match:
                for (size_t j = 0; j < len; j++) {
                    if (tseq[tpos + j] == qseq[qpos + j]) m++;
                    else s++;
                }
                tpos += len;
                qpos += len;
                break;
            case Mismatch: panic("unreachable");
            case Insertion:
                i += len;
                qpos += len;
                break;
            case Deletion:
                d += len;
                tpos += len;
                break;
        }
    }
    return (float)m / (float)(m + s + i + d);  // :678
}

// src/features.rs:681-722
static std::vector<SupportedPos> get_supported(const Mat& bases) {
    std::map<uint8_t, size_t> counter = {{'A', 0}, {'C', 0}, {'G', 0}, {'T', 0}, {'*', 0}};
    std::vector<SupportedPos> supported;
    int16_t tpos = -1;
    uint8_t ins = 0;
    for (size_t r = 0; r < bases.rows; r++) {
        if (bases.at(r, 0) == '*') {
            ins = (uint8_t)(ins + 1);  // u8, wraps in release (H13)
        } else {
            tpos = (int16_t)(tpos + 1);
            ins = 0;
        }
        for (auto& kv : counter) kv.second = 0;
        for (size_t c = 0; c < bases.cols; c++) {
            uint8_t b = bases.at(r, c);
            if (b == '.') continue;
            auto it = counter.find(b < 128 ? base_forward(b) : 255);
            if (it == counter.end()) panic("called `Option::unwrap()` on a `None` value (counter)");
            it->second += 1;
        }
        size_t thresh = (size_t)((double)bases.cols * 0.1);  // :712
        uint8_t n_supported = 0;
        for (auto& kv : counter)
            if (kv.second >= thresh) n_supported++;
        if (n_supported >= 2) supported.push_back(SupportedPos{(uint16_t)tpos, ins});
    }
    return supported;
}

// One emitted window = the arguments of FeaturesOutput::update (src/features.rs:571-579)
struct WindowFeatures {
    uint32_t rid;
    uint16_t wid;
    Mat bases;  // ASCII, [L', 31]
    Mat quals;  // raw Phred+33 bytes, [L', 31]
    std::vector<SupportedPos> supported;
    std::vector<uint32_t> qids;  // read indices of ALL n overlaps in final rank order
    uint16_t n_wids;
};

// src/features.rs:326-583
static std::vector<WindowFeatures> extract_features(uint32_t rid, const std::vector<HAECRecord>& reads,
                                                    const std::vector<Alignment>& overlaps,
                                                    uint32_t window_size, std::vector<uint8_t>& tbuf,
                                                    std::vector<uint8_t>& qbuf) {
    const HAECRecord& read = reads[rid];
    decode(read.seq, 0, read.seq.length, false, tbuf.data());

    size_t n_windows = (read.seq.length + window_size - 1) / window_size;
    Windows windows(n_windows);

    std::vector<const std::string*> cigar_of_ovl(overlaps.size());
    for (size_t a = 0; a < overlaps.size(); a++) {
        const Alignment& al = overlaps[a];
        bool is_target = al.overlap.tid == rid;
        extract_windows(windows, &al.overlap, (uint32_t)a, al.cigar, 0, 0, is_target, window_size);
        cigar_of_ovl[a] = &al.cigar;
    }

    struct First {
        uint16_t i;
        Mat bases, quals;
        std::vector<SupportedPos> supported;
        std::vector<uint32_t> qids;
    };
    std::vector<First> all_features;

    for (size_t i = 0; i < n_windows; i++) {
        size_t win_len = (i == n_windows - 1) ? read.seq.length - i * window_size : window_size;

        // Filter :376-383
        {
            std::vector<OverlapWindow> kept;
            for (const auto& ow : windows[i]) {
                const std::string& cigar = *cigar_of_ovl[ow.ovl_index];
                if (ow.cigar_end_idx < ow.cigar_start_idx || ow.cigar_end_idx > cigar.size())
                    panic("cigar slice out of range");
                if (overlap_window_filter((const uint8_t*)cigar.data() + ow.cigar_start_idx,
                                          ow.cigar_end_idx - ow.cigar_start_idx))
                    kept.push_back(ow);
            }
            windows[i].swap(kept);
        }

        // Sort :386-409 (stable; key = OrderedFloat(-acc))
        {
            std::vector<float> key(windows[i].size());
            for (size_t k = 0; k < windows[i].size(); k++) {
                const OverlapWindow& ow = windows[i][k];
                const std::string& cigar = *cigar_of_ovl[ow.ovl_index];
                size_t tstart = ow.tstart;
                size_t tend = i * window_size + win_len;
                uint32_t qid = ow.overlap->return_other_id(rid);
                auto qr = get_query_region(ow, rid);
                size_t qlen = (size_t)(qr.second - qr.first);
                decode(reads[qid].seq, qr.first, qr.second, ow.overlap->strand == Reverse, qbuf.data());
                if (tstart > tend) panic("slice index starts after end (tbuf)");
                float acc = calculate_accuracy(ow, cigar, tbuf.data() + tstart, tend - tstart, qbuf.data(), qlen);
                key[k] = -acc;
            }
            std::vector<size_t> ord(windows[i].size());
            for (size_t k = 0; k < ord.size(); k++) ord[k] = k;
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return key[a] < key[b]; });
            std::vector<OverlapWindow> sorted;
            sorted.reserve(ord.size());
            for (size_t k : ord) sorted.push_back(windows[i][k]);
            windows[i].swap(sorted);
        }

        std::vector<uint16_t> max_ins =
            get_max_ins_for_window(windows[i], cigar_of_ovl, i * window_size, win_len);

        // get_features_for_window :268-313
        size_t length = max_ins.size();
        for (uint16_t v : max_ins) length += v;
        size_t ncols = 1 + std::max(windows[i].size(), TOP_K_SORT);
        Mat bases(length, ncols, '.'), quals(length, ncols, '!');
        write_target_for_window(i * window_size, read, max_ins, bases, quals, win_len, tbuf);
        for (size_t k = 0; k < windows[i].size(); k++) {
            const OverlapWindow& ow = windows[i][k];
            uint32_t qid = ow.overlap->return_other_id(rid);
            get_features_for_ol_window(bases, quals, k + 1, ow, *cigar_of_ovl[ow.ovl_index], reads[qid],
                                       (size_t)ow.tstart - i * window_size, rid, max_ins, qbuf);
        }

        std::vector<uint32_t> qids;
        for (const auto& ow : windows[i]) qids.push_back(ow.overlap->return_other_id(rid));

        std::vector<SupportedPos> supported = get_supported(bases);
        all_features.push_back(First{(uint16_t)i, std::move(bases), std::move(quals), std::move(supported),
                                     std::move(qids)});
    }

    // Ratios :461-500 (keyed by query id; names are unique so this equals keying by name)
    std::unordered_map<uint32_t, std::pair<double, double>> ratios;
    for (const First& f : all_features) {
        std::vector<size_t> pos_to_idx;
        for (size_t r = 0; r < f.bases.rows; r++)
            if (f.bases.at(r, 0) != '*') pos_to_idx.push_back(r);
        std::unordered_set<size_t> indices;
        for (const auto& s : f.supported) indices.insert(pos_to_idx.at(s.pos) + s.ins);

        for (size_t k = 0; k < f.qids.size(); k++) {
            size_t col = k + 1;
            for (size_t pos = 0; pos < f.bases.rows; pos++) {
                if (!indices.count(pos)) continue;
                uint8_t t = (uint8_t)std::toupper(f.bases.at(pos, 0));
                uint8_t q = (uint8_t)std::toupper(f.bases.at(pos, col));
                if (t == '*') continue;
                auto& e = ratios[f.qids[k]];
                if (q == t) e.first += 1.;
                else e.second += 1.;
            }
        }
    }

    std::vector<WindowFeatures> out;
    for (First& f : all_features) {
        // :503-513
        std::vector<double> iden;
        iden.push_back(std::numeric_limits<double>::max());
        for (uint32_t q : f.qids) {
            auto it = ratios.find(q);
            double s = 0.;
            if (it != ratios.end()) {
                double n = it->second.first, d = it->second.second;
                s = n / (n + d) * std::log(n + d + 1.);
            }
            iden.push_back(s);
        }
        std::vector<size_t> sr(iden.size());
        for (size_t k = 0; k < sr.size(); k++) sr[k] = k;
        std::stable_sort(sr.begin(), sr.end(), [&](size_t a, size_t b) { return iden[a] > iden[b]; });

        std::vector<size_t> cols;
        for (size_t k = 0; k < sr.size() && k < TOP_K_SORT + 1; k++) cols.push_back(sr[k]);
        for (size_t k = sr.size(); k < TOP_K_SORT + 1; k++) cols.push_back(k);
        if (cols.size() != TOP_K_SORT + 1) panic("assert_eq new_bases.len() == TOP_K_SORT + 1");

        // :530-556
        std::vector<size_t> retain_idx;
        for (size_t r = 0; r < f.bases.rows; r++) {
            bool all_gap = true;
            for (size_t c : cols) {
                uint8_t b = f.bases.at(r, c);
                if (b == '.') continue;
                if (!(b == '*' || b == '#')) {
                    all_gap = false;
                    break;
                }
            }
            if (!all_gap) retain_idx.push_back(r);
        }
        WindowFeatures w;
        w.rid = rid;
        w.wid = f.i;
        w.n_wids = (uint16_t)n_windows;
        w.bases = Mat(retain_idx.size(), TOP_K_SORT + 1, 0);
        w.quals = Mat(retain_idx.size(), TOP_K_SORT + 1, 0);
        for (size_t r = 0; r < retain_idx.size(); r++)
            for (size_t c = 0; c < cols.size(); c++) {
                w.bases.at(r, c) = f.bases.at(retain_idx[r], cols[c]);
                w.quals.at(r, c) = f.quals.at(retain_idx[r], cols[c]);
            }
        w.supported = get_supported(w.bases);
        for (size_t k = 1; k < sr.size(); k++) w.qids.push_back(f.qids[sr[k] - 1]);  // :569
        out.push_back(std::move(w));
    }
    return out;
}

// ---------------------------------------------------------------------------------------
// inference.rs — prepare_examples / collate (src/inference.rs:15-31,73-145,214-268)
// ---------------------------------------------------------------------------------------
static const uint8_t BASE_PADDING = 11;  // :15
static const uint8_t QUAL_MAX_VAL = 126; // :17

static uint8_t bases_map(uint8_t b) {  // src/inference.rs:23-31
    switch (b) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case '*': return 4;
        case 'a': return 5;
        case 'c': return 6;
        case 'g': return 7;
        case 't': return 8;
        case '#': return 9;
        case '.': return 10;
        default: return 255;
    }
}

struct ConsensusWindow {  // src/consensus.rs:22-33
    uint32_t rid;
    uint16_t wid;
    uint8_t n_alns;
    uint16_t n_total_wins;
    Mat bases;  // tokens
    Mat quals;
    std::vector<size_t> indices;  // target rows
    std::vector<SupportedPos> supported;
    bool has_logits = false;
    std::vector<float> info_logits;
    std::vector<float> bases_logits;  // [n_supported, 5]
    std::vector<uint32_t> qids;       // (ours) for id dumps
};

struct InferenceBatch {  // src/inference.rs:33-39
    std::vector<uint32_t> wids;  // index into the owning InferenceData::consensus_data
    size_t B = 0, L = 0, R = 0;
    std::vector<uint8_t> bases, quals;      // [B, L, R] u8
    std::vector<int32_t> lens;              // [B]
    std::vector<std::vector<int32_t>> indices;
};

struct InferenceData {
    std::vector<ConsensusWindow> consensus_data;
    std::vector<InferenceBatch> batches;
};

// src/inference.rs:73-145
static InferenceBatch collate(const std::vector<std::pair<uint32_t, const ConsensusWindow*>>& batch) {
    InferenceBatch ib;
    size_t length = 0;
    for (auto& p : batch) length = std::max(length, p.second->bases.rows);
    ib.B = batch.size();
    ib.L = length;
    ib.R = batch[0].second->bases.cols;
    ib.bases.assign(ib.B * ib.L * ib.R, BASE_PADDING);
    ib.quals.assign(ib.B * ib.L * ib.R, QUAL_MAX_VAL);
    for (size_t idx = 0; idx < batch.size(); idx++) {
        const ConsensusWindow* f = batch[idx].second;
        ib.wids.push_back(batch[idx].first);
        size_t l = f->bases.rows;
        std::memcpy(&ib.bases[idx * ib.L * ib.R], f->bases.d.data(), l * ib.R);
        std::memcpy(&ib.quals[idx * ib.L * ib.R], f->quals.d.data(), l * ib.R);
        ib.lens.push_back((int32_t)f->supported.size());
        std::vector<int32_t> tidx;
        for (const auto& sp : f->supported) tidx.push_back((int32_t)(f->indices.at(sp.pos) + sp.ins));
        ib.indices.push_back(std::move(tidx));
    }
    return ib;
}

// src/inference.rs:214-268
static InferenceData prepare_examples(std::vector<WindowFeatures>&& features, size_t batch_size) {
    InferenceData data;
    for (auto& ex : features) {
        ConsensusWindow cw;
        cw.rid = ex.rid;
        cw.wid = ex.wid;
        cw.n_alns = (uint8_t)std::min(ex.qids.size(), TOP_K_SORT);  // src/features.rs:877
        cw.n_total_wins = ex.n_wids;
        cw.bases = std::move(ex.bases);
        for (auto& b : cw.bases.d) b = bases_map(b);  // :222
        cw.quals = std::move(ex.quals);
        for (size_t r = 0; r < cw.bases.rows; r++)  // get_target_indices :255-268
            if (cw.bases.at(r, 0) != bases_map('*')) cw.indices.push_back(r);
        cw.supported = std::move(ex.supported);
        cw.qids = std::move(ex.qids);
        data.consensus_data.push_back(std::move(cw));
    }
    std::vector<std::pair<uint32_t, const ConsensusWindow*>> cur;
    for (uint32_t k = 0; k < data.consensus_data.size(); k++) {
        if (data.consensus_data[k].supported.empty()) continue;
        cur.emplace_back(k, &data.consensus_data[k]);
        if (cur.size() == batch_size) {
            data.batches.push_back(collate(cur));
            cur.clear();
        }
    }
    if (!cur.empty()) data.batches.push_back(collate(cur));
    return data;
}

// ---------------------------------------------------------------------------------------
// consensus.rs — consensus (src/consensus.rs:18-19,86-227)
// ---------------------------------------------------------------------------------------
static const uint8_t BASES_UPPER[10] = {'A', 'C', 'G', 'T', '*', 'A', 'C', 'G', 'T', '*'};
static const size_t BASES_UPPER_COUNTER[10] = {0, 1, 2, 3, 4, 0, 1, 2, 3, 4};

// OrderedFloat total order: NaN greatest, -0 == +0
static bool of_less(float a, float b) {
    bool an = std::isnan(a), bn = std::isnan(b);
    if (an) return false;
    if (bn) return true;
    return a < b;
}

// returns false for None
static bool consensus(const std::vector<const ConsensusWindow*>& data, std::vector<std::vector<uint8_t>>& out) {
    out.clear();
    std::vector<uint8_t> corrected;
    bool any = false;
    size_t st = 0, en = 0;
    for (size_t idx = 0; idx < data.size(); idx++) {
        if (data[idx]->n_alns > 1) {
            if (!any) st = idx;
            en = idx;
            any = true;
        }
    }
    if (!any) return false;
    size_t wid_st = st, wid_en = en + 1;

    uint8_t counts[5];
    for (size_t wi = wid_st; wi < wid_en; wi++) {
        const ConsensusWindow& window = *data[wi];
        if (window.n_alns < 2) {
            if (!corrected.empty()) {
                out.push_back(corrected);
                corrected.clear();
            }
            continue;
        }
        size_t n_rows = (size_t)window.n_alns + 1;

        std::map<std::pair<uint16_t, uint8_t>, const float*> maybe_info;
        if (!window.supported.empty()) {
            if (!window.has_logits) panic("called `Option::unwrap()` on a `None` value (logits)");
            size_t n = std::min(window.supported.size(), window.info_logits.size());
            n = std::min(n, window.bases_logits.size() / 5);
            for (size_t k = 0; k < n; k++)  // later duplicates overwrite (HashMap collect)
                maybe_info[{window.supported[k].pos, window.supported[k].ins}] = &window.bases_logits[k * 5];
        }

        int32_t pos = -1;
        uint8_t ins = 0;
        for (size_t r = 0; r < window.bases.rows; r++) {
            if (window.bases.at(r, 0) == bases_map('*')) ins = (uint8_t)(ins + 1);
            else {
                pos += 1;
                ins = 0;
            }
            auto it = maybe_info.find({(uint16_t)pos, ins});
            if (it != maybe_info.end()) {
                const float* b = it->second;
                // max_by_key returns the LAST maximal element (H9)
                size_t argmax = 0;
                for (size_t k = 1; k < 5; k++)
                    if (!of_less(b[k], b[argmax])) argmax = k;
                static const uint8_t dec[5] = {'A', 'C', 'G', 'T', '*'};
                uint8_t base = dec[argmax];
                if (base != '*') corrected.push_back(base);
            } else {
                for (auto& c : counts) c = 0;
                for (size_t c = 0; c < n_rows; c++) {
                    uint8_t b = window.bases.at(r, c);
                    if (b != bases_map('.')) {
                        if (b >= 10) panic("index out of bounds: BASES_UPPER_COUNTER");
                        counts[BASES_UPPER_COUNTER[b]] = (uint8_t)(counts[BASES_UPPER_COUNTER[b]] + 1);
                    }
                }
                // sorted_by_key(Reverse(count)) — stable — take(2)
                size_t order[5] = {0, 1, 2, 3, 4};
                std::stable_sort(order, order + 5, [&](size_t a, size_t b) { return counts[a] > counts[b]; });
                uint8_t mc0c = counts[order[0]], mc0b = BASES_UPPER[order[0]];
                uint8_t mc1c = counts[order[1]], mc1b = BASES_UPPER[order[1]];
                uint8_t t0 = window.bases.at(r, 0);
                if (t0 >= 10) panic("index out of bounds: BASES_UPPER");
                uint8_t tbase = BASES_UPPER[t0];
                uint8_t base = (mc0c < 2 || (mc0c == mc1c && (mc0b == tbase || mc1b == tbase))) ? tbase : mc0b;
                if (base != '*') corrected.push_back(base);
            }
        }
    }
    if (!corrected.empty()) out.push_back(corrected);
    return true;
}

// src/lib.rs:294-317 (+ :282-288 for the idx rule)
static void write_fasta(std::string& out, const HAECRecord& read, const std::vector<std::vector<uint8_t>>& seqs) {
    for (size_t i = 0; i < seqs.size(); i++) {
        out += '>';
        out += read.id;
        if (seqs.size() == 1) out += ' ';
        else {
            out += ':';
            out += std::to_string(i);
            out += ' ';
        }
        if (read.has_description) out += read.description;
        out += '\n';
        out.append((const char*)seqs[i].data(), seqs[i].size());
        out += '\n';
    }
}

}  // namespace ho

// =======================================================================================
// C ABI (ctypes-friendly).  All functions return 0 on success, <0 on a reference "panic"
// (message via ho_last_error); handles are opaque.
// =======================================================================================
using namespace ho;

static thread_local std::string g_err;
#define HO_TRY try {
#define HO_CATCH                         \
    }                                    \
    catch (const std::exception& e) {    \
        g_err = e.what();                \
        return -1;                       \
    }

struct ho_reads {
    std::vector<HAECRecord> reads;
    size_t max_len = 0;
};

struct ho_target {
    const ho_reads* reads = nullptr;
    uint32_t rid = 0;
    // one InferenceData per flush of InferenceOutput (src/features.rs:884-893)
    std::vector<InferenceData> datas;
    // flat view in window order
    std::vector<ConsensusWindow*> wins;
    std::vector<std::pair<size_t, size_t>> batch_ref;  // (data idx, batch idx)
    bool has_result = false;
    std::vector<std::vector<uint8_t>> segs;
};

extern "C" {

const char* ho_last_error() { return g_err.c_str(); }

// ---- codec ----
int ho_encode(const uint8_t* seq, uint64_t n, uint64_t* out_words /* (n+31)/32 */) {
    HO_TRY
    HAECSeq s = encode(seq, n);
    for (size_t i = 0; i < s.data.size(); i++) out_words[i] = s.data[i];
    return 0;
    HO_CATCH
}
int ho_decode(const uint64_t* words, uint64_t length, uint64_t start, uint64_t end, int rc, uint8_t* out) {
    HO_TRY
    HAECSeq s;
    s.length = length;
    s.data.assign(words, words + (length + 31) / 32);
    decode(s, start, end, rc != 0, out);
    return 0;
    HO_CATCH
}

// ---- token tables (for the table tests) ----
int ho_bases_map(int b) { return bases_map((uint8_t)b); }
int ho_base_lower(int b) { return base_lower((uint8_t)b); }
int ho_base_forward(int b) { return base_forward((uint8_t)b); }

// ---- CigarIter ----
// writes up to cap items: kind(0 M,2 I,3 D), len, start, end ; returns count or -1
int ho_cigar_iter(const uint8_t* cigar, uint64_t n, uint32_t* out4, uint64_t cap) {
    HO_TRY
    CigarIter it(cigar, n);
    CigarItem ci;
    uint64_t k = 0;
    while (it.next(ci)) {
        if (k < cap) {
            out4[4 * k + 0] = ci.op.kind;
            out4[4 * k + 1] = ci.op.len;
            out4[4 * k + 2] = (uint32_t)ci.start;
            out4[4 * k + 3] = (uint32_t)ci.end;
        }
        k++;
    }
    return (int)k;
    HO_CATCH
}

// ---- extract_windows on one alignment ----
// ovl9 = qid,qlen,qstart,qend,strand,tid,tlen,tstart,tend ; out rows of 8 u32:
// window_idx,tstart,qstart,qend,cigar_start_idx,cigar_start_offset,cigar_end_idx,cigar_end_offset
int ho_extract_windows(const uint32_t* ovl9, const uint8_t* cigar, uint64_t cigar_len, int is_target,
                       uint32_t window_size, uint32_t n_windows, uint32_t* out8, uint64_t cap) {
    HO_TRY
    Overlap o{ovl9[0], ovl9[1], ovl9[2], ovl9[3], (Strand)ovl9[4], ovl9[5], ovl9[6], ovl9[7], ovl9[8]};
    Windows w(n_windows);
    std::string cg((const char*)cigar, cigar_len);
    extract_windows(w, &o, 0, cg, 0, 0, is_target != 0, window_size);
    uint64_t k = 0;
    for (uint32_t i = 0; i < n_windows; i++)
        for (auto& ow : w[i]) {
            if (k < cap) {
                uint32_t* r = out8 + 8 * k;
                r[0] = i; r[1] = ow.tstart; r[2] = ow.qstart; r[3] = ow.qend;
                r[4] = (uint32_t)ow.cigar_start_idx; r[5] = ow.cigar_start_offset;
                r[6] = (uint32_t)ow.cigar_end_idx; r[7] = ow.cigar_end_offset;
            }
            k++;
        }
    return (int)k;
    HO_CATCH
}

// ---- read store ----
// seqs/quals are concatenated ASCII; off[n+1]; ids concatenated with id_off[n+1];
// desc_off may be NULL (no descriptions); a read with desc_off[i]==desc_off[i+1] and
// has_desc[i]==0 has None.
ho_reads* ho_reads_new(uint32_t n, const uint8_t* seqs, const uint8_t* quals, const uint64_t* off,
                       const uint8_t* ids, const uint64_t* id_off, const uint8_t* descs,
                       const uint64_t* desc_off, const uint8_t* has_desc) {
    try {
        auto* r = new ho_reads();
        r->reads.resize(n);
        for (uint32_t i = 0; i < n; i++) {
            HAECRecord& rec = r->reads[i];
            size_t len = off[i + 1] - off[i];
            rec.seq = encode(seqs + off[i], len);
            rec.qual.assign(quals + off[i], quals + off[i + 1]);
            rec.id.assign((const char*)ids + id_off[i], id_off[i + 1] - id_off[i]);
            if (descs && has_desc && has_desc[i]) {
                rec.has_description = true;
                rec.description.assign((const char*)descs + desc_off[i], desc_off[i + 1] - desc_off[i]);
            }
            r->max_len = std::max(r->max_len, len);
        }
        return r;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ho_reads_free(ho_reads* r) { delete r; }

// ---- features for one target (extract_features with the InferenceOutput sink) ----
// ovl9[n_ovl*9]; cigars concatenated with cig_off[n_ovl+1]
ho_target* ho_features(const ho_reads* reads, uint32_t rid, uint32_t n_ovl, const uint32_t* ovl9,
                       const uint8_t* cigars, const uint64_t* cig_off, uint32_t window_size,
                       uint32_t batch_size) {
    try {
        std::vector<Alignment> alns(n_ovl);
        for (uint32_t a = 0; a < n_ovl; a++) {
            const uint32_t* o = ovl9 + 9 * a;
            alns[a].overlap = Overlap{o[0], o[1], o[2], o[3], (Strand)o[4], o[5], o[6], o[7], o[8]};
            alns[a].cigar.assign((const char*)cigars + cig_off[a], cig_off[a + 1] - cig_off[a]);
        }
        std::vector<uint8_t> tbuf(reads->max_len + 1), qbuf(reads->max_len + 1);
        std::vector<WindowFeatures> feats = extract_features(rid, reads->reads, alns, window_size, tbuf, qbuf);

        auto* t = new ho_target();
        t->reads = reads;
        t->rid = rid;
        // InferenceOutput::update/emit  (src/features.rs:864-893)
        std::vector<WindowFeatures> pending;
        for (auto& f : feats) {
            pending.push_back(std::move(f));
            if (pending.size() == batch_size) {
                t->datas.push_back(prepare_examples(std::move(pending), batch_size));
                pending.clear();
            }
        }
        t->datas.push_back(prepare_examples(std::move(pending), batch_size));  // emit()
        for (size_t d = 0; d < t->datas.size(); d++) {
            for (auto& cw : t->datas[d].consensus_data) t->wins.push_back(&cw);
            for (size_t b = 0; b < t->datas[d].batches.size(); b++) t->batch_ref.emplace_back(d, b);
        }
        return t;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ho_target_free(ho_target* t) { delete t; }

uint32_t ho_n_windows(const ho_target* t) { return (uint32_t)t->wins.size(); }
// info6: L', n_alns, n_supported, n_qids, wid, n_total_wins
int ho_window_info(const ho_target* t, uint32_t w, uint32_t* info6) {
    if (w >= t->wins.size()) return -1;
    const ConsensusWindow& cw = *t->wins[w];
    info6[0] = (uint32_t)cw.bases.rows;
    info6[1] = cw.n_alns;
    info6[2] = (uint32_t)cw.supported.size();
    info6[3] = (uint32_t)cw.qids.size();
    info6[4] = cw.wid;
    info6[5] = cw.n_total_wins;
    return 0;
}
// bases (tokens) / quals [L',31]; supported as (pos,ins) u32 pairs; sup_rows = indices[pos]+ins;
// qids in final rank order.  Any pointer may be NULL.
int ho_window_get(const ho_target* t, uint32_t w, uint8_t* bases, uint8_t* quals, uint32_t* supported2,
                  uint32_t* sup_rows, uint32_t* qids) {
    if (w >= t->wins.size()) return -1;
    const ConsensusWindow& cw = *t->wins[w];
    if (bases) std::memcpy(bases, cw.bases.d.data(), cw.bases.d.size());
    if (quals) std::memcpy(quals, cw.quals.d.data(), cw.quals.d.size());
    for (size_t k = 0; k < cw.supported.size(); k++) {
        if (supported2) {
            supported2[2 * k] = cw.supported[k].pos;
            supported2[2 * k + 1] = cw.supported[k].ins;
        }
        if (sup_rows) sup_rows[k] = (uint32_t)(cw.indices.at(cw.supported[k].pos) + cw.supported[k].ins);
    }
    if (qids)
        for (size_t k = 0; k < cw.qids.size(); k++) qids[k] = cw.qids[k];
    return 0;
}

// ---- reference batches (collate) ----
uint32_t ho_n_batches(const ho_target* t) { return (uint32_t)t->batch_ref.size(); }
// shape3 = B, Lmax, R
int ho_batch_shape(const ho_target* t, uint32_t b, uint32_t* shape3) {
    if (b >= t->batch_ref.size()) return -1;
    const InferenceBatch& ib = t->datas[t->batch_ref[b].first].batches[t->batch_ref[b].second];
    shape3[0] = (uint32_t)ib.B; shape3[1] = (uint32_t)ib.L; shape3[2] = (uint32_t)ib.R;
    return 0;
}
// bases/quals [B,Lmax,R] u8; lens [B] i32; win_index[B] = flat window index (ho_window_*);
// indices_flat [sum lens] i32
int ho_batch_get(const ho_target* t, uint32_t b, uint8_t* bases, uint8_t* quals, int32_t* lens,
                 uint32_t* win_index, int32_t* indices_flat) {
    if (b >= t->batch_ref.size()) return -1;
    size_t d = t->batch_ref[b].first;
    const InferenceBatch& ib = t->datas[d].batches[t->batch_ref[b].second];
    size_t base = 0;
    for (size_t k = 0; k < d; k++) base += t->datas[k].consensus_data.size();
    if (bases) std::memcpy(bases, ib.bases.data(), ib.bases.size());
    if (quals) std::memcpy(quals, ib.quals.data(), ib.quals.size());
    size_t o = 0;
    for (size_t k = 0; k < ib.B; k++) {
        if (lens) lens[k] = ib.lens[k];
        if (win_index) win_index[k] = (uint32_t)(base + ib.wids[k]);
        for (int32_t v : ib.indices[k]) {
            if (indices_flat) indices_flat[o] = v;
            o++;
        }
    }
    return 0;
}

// scatter logits back (src/inference.rs:196-207)
int ho_set_logits(ho_target* t, uint32_t w, const float* info, const float* bases5, uint32_t n) {
    if (w >= t->wins.size()) return -1;
    ConsensusWindow& cw = *t->wins[w];
    cw.info_logits.assign(info, info + n);
    cw.bases_logits.assign(bases5, bases5 + 5 * (size_t)n);
    cw.has_logits = true;
    return 0;
}

// ---- consensus (consensus_worker: sort by wid, consensus) ----
// returns n_segs >= 0 (Some), -2 for None, -1 error
int ho_consensus(ho_target* t) {
    HO_TRY
    std::vector<const ConsensusWindow*> wins(t->wins.begin(), t->wins.end());
    std::stable_sort(wins.begin(), wins.end(),
                     [](const ConsensusWindow* a, const ConsensusWindow* b) { return a->wid < b->wid; });
    bool some = consensus(wins, t->segs);
    t->has_result = some;
    if (!some) return -2;
    return (int)t->segs.size();
    HO_CATCH
}
uint64_t ho_seg_len(const ho_target* t, uint32_t s) { return s < t->segs.size() ? t->segs[s].size() : 0; }
int ho_seg_get(const ho_target* t, uint32_t s, uint8_t* out) {
    if (s >= t->segs.size()) return -1;
    std::memcpy(out, t->segs[s].data(), t->segs[s].size());
    return 0;
}
// FASTA text for this target (empty when consensus was None); returns length, copies up to cap
int64_t ho_fasta(const ho_target* t, uint8_t* out, uint64_t cap) {
    std::string s;
    if (t->has_result) write_fasta(s, t->reads->reads[t->rid], t->segs);
    if (out && cap >= s.size()) std::memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
}

}  // extern "C"
