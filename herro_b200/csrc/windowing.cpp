// windowing.cpp — host-side windowing for hb_submit_alignments().
//
// In the deployed layout the Rust host keeps windowing::extract_windows
// (src/windowing.rs:44-273) and calls hb_submit_target(); this file exists for hosts that
// hand over raw alignments and for the C++/Python harness that stands in for the Rust binary
// (no Rust toolchain offline).  Semantics follow SURVEY.md App. G; the CIGAR is tokenised
// once into arrays and a small cursor walks window boundaries.
#include "windowing.h"

#include <vector>

namespace hb {

namespace {
struct Op {
    char kind;       // 'M' 'I' 'D'
    uint32_t len;
    uint32_t b0, b1;  // byte range in the CIGAR text
};

bool tokenize(const uint8_t* cg, uint32_t n, std::vector<Op>& ops) {
    uint32_t i = 0;
    while (i < n) {
        const uint32_t b0 = i;
        uint64_t v = 0;
        uint32_t nd = 0;
        while (i < n && cg[i] >= '0' && cg[i] <= '9') {
            v = v * 10 + (cg[i] - '0');
            if (v > 0xffffffffull) return false;
            i++;
            nd++;
        }
        if (i >= n || nd == 0 || v == 0) return false;
        const char k = (char)cg[i++];
        if (k != 'M' && k != 'I' && k != 'D') return false;
        ops.push_back(Op{k, (uint32_t)v, b0, i});
    }
    return !ops.empty();
}
}  // namespace

int host_extract_windows(const hb_overlap& o, uint32_t overlap_idx, uint32_t W, uint32_t n_windows,
                         std::vector<hb_overlap_window>& out) {
    // admission (src/windowing.rs:53-57) — the inference path always has is_target == true
    if (o.tend < o.tstart || o.qend < o.qstart) return -1;
    if (o.tend - o.tstart < W || o.qend - o.qstart < W) return 0;
    std::vector<Op> ops;
    if (!o.cigar || !tokenize(o.cigar, o.cigar_len, ops)) return -1;

    const uint32_t edge = (uint32_t)(0.1f * (float)W);  // :65
    if (o.tlen < edge) return -1;
    const uint32_t tail_thresh = o.tlen - edge;
    const uint32_t first_w = o.tstart < edge ? 0 : (o.tstart + W - 1) / W;                     // :75-79
    const uint32_t last_w = o.tend > tail_thresh ? (o.tend - 1) / W + 1 : o.tend / W;          // :81-85
    if (last_w <= first_w) return 0;                                                          // :106

    // state of the window currently being filled
    bool open = (o.tstart % W == 0) || (o.tstart < edge);  // :120-125
    uint32_t w_t = o.tstart, w_q = 0, w_ci = 0, w_co = 0;
    uint32_t t = o.tstart, q = 0;

    auto emit = [&](uint32_t window, uint32_t qend, uint32_t cei, uint32_t ceo) -> bool {
        if (window >= n_windows) return false;
        out.push_back(hb_overlap_window{overlap_idx, window, w_t, w_q, qend, w_ci, w_co, cei, ceo});
        return true;
    };

    for (size_t k = 0; k < ops.size(); k++) {
        const Op& op = ops[k];
        if (op.kind == 'I') { q += op.len; continue; }       // :132-135
        const bool consumes_q = op.kind == 'M';
        const uint32_t t_end = t + op.len;
        const uint32_t w_cur = t / W, w_new = t_end / W;
        // every boundary b in (t, t_end] closes a window
        for (uint32_t wb = w_cur + 1; wb <= w_new; wb++) {
            const uint32_t off = wb * W - t;                 // bases of this op before the boundary
            uint32_t qend = consumes_q ? q + off : q;
            uint32_t cei = op.b1, ceo = off, nci = op.b0, nco = off;
            if (wb == w_new && t_end == wb * W) {
                // op ends exactly on the boundary: a following insertion stays with this window
                // (:210-223); the next window starts on the op after it
                if (k + 1 < ops.size() && ops[k + 1].kind == 'I') {
                    qend += ops[k + 1].len;
                    cei = ops[k + 1].b1;
                    ceo = ops[k + 1].len;
                } else {
                    ceo = op.len;
                }
                nci = cei;
                nco = 0;
            }
            if (open) {
                if (!emit(wb - 1, qend, cei, ceo)) return -1;
            }
            open = true;
            w_t = wb * W;
            w_q = qend;
            w_ci = nci;
            w_co = nco;
        }
        t = t_end;
        if (consumes_q) q += op.len;
    }
    // trailing partial window at the end of the read (:260-272)
    if (t > tail_thresh && t % W != 0) {
        if (!open) return -1;  // the reference unwraps a None here
        if (!emit(last_w - 1, q, o.cigar_len, ops.back().len)) return -1;
    }
    return 0;
}

}  // namespace hb
