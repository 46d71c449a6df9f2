// windowing_dev.cu — windowing::extract_windows (src/windowing.rs:44-273) on the device, for alignments that the host
// submits raw (hb_submit_alignments): SURVEY.md §8f-1.
//
// The reference walks an alignment's CIGAR once and emits an OverlapWindow every time the walk crosses a multiple
// of the window size (state machine of SURVEY.md App. G).  WHICH windows an alignment contributes to depends only on
// its PAF coordinates (first / last window rules of src/windowing.rs:65-125, the 0.1·W edge rule, the trailing partial
// window :260-272), so the host lays out the overlap-window skeleton (overlap, window, per-window CSR) without touching
// a CIGAR byte (ctx.cu: skeleton_for_alignment).  WHERE each window starts and ends inside the CIGAR is found here:
//
//   k_parse_cigars   one warp per alignment: the CIGAR text -> raw ops (kind | len), with the inclusive prefix sums of
//                    target and query bases consumed.  Validates the text (the reference's CigarIter panics) and that
//                    it spans exactly the PAF coordinates (the reference would silently mis-window or index out of
//                    bounds later).
//   k_windows        one thread per overlap-window: binary search of the two window boundaries in the target prefix;
//                    boundary inside an op -> the op is shared (offsets), boundary at an op end -> a following
//                    insertion stays with the earlier window (src/windowing.rs:210-223).  Produces the OverlapWindow
//                    fields (tstart, qstart, qend, first / last op + offsets) and the window's op count.
//                    Op slots are handed out with a warp-aggregated atomic counter (any disjoint region will do).
//   k_tokenize<true> (features.cu) clips / prefix-sums / scores the ops exactly as for host-provided windows.
#include "common.cuh"
#include "forward.h"

namespace hb {

// ---- one warp per alignment: tokenise the whole CIGAR ---------------------------------------------------------------
__global__ void __launch_bounds__(128) k_parse_cigars(BatchView b) {
    const int lane = threadIdx.x & 31;
    const uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (a >= b.n_ovl) return;
    const DevOverlap ov = b.ovl[a];
    if (ov.raw_base == RAW_NONE) return;
    const uint8_t* __restrict__ cg = b.cig + ov.cig_off;
    const int slen = (int)ov.cig_len;
    uint32_t* __restrict__ okl = b.raw_kl + ov.raw_base;
    uint32_t* __restrict__ ot = b.raw_t + ov.raw_base;
    uint32_t* __restrict__ oq = b.raw_q + ov.raw_base;
    uint32_t flags = slen <= 0 ? OWF_BAD : 0u;
    uint32_t nops = 0, tc = 0, qc = 0;
    // lanes 0..9 look back, lanes 10..31 are the 22 active bytes of a step (same scheme as k_tokenize)
    for (int base = 0; base < slen; base += 22) {
        const int idx = base - 10 + lane;
        const bool inrange = idx >= 0 && idx < slen;
        const int c = inrange ? (int)__ldg(cg + idx) : 0;
        const bool active = lane >= 10 && inrange;
        const bool is_digit = inrange && c >= '0' && c <= '9';
        const bool is_letter = active && !is_digit;
        uint32_t num = 0, mul = 1;
        bool stop = false;
        int ndig = 0;
#pragma unroll
        for (int s = 1; s <= 10; s++) {
            const int pc = __shfl_up_sync(HB_FULL, c, s);
            const bool ok = (lane >= s) && pc >= '0' && pc <= '9';
            if (!stop && ok) { num += (uint32_t)(pc - '0') * mul; mul *= 10u; ndig++; } else { stop = true; }
            if (s >= 2 && !__any_sync(HB_FULL, is_letter && !stop)) break;
        }
        const uint32_t mask = __ballot_sync(HB_FULL, is_letter);
        uint32_t kind = 1u;
        if (is_letter) {
            kind = (c == 'M') ? OP_M : (c == 'I') ? OP_I : (c == 'D') ? OP_D : 1u;
            if (kind == 1u || num == 0 || ndig == 0 || ndig >= 10) flags |= OWF_BAD;
        }
        const uint32_t dt = (is_letter && kind != OP_I) ? num : 0u, dq = (is_letter && kind != OP_D) ? num : 0u;
        const uint32_t it = warp_incl_scan(dt, lane), iq = warp_incl_scan(dq, lane);
        if (is_letter) {
            const uint32_t k = nops + __popc(mask & ((1u << lane) - 1u));
            if (k < (uint32_t)slen / 2u + 1u) {  // the alignment's region holds cig_len / 2 + 1 ops: every valid op is >= 2 bytes
                okl[k] = (kind & 3u) | (num << 2);
                ot[k] = tc + it;
                oq[k] = qc + iq;
            } else {
                flags |= OWF_BAD;  // more letters than digits
            }
        }
        nops += __popc(mask);
        tc += __shfl_sync(HB_FULL, it, 31);
        qc += __shfl_sync(HB_FULL, iq, 31);
    }
    if (lane == 0 && slen > 0) {
        const int lc = __ldg(cg + slen - 1);
        if (lc >= '0' && lc <= '9') flags |= OWF_BAD;  // the text must end on an op letter
    }
    flags = __reduce_or_sync(HB_FULL, flags);
    if (nops == 0) flags |= OWF_BAD;
    // the CIGAR must span exactly the PAF coordinates the window skeleton was derived from
    if (tc != ov.tend - ov.tstart || qc != ov.qend - ov.qstart) flags |= OWF_BAD;
    if (lane == 0) { b.aln_nops[a] = nops; b.aln_flags[a] = flags; }
}

// ---- one thread per overlap-window: where the window starts and ends inside the alignment ------------------------------
struct Boundary {
    uint32_t k;        // the M / D op that reaches the boundary
    uint32_t off;      // bases of that op before the boundary (1 .. len)
    uint32_t len;
    uint32_t qend;     // query bases consumed up to the boundary (a following insertion included when the op ends on it)
    uint32_t ins_len;  // length of that insertion, 0 if none
    bool exact;        // the op ends exactly on the boundary
    bool ok;
};
__device__ __forceinline__ Boundary locate(const uint32_t* __restrict__ kl, const uint32_t* __restrict__ T,
                                           const uint32_t* __restrict__ Q, uint32_t n, uint32_t bnd) {
    Boundary r{};
    uint32_t lo = 0, hi = n;  // first op whose inclusive target prefix reaches bnd
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (T[mid] < bnd) lo = mid + 1; else hi = mid;
    }
    if (lo >= n) return r;
    const uint32_t w = kl[lo], kind = w & 3u, len = w >> 2;
    if (kind == OP_I) return r;  // cannot happen: the target prefix only grows at M / D ops
    const uint32_t tstart = T[lo] - len;
    r.k = lo; r.len = len; r.off = bnd - tstart;
    const uint32_t qstart = Q[lo] - (kind == OP_M ? len : 0u);
    r.qend = kind == OP_M ? qstart + r.off : qstart;
    r.exact = r.off == len;
    if (r.exact && lo + 1 < n && (kl[lo + 1] & 3u) == OP_I) {  // src/windowing.rs:210-223: the insertion stays with the earlier window
        r.ins_len = kl[lo + 1] >> 2;
        r.qend += r.ins_len;
    }
    r.ok = r.off >= 1 && r.off <= len;
    return r;
}

__global__ void __launch_bounds__(128) k_windows(BatchView b) {
    const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = wi < b.n_ow;
    DevOW ow{};
    DevOverlap ov{};
    if (valid) { ow = b.ow[wi]; ov = b.ovl[ow.ovl]; }
    const bool raw = valid && ov.raw_base != RAW_NONE;  // host-windowed overlap-windows: k_tokenize<false> counts their ops later
    uint32_t nops = 0;
    if (raw) {
        const DevWin win = b.win[ow.win];
        const uint32_t n = b.aln_nops[ow.ovl];
        uint32_t flags = (b.aln_flags[ow.ovl] & OWF_BAD) ? OWF_BAD : 0u;
        const uint32_t* __restrict__ kl = b.raw_kl + ov.raw_base;
        const uint32_t* __restrict__ T = b.raw_t + ov.raw_base;
        const uint32_t* __restrict__ Q = b.raw_q + ov.raw_base;
        const uint32_t ws = win.tstart, we = win.tstart + b.W;  // window [ws, we) in target coordinates (the last one may be shorter)
        uint32_t ks = 0, cso = 0, ke = 0, ceo = 0, w_t = ov.tstart, w_q = 0, qend = 0;
        if (!flags) {
            if (ws > ov.tstart) {  // the window starts on a boundary the walk crossed
                const Boundary s = locate(kl, T, Q, n, ws - ov.tstart);
                if (!s.ok) flags |= OWF_BAD;
                w_t = ws;
                w_q = s.qend;
                if (s.exact) { ks = s.k + (s.ins_len ? 2u : 1u); cso = 0; } else { ks = s.k; cso = s.off; }
            }
            if (we <= ov.tend) {   // ... and ends on the next one
                const Boundary e = locate(kl, T, Q, n, we - ov.tstart);
                if (!e.ok) flags |= OWF_BAD;
                qend = e.qend;
                if (e.exact) {
                    if (e.ins_len) { ke = e.k + 1u; ceo = e.ins_len; } else { ke = e.k; ceo = e.len; }
                } else { ke = e.k; ceo = e.off; }
            } else {               // trailing partial window (src/windowing.rs:260-272): up to the end of the CIGAR
                ke = n - 1u;
                ceo = kl[ke] >> 2;
                qend = Q[n - 1u];
            }
            if (ke < ks || ke >= n) flags |= OWF_BAD;
        }
        DevOW* o = b.ow_mut + wi;
        o->tstart = w_t; o->qstart = w_q; o->qend = qend;
        o->csi = ks; o->cso = cso; o->cei = ke; o->ceo = ceo;
        nops = flags ? 0u : ke - ks + 1u;
        b.ow_flags[wi] = flags;
    }
    if (valid) b.ow_nops[wi] = nops;
    // op slots of the window: any disjoint region will do, so a warp-aggregated atomic counter replaces an ordered scan (a one-block
    // scan over the 240 k overlap-windows of a 2 000-target launch took 0.33 ms).  Every lane of the warp takes part.
    const int lane = threadIdx.x & 31;
    const uint32_t inc = warp_incl_scan(nops, lane);
    uint32_t base = 0;
    if (lane == 31 && inc) base = atomicAdd(&b.counters[CNT_DEV_OPS], inc);
    base = __shfl_sync(HB_FULL, base, 31);
    if (raw) b.ow_opoff[wi] = (uint64_t)(base + inc - nops);
}

int launch_windowing(const BatchView& b, cudaStream_t st) {
    k_parse_cigars<<<(b.n_ovl * 32 + 127) / 128, 128, 0, st>>>(b);
    k_windows<<<(b.n_ow + 127) / 128, 128, 0, st>>>(b);
    return 2;
}

}  // namespace hb
