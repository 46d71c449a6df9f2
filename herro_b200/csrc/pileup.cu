// pileup.cu — the pileup-build kernel (second pass of extract_features + the row-wise work that
// follows it), rewritten bit-parallel.
//
// Reference semantics: get_features_for_ol_window / write_target_for_window
// (src/features.rs:110-266) restricted to the <= 30 kept columns, the second
// get_supported (src/features.rs:681-722 on [L',31], thresh = floor(3.1) = 3), SupportedPos
// (pos, ins) and the majority vote of consensus (src/consensus.rs:176-200).
//
// Formulation.  A column (one query read in one window) walks its CIGAR and, row by row, either
// consumes the next base of its (strand-oriented) query slice or leaves a gap / '.' cell.  So the
// whole column is described by ONE BIT PER ROW ("this row consumes a query base"): the base and
// quality of row r are those at query offset popcount(consume bits before r).  The bitmap is
// built from the ops without walking positions:
//     M op over target positions [t0, t0+n)  ->  base rows inside [row(t0), row(t0+n-1)]
//     I op of n bases after position p       ->  rows row(p)+1 .. row(p)+n
// Each range contributes two toggle bits (start, end); a prefix-XOR over the toggle bitmap turns
// them into "inside a range" masks (M ranges are additionally ANDed with the window's base-row
// mask).  One CTA per window; a thread then owns FOUR consecutive rows: per column it reads the
// 4 consume bits, fetches the <= 4 consumed bases (one funnel-shifted word pair of the 2-bit
// store) and qualities (one unaligned word), expands them with PRMT through a 16-entry selector
// table, and after every 4 columns transposes the 4x4 byte block so that a row's 32 tokens /
// qualities end up in 8+8 registers that are stored row-major with 16-byte stores.  Class counts
// for get_supported / the majority vote are five 6-bit fields of one word per row.
// (Measured alternative, reverted: a group's columns split over 4 lanes — 64 registers, 4 CTAs per
// SM — was 20 % slower: the shared-memory pipe, not occupancy, limits the compose loop.)
//
// The former kernel (features.cu: k_pass2b) walked every target position of every column with
// scalar byte stores into a shared-memory tile and was instruction bound (1.28 G
// warp instructions per launch, 0.059 of the HBM roofline); it stays selectable with
// HERRO_B200_PILEUP_V1=1 for the A-B parity test.
#include <cstdlib>

#include "common.cuh"
#include "forward.h"

namespace hb {

constexpr int P_CH = 5120;          // rows per chunk (a window with more rows is processed in several chunks)
constexpr int P_CW = P_CH / 32;     // words per bitmap
constexpr int P_WPL = P_CW / 32;    // bitmap words per lane in the warp-wide scans
static_assert(P_CW % 32 == 0, "bitmap words must split evenly over a warp");

struct __align__(16) ColA { const uint32_t* w32; const uint8_t* qp; };
struct __align__(16) ColB { int32_t x0, sgn; uint32_t rs, re; };
struct __align__(16) ColC { uint32_t gap4, xr4, add4, rsel; };

// selector table for PRMT: nibble k = index of the consumed base that row k of the group takes (b_k set),
// or 4 + k (byte k of the fill word) when the row does not consume
__device__ __constant__ uint16_t c_sel16[16] = {
    0x7654, 0x7650, 0x7604, 0x7610, 0x7054, 0x7150, 0x7104, 0x7210,
    0x0654, 0x1650, 0x1604, 0x2610, 0x1054, 0x2150, 0x2104, 0x3210};
// class-count increments: five 6-bit counters (A C G T gap; a count is <= 31, bit 5 of a field is headroom for the threshold
// test) in one word; tokens 10 ('.') and 11 (pad) count nothing.  The table is indexed by a PAIR of tokens (t0 | t1 << 4).
__host__ __device__ constexpr uint32_t cls_inc(uint32_t t) { return t < 10u ? 1u << (6u * (t < 5u ? t : t - 5u)) : 0u; }

// toggle the two ends of the row range [a, e) clipped to the chunk [c0, c0 + P_CH)
__device__ __forceinline__ void toggle_range(uint32_t* T, uint32_t a, uint32_t e, uint32_t c0) {
    const uint32_t c1 = c0 + P_CH;
    if (a >= c1 || e <= c0 || e <= a) return;
    const uint32_t s = (a > c0 ? a : c0) - c0;
    atomicXor(&T[s >> 5], 1u << (s & 31u));
    if (e < c1) {
        const uint32_t t = e - c0;
        atomicXor(&T[t >> 5], 1u << (t & 31u));
    }
}

__device__ __forceinline__ uint32_t prefix_xor32(uint32_t x) {  // bit i = parity of bits 0..i
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
    return x;
}

__global__ void __launch_bounds__(256, 2) k_pileup(BatchView b) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* TM = (uint32_t*)smem_raw;                 // [32][P_CW] toggles -> "inside an M range" -> consume bitmap
    uint32_t* TI = TM + 32 * P_CW;                      // [32][P_CW] toggles -> "inside an insertion"
    uint32_t* ins = TI + 32 * P_CW;                     // [P_CW] toggles -> rows that are insertion slots (not base rows)
    uint16_t* pref = (uint16_t*)(ins + P_CW);           // [32][P_CW] consumed bases of the column before each word (chunk-local)
    uint16_t* bpref = pref + 32 * P_CW;                 // [P_CW] base rows before each word (chunk-local)
    ColA* colA = (ColA*)(bpref + P_CW);                 // [32]
    ColB* colB = (ColB*)(colA + 32);
    ColC* colC = (ColC*)(colB + 32);
    uint32_t* sel_s = (uint32_t*)(colC + 32);           // [16]
    uint32_t* cls_s = sel_s + 16;                       // [256]
    __shared__ uint32_t s_x0[32], s_carry[32], s_tot[33], s_warp[8];
    __shared__ uint32_t s_opb[32], s_opoff[33];  // op arrays of the columns: first slot, and the exclusive prefix of their op counts
    __shared__ const uint8_t* s_qp0[32];
    __shared__ uint32_t s_nsup, s_pcarry;

    const uint32_t w = blockIdx.x;
    const DevWin win = b.win[w];
    const uint32_t W = b.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nsel = b.w_nsel[w];
    const uint32_t L = b.w_L[w];
    const uint64_t rowbase = b.w_rowbase[w];
    if (b.counters[CNT_OVERFLOW]) return;  // arena too small: host grows it and re-launches
    const uint32_t* __restrict__ rm = b.rowmap + (size_t)w * (W + 1);

    // ---- per-column constants (column 0 = the target read, forward, every base row consumes)
    if (tid < 32) {
        const uint32_t c = tid;
        ColA a; ColB bb; ColC cc;
        uint32_t owi = 0, x0 = 0, nops_c = 0, opb_c = 0;
        const uint8_t* qp0;
        const uint64_t* words;
        bb.sgn = 0; bb.rs = 0; bb.re = 0; bb.x0 = 0;
        cc.gap4 = TOK_NONE * 0x01010101u; cc.xr4 = 0; cc.add4 = 0; cc.rsel = 0x3210u;
        words = b.rs.words + b.rs.word_off[win.rid];
        qp0 = b.rs.qual + b.rs.qual_off[win.rid];
        if (c == 0) {
            bb.sgn = 1; bb.rs = 0; bb.re = L;
            x0 = win.tstart;
            qp0 += win.tstart;
            cc.gap4 = TOK_GAP_F * 0x01010101u;  // bases.fill('*') (src/features.rs:248)
        } else if (c <= nsel) {
            owi = b.sel_ow[w * TOP_K + c - 1];
            const DevOW ow = b.ow[owi];
            const DevOverlap ov = b.ovl[ow.ovl];
            nops_c = b.ow_nops[owi];
            opb_c = ow.op_base;
            words = b.rs.words + b.rs.word_off[ov.qid];
            qp0 = b.rs.qual + b.rs.qual_off[ov.qid];
            bb.rs = rm[ow.tstart - win.tstart];  // rows before are '.' (src/features.rs:166-171)
            bb.re = rm[b.ow_tend[owi]];          // rows from here on are '.' (:233-236)
            if (!ov.strand) {
                bb.sgn = 1;
                x0 = ov.qstart + ow.qstart;      // oriented offset 0 <-> this base (src/features.rs:97-108)
                qp0 += x0;
                cc.gap4 = TOK_GAP_F * 0x01010101u;
            } else {
                bb.sgn = -1;
                x0 = ov.qend - ow.qstart - 1u;   // reverse strand: oriented offset x <-> base x0 - x, complemented
                qp0 += x0;
                cc.gap4 = TOK_GAP_R * 0x01010101u;
                cc.xr4 = 0x03030303u; cc.add4 = 0x05050505u; cc.rsel = 0x0123u;
            }
        }
        a.w32 = (const uint32_t*)words;
        a.qp = qp0;
        colA[c] = a; colB[c] = bb; colC[c] = cc;
        s_x0[c] = x0; s_qp0[c] = qp0; s_carry[c] = 0;
        s_opb[c] = opb_c;
        const uint32_t inc = warp_incl_scan(nops_c, lane);
        s_opoff[c] = inc - nops_c;
        if (c == 31) s_opoff[32] = inc;
    }
    if (tid < 16) sel_s[tid] = c_sel16[tid];
    cls_s[tid] = cls_inc((uint32_t)tid & 15u) + cls_inc((uint32_t)tid >> 4);
    if (tid == 0) { s_nsup = 0; s_pcarry = 0; }
    __syncthreads();

    for (uint32_t c0 = 0; c0 < L; c0 += P_CH) {
        const uint32_t nrows = min((uint32_t)P_CH, L - c0);
        const uint32_t cw = (nrows + 31) >> 5;
        // ---- A: clear the toggle bitmaps; per-chunk column offsets (x0 / quality pointer advanced by the bases consumed so far)
        for (uint32_t i = tid; i < 32u * cw; i += 256) {
            const uint32_t c = i / cw, j = i - c * cw;
            TM[c * P_CW + j] = 0; TI[c * P_CW + j] = 0;
        }
        for (uint32_t j = tid; j < cw; j += 256) ins[j] = 0;
        if (tid < 32) {
            const int32_t sg = colB[tid].sgn, car = (int32_t)s_carry[tid];
            // the 4 bases a group may consume are fetched as the 4 consecutive store positions starting at xs:
            // forward xs = x, reverse xs = x - 3 (consumption order is then the byte-reversed word)
            colB[tid].x0 = (int32_t)s_x0[tid] + sg * car - (sg < 0 ? 3 : 0);
            colA[tid].qp = s_qp0[tid] + (ptrdiff_t)(sg * car) - (sg < 0 ? 3 : 0);
        }
        __syncthreads();
        // ---- B: toggles.  insertion-slot rows of the window; column 0; the ops of the selected columns
        for (uint32_t p = tid; p < win.len; p += 256) {
            const uint32_t a = rm[p] + 1u, e = rm[p + 1];
            if (e > a) toggle_range(ins, a, e, c0);
        }
        if (tid == 0) toggle_range(TM, 0, L, c0);
        // all ops of all columns, flattened over the CTA; 4 ops per thread and pass so that the dependent loads
        // (op words, then row(t) of the op's ends) of several ops are in flight together
        {
            const uint32_t total = s_opoff[32];
            for (uint32_t i0 = tid; i0 < total; i0 += 4 * 256) {
                uint32_t col[4], kl[4], t0[4], ra[4], re[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + u * 256;
                    col[u] = 0xffffffffu;
                    if (i < total) {
                        uint32_t lo_c = 1, hi_c = nsel;  // largest column c with s_opoff[c] <= i
#pragma unroll
                        for (int it = 0; it < 5; it++) {
                            const uint32_t mid = (lo_c + hi_c + 1) >> 1;
                            if (s_opoff[mid] <= i) lo_c = mid; else hi_c = mid - 1;
                        }
                        col[u] = lo_c;
                        const uint32_t slot = s_opb[lo_c] + (i - s_opoff[lo_c]);
                        kl[u] = b.op_kl[slot]; t0[u] = b.op_t[slot];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    ra[u] = re[u] = 0;
                    if (col[u] != 0xffffffffu) {
                        const uint32_t kind = kl[u] & 3u, eff = kl[u] >> 2;
                        if (kind == OP_M) { ra[u] = rm[t0[u]]; re[u] = rm[t0[u] + eff - 1u] + 1u; }
                        else if (kind == OP_I) { ra[u] = rm[t0[u] - 1u] + 1u; re[u] = ra[u] + eff; }  // never the first op: t0 >= 1
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (col[u] != 0xffffffffu && re[u] > ra[u])
                        toggle_range(((kl[u] & 3u) == OP_M ? TM : TI) + col[u] * P_CW, ra[u], re[u], c0);
            }
        }
        __syncthreads();
        // ---- D1: prefix-XOR every toggle bitmap in place (65 bitmaps: ins, TM[32], TI[32]); a warp per bitmap
        for (uint32_t item = warp; item < 65; item += 8) {
            uint32_t* B = item == 0 ? ins : (item <= 32 ? TM + (item - 1) * P_CW : TI + (item - 33) * P_CW);
            if (item >= 1 && ((item - 1) & 31u) > nsel) continue;  // padding columns hold no toggles
            uint32_t x[P_WPL];
            uint32_t par = 0;
#pragma unroll
            for (int i = 0; i < P_WPL; i++) {
                const uint32_t j = lane * P_WPL + i;
                uint32_t v = j < cw ? B[j] : 0u;
                v = prefix_xor32(v);
                x[i] = par ? ~v : v;
                par ^= v >> 31;
            }
            const uint32_t bal = __ballot_sync(HB_FULL, par != 0);
            const uint32_t cin = __popc(bal & ((1u << lane) - 1u)) & 1u;
#pragma unroll
            for (int i = 0; i < P_WPL; i++) {
                const uint32_t j = lane * P_WPL + i;
                if (j < cw) B[j] = cin ? ~x[i] : x[i];
            }
        }
        __syncthreads();
        // ---- D2: consume bitmap = (inside-M & base rows) | inside-I, and its popcount prefix; item 32 = base rows
        for (uint32_t item = warp; item < 33; item += 8) {
            uint32_t cnt[P_WPL], val[P_WPL];
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < P_WPL; i++) {
                const uint32_t j = lane * P_WPL + i;
                uint32_t v = 0;
                if (j < cw) {
                    const uint32_t base = ~ins[j];
                    if (item < 32) v = (item <= nsel) ? ((TM[item * P_CW + j] & base) | TI[item * P_CW + j]) : 0u;
                    else v = (j + 1 < cw || (nrows & 31u) == 0) ? base : (base & ((1u << (nrows & 31u)) - 1u));
                }
                val[i] = v; cnt[i] = sum; sum += __popc(v);
            }
            const uint32_t inc = warp_incl_scan(sum, lane);
            const uint32_t off = inc - sum;
#pragma unroll
            for (int i = 0; i < P_WPL; i++) {
                const uint32_t j = lane * P_WPL + i;
                if (j < cw) {
                    if (item < 32) { TM[item * P_CW + j] = val[i]; pref[item * P_CW + j] = (uint16_t)(off + cnt[i]); }
                    else bpref[j] = (uint16_t)(off + cnt[i]);
                }
            }
            if (lane == 31) s_tot[item] = inc;
        }
        __syncthreads();
        // ---- E: compose rows.  A thread owns 4 consecutive rows (one nibble of every bitmap word).
        const uint32_t ngroups = (nrows + 3) >> 2;
        for (uint32_t g0 = 0; g0 < ngroups; g0 += 256) {
            const uint32_t g = g0 + tid;
            const bool act = g < ngroups;
            uint32_t tw[4][8], qw[4][8], acc[4] = {0, 0, 0, 0};
            const uint32_t r0 = act ? 4u * g : 0u;
            const uint32_t wi = r0 >> 5, bsh = r0 & 31u, lowmask = (1u << bsh) - 1u;
            const uint32_t row0 = c0 + r0;
            if (act) {
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {
                    uint32_t tcol[4], qcol[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int c = c4 * 4 + e;
                        if (c == 31) { tcol[e] = TOK_NONE * 0x01010101u; qcol[e] = QUAL_EMPTY * 0x01010101u; continue; }  // pad byte of the 32-byte row
                        const ColA ca = colA[c];
                        const ColB cb = colB[c];
                        const ColC cc = colC[c];
                        const uint32_t word = TM[c * P_CW + wi];
                        const uint32_t nib = (word >> bsh) & 15u;
                        const int32_t qi0 = (int32_t)((uint32_t)pref[c * P_CW + wi] + __popc(word & lowmask));
                        const int32_t d = cb.sgn * qi0;
                        // bases: 4 consecutive 2-bit codes of the packed store starting at position xs
                        const int32_t xs = cb.x0 + d;
                        const uint32_t* wp = ca.w32 + (xs >> 4);
                        uint32_t v = __funnelshift_r(__ldg(wp), __ldg(wp + 1), (uint32_t)(xs & 15) << 1) & 0xffu;
                        v = (v | (v << 12)) & 0x000f000fu;
                        v = (v | (v << 6)) & 0x03030303u;
                        v = __byte_perm((v ^ cc.xr4) + cc.add4, 0u, cc.rsel);
                        // qualities: the 4 bytes at the same positions (unaligned word = two aligned loads + funnel shift)
                        const uint8_t* qa = ca.qp + d;
                        const uint32_t* qp4 = (const uint32_t*)((uintptr_t)qa & ~(uintptr_t)3);
                        uint32_t qv = __funnelshift_r(__ldg(qp4), __ldg(qp4 + 1), (uint32_t)((uintptr_t)qa & 3u) << 3);
                        qv = __byte_perm(qv, 0u, cc.rsel);
                        // fill: gap inside the column's aligned row range [rs, re), '.' outside
                        const int32_t lo = max(0, min(4, (int32_t)(cb.rs - row0))), hi = max(0, min(4, (int32_t)(cb.re - row0)));
                        const uint32_t m = __funnelshift_lc(0u, 0xffffffffu, (uint32_t)lo << 3) & ~__funnelshift_lc(0u, 0xffffffffu, (uint32_t)hi << 3);
                        const uint32_t fill = (cc.gap4 & m) | ((TOK_NONE * 0x01010101u) & ~m);
                        const uint32_t sel = sel_s[nib];
                        tcol[e] = __byte_perm(v, fill, sel);
                        qcol[e] = __byte_perm(qv, QUAL_EMPTY * 0x01010101u, sel);
                    }
                    // 4x4 byte transpose: tcol[e] byte k = (row k, column 4*c4+e)  ->  tw[k][c4] byte e
                    {
                        const uint32_t t0 = __byte_perm(tcol[0], tcol[1], 0x5140u), t1 = __byte_perm(tcol[2], tcol[3], 0x5140u);
                        const uint32_t t2 = __byte_perm(tcol[0], tcol[1], 0x7362u), t3 = __byte_perm(tcol[2], tcol[3], 0x7362u);
                        tw[0][c4] = __byte_perm(t0, t1, 0x5410u); tw[1][c4] = __byte_perm(t0, t1, 0x7632u);
                        tw[2][c4] = __byte_perm(t2, t3, 0x5410u); tw[3][c4] = __byte_perm(t2, t3, 0x7632u);
                        const uint32_t q0 = __byte_perm(qcol[0], qcol[1], 0x5140u), q1 = __byte_perm(qcol[2], qcol[3], 0x5140u);
                        const uint32_t q2 = __byte_perm(qcol[0], qcol[1], 0x7362u), q3 = __byte_perm(qcol[2], qcol[3], 0x7362u);
                        qw[0][c4] = __byte_perm(q0, q1, 0x5410u); qw[1][c4] = __byte_perm(q0, q1, 0x7632u);
                        qw[2][c4] = __byte_perm(q2, q3, 0x5410u); qw[3][c4] = __byte_perm(q2, q3, 0x7632u);
                    }
                    // class counts of the row's 4 new tokens: two pair look-ups (tokens are < 16, so a pair packs into a byte)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t pr = (tw[k][c4] | (tw[k][c4] >> 4)) & 0x00ff00ffu;
                        acc[k] += cls_s[pr & 0xffu] + cls_s[pr >> 16];
                    }
                }
            }
            // ---- per-row work: second get_supported (thresh 3), majority vote, row-major stores
            uint32_t supm = 0;  // bit k: row k of this group is supported
            if (act) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t row = row0 + k;
                    if (row >= L) break;
                    // second get_supported: >= 2 classes with >= 3 reads (adding 29 to a 6-bit field sets its bit 5 iff count >= 3)
                    const uint32_t a6 = acc[k];
                    const bool sup = __popc((a6 + 29u * 0x01041041u) & 0x20820820u) >= 2;
                    // two most common classes, stable on ties (A<C<G<T<*): keys (count << 3 | 7 - class), top two by a max/min chain
                    uint32_t ka = ((a6 & 63u) << 3) | 7u, kb = 0u;
#pragma unroll
                    for (int q = 1; q < 5; q++) {
                        const uint32_t kq = (((a6 >> (6 * q)) & 63u) << 3) | (uint32_t)(7 - q);
                        kb = max(kb, min(ka, kq));
                        ka = max(ka, kq);
                    }
                    const uint32_t b0 = 7u - (ka & 7u), b1 = 7u - (kb & 7u), m0 = ka >> 3, m1 = kb >> 3;
                    const uint32_t tb = tw[k][0] & 0xffu;  // target column, token 0..4
                    const uint32_t base = (m0 < 2u || (m0 == m1 && (b0 == tb || b1 == tb))) ? tb : b0;
                    const uint32_t emit = nsel >= 2 ? base : 4u;  // n_alns < 2: window dropped (src/consensus.rs:104-111)
                    b.row_emit[rowbase + row] = (uint8_t)(emit | (sup ? 0x80u : 0u));
                    supm |= (sup ? 1u : 0u) << k;
                    uint4* gb = (uint4*)(b.mat_bases + (rowbase + row) * ROW_BYTES);
                    uint4* gq = (uint4*)(b.mat_quals + (rowbase + row) * ROW_BYTES);
                    gb[0] = make_uint4(tw[k][0], tw[k][1], tw[k][2], tw[k][3]); gb[1] = make_uint4(tw[k][4], tw[k][5], tw[k][6], tw[k][7]);
                    gq[0] = make_uint4(qw[k][0], qw[k][1], qw[k][2], qw[k][3]); gq[1] = make_uint4(qw[k][4], qw[k][5], qw[k][6], qw[k][7]);
                }
            }
            // ---- ordered list of supported rows: (row, pos << 8 | ins) appended in row order.  The three barriers per 1024 rows also
            //      keep the CTA's warps on the same stretch of every column's query data: a barrier-free variant (supported rows
            //      flagged in a bitmap, list built once per chunk) measured 20 % SLOWER (0.80 vs 0.67 ms per 2 079 windows).
            {
                const uint32_t nf = __popc(supm);
                const uint32_t inc = warp_incl_scan(nf, lane);
                if (lane == 31) s_warp[warp] = inc;
                __syncthreads();
                uint32_t off = s_nsup + inc - nf;
                for (int k = 0; k < warp; k++) off += s_warp[k];
                if (nf) {
                    const uint32_t insw = ins[wi];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (!((supm >> k) & 1u)) continue;
                        const uint32_t rl = r0 + k;  // chunk-local row; base rows at or before it give the target position
                        const uint32_t basew = ~insw;
                        const uint32_t upto = (rl & 31u) == 31u ? basew : (basew & ((2u << (rl & 31u)) - 1u));
                        const uint32_t p = s_pcarry + (uint32_t)bpref[wi] + __popc(upto) - 1u;
                        const uint32_t kk = (c0 + rl) - rm[p];
                        b.sup_row[rowbase + off] = c0 + rl;
                        b.sup_pk[rowbase + off] = (p << 8) | (kk & 0xffu);  // SupportedPos.ins is a u8 (H13)
                        off++;
                    }
                }
                __syncthreads();
                if (tid == 0) { uint32_t t = 0; for (int k = 0; k < 8; k++) t += s_warp[k]; s_nsup += t; }
                __syncthreads();
            }
        }
        // ---- carry the per-column consumed-base counts and the base-row count into the next chunk
        if (tid < 32) s_carry[tid] += s_tot[tid];
        if (tid == 32) s_pcarry += s_tot[32];
        __syncthreads();
    }
    if (tid == 0) b.w_nsup[w] = s_nsup;
}

size_t pileup_smem() {
    return (size_t)2 * 32 * P_CW * 4 + (size_t)P_CW * 4 + (size_t)32 * P_CW * 2 + (size_t)P_CW * 2 +
           32 * (sizeof(ColA) + sizeof(ColB) + sizeof(ColC)) +
           (16 + 256) * 4 + 64;
}

cudaError_t pileup_configure() {
    return cudaFuncSetAttribute(k_pileup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pileup_smem());
}

void launch_pileup_v2(const BatchView& b, cudaStream_t st) { k_pileup<<<b.n_win, 256, pileup_smem(), st>>>(b); }

}  // namespace hb
