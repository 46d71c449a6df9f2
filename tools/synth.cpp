// synth.cpp — deterministic synthetic ONT-like read set with truth-derived read-vs-read
// alignments (PAF fields + M/I/D CIGAR), standing in for `minimap2 -cx ava-ont` output
// (src/mm2.rs:15-37) because neither demo data nor minimap2 nor a network exist offline
// (SURVEY.md §7.1, §8d "Synthetic generator spec").
//
// Model: random diploid genome (het SNPs / het 1-bp indels on one haplotype), reads sampled
// uniformly with random strand and haplotype, per-base sub/ins/del errors with a
// homopolymer boost, occasional long (>50 bp) deletions so the indel filter
// (src/features.rs:315-324) is exercised.  For every pair of reads sharing a genome
// interval the two read->genome edit scripts are composed into a target-vs-query CIGAR
// (PAF convention: target forward, query reverse-complemented for '-'; both directions
// emitted like `--dual=yes`).
//
// This is bench/test input generation, not part of the hot path.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64 / xoshiro-free: small, deterministic, seedable per item
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double normal() {
        double u1 = uni() + 1e-300, u2 = uni();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char BASES[4] = {'A', 'C', 'G', 'T'};
inline int code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }
inline char comp(char c) { return BASES[3 - code(c)]; }

struct Params {
    uint64_t seed;
    uint64_t genome_len;
    uint32_t n_reads;
    uint32_t mean_len, sd_len, min_len, max_len;
    double sub, ins, del;          // sequencing error rates
    double hp_boost;               // indel multiplier per extra homopolymer base (capped)
    double het_snp, het_indel;     // per-base haplotype differences
    double long_del;               // per-base probability of a 51..80 bp deletion in a read
    uint32_t min_ovl;              // minimum shared genome interval to emit an alignment
};

struct Read {
    uint64_t gs, ge;               // genome interval [gs, ge)
    uint8_t strand, hap;
    std::string seq, qual;         // as stored (reverse-complemented when strand==1)
    std::vector<uint8_t> ev;       // per genome pos: bit0 present, bits1.. inserted-after count
    std::vector<uint32_t> ckpt;    // read-forward(genome orientation) offset at every 32nd genome pos
    uint32_t len() const { return (uint32_t)seq.size(); }
    // number of read bases (genome orientation) before genome position g (gs <= g <= ge)
    uint32_t offset_at(uint64_t g) const {
        uint64_t k = g - gs;
        uint32_t o = ckpt[k >> 5];
        for (uint64_t j = (k >> 5) << 5; j < k; j++) o += (ev[j] & 1) + (ev[j] >> 1);
        return o;
    }
};

struct Set {
    Params p;
    std::vector<Read> reads;
    std::vector<std::string> ids;
    // alignments grouped by target
    std::vector<uint64_t> aln_off;        // [n_reads+1]
    std::vector<uint32_t> ovl9;           // [n_aln*9]
    std::vector<uint64_t> cig_off;        // [n_aln+1]
    std::string cigars;
};

void append_op(std::string& s, char op, uint32_t n) {
    if (!n) return;
    char buf[16];
    int k = 0;
    while (n) { buf[k++] = (char)('0' + n % 10); n /= 10; }
    while (k) s.push_back(buf[--k]);
    s.push_back(op);
}

struct OpList {
    std::vector<std::pair<char, uint32_t>> ops;
    void add(char op, uint32_t n) {
        if (!n) return;
        if (!ops.empty() && ops.back().first == op) ops.back().second += n;
        else ops.emplace_back(op, n);
    }
};

}  // namespace

extern "C" {

struct synth_set;  // opaque = Set
void synth_make_alns(void* h, uint32_t tgt_begin, uint32_t tgt_end, uint32_t tgt_stride, uint32_t tgt_phase, uint32_t n_threads);

void* synth_generate(uint64_t seed, uint64_t genome_len, uint32_t n_reads, uint32_t mean_len, uint32_t sd_len,
                     uint32_t min_len, uint32_t max_len, double sub, double ins, double del, double hp_boost,
                     double het_snp, double het_indel, double long_del, uint32_t min_ovl, uint32_t n_threads,
                     uint32_t tgt_begin, uint32_t tgt_end, uint32_t tgt_stride, uint32_t tgt_phase) {
    // alignments are produced only for targets t in [tgt_begin, tgt_end) with t % tgt_stride == tgt_phase (every read is still
    // generated: queries come from the whole set) — a rank of a sharded run builds just its own targets' alignments
    Set* S = new Set();
    S->p = Params{seed, genome_len, n_reads, mean_len, sd_len, min_len, max_len, sub, ins, del,
                  hp_boost, het_snp, het_indel, long_del, min_ovl};
    const uint64_t G = genome_len;
    if (max_len > G) max_len = (uint32_t)G;
    if (min_len > max_len) min_len = max_len;

    // ---- genome + haplotype events -------------------------------------------------
    std::string genome(G, 'A');
    std::vector<uint8_t> hp(G, 1);  // homopolymer run length (capped) containing g
    {
        Rng r(seed ^ 0x67656e6full);
        for (uint64_t g = 0; g < G; g++) {
            // mild homopolymer enrichment: 30% chance to repeat the previous base
            if (g > 0 && r.uni() < 0.30) genome[g] = genome[g - 1];
            else genome[g] = BASES[r.below(4)];
        }
        uint64_t g = 0;
        while (g < G) {
            uint64_t e = g;
            while (e < G && genome[e] == genome[g]) e++;
            uint8_t l = (uint8_t)std::min<uint64_t>(e - g, 8);
            for (uint64_t k = g; k < e; k++) hp[k] = l;
            g = e;
        }
    }
    // hap event per position: 0 none, 1..3 sub to (base+e)%4, 4 del, 5 ins-after (base = (g*7)%4)
    std::vector<uint8_t> hev[2] = {std::vector<uint8_t>(G, 0), std::vector<uint8_t>(G, 0)};
    {
        Rng r(seed ^ 0x686170ull);
        for (uint64_t g = 0; g < G; g++) {
            double u = r.uni();
            int h = (int)r.below(2);
            if (u < het_snp) hev[h][g] = (uint8_t)(1 + r.below(3));
            else if (u < het_snp + het_indel * 0.5) hev[h][g] = 4;
            else if (u < het_snp + het_indel) hev[h][g] = 5;
        }
    }

    // ---- reads -------------------------------------------------------------------------
    S->reads.resize(n_reads);
    S->ids.resize(n_reads);
    {
        std::atomic<uint32_t> next{0};
        auto work = [&]() {
            for (;;) {
                uint32_t i = next.fetch_add(1);
                if (i >= n_reads) break;
                Rng r(seed * 1000003ull + i * 2ull + 1);
                Read& R = S->reads[i];
                double l = (double)mean_len + (double)sd_len * r.normal();
                uint64_t glen = (uint64_t)std::min<double>(std::max<double>(l, min_len), max_len);
                R.gs = (G > glen) ? (uint64_t)(r.uni() * (double)(G - glen)) : 0;
                R.ge = R.gs + glen;
                R.strand = (uint8_t)r.below(2);
                R.hap = (uint8_t)r.below(2);
                R.ev.assign(glen, 0);
                std::string seq, qual;
                seq.reserve(glen + glen / 50);
                qual.reserve(glen + glen / 50);
                auto put = [&](char b, bool err) {
                    seq.push_back(b);
                    int q = err ? 3 + (int)r.below(12) : 12 + (int)r.below(38);
                    qual.push_back((char)(33 + q));
                };
                uint64_t skip_until = 0;
                for (uint64_t g = R.gs; g < R.ge; g++) {
                    uint64_t k = g - R.gs;
                    uint8_t he = hev[R.hap][g];
                    char b = genome[g];
                    bool present = true, herr = false;
                    if (he >= 1 && he <= 3) b = BASES[(code(b) + he) & 3];
                    else if (he == 4) present = false;
                    double boost = 1.0 + hp_boost * (double)(hp[g] - 1);
                    if (g < skip_until) present = false;
                    else if (present) {
                        if (long_del > 0 && r.uni() < long_del && g + 100 < R.ge && k > 100) {
                            skip_until = g + 51 + r.below(30);
                            present = false;
                        } else if (r.uni() < del * boost) present = false;
                        else if (r.uni() < sub) {
                            b = BASES[(code(b) + 1 + r.below(3)) & 3];
                            herr = true;
                        }
                    }
                    uint32_t nins = 0;
                    if (present) {
                        put(b, herr);
                        R.ev[k] |= 1;
                    }
                    if (g >= skip_until) {
                        if (he == 5) {
                            put(BASES[(g * 7) & 3], false);
                            nins++;
                        }
                        if (r.uni() < ins * boost) {
                            uint32_t n = 1 + (r.uni() < 0.2 ? 1 + r.below(3) : 0);
                            for (uint32_t t = 0; t < n; t++) {
                                // homopolymer extension 60%, random base 40%
                                put(r.uni() < 0.6 ? genome[g] : BASES[r.below(4)], true);
                                nins++;
                            }
                        }
                    }
                    R.ev[k] |= (uint8_t)(std::min<uint32_t>(nins, 100) << 1);
                }
                // checkpoints
                R.ckpt.assign(glen / 32 + 1, 0);
                uint32_t o = 0;
                for (uint64_t k = 0; k < glen; k++) {
                    if ((k & 31) == 0) R.ckpt[k >> 5] = o;
                    o += (R.ev[k] & 1) + (R.ev[k] >> 1);
                }
                if ((glen & 31) == 0) R.ckpt[glen >> 5] = o;
                if (R.strand) {
                    std::reverse(seq.begin(), seq.end());
                    for (auto& c : seq) c = comp(c);
                    std::reverse(qual.begin(), qual.end());
                }
                R.seq.swap(seq);
                R.qual.swap(qual);
                char nm[32];
                snprintf(nm, sizeof nm, "read_%06u", i);
                S->ids[i] = nm;
            }
        };
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < std::max(1u, n_threads); t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }

    S->p.min_len = min_len; S->p.max_len = max_len;
    if (tgt_end > tgt_begin) synth_make_alns(S, tgt_begin, tgt_end, tgt_stride, tgt_phase, n_threads);
    else { S->aln_off.assign(n_reads + 1, 0); S->cig_off.assign(1, 0); }
    return S;
}

// Alignments (grouped by target) for targets t in [tgt_begin, tgt_end) with t % tgt_stride == tgt_phase, replacing whatever
// alignments the set held.  Deterministic: the same target always gets the same alignments, whichever subset is asked for.
void synth_make_alns(void* h, uint32_t tgt_begin, uint32_t tgt_end, uint32_t tgt_stride, uint32_t tgt_phase, uint32_t n_threads) {
    Set* S = (Set*)h;
    const uint32_t n_reads = (uint32_t)S->reads.size();
    const uint32_t min_ovl = S->p.min_ovl;
    S->ovl9.clear(); S->cigars.clear(); S->cig_off.clear();
    // ---- overlaps ----------------------------------------------------------------------
    std::vector<uint32_t> order(n_reads);
    for (uint32_t i = 0; i < n_reads; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return S->reads[a].gs != S->reads[b].gs ? S->reads[a].gs < S->reads[b].gs : a < b;
    });
    std::vector<uint32_t> rank(n_reads);
    for (uint32_t k = 0; k < n_reads; k++) rank[order[k]] = k;
    uint64_t max_glen = 0;
    for (auto& R : S->reads) max_glen = std::max(max_glen, R.ge - R.gs);

    struct PerTarget {
        std::vector<uint32_t> ovl9;
        std::vector<uint64_t> coff;
        std::string cig;
    };
    std::vector<PerTarget> per(n_reads);
    {
        std::atomic<uint32_t> next{0};
        auto work = [&]() {
            for (;;) {
                uint32_t t = next.fetch_add(1);
                if (t >= n_reads) break;
                const Read& T = S->reads[t];
                PerTarget& P = per[t];
                P.coff.push_back(0);
                if (t < tgt_begin || t >= tgt_end || (tgt_stride > 1 && t % tgt_stride != tgt_phase)) continue;
                // candidates: reads whose start lies in (T.gs - max_glen, T.ge)
                int64_t k0 = (int64_t)rank[t];
                while (k0 > 0 && S->reads[order[k0 - 1]].gs + max_glen > T.gs) k0--;
                for (uint32_t k = (uint32_t)k0; k < n_reads; k++) {
                    uint32_t q = order[k];
                    const Read& Q = S->reads[q];
                    if (Q.gs >= T.ge) break;
                    if (q == t) continue;
                    uint64_t a = std::max(T.gs, Q.gs), b = std::min(T.ge, Q.ge);
                    if (b <= a || b - a < min_ovl) continue;
                    // trim so the alignment starts/ends on a base present in both reads
                    while (a < b && !((T.ev[a - T.gs] & 1) && (Q.ev[a - Q.gs] & 1))) a++;
                    while (b > a && !((T.ev[b - 1 - T.gs] & 1) && (Q.ev[b - 1 - Q.gs] & 1))) b--;
                    if (b <= a || b - a < min_ovl) continue;
                    OpList ol;
                    uint32_t tspan = 0, qspan = 0;
                    for (uint64_t g = a; g < b; g++) {
                        uint8_t et = T.ev[g - T.gs], eq = Q.ev[g - Q.gs];
                        bool tp = et & 1, qp = eq & 1;
                        if (tp && qp) { ol.add('M', 1); tspan++; qspan++; }
                        else if (tp) { ol.add('D', 1); tspan++; }
                        else if (qp) { ol.add('I', 1); qspan++; }
                        if (g + 1 < b) {
                            uint32_t it = et >> 1, iq = eq >> 1, m = std::min(it, iq);
                            if (m) { ol.add('M', m); tspan += m; qspan += m; }
                            if (it > m) { ol.add('D', it - m); tspan += it - m; }
                            if (iq > m) { ol.add('I', iq - m); qspan += iq - m; }
                        }
                    }
                    uint32_t tfs = T.offset_at(a), qfs = Q.offset_at(a);  // genome-orientation offsets
                    uint32_t ts, te, qs, qe;
                    if (!T.strand) { ts = tfs; te = tfs + tspan; }
                    else { te = T.len() - tfs; ts = te - tspan; }
                    if (!Q.strand) { qs = qfs; qe = qfs + qspan; }
                    else { qe = Q.len() - qfs; qs = qe - qspan; }
                    if (T.strand) std::reverse(ol.ops.begin(), ol.ops.end());
                    uint32_t rel = (uint32_t)(T.strand ^ Q.strand);
                    uint32_t o9[9] = {q, Q.len(), qs, qe, rel, t, T.len(), ts, te};
                    P.ovl9.insert(P.ovl9.end(), o9, o9 + 9);
                    for (auto& op : ol.ops) append_op(P.cig, op.first, op.second);
                    P.coff.push_back(P.cig.size());
                }
            }
        };
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < std::max(1u, n_threads); t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    S->aln_off.assign(n_reads + 1, 0);
    S->cig_off.push_back(0);
    for (uint32_t t = 0; t < n_reads; t++) {
        PerTarget& P = per[t];
        uint64_t na = P.ovl9.size() / 9;
        S->aln_off[t + 1] = S->aln_off[t] + na;
        S->ovl9.insert(S->ovl9.end(), P.ovl9.begin(), P.ovl9.end());
        uint64_t base = S->cigars.size();
        S->cigars += P.cig;
        for (uint64_t k = 1; k <= na; k++) S->cig_off.push_back(base + P.coff[k]);
        PerTarget().cig.swap(P.cig);
        std::vector<uint32_t>().swap(P.ovl9);
    }
}

// read lengths (so that a caller can shard targets before asking for alignments)
void synth_read_lens(void* h, uint32_t* out) {
    Set* S = (Set*)h;
    for (size_t i = 0; i < S->reads.size(); i++) out[i] = S->reads[i].len();
}

void synth_free(void* h) { delete (Set*)h; }

// sizes: n_reads, total_bases, n_alns, cigar_bytes
void synth_sizes(void* h, uint64_t* out4) {
    Set* S = (Set*)h;
    uint64_t tb = 0;
    for (auto& R : S->reads) tb += R.seq.size();
    out4[0] = S->reads.size();
    out4[1] = tb;
    out4[2] = S->ovl9.size() / 9;
    out4[3] = S->cigars.size();
}

// reads: seqs/quals concatenated, off[n+1]; strand/hap/gs per read
void synth_get_reads(void* h, uint8_t* seqs, uint8_t* quals, uint64_t* off, uint8_t* strand, uint8_t* hap,
                     uint64_t* gs) {
    Set* S = (Set*)h;
    uint64_t o = 0;
    for (size_t i = 0; i < S->reads.size(); i++) {
        const Read& R = S->reads[i];
        off[i] = o;
        std::memcpy(seqs + o, R.seq.data(), R.seq.size());
        std::memcpy(quals + o, R.qual.data(), R.qual.size());
        o += R.seq.size();
        if (strand) strand[i] = R.strand;
        if (hap) hap[i] = R.hap;
        if (gs) gs[i] = R.gs;
    }
    off[S->reads.size()] = o;
}

void synth_get_alns(void* h, uint64_t* aln_off, uint32_t* ovl9, uint64_t* cig_off, uint8_t* cigars) {
    Set* S = (Set*)h;
    std::memcpy(aln_off, S->aln_off.data(), S->aln_off.size() * 8);
    std::memcpy(ovl9, S->ovl9.data(), S->ovl9.size() * 4);
    std::memcpy(cig_off, S->cig_off.data(), S->cig_off.size() * 8);
    std::memcpy(cigars, S->cigars.data(), S->cigars.size());
}

}  // extern "C"
