/*
 * gg_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A single-threaded C++ restatement of the reference's per-fragment algorithm
 * (grenaud/gargammel: src/fragSim.cpp, src/deamSim.cpp, src/adptSim.cpp), written from
 * the reference's behaviour, with every random draw taken from the counter-based
 * Philox4x32-10 stream specified in DESIGN.md ("Draw specification") instead of the
 * reference's rand()/libgab helpers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (gargammel_b200/csrc) never links or calls it.
 *
 * PARITY STATUS: the deterministic transforms (reverse complement, trimming, record grammar,
 * matrix row selection, cumulative-probability pick) are pinned against the reference sources
 * compiled with our shim (oracle/_ref, see oracle/Makefile) by tests/test_cli.py;
 * everything stochastic is pinned STATISTICALLY against those binaries (the reference is
 * time-seeded and its uniform helpers live in un-vendored libgab: bit-level parity of the
 * random stream is unpinnable).
 *
 * It reads bases THROUGH the FASTA line structure exactly like IndexedGenome::fetchSeq
 * (fragSim.cpp:305-329); the product instead packs the genome to 2 bits once.  The two
 * implementations share no code.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>
#include <string>
#include <vector>
#include <algorithm>
#include <map>
#include <math.h>

#include "../include/gargammel_b200.h"   /* struct layouts + enums only (gg_fai_entry, gg_range, modes) */

/* ------------------------------------------------------------------ Philox4x32-10 */
/* Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11). */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream ids (counter word 3) -- DESIGN.md "Draw specification" */
enum { ST_FRAG = 0, ST_DEAM = 1, ST_ADPT_F = 2, ST_DEAM2 = 3, ST_ADPT_R = 4, ST_FRAG2 = 5, ST_GCPROBE = 6, ST_FRAG3 = 7 };

struct Rng {
    uint32_t k0, k1;
    void block(uint64_t id, uint32_t blk, uint32_t stream, uint32_t out[4]) const {
        philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), blk, stream, k0, k1, out);
    }
    /* word w of the stream: block w/4, lane w%4 */
    uint32_t word(uint64_t id, uint32_t stream, uint32_t w) const {
        uint32_t o[4]; block(id, w >> 2, stream, o); return o[w & 3];
    }
};
/* the uniform in (0,1) used wherever the reference calls randomProb(): 31 bits, never 0 or 1 */
static inline double u01(uint32_t x) { return ((double)(x >> 1) + 0.5) * (1.0 / 2147483648.0); }

/* ------------------------------------------------------------------ helpers (libgab semantics inferred from call sites) */
static inline bool isResolvedDNA(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static inline int  base2int(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }
/* lower-case letters complement to lower-case letters: "methylated cytosines on the - strand can be specified using 'g'" (README.md:245)
 * only works if reverseComplement keeps the case that fragSim --case preserved */
static inline char complementBase(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
                 case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; default: return 'N'; }
}
static std::string reverseComplement(const std::string& s) {
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); i++) r[i] = complementBase(s[s.size() - 1 - i]);
    return r;
}

/* ------------------------------------------------------------------ state */
struct Chr { std::string name; int64_t len; uint64_t offset; int32_t line_blen, line_len; uint64_t start, end; /* 1-based concatenated, fragSim.cpp:1160-1162 */ };
struct Genome { std::vector<uint8_t> fasta; std::vector<Chr> chrs; /* name-sorted like std::map */ uint64_t G; int circ; /* --circ: index in chrs, -1 none */ };
struct Damage {
    int mode; int clamp_last;
    std::vector<double> sub5, sub3;  /* rows x 12 */
    double v, l, d, s;
    double s5, s3, l5, l3;           /* -damagess */
    std::vector<double> sub5m, sub3m; /* -matfilemeth (methylated = lower-case bases); sub5/sub3 then hold -matfilenonmeth */
    Damage() : mode(GG_DMG_NONE), clamp_last(1), v(0), l(0), d(0), s(0), s5(0), s3(0), l5(0), l3(0) {}
};
struct Comp { int dist; std::vector<double> s5p, s5m, s3p, s3m; double maxP, maxM; Comp() : dist(0), maxP(1), maxM(1) {} };
struct GcBias { bool on; int dist; double bias, mean, sd, maxNorm; GcBias() : on(false), dist(0), bias(0), mean(0.4), sd(0.1), maxNorm(1) {} };
struct orc_ctx {
    Rng rng;
    GcBias gc[GG_MAX_GC_SLOTS];
    Comp comp[GG_MAX_COMP_SLOTS];
    std::vector<Genome> genomes;
    int size_mode, fixed_len, minlen, maxlen, norev;
    int keep_case;                   /* fragSim --case (fragSim.cpp:548-551) */
    std::vector<int32_t> size_len; std::vector<double> size_cum;
    Damage dmg[GG_MAX_DAMAGE_SLOTS];
    std::string adF, adS; int read_len;
    int add_in_name; std::string adpt_tag; int deam_name;
    std::vector<gg_range> plan;
    std::string err;
    uint64_t attempts, gather_bytes;
    orc_ctx() : size_mode(GG_SIZE_FIXED), fixed_len(20), minlen(0), maxlen(1000), norev(0), keep_case(0),
        adF("AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG"),   /* adptSim.cpp:28 */
        adS("AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT"),         /* adptSim.cpp:29 */
        read_len(100), add_in_name(0), deam_name(0), attempts(0), gather_bytes(0) {}
};

/* ------------------------------------------------------------------ fragSim */
/* IndexedGenome::fetchSeq, fragSim.cpp:305-329 (uppercase branch; with --case the bytes as they are, :318-326).  Positions that run
 * past the file are returned as '\0' (the reference would read past the mmap: undefined). */
static std::string fetchSeq(const Genome& g, const Chr& c, int64_t index, int n, bool keep_case = false) {
    std::string s; if (n > 0) s.reserve(n);              /* n <= 0 gives "" like the reference loop (a --circ corner) */
    int64_t index2 = index;
    for (int j = 0; j < n; j++) {
        uint64_t pos = c.offset + (uint64_t)(index2 / c.line_blen) * c.line_len + (uint64_t)(index2 % c.line_blen);
        char ch = pos < g.fasta.size() ? (keep_case ? (char)g.fasta[pos] : (char)toupper(g.fasta[pos])) : '\0';
        s += ch;
        index2++;
    }
    return s;
}

/* selectFragmentLength, fragSim.cpp:137-187.  x0 = length word, x1 = fallback word (bits 31..1). */
static int selectFragmentLength(const orc_ctx* c, uint32_t x0, uint32_t x1) {
    if (c->size_mode == GG_SIZE_LIST) {                      /* :149-152  randomInt(0,n-1) */
        uint64_t n = c->size_len.size();
        return c->size_len[(size_t)(((uint64_t)x0 * n) >> 32)];
    }
    if (c->size_mode == GG_SIZE_FREQ) {                      /* :156-178 */
        double p = u01(x0);
        for (size_t i = 0; i < c->size_cum.size(); i++)
            if (0.0 <= p && p <= c->size_cum[i]) return c->size_len[i];   /* cProbLower stays 0.0 (:159,:163) */
        uint64_t n = c->size_len.size();                     /* :176 fallback: uniform row */
        return c->size_len[(size_t)(((uint64_t)(x1 >> 1) * n) >> 31)];
    }
    return c->fixed_len;                                     /* :181 */
}

struct FragResult { const Chr* chr; uint64_t idx; int length; bool plus; std::string frag; };

/* One fragment = the first accepted attempt; mirrors the while-loop body fragSim.cpp:1361-1507
 * with distFromEnd == 0 (no --comp: fragSim.cpp:1019-1021), no --circ, no -gc. */
/* pdfNorm, fragSim.cpp:96-102 */
static double pdfNorm(double x, double m, double s) { static const double inv_sqrt_2pi = 0.3989422804014327; double a = (x - m) / s; return inv_sqrt_2pi / s * exp(-0.5 * a * a); }
/* survival probability of a window with gcCount G/C and atCount A/T bases, fragSim.cpp:1524-1526 */
static double gcSurvival(const GcBias& B, unsigned gcCount, unsigned atCount) {
    double gcContent = double(gcCount) / double(gcCount + atCount);
    double pSurv = 1 / (1 + exp(B.bias * (gcContent - B.mean)));
    pSurv = pSurv * (pdfNorm(gcContent, B.mean, B.sd) / B.maxNorm);
    return pSurv;
}
static int sampleFragment(orc_ctx* c, const Genome& g, uint64_t id, int comp_slot, int gc_slot, FragResult* out) {
    const Comp* cp = (comp_slot >= 0 && c->comp[comp_slot].dist > 0) ? &c->comp[comp_slot] : NULL;
    const int d = cp ? cp->dist : 0;                         /* distFromEnd: 0 without --comp (fragSim.cpp:1019-1021) */
    for (uint32_t attempt = 0; attempt < GG_MAX_ATTEMPTS; attempt++) {
        uint32_t x[4];
        c->rng.block(id, attempt, ST_FRAG, x);
        c->attempts++;
        int length = selectFragmentLength(c, x[0], x[1]);
        if (length < d) continue;                            /* :1375 */
        if (length < c->minlen) continue;                    /* :1377 */
        if (length > c->maxlen) continue;                    /* :1380 */
        /* randGenomicCoord, fragSim.cpp:80-93: uniform in [0,G).  (multiply-shift instead of %: DESIGN.md) */
        uint64_t r64 = ((uint64_t)x[3] << 32) | x[2];
        uint64_t coord = (uint64_t)(((unsigned __int128)r64 * g.G) >> 64);
        /* chromosome scan fragSim.cpp:1397-1403: first chr with start+d <= coord <= end-d+1 */
        const Chr* found = NULL;
        for (size_t i = 0; i < g.chrs.size(); i++) {
            if (g.chrs[i].start + (uint64_t)d <= coord && coord + (uint64_t)d <= g.chrs[i].end + 1) { found = &g.chrs[i]; break; }
        }
        if (!found) continue;   /* the reference redraws only the coordinate (:1390); we redraw the attempt */
        uint64_t idx = coord - found->start;                 /* :1404 */
        std::string temp;
        const bool circular = g.circ >= 0 && found == &g.chrs[(size_t)g.circ];       /* :1407-1408 */
        /* :1410, in the reference's own uint64_t arithmetic (end-length-d+1 may wrap around for the very first contigs) */
        if (circular && !(coord <= (found->end - (uint64_t)length - (uint64_t)d + 1))) {
            /* need to circularize, :1412-1426: temp1 stops one base short of the contig end, temp2 restarts at base 0 */
            int n1 = (int)(found->end - (coord - (uint64_t)d));
            int l2 = (int)((uint64_t)length + 2 * (uint64_t)d - (found->end - (coord - (uint64_t)d)));
            temp = fetchSeq(g, *found, (int64_t)idx - d, n1, c->keep_case) + fetchSeq(g, *found, 0, l2, c->keep_case);
        } else {
            if (idx + (uint64_t)length + (uint64_t)d > (uint64_t)found->len) {
                /* run-off: the reference fetches past the contig and always hits '\n', '>' or EOF, so
                 * isResolvedDNAstring rejects it (:1472); with --circ on another contig it redraws the coordinate (:1428-1432).
                 * Same outcome, stated directly. */
                continue;
            }
            temp = fetchSeq(g, *found, (int64_t)idx - d, length + 2 * d, c->keep_case);   /* :1443 */
        }
        bool plus = c->norev ? true : ((x[1] & 1u) != 0);    /* :1462-1467 randomBool */
        bool resolved = true;                                /* :1472 isResolvedDNAstring; with --case a lower-case acgt is a (methylated) resolved base */
        for (size_t i = 0; i < temp.size(); i++) if (!isResolvedDNA((char)toupper(temp[i]))) { resolved = false; break; }
        if (!resolved) continue;
        if (!plus) temp = reverseComplement(temp);           /* :1484 */
        if (gc_slot >= 0 && c->gc[gc_slot].on) {             /* -gc, fragSim.cpp:1509-1534 (and again :1548-1576 without --comp) */
            unsigned gcCount = 0, atCount = 0;
            for (size_t i = 0; i < temp.size(); i++) { const char u = (char)toupper(temp[i]); if (u == 'G' || u == 'C') gcCount++; else atCount++; }
            const double pSurv = gcSurvival(c->gc[gc_slot], gcCount, atCount);
            uint32_t y[4]; c->rng.block(id, attempt, ST_FRAG3, y);
            if (u01(y[0]) > pSurv) continue;                 /* killed = (rP > pSurv) */
            if (!cp && u01(y[1]) > pSurv) continue;          /* the reference repeats the test when no composition file is given */
        }
        std::string preFrag = temp.substr(0, d), frag = temp.substr(d, length), posFrag = temp.substr(length + d, d);   /* :1485-1496 */
        if (cp) {                                            /* --comp acceptance, fragSim.cpp:1609-1759 */
            const std::vector<double>& s5 = plus ? cp->s5p : cp->s5m;
            const std::vector<double>& s3 = plus ? cp->s3p : cp->s3m;
            double probAcc = 1.0;
            for (int i = 0; i < d; i++) probAcc *= s5[(size_t)i * 4 + base2int((char)toupper(preFrag[i]))];                 /* :1614-1616 */
            for (int i = 0; i < d; i++) probAcc *= s5[(size_t)(i + d) * 4 + base2int((char)toupper(frag[i]))];              /* :1619-1621 */
            for (int i = 0; i < d; i++) probAcc *= s3[(size_t)i * 4 + base2int((char)toupper(frag[length - d + i]))];       /* :1624-1626 */
            for (int i = 0; i < d; i++) probAcc *= s3[(size_t)(i + d) * 4 + base2int((char)toupper(posFrag[i]))];           /* :1629-1631 */
            probAcc = probAcc / (plus ? cp->maxP : cp->maxM);                                                 /* :1633 */
            uint32_t y[4]; c->rng.block(id, attempt, ST_FRAG2, y);
            if (!(u01(y[0]) < probAcc)) continue;                                                             /* :1639 */
        }
        out->chr = found; out->idx = idx; out->length = length; out->plus = plus; out->frag = frag;
        const int w = length + 2 * d;
        c->gather_bytes += (uint64_t)((w + 3) / 4 + (w + 7) / 8);
        return 0;
    }
    return GG_ERR_GIVEUP;
}

static std::string u64str(uint64_t v) { char b[32]; snprintf(b, sizeof b, "%llu", (unsigned long long)v); return b; }

/* defline, fragSim.cpp:1489/1498 + tag without separator :1505-1507; printed with '>' :1603 */
static std::string fragDefline(const FragResult& f, const char* tag) {
    return ">" + f.chr->name + ":" + (f.plus ? "+" : "-") + ":" + u64str(f.idx) + ":" + u64str(f.idx + f.length) + ":" + u64str(f.length) + tag;
}

/* ------------------------------------------------------------------ deamSim */
/* cumulative table of a geometric_distribution<int>(p) (support 0,1,...; deamSim.cpp:297): CDF(k) = 1-(1-p)^(k+1),
 * accumulated as a running product so that product and oracle can build it identically on the host. */
static std::vector<double> geomCdf(double p, int n) {
    std::vector<double> cdf(n); double q = 1.0;
    for (int k = 0; k < n; k++) { q *= (1.0 - p); cdf[k] = 1.0 - q; }
    return cdf;
}
/* the same table, built once per parameter (the oracle is single-threaded) */
static const std::vector<double>& geomCdfCached(double p) {
    static std::map<double, std::vector<double> > cache;
    std::map<double, std::vector<double> >::iterator it = cache.find(p);
    if (it == cache.end()) it = cache.insert(std::make_pair(p, geomCdf(p, 4096))).first;
    return it->second;
}
static int geomDraw(const std::vector<double>& cdf, uint32_t x) {
    double p = u01(x);
    for (size_t k = 0; k < cdf.size(); k++) if (p <= cdf[k]) return (int)k;
    return (int)cdf.size();
}
#define GEOM_TABLE 4096

/* ---- draw specification v2 (DESIGN.md "Draw specification"): -matfile and -damage are drawn as EVENTS, not per base.
 * The reference draws one uniform per base and changes the base with the probability its table gives (deamSim.cpp:1847-1863,
 * 1488-1592).  Per-base draws are independent, so the same law is obtained by thinning: positions become "candidates" with a
 * probability m >= (the total rate of any base there), found by geometric skips, and a candidate with base b becomes a with
 * probability rate(b->a)/m.  The few terminal rows whose rates are far above the rest keep a per-base categorical draw. */
#define NT_MAX 8                     /* at most this many terminal rows per end are drawn per base */
#define NT_KAPPA 1.25                /* a row is terminal when its total rate exceeds NT_KAPPA x (largest rate of the rows >= NT_MAX) */
#define GEOM_MISS (1 << 30)          /* "beyond any record": an overhang / nick draw that ran off the 4096-entry table */

/* total substitution rate of base b in a row of 12, summed in the reference's order (a ascending, deamSim.cpp:1820-1830) */
static double rowRate(const double* row, int b) {
    double sum = 0.0;
    for (int a = 0; a < 4; a++) { if (a == b) continue; sum += row[(a < b) ? b * 3 + a : b * 3 + (a - 1)]; }
    return sum;
}
static double rowMax(const double* row) { double m = 0.0; for (int b = 0; b < 4; b++) m = std::max(m, rowRate(row, b)); return m; }
/* row of `sub` (n rows) that applies at distance `dist` from its end; NULL = position left untouched (-last, deamSim.cpp:1814-1816) */
static const double* effRow(const std::vector<double>& sub, int clamp_last, int dist) {
    int n = (int)(sub.size() / 12);
    if (dist >= n) { if (!clamp_last) return NULL; dist = n - 1; }          /* :1811-1813 */
    return &sub[(size_t)dist * 12];
}
struct MatrixPlan { int n5t, n3t; double M; std::vector<double> cdfM; };
static MatrixPlan matrixPlan(const Damage& D) {
    MatrixPlan P; P.n5t = P.n3t = 0;
    double tail = 0.0;                                                      /* largest total rate at distances >= NT_MAX */
    for (int e = 0; e < 2; e++) {
        const std::vector<double>& sub = e ? D.sub3 : D.sub5; const int n = (int)(sub.size() / 12);
        for (int r = NT_MAX; r < n; r++) tail = std::max(tail, rowMax(&sub[(size_t)r * 12]));
        if (D.clamp_last) tail = std::max(tail, rowMax(&sub[(size_t)(n - 1) * 12]));
    }
    double M = tail;
    for (int e = 0; e < 2; e++) {
        const std::vector<double>& sub = e ? D.sub3 : D.sub5;
        int nt = 0;
        for (int dist = 0; dist < NT_MAX; dist++) { const double* row = effRow(sub, D.clamp_last, dist); if (row && rowMax(row) > NT_KAPPA * tail) nt = dist + 1; }
        for (int dist = nt; dist < NT_MAX; dist++) { const double* row = effRow(sub, D.clamp_last, dist); if (row) M = std::max(M, rowMax(row)); }
        if (e) P.n3t = nt; else P.n5t = nt;
    }
    P.M = M;
    return P;
}
/* the reference's categorical pick for one base (deamSim.cpp:1797-1863): cumulative sums in its operation order, inclusive compares */
static int categorical(const double* row, int b, double p) {
    double sumProb[5], _sumProb[5]; sumProb[0] = 0.0; _sumProb[0] = 0.0;
    double sum = 0.0;
    for (int a = 0; a < 4; a++) {                                           /* :1820-1830 */
        if (a == b) continue;
        int idx = (a < b) ? b * 3 + a : b * 3 + (a - 1);
        sum += row[idx]; _sumProb[a + 1] = row[idx];
    }
    _sumProb[b + 1] = 1.0 - sum;                                            /* :1831 */
    double s = 0;
    for (int a = 0; a < 5; a++) { s += _sumProb[a]; sumProb[a] = s; }       /* :1833-1837 */
    for (int a = 0; a < 4; a++) if (sumProb[a] <= p && sumProb[a + 1] >= p) return a;   /* :1849-1863 */
    return b;
}

/* -matfile, deamSim.cpp:1794-1930, drawn as events.  Stream DEAM of the fragment:
 *   block t/4, word t%4 (t < n5t): the categorical draw of the 5' terminal position t           (exists while t < len/2)
 *   block B3 + t/4, word t%4 (t < n3t), B3 = ceil(n5t/4): the 3' terminal position len-1-t       (exists while len-1-t >= len/2)
 *   block BG + n/2, BG = B3 + ceil(n3t/4): round n of the interior [min(n5t,len/2), max(len/2,len-n3t)) uses words 2(n%2) (geometric
 *   skip under the majorant M) and 2(n%2)+1 (which substitution, if any, the candidate takes). */
static void deamMatrix(const orc_ctx* c, const Damage& D, uint64_t id, std::string& seq, std::vector<int>* deamPos) {
    const int len = (int)seq.size(), half = len / 2;                        /* i < len/2 is the 5' side (:1805) */
    const MatrixPlan P = matrixPlan(D);
    const int n5 = std::min(P.n5t, half), n3 = std::min(P.n3t, len - half);
    const int B3 = (P.n5t + 3) / 4, BG = B3 + (P.n3t + 3) / 4;
    std::vector<std::pair<int, int> > ev;                                   /* (position, new base) */
    for (int t = 0; t < n5; t++) {
        const int i = t;
        if (!isResolvedDNA(seq[i])) continue;                               /* :1795 */
        const double* row = effRow(D.sub5, D.clamp_last, t); if (!row) continue;
        const int b = base2int(seq[i]), a = categorical(row, b, u01(c->rng.word(id, ST_DEAM, (uint32_t)t)));
        if (a != b) ev.push_back(std::make_pair(i, a));
    }
    for (int t = 0; t < n3; t++) {
        const int i = len - 1 - t;
        if (!isResolvedDNA(seq[i])) continue;
        const double* row = effRow(D.sub3, D.clamp_last, t); if (!row) continue;   /* distance from the 3' end, :1873 */
        const int b = base2int(seq[i]), a = categorical(row, b, u01(c->rng.word(id, ST_DEAM, (uint32_t)(4 * B3 + t))));
        if (a != b) ev.push_back(std::make_pair(i, a));
    }
    if (P.M > 0.0) {
        const std::vector<double>& cdfM = geomCdfCached(P.M);
        const int end = len - n3;
        int pos = n5;
        for (uint32_t n = 0; pos < end; n++) {
            const uint32_t w0 = (uint32_t)(4 * BG) + 2 * n;
            const int j = geomDraw(cdfM, c->rng.word(id, ST_DEAM, w0));
            pos += j;                                                       /* a table miss is "at least GEOM_TABLE": no candidate, the next round goes on from there */
            if (j == GEOM_TABLE || pos >= end) continue;
            const int i = pos++;
            if (!isResolvedDNA(seq[i])) continue;
            const double* row = (i < half) ? effRow(D.sub5, D.clamp_last, i) : effRow(D.sub3, D.clamp_last, len - i - 1);
            if (!row) continue;
            const int b = base2int(seq[i]);
            const double p = u01(c->rng.word(id, ST_DEAM, w0 + 1));
            double cum = 0.0;
            for (int a = 0; a < 4; a++) {                                   /* rate(b->a)/M, a ascending */
                if (a == b) continue;
                cum += row[(a < b) ? b * 3 + a : b * 3 + (a - 1)];
                if (p < cum / P.M) { ev.push_back(std::make_pair(i, a)); break; }
            }
        }
    }
    std::sort(ev.begin(), ev.end());                                        /* the reference walks i upwards (:1794) */
    for (size_t k = 0; k < ev.size(); k++) {
        const int i = ev[k].first;
        seq[i] = "ACGT"[ev[k].second];
        if (deamPos) deamPos->push_back(i < half ? i + 1 : -(len - i));     /* :1856-1859, :1918-1921 */
    }
}

/* -damage v,l,d,s, deamSim.cpp:1477-1613, drawn as events.  Stream DEAM: block 0 words 0,1,2 = 5' overhang, 3' overhang, nick
 * position (geometric: first success of the per-base Bernoulli(v) scan :1503-1507); block 1+n: words 0 and 2 are geometric skips
 * (rate s) over the single-stranded positions (5' overhang, then 3' overhang; the whole fragment when the overhangs meet, :1483),
 * words 1 and 3 geometric skips (rate d) over the double-stranded middle.  A candidate changes iff its base is the one the region
 * deaminates (C->T before the nick / in the 5' overhang, G->A behind it / in the 3' overhang). */
static void deamBriggs(const orc_ctx* c, const Damage& D, uint64_t id, std::string& seq, std::vector<int>* deamPos) {
    const int len = (int)seq.size();
    const std::vector<double> &cdfL = geomCdfCached(D.l), &cdfV = geomCdfCached(D.v), &cdfS = geomCdfCached(D.s), &cdfD = geomCdfCached(D.d);
    uint32_t h[4]; c->rng.block(id, 0, ST_DEAM, h);
    long o5 = geomDraw(cdfL, h[0]), o3 = geomDraw(cdfL, h[1]), nick = geomDraw(cdfV, h[2]);   /* :1478, :1479, :1503-1507 */
    if (o5 == GEOM_TABLE) o5 = GEOM_MISS;                               /* off the table: longer than any record */
    if (o3 == GEOM_TABLE) o3 = GEOM_MISS;
    if (nick == GEOM_TABLE) nick = GEOM_MISS;
    const bool allss = o5 + o3 >= len;                                      /* :1483 */
    const long nS = allss ? len : o5 + o3, qend = allss ? len : len - o3;
    long v = 0, q = allss ? len : o5;
    std::vector<int> hits;
    for (uint32_t n = 0; v < nS || q < qend; n++) {
        uint32_t x[4]; c->rng.block(id, 1 + n, ST_DEAM, x);
        for (int sub = 0; sub < 2; sub++) {
            if (v < nS) {
                const int j = geomDraw(cdfS, x[2 * sub]);
                v += j;
                if (j != GEOM_TABLE && v < nS) {
                    const bool five = allss || v < o5;                      /* :1486, :1513 / :1527 */
                    const long i = five ? v : len - o3 + (v - o5);
                    if (five) { if (seq[i] == 'C') { seq[i] = 'T'; hits.push_back((int)i); } }
                    else      { if (seq[i] == 'G') { seq[i] = 'A'; hits.push_back((int)i); } }
                    v++;
                }
            }
            if (q < qend) {
                const int j = geomDraw(cdfD, x[2 * sub + 1]);
                q += j;
                if (j != GEOM_TABLE && q < qend) {
                    if (nick <= q) { if (seq[q] == 'G') { seq[q] = 'A'; hits.push_back((int)q); } }      /* :1542 */
                    else           { if (seq[q] == 'C') { seq[q] = 'T'; hits.push_back((int)q); } }      /* :1586 */
                    q++;
                }
            }
        }
    }
    if (deamPos) { std::sort(hits.begin(), hits.end()); for (size_t k = 0; k < hits.size(); k++) deamPos->push_back(hits[k] + 1); }
}

/* -damagess d,s5,s3,l5,l3 loop, deamSim.cpp:1337-1468 (single-strand libraries: C->T only).  Draw layout as for -damage:
 * DEAM words 0,1 = 5' / 3' overhang, per-base word 4+i. */
static void deamBriggsSS(const orc_ctx* c, const Damage& D, uint64_t id, std::string& seq, std::vector<int>* deamPos) {
    int len = (int)seq.size();
    std::vector<double> cdf5 = geomCdf(D.l5, GEOM_TABLE), cdf3 = geomCdf(D.l3, GEOM_TABLE);
    uint32_t h[4]; c->rng.block(id, 0, ST_DEAM, h);
    int overhang5p = geomDraw(cdf5, h[0]);                          /* :1342 */
    int overhang3p = geomDraw(cdf3, h[1]);                          /* :1343 */
    for (int i = 0; i < len; i++) {
        double rate;
        if ((overhang5p + overhang3p) >= len) {                     /* :1345 all single strand */
            if (((i + 1) <= overhang5p) && ((len - i) <= overhang3p)) rate = (i < int(double(len) * 0.5)) ? D.s5 : D.s3;   /* :1351-1371 */
            else if (!((i + 1) <= overhang5p) && ((len - i) <= overhang3p)) rate = D.s3;                                   /* :1372-1380 */
            else rate = D.s5;                                                                                              /* :1384-1393; "neither" cannot happen */
        } else {
            if ((i + 1) <= overhang5p) rate = D.s5;                 /* :1414-1424 */
            else if ((len - i) <= overhang3p) rate = D.s3;          /* :1428-1438 */
            else rate = D.d;                                        /* :1439-1449 */
        }
        if (seq[i] == 'C' && u01(c->rng.word(id, ST_DEAM, 4 + (uint32_t)i)) < rate) { seq[i] = 'T'; if (deamPos) deamPos->push_back(i + 1); }
    }
}

/* -mat single|double / -mapdamage loops, deamSim.cpp:1941-2002.  Row index clamped to the last row
 * (the reference indexes out of bounds for len > rows: DESIGN.md deviation). */
static void deamSingleDouble(const orc_ctx* c, const Damage& D, uint64_t id, std::string& seq, std::vector<int>* deamPos) {
    int n5 = (int)(D.sub5.size() / 12), n3 = (int)(D.sub3.size() / 12);
    int len = (int)seq.size();
    for (int i = 0; i < len; i++) {                                 /* :1944-1954 / :1976-1986 */
        if (seq[i] == 'C') {
            int r = std::min(i, n5 - 1);
            if (u01(c->rng.word(id, ST_DEAM, (uint32_t)i)) < D.sub5[(size_t)r * 12 + 5]) { seq[i] = 'T'; if (deamPos) deamPos->push_back(i + 1); }
        }
    }
    for (int i = 0; i < len; i++) {
        int r = std::min(len - i - 1, n3 - 1);
        if (D.mode == GG_DMG_SINGLE) {                              /* :1957-1967 */
            if (seq[i] == 'C' && u01(c->rng.word(id, ST_DEAM2, (uint32_t)i)) < D.sub3[(size_t)r * 12 + 5]) { seq[i] = 'T'; if (deamPos) deamPos->push_back(-(len - i)); }
        } else {                                                    /* :1989-2000 */
            if (seq[i] == 'G' && u01(c->rng.word(id, ST_DEAM2, (uint32_t)i)) < D.sub3[(size_t)r * 12 + 6]) { seq[i] = 'A'; if (deamPos) deamPos->push_back(-(len - i)); }
        }
    }
}

/* -matfilemeth / -matfilenonmeth, deamSim.cpp:1630-1787: a lower-case base is methylated and takes the "with methylation" matrix.
 * The sequence is NOT upper-cased first (:1319-1321); every resolved base that has a row is rewritten in upper case (:1707, :1769),
 * unresolved bases and positions beyond the rows of a -last table keep their bytes.  Drawn per base: word i of stream DEAM. */
static void deamMeth(const orc_ctx* c, const Damage& D, uint64_t id, std::string& seq, std::vector<int>* deamPos) {
    const int len = (int)seq.size();
    for (int i = 0; i < len; i++) {
        const char originalBase = (char)toupper(seq[i]);
        const bool isMethylated = originalBase != seq[i];                   /* :1638-1639 */
        if (!isResolvedDNA(originalBase)) continue;                         /* :1651-1652 */
        const bool five = i < len / 2;                                      /* :1661 */
        const std::vector<double>& sub = five ? (isMethylated ? D.sub5m : D.sub5) : (isMethylated ? D.sub3m : D.sub3);
        const double* row = effRow(sub, D.clamp_last, five ? i : len - i - 1);   /* :1665-1672, :1724-1731 */
        if (!row) continue;
        const int b = base2int(originalBase), a = categorical(row, b, u01(c->rng.word(id, ST_DEAM, (uint32_t)i)));
        seq[i] = "ACGT"[a];
        if (a != b && deamPos) deamPos->push_back(five ? i + 1 : -(len - i));   /* :1709-1711, :1770-1773 */
    }
}

static void deaminate(const orc_ctx* c, int slot, uint64_t id, std::string& seq, std::vector<int>* deamPos = NULL) {
    if (slot >= 0 && c->dmg[slot].mode == GG_DMG_METH) { deamMeth(c, c->dmg[slot], id, seq, deamPos); return; }   /* no toupper: deamSim.cpp:1319 */
    for (size_t i = 0; i < seq.size(); i++) seq[i] = (char)toupper(seq[i]);   /* deamSim.cpp:1319-1321 */
    if (slot < 0) return;
    const Damage& D = c->dmg[slot];
    switch (D.mode) {
        case GG_DMG_MATRIX: deamMatrix(c, D, id, seq, deamPos); break;
        case GG_DMG_BRIGGS: deamBriggs(c, D, id, seq, deamPos); break;
        case GG_DMG_BRIGGS_SS: deamBriggsSS(c, D, id, seq, deamPos); break;
        case GG_DMG_SINGLE: case GG_DMG_DOUBLE: deamSingleDouble(c, D, id, seq, deamPos); break;
        default: break;
    }
}

/* ------------------------------------------------------------------ adptSim */
/* randomDNASeq(n) (libgab): n uniform bases; word q/16 of the stream, 2 bits each */
static std::string randomDNASeq(const orc_ctx* c, uint64_t id, uint32_t stream, int n) {
    std::string s(n, 'A');
    for (int q = 0; q < n; q++) {
        uint32_t w = c->rng.word(id, stream, (uint32_t)(q >> 4));
        s[q] = "ACGT"[(w >> (2 * (q & 15))) & 3];
    }
    return s;
}
/* adptSim.cpp:288-322 */
static void adaptReads(const orc_ctx* c, uint64_t id, const std::string& seq_in, std::string& seqf, std::string& seqr, bool* addedAdapter = NULL, bool* addedRandom = NULL) {
    seqf = seq_in;
    for (size_t i = 0; i < seqf.size(); i++) seqf[i] = (char)toupper(seqf[i]);       /* :288 */
    seqr = reverseComplement(seqf);                                                   /* :289 */
    int L = c->read_len;
    if (addedAdapter) *addedAdapter = false;
    if (addedRandom) *addedRandom = false;
    if ((int)seqf.size() < L) { seqf = seqf + c->adF; seqf = seqf.substr(0, L); if (addedAdapter) *addedAdapter = true; }      /* :298-301 */
    else seqf = seqf.substr(0, L);                                                    /* :303 */
    if ((int)seqf.size() < L) { seqf = seqf + randomDNASeq(c, id, ST_ADPT_F, L - (int)seqf.size()); if (addedRandom) *addedRandom = true; }   /* :307-310 */
    if ((int)seqr.size() < L) { seqr = seqr + c->adS; seqr = seqr.substr(0, L); }      /* :313-315 */
    else seqr = seqr.substr(0, L);                                                    /* :317 */
    if ((int)seqr.size() < L) seqr = seqr + randomDNASeq(c, id, ST_ADPT_R, L - (int)seqr.size());   /* :320-322 */
}
/* adptSim.cpp:396-411 */
static void emitAdpt(int emit, const std::string& name_in, const std::string& seqf, const std::string& seqr, std::string& out, const std::string& suffix = "") {
    std::string namef = name_in + suffix, namer = name_in + "_r" + suffix;             /* :286, :375-393 */
    switch (emit) {
        case GG_EMIT_STDOUT4: out += namef + "\n" + seqf + "\n" + namer + "\n" + seqr + "\n"; break;   /* :408 */
        case GG_EMIT_ARTS: case GG_EMIT_FR: out += namef + "\n" + seqf + "\n"; break;                  /* :402 / :397 */
        case GG_EMIT_RR: out += namer + "\n" + seqr + "\n"; break;                                     /* :398 */
        case GG_EMIT_ARTP: out += namef + "\n" + seqf + reverseComplement(seqr) + "\n"; break;         /* :404-405 */
        default: break;
    }
}

/* ------------------------------------------------------------------ extern "C" surface (mirrors include/gargammel_b200.h with orc_ prefix) */
extern "C" {

int orc_ctx_create(uint64_t seed, orc_ctx** out) {
    orc_ctx* c = new orc_ctx(); c->rng.k0 = (uint32_t)seed; c->rng.k1 = (uint32_t)(seed >> 32); *out = c; return 0;
}
int orc_ctx_destroy(orc_ctx* c) { delete c; return 0; }
int orc_ctx_set_seed(orc_ctx* c, uint64_t seed) { c->rng.k0 = (uint32_t)seed; c->rng.k1 = (uint32_t)(seed >> 32); return 0; }
const char* orc_last_error(orc_ctx* c) { return c->err.c_str(); }

void orc_philox(uint64_t seed, uint64_t id, uint32_t blk, uint32_t stream, uint32_t* out4) {
    Rng r; r.k0 = (uint32_t)seed; r.k1 = (uint32_t)(seed >> 32); r.block(id, blk, stream, out4);
}
void orc_philox_raw(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
    philox4x32_10(ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1], out4);
}

int orc_genome_upload_circ(orc_ctx* c, const uint8_t* fasta, size_t n, const gg_fai_entry* fai, int n_chr, int circ_chr, int* id_out);
int orc_genome_upload(orc_ctx* c, const uint8_t* fasta, size_t n, const gg_fai_entry* fai, int n_chr, int* id_out) {
    return orc_genome_upload_circ(c, fasta, n, fai, n_chr, -1, id_out);
}
int orc_genome_upload_circ(orc_ctx* c, const uint8_t* fasta, size_t n, const gg_fai_entry* fai, int n_chr, int circ_chr, int* id_out) {
    Genome g; g.fasta.assign(fasta, fasta + n); g.circ = -1;
    std::map<std::string, Chr> byname;                      /* std::map ordering, fragSim.cpp:215 */
    for (int i = 0; i < n_chr; i++) {
        Chr ch; ch.name = fai[i].name; ch.len = fai[i].len; ch.offset = fai[i].offset;
        ch.line_blen = fai[i].line_blen; ch.line_len = fai[i].line_len; ch.start = ch.end = 0;
        if (ch.len <= 0 || ch.line_blen <= 0 || ch.line_len < ch.line_blen) { c->err = "bad fai entry"; return GG_ERR_ARG; }
        byname.insert(std::make_pair(ch.name, ch));         /* duplicates: first wins, like map::insert (:266) */
    }
    uint64_t G = 0;
    for (std::map<std::string, Chr>::iterator it = byname.begin(); it != byname.end(); ++it) {   /* :1152-1166 */
        Chr ch = it->second; ch.start = G + 1; ch.end = G + (uint64_t)ch.len; G += (uint64_t)ch.len;
        g.chrs.push_back(ch);
    }
    g.G = G;
    if (circ_chr >= 0) {                                    /* name2index[circName], fragSim.cpp:1173-1178 */
        if (circ_chr >= n_chr) { c->err = "bad circular contig"; return GG_ERR_ARG; }
        for (size_t i = 0; i < g.chrs.size(); i++) if (g.chrs[i].name == fai[circ_chr].name) g.circ = (int)i;
    }
    c->genomes.push_back(g);
    if (id_out) *id_out = (int)c->genomes.size() - 1;
    return 0;
}

int orc_tables_set_sizes(orc_ctx* c, int mode, const int32_t* len, const double* cum, int n, int fixed_len, int minlen, int maxlen) {
    c->size_mode = mode; c->fixed_len = fixed_len; c->minlen = minlen; c->maxlen = maxlen;
    c->size_len.clear(); c->size_cum.clear();
    if (mode != GG_SIZE_FIXED) {
        if (n <= 0) { c->err = "empty size table"; return GG_ERR_ARG; }
        c->size_len.assign(len, len + n);
        if (mode == GG_SIZE_FREQ) c->size_cum.assign(cum, cum + n);
    }
    return 0;
}
int orc_tables_set_norev(orc_ctx* c, int norev) { c->norev = norev; return 0; }
int orc_tables_set_matrix(orc_ctx* c, int slot, const double* sub5, int rows5, const double* sub3, int rows3, int clamp_last) {
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS || rows5 <= 0 || rows3 <= 0) return GG_ERR_ARG;
    Damage& D = c->dmg[slot]; D.mode = GG_DMG_MATRIX; D.clamp_last = clamp_last;
    D.sub5.assign(sub5, sub5 + (size_t)rows5 * 12); D.sub3.assign(sub3, sub3 + (size_t)rows3 * 12);
    return 0;
}
int orc_tables_set_case(orc_ctx* c, int keep_case) { c->keep_case = keep_case ? 1 : 0; return 0; }
int orc_tables_set_matrix_meth(orc_ctx* c, int slot, const double* no5, int rows_no5, const double* no3, int rows_no3,
                               const double* wi5, int rows_wi5, const double* wi3, int rows_wi3, int clamp_last) {
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS || rows_no5 <= 0 || rows_no3 <= 0 || rows_wi5 <= 0 || rows_wi3 <= 0) return GG_ERR_ARG;
    Damage& D = c->dmg[slot]; D = Damage(); D.mode = GG_DMG_METH; D.clamp_last = clamp_last;
    D.sub5.assign(no5, no5 + (size_t)rows_no5 * 12); D.sub3.assign(no3, no3 + (size_t)rows_no3 * 12);
    D.sub5m.assign(wi5, wi5 + (size_t)rows_wi5 * 12); D.sub3m.assign(wi3, wi3 + (size_t)rows_wi3 * 12);
    return 0;
}
int orc_tables_set_singledouble(orc_ctx* c, int slot, int protocol, const double* sub5, int rows5, const double* sub3, int rows3) {
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS || rows5 <= 0 || rows3 <= 0) return GG_ERR_ARG;
    if (protocol != GG_DMG_SINGLE && protocol != GG_DMG_DOUBLE) return GG_ERR_ARG;
    Damage& D = c->dmg[slot]; D.mode = protocol; D.clamp_last = 1;
    D.sub5.assign(sub5, sub5 + (size_t)rows5 * 12); D.sub3.assign(sub3, sub3 + (size_t)rows3 * 12);
    return 0;
}
/* maxProb as fragSim.cpp:1030-1047 */
int orc_tables_set_composition(orc_ctx* c, int slot, int dist, const double* s5p, const double* s5m, const double* s3p, const double* s3m) {
    if (slot < 0 || slot >= GG_MAX_COMP_SLOTS || dist < 1 || dist > GG_MAX_COMP_DIST) return GG_ERR_ARG;
    Comp& C = c->comp[slot]; C.dist = dist;
    const size_t n = (size_t)2 * dist * 4;
    C.s5p.assign(s5p, s5p + n); C.s5m.assign(s5m, s5m + n); C.s3p.assign(s3p, s3p + n); C.s3m.assign(s3m, s3m + n);
    double maxP = 1.0, maxM = 1.0;
    for (int i = 0; i < 2 * dist; i++) maxP = maxP * *std::max_element(s5p + 4 * i, s5p + 4 * i + 4);
    for (int i = 0; i < 2 * dist; i++) maxM = maxM * *std::max_element(s5m + 4 * i, s5m + 4 * i + 4);
    for (int i = 0; i < 2 * dist; i++) maxP = maxP * *std::max_element(s3p + 4 * i, s3p + 4 * i + 4);
    for (int i = 0; i < 2 * dist; i++) maxM = maxM * *std::max_element(s3m + 4 * i, s3m + 4 * i + 4);
    C.maxP = maxP; C.maxM = maxM;
    return 0;
}
int orc_tables_clear_composition(orc_ctx* c, int slot) { if (slot < 0 || slot >= GG_MAX_COMP_SLOTS) return GG_ERR_ARG; c->comp[slot] = Comp(); return 0; }
/* GC statistics of GG_GC_PROBES random 50-bp windows, fragSim.cpp:1190-1269 (probe p draws from stream GCPROBE, id = p) */
int orc_tables_set_gcbias(orc_ctx* c, int slot, int genome_id, int comp_dist, double gc_bias, double* mean_out, double* sd_out) {
    if (slot < 0 || slot >= GG_MAX_GC_SLOTS || genome_id < 0 || genome_id >= (int)c->genomes.size() || comp_dist < 0) return GG_ERR_ARG;
    const Genome& g = c->genomes[genome_id];
    const int d = comp_dist, length = GG_GC_PROBE_LEN, W = length + 2 * d;
    uint64_t S1 = 0, S2 = 0;
    for (uint64_t p = 0; p < GG_GC_PROBES; p++) {
        bool done = false;
        for (uint32_t attempt = 0; attempt < GG_MAX_ATTEMPTS && !done; attempt++) {
            uint32_t x[4]; c->rng.block(p, attempt, ST_GCPROBE, x);
            uint64_t r64 = ((uint64_t)x[3] << 32) | x[2];
            uint64_t coord = (uint64_t)(((unsigned __int128)r64 * g.G) >> 64);
            const Chr* found = NULL;
            for (size_t i = 0; i < g.chrs.size(); i++)     /* :1217-1219 */
                if (g.chrs[i].start + (uint64_t)d <= coord && coord + (uint64_t)length + (uint64_t)d <= g.chrs[i].end + 1) { found = &g.chrs[i]; break; }
            if (!found) continue;
            std::string temp = fetchSeq(g, *found, (int64_t)(coord - found->start) - d, W);
            bool resolved = true;
            for (size_t i = 0; i < temp.size(); i++) if (!isResolvedDNA(temp[i])) { resolved = false; break; }
            if (!resolved) continue;                         /* :1241-1242 */
            uint64_t gcCount = 0;
            for (size_t i = 0; i < temp.size(); i++) if (temp[i] == 'G' || temp[i] == 'C') gcCount++;
            S1 += gcCount; S2 += gcCount * gcCount; done = true;
        }
        if (!done) { c->err = "GC probe exhausted its attempts"; return GG_ERR_GIVEUP; }
    }
    GcBias& B = c->gc[slot];
    const double n = (double)GG_GC_PROBES, w = (double)W;
    B.mean = (double)S1 / (n * w);
    const double var = (double)S2 / (n * w * w) - B.mean * B.mean;
    B.sd = sqrt(var > 0 ? var : 0);
    B.maxNorm = pdfNorm(B.mean, B.mean, B.sd);               /* :1266 */
    B.bias = gc_bias; B.dist = d; B.on = true;
    if (mean_out) *mean_out = B.mean;
    if (sd_out) *sd_out = B.sd;
    return 0;
}
int orc_tables_clear_gcbias(orc_ctx* c, int slot) { if (slot < 0 || slot >= GG_MAX_GC_SLOTS) return GG_ERR_ARG; c->gc[slot] = GcBias(); return 0; }
int orc_tables_set_briggs(orc_ctx* c, int slot, double v, double l, double d, double s) {
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS) return GG_ERR_ARG;
    Damage& D = c->dmg[slot]; D.mode = GG_DMG_BRIGGS; D.v = v; D.l = l; D.d = d; D.s = s; return 0;
}
int orc_tables_set_briggs_ss(orc_ctx* c, int slot, double d, double s5, double s3, double l5, double l3) {
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS) return GG_ERR_ARG;
    Damage& D = c->dmg[slot]; D.mode = GG_DMG_BRIGGS_SS; D.d = d; D.s5 = s5; D.s3 = s3; D.l5 = l5; D.l3 = l3; return 0;
}
int orc_tables_clear_damage(orc_ctx* c, int slot) {
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS) return GG_ERR_ARG;
    c->dmg[slot] = Damage(); return 0;
}
int orc_tables_set_adapters(orc_ctx* c, const char* fwd, const char* rev, int read_len) {
    c->adF = fwd; c->adS = rev; c->read_len = read_len; return 0;
}
int orc_tables_set_deam_naming(orc_ctx* c, int on) { c->deam_name = on; return 0; }
int orc_tables_set_adpt_naming(orc_ctx* c, int add_in_name, const char* tag) { c->add_in_name = add_in_name; c->adpt_tag = tag ? tag : ""; return 0; }
int orc_plan_set(orc_ctx* c, const gg_range* r, int n) { c->plan.assign(r, r + n); return 0; }

/* fused: for each id: fragSim record -> deamSim -> adptSim, i.e. the literal composition of the three loops */
int orc_run_fused(orc_ctx* c, uint64_t first, uint64_t count, int emit, uint8_t* dst, size_t cap, size_t* n_out, gg_out* info) {
    std::string out;
    c->attempts = 0; c->gather_bytes = 0;
    for (uint64_t id = first; id < first + count; id++) {
        const gg_range* rg = NULL;
        for (size_t k = 0; k < c->plan.size(); k++)
            if (id >= c->plan[k].first_id && id < c->plan[k].first_id + c->plan[k].count) { rg = &c->plan[k]; break; }
        if (!rg) { c->err = "fragment id outside the plan"; return GG_ERR_STATE; }
        if (rg->genome < 0 || rg->genome >= (int)c->genomes.size()) { c->err = "bad genome id"; return GG_ERR_ARG; }
        FragResult f;
        int rc = sampleFragment(c, c->genomes[rg->genome], id, rg->comp_slot, rg->gc_slot, &f);
        if (rc) { c->err = "fragment exhausted its attempts"; return rc; }
        std::string name = fragDefline(f, rg->tag);
        std::string seq = f.frag;
        if (emit == GG_EMIT_FRAG) { out += name + "\n" + seq + "\n"; continue; }      /* fragSim.cpp:1603 */
        deaminate(c, rg->damage_slot, id, seq);
        if (emit == GG_EMIT_DEAM) { out += name + "\n" + seq + "\n"; continue; }      /* deamSim.cpp:2038 */
        std::string seqf, seqr;
        adaptReads(c, id, seq, seqf, seqr);
        emitAdpt(emit, name, seqf, seqr, out);
    }
    if (info) { memset(info, 0, sizeof *info); info->bytes = out.size(); info->records = count; info->attempts = c->attempts; info->gather_bytes = c->gather_bytes; }
    *n_out = out.size();
    if (out.size() > cap) return GG_ERR_CAPACITY;
    memcpy(dst, out.data(), out.size());
    return 0;
}

/* split a 2-line FASTA record stream (what fragSim/deamSim write; FastQParser(file,true) in deamSim.cpp:1288) */
static int splitRecords(orc_ctx* c, const uint8_t* in, size_t n, std::vector<std::pair<std::string, std::string> >& recs) {
    size_t p = 0;
    while (p < n) {
        const uint8_t* e = (const uint8_t*)memchr(in + p, '\n', n - p);
        size_t le = e ? (size_t)(e - in) : n;
        std::string id((const char*)in + p, le - p);
        p = le + 1;
        if (id.empty() && p >= n) break;
        if (id.empty() || id[0] != '>') { c->err = "record does not start with '>'"; return GG_ERR_PARSE; }
        if (p > n) p = n;
        e = p < n ? (const uint8_t*)memchr(in + p, '\n', n - p) : NULL;
        le = e ? (size_t)(e - in) : n;
        std::string seq((const char*)in + p, le - p);
        p = le + 1;
        recs.push_back(std::make_pair(id, seq));
    }
    return 0;
}

int orc_run_deam(orc_ctx* c, const uint8_t* in, size_t n, int slot, uint64_t first_id, uint8_t* dst, size_t cap, size_t* n_out, gg_out* info) {
    std::vector<std::pair<std::string, std::string> > recs;
    int rc = splitRecords(c, in, n, recs); if (rc) return rc;
    std::string out;
    for (size_t r = 0; r < recs.size(); r++) {
        std::string seq = recs[r].second;
        std::vector<int> deamPos;
        deaminate(c, slot, first_id + r, seq, &deamPos);
        std::string idl = recs[r].first;
        if (c->deam_name && !deamPos.empty()) {                                       /* deamSim.cpp:2035-2036: ID + "_DEAM:" + vectorToString(deamPos) */
            idl += "_DEAM:";
            for (size_t k = 0; k < deamPos.size(); k++) { if (k) idl += ","; char b[16]; snprintf(b, sizeof b, "%d", deamPos[k]); idl += b; }
        }
        out += idl + "\n" + seq + "\n";                                              /* deamSim.cpp:2038 */
    }
    if (info) { memset(info, 0, sizeof *info); info->bytes = out.size(); info->records = recs.size(); info->in_bytes = n; }
    *n_out = out.size();
    if (out.size() > cap) return GG_ERR_CAPACITY;
    memcpy(dst, out.data(), out.size());
    return 0;
}

int orc_run_adpt(orc_ctx* c, const uint8_t* in, size_t n, int emit, uint64_t first_id, uint8_t* dst, size_t cap, size_t* n_out, gg_out* info) {
    std::vector<std::pair<std::string, std::string> > recs;
    int rc = splitRecords(c, in, n, recs); if (rc) return rc;
    std::string out;
    for (size_t r = 0; r < recs.size(); r++) {
        std::string seqf, seqr; bool aA, aR;
        adaptReads(c, first_id + r, recs[r].second, seqf, seqr, &aA, &aR);
        std::string suffix;                                                           /* adptSim.cpp:375-393 */
        if (c->add_in_name) { if (aA) suffix += "_adpt"; if (aR) suffix += "_rand"; }
        suffix += c->adpt_tag;
        emitAdpt(emit, recs[r].first, seqf, seqr, out, suffix);
    }
    if (info) { memset(info, 0, sizeof *info); info->bytes = out.size(); info->records = recs.size(); info->in_bytes = n; }
    *n_out = out.size();
    if (out.size() > cap) return GG_ERR_CAPACITY;
    memcpy(dst, out.data(), out.size());
    return 0;
}

/* deterministic transforms exposed one by one for level-2 parity tests */
int orc_revcomp(const char* in, int n, char* out) {
    std::string r = reverseComplement(std::string(in, n)); memcpy(out, r.data(), n); return 0;
}
int orc_fetch(orc_ctx* c, int genome, int chr_sorted, int64_t start, int len, char* dst) {
    if (genome < 0 || genome >= (int)c->genomes.size()) return GG_ERR_ARG;
    const Genome& g = c->genomes[genome];
    if (chr_sorted < 0 || chr_sorted >= (int)g.chrs.size()) return GG_ERR_ARG;
    std::string s = fetchSeq(g, g.chrs[chr_sorted], start, len); memcpy(dst, s.data(), len); return 0;
}

} /* extern "C" */
