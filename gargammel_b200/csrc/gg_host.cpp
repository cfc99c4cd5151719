// gg_host.cpp -- file-format layer + fragment apportioning (host, C++11, no CUDA).
#include "gg_host.h"
#include <zlib.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <errno.h>
#include <algorithm>
#include <map>
#include <thread>
#include <queue>

namespace ggh {

bool read_file(const std::string& path, std::vector<uint8_t>& out, std::string& err) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) { err = "Unable to open file " + path; return false; }
    gzbuffer(f, 1 << 20);
    out.clear();
    std::vector<uint8_t> buf(1 << 22);
    while (true) {
        int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) { err = "Read error in " + path; gzclose(f); return false; }
        if (n == 0) break;
        out.insert(out.end(), buf.begin(), buf.begin() + n);
    }
    gzclose(f);
    return true;
}

std::vector<std::string> all_tokens(const std::string& s, char delim) {
    std::vector<std::string> out; std::string cur;
    for (size_t i = 0; i < s.size(); i++) { if (s[i] == delim) { out.push_back(cur); cur.clear(); } else cur += s[i]; }
    out.push_back(cur);
    return out;
}

static void split_lines(const std::vector<uint8_t>& b, std::vector<std::string>& lines) {
    size_t p = 0;
    while (p < b.size()) {
        const uint8_t* e = (const uint8_t*)memchr(b.data() + p, '\n', b.size() - p);
        size_t le = e ? (size_t)(e - b.data()) : b.size();
        lines.push_back(std::string((const char*)b.data() + p, le - p));
        p = le + 1;
    }
}
// destringify<T> of the reference = operator>> on an istringstream: leading blanks skipped, trailing junk ignored
static bool to_double(const std::string& s, double* v) { char* e; errno = 0; *v = strtod(s.c_str(), &e); return e != s.c_str(); }
static bool to_i64(const std::string& s, long long* v) { char* e; errno = 0; *v = strtoll(s.c_str(), &e, 10); return e != s.c_str(); }

bool parse_fai(const std::string& path, std::vector<FaiEntry>& out, std::string& err) {
    std::vector<uint8_t> b;
    if (!read_file(path, b, err)) { err = "Cannot open faidx: " + path; return false; }
    std::vector<std::string> lines; split_lines(b, lines);
    out.clear();
    for (size_t i = 0; i < lines.size(); i++) {
        const std::string& line = lines[i];
        if (line.empty()) continue;                                   // fragSim.cpp:239
        if (line.find('\t') == std::string::npos) continue;           // fragSim.cpp:243
        std::vector<std::string> t = all_tokens(line, '\t');
        if (t.size() != 5) { err = "Parse error in " + line; return false; }   // fragSim.cpp:250-253
        FaiEntry e; long long a, o, bl, ll;
        if (!to_i64(t[1], &a) || !to_i64(t[2], &o) || !to_i64(t[3], &bl) || !to_i64(t[4], &ll)) { err = "Parse error in " + line; return false; }
        e.name = t[0]; e.len = a; e.offset = (uint64_t)o; e.line_blen = (int32_t)bl; e.line_len = (int32_t)ll;
        out.push_back(e);
    }
    if (out.empty()) { err = "Empty faidx: " + path; return false; }
    return true;
}

bool build_fai(const std::vector<uint8_t>& fa, std::vector<FaiEntry>& out, std::string& err) {
    out.clear();
    size_t p = 0, n = fa.size();
    FaiEntry cur; bool have = false; int64_t seen_short = 0;
    while (p < n) {
        const uint8_t* e = (const uint8_t*)memchr(fa.data() + p, '\n', n - p);
        size_t le = e ? (size_t)(e - fa.data()) : n;
        size_t llen = le - p + (e ? 1 : 0);
        size_t blen = le - p;
        if (blen && fa[le - 1] == '\r') blen--;
        if (fa[p] == '>') {
            if (have) out.push_back(cur);
            size_t q = p + 1; while (q < le && !isspace(fa[q])) q++;
            cur = FaiEntry(); cur.name = std::string((const char*)fa.data() + p + 1, q - p - 1);
            cur.len = 0; cur.offset = le + 1; cur.line_blen = 0; cur.line_len = 0; have = true; seen_short = 0;
        } else if (have && blen) {
            if (cur.line_blen == 0) { cur.line_blen = (int32_t)blen; cur.line_len = (int32_t)llen; }
            else {
                if (seen_short) { err = "build_fai: different line length in sequence " + cur.name; return false; }
                if ((int32_t)blen != cur.line_blen) { if ((int32_t)blen > cur.line_blen) { err = "build_fai: different line length in sequence " + cur.name; return false; } seen_short = 1; }
            }
            cur.len += (int64_t)blen;
        }
        p = le + 1;
    }
    if (have) out.push_back(cur);
    if (out.empty()) { err = "build_fai: no sequence found"; return false; }
    return true;
}

bool load_size_list(const std::string& path, std::vector<int32_t>& len, std::string& err) {
    std::vector<uint8_t> b;
    if (!read_file(path, b, err)) { err = "Unable to open size distribution file " + path; return false; }   // fragSim.cpp:742
    std::vector<std::string> lines; split_lines(b, lines);
    len.clear();
    for (size_t i = 0; i < lines.size(); i++) {
        if (all_tokens(lines[i], '\t').size() != 1) { err = "The following line " + lines[i] + " in file " + path + " does not have 1 field"; return false; }   // :730-733
        long long t;
        if (!to_i64(lines[i], &t)) { err = "The following line " + lines[i] + " in file " + path + " is not a number"; return false; }
        if (t >= 1) len.push_back((int32_t)t);                        // :736-737
    }
    if (len.empty()) { err = "Size distribution file " + path + " holds no usable length"; return false; }
    return true;
}

bool load_size_freq(const std::string& path, std::vector<int32_t>& len, std::vector<double>& cum, std::string& err) {
    std::vector<uint8_t> b;
    if (!read_file(path, b, err)) { err = "Unable to open size frequency file " + path; return false; }      // fragSim.cpp:774
    std::vector<std::string> lines; split_lines(b, lines);
    len.clear(); cum.clear();
    double cumul = 0.0;
    for (size_t i = 0; i < lines.size(); i++) {
        std::vector<std::string> f = all_tokens(lines[i], '\t');
        if (f.size() != 2) { err = "The following line " + lines[i] + " in file " + path + " does not have 2 fields"; return false; }   // :760-763
        double fr; long long l;
        if (!to_double(f[1], &fr) || !to_i64(f[0], &l)) { err = "The following line " + lines[i] + " in file " + path + " cannot be parsed"; return false; }
        cumul += fr;                                                  // :765
        len.push_back((int32_t)l); cum.push_back(cumul);              // :766-769
    }
    if (cumul < 0.99 || cumul > 1.01) {                               // :778-782
        char t[64]; snprintf(t, sizeof t, "%g", cumul);
        err = "Problem in file: " + path + ", the sum of frequencies does not sum to 1, sum=" + t; return false;
    }
    return true;
}

void lognormal_table(double loc, double scale, int cap, std::vector<int32_t>& len, std::vector<double>& cum) {
    len.clear(); cum.clear();
    for (int l = 0; l <= cap; l++) {
        const double z = (log((double)l + 1.0) - loc) / scale;
        len.push_back(l); cum.push_back(0.5 * erfc(-z * 0.70710678118654752440));
    }
    len.push_back(cap + 1); cum.push_back(1.0);
}

bool load_dnacomp(const std::string& path, int dist, CompTables& out, std::string& err) {
    std::vector<uint8_t> b;
    if (dist < 1 || dist > GG_MAX_COMP_DIST) { err = "--dist must be between 1 and 32"; return false; }
    if (!read_file(path, b, err)) { err = "Unable to open composition file " + path; return false; }       // fragSim.cpp:1014
    std::vector<std::string> lines; split_lines(b, lines);
    out.dist = dist; out.s5p.clear(); out.s5m.clear(); out.s3p.clear(); out.s3m.clear();
    int first[4] = {-1000, -1000, -1000, -1000};                        // sub5pPlusi, sub5pMinusi, sub3pPlusi, sub3pMinusi (:793-796)
    unsigned int tot[5] = {0, 0, 0, 0, 0};
    for (size_t li = 0; li < lines.size(); li++) {
        const std::string& line = lines[li];
        if (line.compare(0, 1, "#") == 0 || line.compare(0, 3, "Chr") == 0) continue;          // :817-819
        std::vector<std::string> f = all_tokens(line, '\t');
        if (f.size() <= 7) continue;                                     // :825
        if (f.size() < 9) { err = "Parse error in composition file " + path + ": " + line; return false; }
        long long d, c[5];
        if (!to_i64(f[3], &d)) { err = "Parse error in composition file " + path + ": " + line; return false; }
        for (int k = 0; k < 5; k++) if (!to_i64(f[4 + k], &c[k])) { err = "Parse error in composition file " + path + ": " + line; return false; }
        if (d >= -(long long)dist && d <= (long long)dist) {             // abs(d) <= dist, :837 (without llabs: LLONG_MIN in a mangled file)
            const bool is3 = f[1] == "3p", is5 = f[1] == "5p", plus = f[2] == "+", minus = f[2] == "-";
            if (!(is3 || is5) || !(plus || minus)) continue;
            const int which = (is5 ? 0 : 2) + (plus ? 0 : 1);
            if (first[which] == -1000) first[which] = (int)d;
            else if (d < first[which]) { err = "Parse error, wrong order of distance to fragment end  for line " + line; return false; }   // :846-849
            std::vector<double>& v = which == 0 ? out.s5p : which == 1 ? out.s5m : which == 2 ? out.s3p : out.s3m;
            for (int k = 0; k < 4; k++) v.push_back((double)(unsigned int)c[k] / (double)(unsigned int)c[4]);      // :853-857
        } else for (int k = 0; k < 5; k++) tot[k] += (unsigned int)c[k];  // background, :922-929
    }
    const size_t need = (size_t)2 * dist * 4;
    if (out.s5p.size() != need || out.s5m.size() != need || out.s3p.size() != need || out.s3m.size() != need) {
        err = "Composition file " + path + " does not hold " + std::to_string(2 * dist) + " positions within --dist of every end/strand"; return false;
    }
    double bg[4]; for (int k = 0; k < 4; k++) bg[k] = (double)tot[k] / (double)tot[4];             // :937-940
    std::vector<double>* all[4] = {&out.s5p, &out.s5m, &out.s3p, &out.s3m};
    for (int t = 0; t < 4; t++) for (size_t i = 0; i < need; i++) { const double x = (*all[t])[i] - bg[i & 3]; (*all[t])[i] = 0.25 + x; }   // :948-1008
    return true;
}

// one .dat/.prof file: header line skipped; per line: tab fields (dropping the first `skip` columns), token before ' ' -> double
static bool load_rate_file(const std::string& path, int skip, Rates& sub, std::string& err, const char* which) {
    std::vector<uint8_t> b;
    if (!read_file(path, b, err)) { err = std::string("Unable to open ") + which + " deamination file " + path; return false; }   // deamSim.cpp:1165-1168
    std::vector<std::string> lines; split_lines(b, lines);
    if (lines.empty()) { err = std::string("Parsing error for ") + which + " deamination file " + path; return false; }          // :1121-1124
    sub.clear();
    for (size_t i = 1; i < lines.size(); i++) {
        std::vector<std::string> fields = all_tokens(lines[i], '\t');
        std::vector<std::string> first;
        for (size_t k = (size_t)skip; k < fields.size(); k++) first.push_back(all_tokens(fields[k], ' ')[0]);   // :1127-1132
        if (first.size() != 12) { char t[32]; snprintf(t, sizeof t, "%zu", first.size()); err = "line from deamination does not have 12 fields " + lines[i] + " " + t; return false; }   // :1134-1137
        double r[12];
        for (int k = 0; k < 12; k++) if (!to_double(first[k], &r[k])) { err = "cannot parse rate \"" + first[k] + "\" in " + path; return false; }
        for (int bb = 0; bb < 4; bb++) {                                 // :1148-1158
            double sum = 0.0;
            for (int a = 0; a < 3; a++) sum += r[bb * 3 + a];
            if (sum > 1.0) { err = "Problem with line " + lines[i] + " in file " + path + " the sum of the substitution probabilities exceeds 1"; return false; }
        }
        sub.insert(sub.end(), r, r + 12);
    }
    if (sub.empty()) { err = std::string("No rows in ") + which + " deamination file " + path; return false; }
    return true;
}
static void reverse_rows(Rates& s) {
    const size_t rows = s.size() / 12;
    for (size_t i = 0; i < rows / 2; i++) for (int k = 0; k < 12; k++) std::swap(s[i * 12 + k], s[(rows - 1 - i) * 12 + k]);
}
bool load_matrix_dat(const std::string& prefix, Rates& sub5, Rates& sub3, std::string& err) {
    if (!load_rate_file(prefix + "5.dat", 1, sub5, err, "5p")) return false;     // deamSim.cpp:1102
    if (!load_rate_file(prefix + "3.dat", 1, sub3, err, "3p")) return false;     // :1103
    reverse_rows(sub3);                                                           // :1236
    return true;
}
bool load_profile(const std::string& prefix, Rates& sub5, Rates& sub3, std::string& err) {
    if (!load_rate_file(prefix + "5p.prof", 0, sub5, err, "5p")) return false;   // deamSim.cpp:960
    if (!load_rate_file(prefix + "3p.prof", 0, sub3, err, "3p")) return false;   // :961  (not reversed, :1090)
    return true;
}

bool load_mapdamage(const std::string& path, Rates& sub5, Rates& sub3, std::string& err) {
    std::vector<uint8_t> b;
    if (!read_file(path, b, err)) { err = "Unable to open misincorporation file " + path; return false; }    // deamSim.cpp:611-613
    std::vector<std::string> lines; split_lines(b, lines);
    if (lines.size() < 4) { err = "Parsing error for misincorporation file " + path; return false; }        // :471-481
    std::vector<std::string> hdr = all_tokens(lines[3], '\t');
    if (hdr.size() < 21) { err = "Parsing error for misincorporation file " + path; return false; }
    int map12[12];
    for (int i = 0; i < 12; i++) {                                      // :486-536
        const std::string& h = hdr[9 + i];
        if (h.size() != 3 || h[1] != '>') { err = "Parsing error for misincorporation file " + path + " wrong header field " + h; return false; }
        const char* B = "ACGT";
        const char* f = strchr(B, h[0]); const char* t = strchr(B, h[2]);
        if (!f || !t || !h[0] || !h[2] || f == t) { err = "Parsing error for misincorporation file " + path; return false; }
        const int from = (int)(f - B), to = (int)(t - B);
        map12[i] = from * 3 + (to < from ? to : to - 1);
    }
    struct Cnt { double sub[12]; double base[4]; };
    std::vector<Cnt> c5, c3; bool first5 = true, first3 = true;
    for (size_t li = 4; li < lines.size(); li++) {                      // :539-609
        if (lines[li].empty() && li + 1 == lines.size()) break;
        std::vector<std::string> f = all_tokens(lines[li], '\t');
        if (f.size() < 21) { err = "line from misincorporation.txt does not have a 5p or 3p in the 2nd column: " + lines[li]; return false; }
        const bool is3 = f[1] == "3p", is5 = f[1] == "5p";
        if (!is3 && !is5) { err = "line from misincorporation.txt does not have a 5p or 3p in the 2nd column: " + lines[li]; return false; }   // :604
        std::vector<Cnt>& C = is3 ? c3 : c5; bool& first = is3 ? first3 : first5;
        if (first && f[2] == "-") first = false;                        // :544-546 / :576-578
        long long v;
        if (first) {
            Cnt c; memset(&c, 0, sizeof c);
            for (int i = 0; i < 4; i++) { if (!to_i64(f[4 + i], &v)) { err = "bad count in " + lines[li]; return false; } c.base[i] = (double)(uint32_t)v; }
            for (int i = 0; i < 12; i++) { if (!to_i64(f[9 + i], &v)) { err = "bad count in " + lines[li]; return false; } c.sub[map12[i]] = (double)(uint32_t)v; }
            C.push_back(c);
        } else {
            long long pos; if (!to_i64(f[3], &pos) || pos < 1 || (size_t)pos > C.size()) { err = "bad position in " + lines[li]; return false; }
            Cnt& c = C[(size_t)pos - 1];                                // :561 / :592
            for (int i = 0; i < 4; i++) { if (!to_i64(f[4 + i], &v)) { err = "bad count in " + lines[li]; return false; } c.base[i] = (double)(uint32_t)((uint32_t)c.base[i] + (uint32_t)v); }
            for (int i = 0; i < 12; i++) { if (!to_i64(f[9 + i], &v)) { err = "bad count in " + lines[li]; return false; } c.sub[map12[i]] = (double)(uint32_t)((uint32_t)c.sub[map12[i]] + (uint32_t)v); }
        }
    }
    if (c5.size() != c3.size()) { err = "Error in misincorporation file " + path + " the number of defined bases for the 5' and 3' ends differ"; return false; }   // :616-619
    sub5.clear(); sub3.clear();
    for (size_t i = 0; i < c5.size(); i++) for (int bb = 0; bb < 4; bb++) for (int a = 0; a < 3; a++) sub5.push_back(c5[i].sub[bb * 3 + a] / c5[i].base[bb]);   // :622-632
    for (size_t i = 0; i < c3.size(); i++) for (int bb = 0; bb < 4; bb++) for (int a = 0; a < 3; a++) sub3.push_back(c3[i].sub[bb * 3 + a] / c3[i].base[bb]);   // :641-649
    if (sub5.empty()) { err = "Parsing error for misincorporation file " + path; return false; }
    return true;
}

bool load_bact_list(const std::string& path, std::vector<std::string>& names, std::vector<double>& w, std::string& err) {
    std::vector<uint8_t> b;
    if (!read_file(path, b, err)) { err = "List file " + path + " does not exist"; return false; }
    std::vector<std::string> lines; split_lines(b, lines);
    names.clear(); w.clear();
    double sum = 0;
    for (size_t i = 0; i < lines.size(); i++) {
        // /^(\S+)\s+(\S+)$/   gargammel.pl:870
        const std::string& l = lines[i];
        size_t a = 0; while (a < l.size() && !isspace((unsigned char)l[a])) a++;
        size_t c = a; while (c < l.size() && isspace((unsigned char)l[c])) c++;
        size_t d = c; while (d < l.size() && !isspace((unsigned char)l[d])) d++;
        double v;
        if (a == 0 || c == a || d == c || d != l.size() || !to_double(l.substr(c, d - c), &v)) { err = "Cannot parse line from bacterial list: " + path + " line:" + l; return false; }
        names.push_back(l.substr(0, a)); w.push_back(v); sum += v;
    }
    if (sum < 0.99 || sum > 1.01) { char t[64]; snprintf(t, sizeof t, "%g", sum); err = "Problem from bacterial list: " + path + " sum is not 1, found: " + t; return false; }   // gargammel.pl:887
    return true;
}

// ---------------------------------------------------------------- apportioning
// host-side Philox4x32-10 (same generator as the device path; used for the per-source counts only)
static void philox(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
struct UStream {                        // uniform doubles in (0,1), 53 bits
    uint64_t seed, stream, n; uint32_t buf[4]; int have;
    UStream(uint64_t s, uint64_t st) : seed(s), stream(st), n(0), have(0) {}
    double next() {
        if (have < 2) { philox(n++, stream | (0xA5ull << 56), seed, buf); have = 4; }
        uint64_t v = ((uint64_t)buf[have - 1] << 32) | buf[have - 2]; have -= 2;
        return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
};
static double stirling_tail(double k) {   // log(k!) - [ (k+0.5)log(k+1) - (k+1) + 0.5 log(2 pi) ]
    static const double t[] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509, 0.0166446911898211,
                               0.0138761288230707, 0.0118967099458917, 0.0104112652619720, 0.00925546218271273, 0.00833056343336287};
    if (k <= 9) return t[(int)k];
    const double kp1sq = (k + 1) * (k + 1);
    return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / (k + 1);
}
static uint64_t binomial_stream(uint64_t n, double p, UStream& U) {
    if (n == 0 || p <= 0.0) return 0;
    if (p >= 1.0) return n;
    if (p > 0.5) return n - binomial_stream(n, 1.0 - p, U);
    const double nd = (double)n;
    if (nd * p < 10.0) {                 // inversion by sequential search
        const double q = 1.0 - p, s = p / q;
        double f = pow(q, nd), u = U.next(); uint64_t k = 0;
        while (u > f && k < n) { u -= f; k++; f *= s * (nd - (double)k + 1.0) / (double)k; if (f <= 0) break; }
        return k;
    }
    // BTRS: Hormann (1993) transformed rejection with squeeze
    const double spq = sqrt(nd * p * (1 - p)), b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * p, c = nd * p + 0.5;
    const double vr = 0.92 - 4.2 / b, alpha = (2.83 + 5.1 / b) * spq, r = p / (1 - p), m = floor((nd + 1) * p);
    while (true) {
        const double u = U.next() - 0.5; double v = U.next();
        const double us = 0.5 - fabs(u), k = floor((2 * a / us + b) * u + c);
        if (k < 0 || k > nd) continue;
        if (us >= 0.07 && v <= vr) return (uint64_t)k;
        v = log(v * alpha / (a / (us * us) + b));
        const double ub = (m + 0.5) * log((m + 1) / (r * (nd - m + 1))) + (nd + 1) * log((nd - m + 1) / (nd - k + 1)) +
                          (k + 0.5) * log(r * (nd - k + 1) / (k + 1)) + stirling_tail(m) + stirling_tail(nd - m) - stirling_tail(k) - stirling_tail(nd - k);
        if (v <= ub) return (uint64_t)k;
    }
}
uint64_t binomial_draw(uint64_t n, double p, uint64_t seed, uint64_t stream) { UStream U(seed, stream); return binomial_stream(n, p, U); }

// Multinomial(n, w/sum w) by conditional binomials: the law of n iid categorical draws (gargammel.pl:1322-1348, 1421-1431)
static void multinomial(uint64_t n, const std::vector<double>& w, UStream& U, std::vector<uint64_t>& out) {
    out.assign(w.size(), 0);
    double rest = 0; for (size_t i = 0; i < w.size(); i++) rest += w[i];
    uint64_t left = n;
    for (size_t i = 0; i < w.size() && left > 0; i++) {
        if (i + 1 == w.size() || rest <= w[i]) { out[i] = left; left = 0; break; }
        const uint64_t k = binomial_stream(left, w[i] / rest, U);
        out[i] = k; left -= k; rest -= w[i];
    }
}

bool apportion(const ApportionIn& in, ApportionOut& out, std::string& err) {
    if (fabs(in.compB + in.compC + in.compE - 1.0) > 1e-9) {           // gargammel.pl:773 (exact != 1 there; tolerance here)
        char t[64]; snprintf(t, sizeof t, "%g", in.compB + in.compC + in.compE);
        err = std::string("The --comp option must have 3 comma-delimited numbers that sum up to 1, current sum ") + t; return false;
    }
    if (in.n_endo != 1 && in.n_endo != 2 && in.compE > 0) { err = "The endogenous directory must have 1 (haploid) or 2 (diploid) files"; return false; }
    if (in.has_nE) {                                                    // -c: gargammel.pl:1255-1258
        const double sum = in.compE + in.compC + in.compB;
        out.nE = in.nE;
        out.nC = (uint64_t)((double)in.n_fragments * in.compC / sum);
        out.nB = (uint64_t)((double)in.n_fragments * in.compB / sum);
    } else {
        out.nC = (uint64_t)((double)in.n_fragments * in.compC);         // gargammel.pl:1261-1263  int()
        out.nB = (uint64_t)((double)in.n_fragments * in.compB);
        out.nE = (uint64_t)((double)in.n_fragments * in.compE);
    }
    if (out.nC && in.cont_len.empty()) { err = "If you want contamination, please have at least one file in the cont/ directory"; return false; }
    if (out.nB && in.bact_w.empty()) { err = "If you want bacterial contamination, please have at least one file in the bact/ directory"; return false; }
    UStream U(in.seed, 1);
    out.endo.assign(in.n_endo > 0 ? in.n_endo : 1, 0);
    if (in.n_endo == 2) {                                               // gargammel.pl:1272-1279: nE coin flips
        out.endo[0] = binomial_stream(out.nE, 0.5, U); out.endo[1] = out.nE - out.endo[0];
    } else out.endo[0] = out.nE;
    out.cell.clear(); out.cell_e1.clear(); out.cell_e2.clear();
    if (in.n_cells > 0) {
        // multi-cell mode, gargammel.pl:1363-1403: every endogenous fragment picks a cell uniformly; with any endogenous fragments at
        // all cell 0 is given one fragment MORE before the draws ("force the first file to have at least a fragment", :1372-1374), and
        // its first file one more on top before the per-cell coin flips (:1383-1385; overwritten in haploid mode, :1397-1398) -- so the
        // script extracts nE+1 (haploid) or nE+2 (diploid) endogenous fragments.  Mirrored as it is.
        std::vector<double> w((size_t)in.n_cells, 1.0 / (double)in.n_cells);
        multinomial(out.nE, w, U, out.cell);
        out.cell_e1.assign((size_t)in.n_cells, 0); out.cell_e2.assign((size_t)in.n_cells, 0);
        if (out.nE > 0) out.cell[0]++;
        for (int i = 0; i < in.n_cells; i++) {
            if (in.n_endo == 2) { const uint64_t h = binomial_stream(out.cell[i], 0.5, U); out.cell_e1[i] = h + ((i == 0 && out.nE > 0) ? 1 : 0); out.cell_e2[i] = out.cell[i] - h; }
            else out.cell_e1[i] = out.cell[i];
        }
    }
    multinomial(out.nB, in.bact_w, U, out.bact);                        // gargammel.pl:1322-1348 (a miss redraws: weights are renormalised)
    // contaminant: randC = int(rand(sumC)); first j with randC <= cumulative length (gargammel.pl:1421-1431)
    std::vector<double> cw(in.cont_len.size());
    uint64_t sumC = 0; for (size_t j = 0; j < in.cont_len.size(); j++) sumC += in.cont_len[j];
    uint64_t prev_hi = 0, run = 0;   // integers r in [0,sumC) with r <= SL[j] not taken by an earlier j
    for (size_t j = 0; j < in.cont_len.size(); j++) {
        run += in.cont_len[j];
        const uint64_t hi = std::min(run + 1, sumC);   // r in [0, SL[j]] -> SL[j]+1 values, capped by sumC
        cw[j] = (double)(hi - prev_hi); prev_hi = hi;
    }
    multinomial(out.nC, cw, U, out.cont);
    return true;
}

// one BGZF block: gzip member with the 'BC' extra field holding the block size (SAM spec 4.1)
static bool bgzf_block(const uint8_t* in, size_t n, int level, std::vector<uint8_t>& out) {
    out.resize(18 + compressBound((uLong)n) + 8);
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    zs.next_in = (Bytef*)in; zs.avail_in = (uInt)n; zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 18 - 8);
    const int rc = deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    const size_t total = 18 + clen + 8;
    if (total > 65536) return false;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out.data(), hdr, 16);
    out[16] = (uint8_t)((total - 1) & 0xff); out[17] = (uint8_t)((total - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), in, (uInt)n), isz = (uint32_t)n;
    uint8_t* t = out.data() + 18 + clen;
    for (int k = 0; k < 4; k++) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)(isz >> (8 * k)); }
    out.resize(total);
    return true;
}
bool bgzf_compress(const uint8_t* in, size_t n, int threads, int level, std::vector<uint8_t>& out, bool eof_marker, std::string& err) {
    const size_t B = 0xff00;
    const size_t nb = (n + B - 1) / B;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(nb, 1));
    std::vector<std::vector<uint8_t> > blocks(nb);
    std::vector<int> ok(threads, 1);
    auto work = [&](int t) {
        for (size_t b = (size_t)t; b < nb; b += (size_t)threads)
            if (!bgzf_block(in + b * B, std::min(B, n - b * B), level, blocks[b])) { ok[t] = 0; return; }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.push_back(std::thread(work, t));
    work(0);
    for (size_t t = 0; t < th.size(); t++) th[t].join();
    for (int t = 0; t < threads; t++) if (!ok[t]) { err = "bgzf_compress: deflate failed"; return false; }
    size_t tot = 0; for (size_t b = 0; b < nb; b++) tot += blocks[b].size();
    out.clear(); out.reserve(tot + 28);
    for (size_t b = 0; b < nb; b++) out.insert(out.end(), blocks[b].begin(), blocks[b].end());
    if (eof_marker) {
        static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        out.insert(out.end(), eof, eof + 28);
    }
    return true;
}


// ---------------------------------------------------------------- unaligned BAM (SAM/BAM specification section 4.2)
namespace {
void put_u32(std::vector<uint8_t>& o, uint32_t v) { for (int k = 0; k < 4; k++) o.push_back((uint8_t)(v >> (8 * k))); }
void put_u16(std::vector<uint8_t>& o, uint16_t v) { o.push_back((uint8_t)v); o.push_back((uint8_t)(v >> 8)); }
uint32_t get_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint8_t base4(char c) {
    static const char* T = "=ACMGRSVTWYHKDBN";
    if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    for (int k = 0; k < 16; k++) if (T[k] == c) return (uint8_t)k;
    return 15;
}
}
void bam_header_bytes(const BamHeader& h, std::vector<uint8_t>& out) {
    out.push_back('B'); out.push_back('A'); out.push_back('M'); out.push_back(1);
    put_u32(out, (uint32_t)h.text.size()); out.insert(out.end(), h.text.begin(), h.text.end());
    put_u32(out, (uint32_t)h.refs.size());
    for (size_t i = 0; i < h.refs.size(); i++) {
        put_u32(out, (uint32_t)h.refs[i].first.size() + 1); out.insert(out.end(), h.refs[i].first.begin(), h.refs[i].first.end()); out.push_back(0);
        put_u32(out, h.refs[i].second);
    }
}
void bam_append_record(const BamRecord& r, std::vector<uint8_t>& out) {
    const size_t name_n = std::min<size_t>(r.name.size(), 254);            // l_read_name is one byte (incl. the NUL)
    const uint32_t l_seq = (uint32_t)r.seq.size(), l_name = (uint32_t)name_n + 1;
    const uint32_t block = 32 + l_name + 4 * (uint32_t)r.cigar.size() + (l_seq + 1) / 2 + l_seq + (uint32_t)r.aux.size();
    put_u32(out, block); put_u32(out, (uint32_t)r.refid); put_u32(out, (uint32_t)r.pos);
    out.push_back((uint8_t)l_name); out.push_back(r.mapq); put_u16(out, r.bin); put_u16(out, (uint16_t)r.cigar.size()); put_u16(out, r.flag);
    put_u32(out, l_seq); put_u32(out, (uint32_t)r.next_refid); put_u32(out, (uint32_t)r.next_pos); put_u32(out, (uint32_t)r.tlen);
    out.insert(out.end(), r.name.begin(), r.name.begin() + name_n); out.push_back(0);
    for (size_t k = 0; k < r.cigar.size(); k++) put_u32(out, r.cigar[k]);
    for (uint32_t i = 0; i < l_seq; i += 2) out.push_back((uint8_t)((base4(r.seq[i]) << 4) | (i + 1 < l_seq ? base4(r.seq[i + 1]) : 0)));
    if (r.qual.size() == l_seq) out.insert(out.end(), r.qual.begin(), r.qual.end()); else out.insert(out.end(), l_seq, (uint8_t)0xFF);
    out.insert(out.end(), r.aux.begin(), r.aux.end());
}
bool bam_read(const std::string& path, BamHeader& h, std::vector<BamRecord>& recs, std::string& err) {
    std::vector<uint8_t> raw;
    if (!read_file(path, raw, err)) return false;                      // gzread inflates the BGZF members one after the other
    const uint8_t* p = raw.data(); const size_t n = raw.size();
    if (n < 12 || memcmp(p, "BAM\1", 4) != 0) { err = "Could not open input BAM file  " + path; return false; }
    size_t q = 4;
    const uint32_t l_text = get_u32(p + q); q += 4;
    if (q + l_text + 4 > n) { err = "truncated BAM header in " + path; return false; }
    h.text.assign((const char*)p + q, l_text); q += l_text;
    while (!h.text.empty() && h.text[h.text.size() - 1] == '\0') h.text.erase(h.text.size() - 1);
    const uint32_t n_ref = get_u32(p + q); q += 4;
    h.refs.clear();
    for (uint32_t i = 0; i < n_ref; i++) {
        if (q + 4 > n) { err = "truncated BAM header in " + path; return false; }
        const uint32_t l = get_u32(p + q); q += 4;
        if (l == 0 || q + l + 4 > n) { err = "truncated BAM header in " + path; return false; }
        h.refs.push_back(std::make_pair(std::string((const char*)p + q, l - 1), get_u32(p + q + l))); q += l + 4;
    }
    static const char* T = "=ACMGRSVTWYHKDBN";
    while (q + 4 <= n) {
        const uint32_t block = get_u32(p + q); q += 4;
        if (block < 32 || q + block > n) { err = "truncated BAM record in " + path; return false; }
        const uint8_t* b = p + q; q += block;
        BamRecord r;
        r.refid = (int32_t)get_u32(b); r.pos = (int32_t)get_u32(b + 4);
        const uint32_t l_name = b[8]; r.mapq = b[9]; r.bin = (uint16_t)(b[10] | (b[11] << 8));
        const uint32_t n_cig = (uint32_t)(b[12] | (b[13] << 8)); r.flag = (uint16_t)(b[14] | (b[15] << 8));
        const uint32_t l_seq = get_u32(b + 16);
        r.next_refid = (int32_t)get_u32(b + 20); r.next_pos = (int32_t)get_u32(b + 24); r.tlen = (int32_t)get_u32(b + 28);
        const size_t need = 32 + (size_t)l_name + 4 * (size_t)n_cig + ((size_t)l_seq + 1) / 2 + l_seq;
        if (l_name == 0 || need > block) { err = "malformed BAM record in " + path; return false; }
        const uint8_t* c = b + 32;
        r.name.assign((const char*)c, l_name - 1); c += l_name;
        for (uint32_t k = 0; k < n_cig; k++) { r.cigar.push_back(get_u32(c)); c += 4; }
        r.seq.resize(l_seq);
        for (uint32_t i = 0; i < l_seq; i++) r.seq[i] = T[(c[i >> 1] >> ((i & 1) ? 0 : 4)) & 15];
        c += (l_seq + 1) / 2;
        r.qual.assign(c, c + l_seq); c += l_seq;
        r.aux.assign(c, b + block);
        recs.push_back(r);
    }
    return true;
}
void bam_add_program(BamHeader& h, const std::string& id, const std::string& cmdline) {
    if (h.text.empty()) h.text = "@HD\tVN:1.4\tSO:unsorted\n";
    else if (h.text[h.text.size() - 1] != '\n') h.text += "\n";
    std::string use = id;
    for (int k = 1; h.text.find("\tID:" + use + "\t") != std::string::npos || h.text.find("\tID:" + use + "\n") != std::string::npos; k++) use = id + "." + std::to_string(k);
    h.text += "@PG\tID:" + use + "\tPN:" + id + "\tCL:" + cmdline + "\tVN:b200\n";
}

// ---------------------------------------------------------------- Huffman-only deflate code (device BGZF encoder)
namespace {
// code lengths of an optimal prefix code, at most `limit` bits: plain Huffman, and while the tree is too deep the counts are
// halved (floor 1) and the tree rebuilt (the classic zlib-style flattening; costs a fraction of a percent in ratio)
void huff_lengths(const std::vector<uint64_t>& cnt_in, int limit, std::vector<int>& len) {
    const int n = (int)cnt_in.size();
    std::vector<uint64_t> cnt(cnt_in);
    len.assign(n, 0);
    int used = 0, last = -1;
    for (int i = 0; i < n; i++) if (cnt[i]) { used++; last = i; }
    if (used == 0) return;
    if (used == 1) { len[last] = 1; return; }
    for (;;) {
        struct Node { uint64_t w; int l, r; };
        std::vector<Node> nodes;
        typedef std::pair<uint64_t, int> QE;              // (weight, node index); ties broken by index: deterministic
        std::priority_queue<QE, std::vector<QE>, std::greater<QE> > q;
        for (int i = 0; i < n; i++) if (cnt[i]) { Node nd = {cnt[i], -1 - i, 0}; nodes.push_back(nd); q.push(QE(cnt[i], (int)nodes.size() - 1)); }
        while (q.size() > 1) {
            QE a = q.top(); q.pop(); QE b = q.top(); q.pop();
            Node nd = {a.first + b.first, a.second, b.second}; nodes.push_back(nd);
            q.push(QE(nd.w, (int)nodes.size() - 1));
        }
        int maxd = 0;
        std::vector<std::pair<int, int> > st; st.push_back(std::make_pair(q.top().second, 0));
        while (!st.empty()) {
            const int v = st.back().first, d = st.back().second; st.pop_back();
            if (nodes[v].l < 0) { len[-1 - nodes[v].l] = d; if (d > maxd) maxd = d; }
            else { st.push_back(std::make_pair(nodes[v].l, d + 1)); st.push_back(std::make_pair(nodes[v].r, d + 1)); }
        }
        if (maxd <= limit) return;
        for (int i = 0; i < n; i++) if (cnt[i]) cnt[i] = (cnt[i] + 1) / 2;
    }
}
// canonical codes of RFC 1951 3.2.2 from lengths
void canonical_codes(const std::vector<int>& len, int maxbits, std::vector<uint32_t>& code) {
    std::vector<int> bl_count(maxbits + 2, 0);
    for (size_t i = 0; i < len.size(); i++) bl_count[len[i]]++;
    bl_count[0] = 0;
    std::vector<uint32_t> next(maxbits + 2, 0);
    uint32_t c = 0;
    for (int b = 1; b <= maxbits; b++) { c = (c + (uint32_t)bl_count[b - 1]) << 1; next[b] = c; }
    code.assign(len.size(), 0);
    for (size_t i = 0; i < len.size(); i++) if (len[i]) code[i] = next[len[i]]++;
}
uint32_t bitrev(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; } return r; }
struct BitWriter {
    std::vector<uint8_t> b; int nbits; BitWriter() : nbits(0) {}
    void put(uint32_t v, int n) { for (int i = 0; i < n; i++) { if ((nbits & 7) == 0) b.push_back(0); if ((v >> i) & 1u) b[nbits >> 3] |= (uint8_t)(1u << (nbits & 7)); nbits++; } }
};
const uint32_t CRC_POLY = 0xEDB88320u;
}
bool build_deflate_code(const uint64_t hist[256], DeflateCode& out, std::string& err) {
    std::vector<uint64_t> cnt(257);
    uint64_t total = 0;
    for (int i = 0; i < 256; i++) { cnt[i] = hist[i] ? hist[i] : 1; total += cnt[i]; }
    cnt[256] = std::max<uint64_t>(1, total / 65280);                       // one end-of-block per BGZF block
    std::vector<int> len; huff_lengths(cnt, 15, len);
    std::vector<uint32_t> code; canonical_codes(len, 15, code);
    for (int s = 0; s < 257; s++) {
        if (len[s] < 1 || len[s] > 15) { err = "build_deflate_code: bad code length"; return false; }
        out.sym[s] = bitrev(code[s], len[s]) | ((uint32_t)len[s] << 16);
    }
    // code-length sequence: 257 literal/length codes + 1 distance code of length 0; runs of equal lengths use symbol 16 (repeat 3-6)
    std::vector<int> seq(len.begin(), len.end()); seq.push_back(0);
    std::vector<std::pair<int, int> > cl;                                  // (symbol, extra-bit value)
    for (size_t i = 0; i < seq.size();) {
        size_t j = i; while (j < seq.size() && seq[j] == seq[i]) j++;
        size_t run = j - i;
        cl.push_back(std::make_pair(seq[i], 0)); run--;
        while (run >= 3) { const int r = (int)std::min<size_t>(run, 6); cl.push_back(std::make_pair(16, r - 3)); run -= (size_t)r; }
        while (run--) cl.push_back(std::make_pair(seq[i], 0));
        i = j;
    }
    std::vector<uint64_t> clcnt(19, 0);
    for (size_t i = 0; i < cl.size(); i++) clcnt[cl[i].first]++;
    std::vector<int> cllen; huff_lengths(clcnt, 7, cllen);
    std::vector<uint32_t> clcode; canonical_codes(cllen, 7, clcode);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19; while (hclen > 4 && cllen[order[hclen - 1]] == 0) hclen--;
    BitWriter w;
    w.put(1, 1); w.put(2, 2);                                              // BFINAL = 1 (one deflate block per BGZF member), BTYPE = 10 dynamic
    w.put(257 - 257, 5); w.put(1 - 1, 5); w.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) w.put((uint32_t)cllen[order[i]], 3);
    for (size_t i = 0; i < cl.size(); i++) {
        w.put(bitrev(clcode[cl[i].first], cllen[cl[i].first]), cllen[cl[i].first]);
        if (cl[i].first == 16) w.put((uint32_t)cl[i].second, 2);
    }
    out.header = w.b; out.header_nbits = w.nbits;
    return true;
}
void crc32_tables(uint32_t t[4][256]) {
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1; t[0][i] = c; }
    for (uint32_t i = 0; i < 256; i++) for (int s = 1; s < 4; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFFu];
}
// product of two polynomials mod P, reflected representation (bit 31 = x^0)
uint32_t crc32_mul(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 31; i >= 0; i--) {
        if ((a >> i) & 1u) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
uint32_t crc32_x8n(uint64_t n_bytes) {
    uint32_t r = 0x80000000u;                   // x^0
    uint32_t sq = crc32_mul(0x40000000u, 0x40000000u);      // x^2 ...
    sq = crc32_mul(sq, sq); sq = crc32_mul(sq, sq);         // ... x^8
    for (uint64_t n = n_bytes; n; n >>= 1) { if (n & 1u) r = crc32_mul(r, sq); sq = crc32_mul(sq, sq); }
    return r;
}

// ---------------------------------------------------------------- LZ77 + Huffman deflate of the device BGZF encoder: code builder and host model
// (parse specification: gg_lz.h)
bool build_lz_code(const uint64_t hist[gglz::N_SYM], LzCode& out, std::string& err) {
    using namespace gglz;
    std::vector<uint64_t> cl_(N_LITLEN), cd_(N_DIST);
    for (int i = 0; i < N_LITLEN; i++) cl_[i] = hist[i] ? hist[i] : 1;             // every symbol encodable
    for (int i = 0; i < N_DIST; i++) cd_[i] = hist[N_LITLEN + i] ? hist[N_LITLEN + i] : 1;
    std::vector<int> ll, dl; huff_lengths(cl_, 15, ll); huff_lengths(cd_, 15, dl);
    std::vector<uint32_t> lc, dc; canonical_codes(ll, 15, lc); canonical_codes(dl, 15, dc);
    for (int s = 0; s < N_LITLEN; s++) {
        if (ll[s] < 1 || ll[s] > 15) { err = "build_lz_code: bad literal/length code length"; return false; }
        out.sym[s] = bitrev(lc[s], ll[s]) | ((uint32_t)ll[s] << 16);
    }
    for (int s = 0; s < N_DIST; s++) {
        if (dl[s] < 1 || dl[s] > 15) { err = "build_lz_code: bad distance code length"; return false; }
        out.sym[N_LITLEN + s] = bitrev(dc[s], dl[s]) | ((uint32_t)dl[s] << 16);
    }
    std::vector<int> seq(ll.begin(), ll.end()); seq.insert(seq.end(), dl.begin(), dl.end());   // 286 + 30 lengths, one sequence (RFC 1951 3.2.7)
    std::vector<std::pair<int, int> > cl;
    for (size_t i = 0; i < seq.size();) {
        size_t j = i; while (j < seq.size() && seq[j] == seq[i]) j++;
        size_t run = j - i;
        cl.push_back(std::make_pair(seq[i], 0)); run--;
        while (run >= 3) { const int r = (int)std::min<size_t>(run, 6); cl.push_back(std::make_pair(16, r - 3)); run -= (size_t)r; }
        while (run--) cl.push_back(std::make_pair(seq[i], 0));
        i = j;
    }
    std::vector<uint64_t> clcnt(19, 0);
    for (size_t i = 0; i < cl.size(); i++) clcnt[cl[i].first]++;
    std::vector<int> cllen; huff_lengths(clcnt, 7, cllen);
    std::vector<uint32_t> clcode; canonical_codes(cllen, 7, clcode);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19; while (hclen > 4 && cllen[order[hclen - 1]] == 0) hclen--;
    BitWriter w;
    w.put(1, 1); w.put(2, 2);                                              // BFINAL = 1, BTYPE = 10
    w.put((uint32_t)(N_LITLEN - 257), 5); w.put((uint32_t)(N_DIST - 1), 5); w.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) w.put((uint32_t)cllen[order[i]], 3);
    for (size_t i = 0; i < cl.size(); i++) {
        w.put(bitrev(clcode[cl[i].first], cllen[cl[i].first]), cllen[cl[i].first]);
        if (cl[i].first == 16) w.put((uint32_t)cl[i].second, 2);
    }
    out.header = w.b; out.header_nbits = w.nbits;
    return true;
}
namespace {
inline uint32_t ld16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t ld32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
struct LzTok { uint16_t len, dist; uint8_t lit; };       // len == 0: literal
// tokens of run w of a member (gg_lz.h)
void lz_run_tokens(const uint8_t* in, int nb, int w, std::vector<LzTok>& toks) {
    using namespace gglz;
    const int start = w * RUN, end = std::min(nb, start + RUN);
    if (start >= end) return;
    uint16_t T[TABLE]; for (int i = 0; i < TABLE; i++) T[i] = 0xFFFFu;
    if (w > 0) for (int p = start - WARM; p < start; p++) if (p + MIN_MATCH <= nb) T[lz_hash(ld32(in + p), ld16(in + p + 4))] = (uint16_t)p;
    int cover = start;
    for (int base = start; base < end; base += 32) {
        int dist[32]; bool m[32];
        for (int l = 0; l < 32; l++) {
            const int p = base + l; m[l] = false; dist[l] = 0;
            if (p + MIN_MATCH > end) continue;
            const uint16_t c = T[lz_hash(ld32(in + p), ld16(in + p + 4))];
            if (c != 0xFFFFu && p - (int)c <= 32768 && memcmp(in + c, in + p, MIN_MATCH) == 0) { m[l] = true; dist[l] = p - (int)c; }
        }
        for (int l = 0; l < 32; l++) { const int p = base + l; if (p + MIN_MATCH <= end) T[lz_hash(ld32(in + p), ld16(in + p + 4))] = (uint16_t)p; }
        for (int l = 0; l < 32; l++) {
            const int p = base + l;
            if (p >= end) break;
            if (p < cover) continue;
            LzTok t; t.len = 0; t.dist = 0; t.lit = in[p];
            if (m[l]) {
                int len = MIN_MATCH; const int maxlen = std::min(MAX_MATCH, end - p);
                while (len < maxlen && in[p + len] == in[p + len - dist[l]]) len++;
                t.len = (uint16_t)len; t.dist = (uint16_t)dist[l]; cover = p + len;
            }
            toks.push_back(t);
        }
    }
}
void lz_count(const std::vector<LzTok>& toks, uint64_t* hist) {
    using namespace gglz;
    for (size_t i = 0; i < toks.size(); i++) {
        if (!toks[i].len) { hist[toks[i].lit]++; continue; }
        int s, nx; uint32_t xv;
        len_symbol(toks[i].len, ilog2((uint32_t)toks[i].len - 3u), &s, &nx, &xv); hist[s]++;
        dist_symbol(toks[i].dist, ilog2((uint32_t)toks[i].dist - 1u), &s, &nx, &xv); hist[N_LITLEN + s]++;
    }
}
}
bool bgzf_lz_model(const uint8_t* in, size_t n, std::vector<uint8_t>& out, bool eof_marker, std::string& err, uint64_t* hist_out) {
    using namespace gglz;
    const size_t B = 65280, n_blocks = (n + B - 1) / B;
    const size_t stride = std::max<size_t>(1, (n_blocks + SAMPLE_MEMBERS - 1) / SAMPLE_MEMBERS);
    uint64_t hist[N_SYM]; memset(hist, 0, sizeof hist);
    std::vector<LzTok> toks;
    for (size_t b = 0; b < n_blocks; b += stride) {
        const int nb = (int)std::min(B, n - b * B);
        for (int w = 0; w < 30; w++) { toks.clear(); lz_run_tokens(in + b * B, nb, w, toks); lz_count(toks, hist); }
        hist[256]++;
    }
    if (hist_out) memcpy(hist_out, hist, sizeof hist);
    LzCode code;
    if (!build_lz_code(hist, code, err)) return false;
    out.clear();
    for (size_t b = 0; b < n_blocks; b++) {
        const uint8_t* src = in + b * B; const int nb = (int)std::min(B, n - b * B);
        BitWriter w; bool overflow = false;
        for (size_t i = 0; i < code.header.size(); i++) w.b.push_back(code.header[i]);
        w.nbits = code.header_nbits;
        for (int r = 0; r < 30; r++) {
            toks.clear(); lz_run_tokens(src, nb, r, toks);
            const int bits0 = w.nbits;
            for (size_t i = 0; i < toks.size(); i++) {
                if (!toks[i].len) { const uint32_t s = code.sym[toks[i].lit]; w.put(s & 0xFFFFu, (int)(s >> 16)); continue; }
                int s, nx; uint32_t xv;
                len_symbol(toks[i].len, ilog2((uint32_t)toks[i].len - 3u), &s, &nx, &xv);
                w.put(code.sym[s] & 0xFFFFu, (int)(code.sym[s] >> 16)); w.put(xv, nx);
                dist_symbol(toks[i].dist, ilog2((uint32_t)toks[i].dist - 1u), &s, &nx, &xv);
                w.put(code.sym[N_LITLEN + s] & 0xFFFFu, (int)(code.sym[N_LITLEN + s] >> 16)); w.put(xv, nx);
            }
            if (w.nbits - bits0 > REGION * 8) overflow = true;       // the kernel stages a run's bits in REGION bytes
        }
        w.put(code.sym[256] & 0xFFFFu, (int)(code.sym[256] >> 16));
        size_t payload = ((size_t)w.nbits + 7) / 8;
        const bool stored = overflow || payload >= (size_t)nb + 5;
        if (stored) payload = (size_t)nb + 5;
        const size_t total = 18 + payload + 8;
        const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)((total - 1) & 0xFF), (uint8_t)((total - 1) >> 8)};
        out.insert(out.end(), hdr, hdr + 18);
        if (stored) {
            const uint8_t sh[5] = {1, (uint8_t)(nb & 0xFF), (uint8_t)(nb >> 8), (uint8_t)(~nb & 0xFF), (uint8_t)((~nb >> 8) & 0xFF)};
            out.insert(out.end(), sh, sh + 5); out.insert(out.end(), src, src + nb);
        } else { w.b.resize(payload, 0); out.insert(out.end(), w.b.begin(), w.b.end()); }
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)nb);
        for (int k = 0; k < 4; k++) out.push_back((uint8_t)(crc >> (8 * k)));
        for (int k = 0; k < 4; k++) out.push_back((uint8_t)((uint32_t)nb >> (8 * k)));
    }
    if (eof_marker) {
        static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        out.insert(out.end(), eof, eof + 28);
    }
    return true;
}

bool load_genome(const std::string& fasta_path, GenomeFile& g, std::string& err) {
    g.path = fasta_path;
    if (!read_file(fasta_path, g.bytes, err)) return false;            // .gz is inflated in memory (the reference writes a temp file, fragSim.cpp:1108-1147)
    std::string e2;
    if (!parse_fai(fasta_path + ".fai", g.fai, e2)) {
        const bool gz = fasta_path.size() > 3 && fasta_path.compare(fasta_path.size() - 3, 3, ".gz") == 0;
        if (gz) { err = e2; return false; }
        if (!build_fai(g.bytes, g.fai, err)) return false;             // samtools faidx equivalent (gargammel.pl:900-903)
    }
    g.total_len = 0;
    for (size_t i = 0; i < g.fai.size(); i++) g.total_len += (uint64_t)g.fai[i].len;
    return true;
}
int upload_genome(gg_ctx* ctx, const GenomeFile& g, int* id_out, int circ_chr) {
    std::vector<gg_fai_entry> f(g.fai.size());
    for (size_t i = 0; i < f.size(); i++) { f[i].name = g.fai[i].name.c_str(); f[i].len = g.fai[i].len; f[i].offset = g.fai[i].offset; f[i].line_blen = g.fai[i].line_blen; f[i].line_len = g.fai[i].line_len; }
    return gg_genome_upload_circ(ctx, g.bytes.data(), g.bytes.size(), f.data(), (int)f.size(), circ_chr, id_out);
}

} // namespace ggh

// ---------------------------------------------------------------- C exports (Python harness)
static int set_err(char* err, int cap, const std::string& m) { if (err && cap > 0) { snprintf(err, cap, "%s", m.c_str()); } return -1; }
extern "C" {
int ggh_load_size_freq(const char* path, int32_t* len, double* cum, int cap, int* n, char* err, int errcap) {
    std::vector<int32_t> l; std::vector<double> c; std::string e;
    if (!ggh::load_size_freq(path, l, c, e)) return set_err(err, errcap, e);
    *n = (int)l.size(); if ((int)l.size() > cap) return set_err(err, errcap, "buffer too small");
    memcpy(len, l.data(), l.size() * 4); memcpy(cum, c.data(), c.size() * 8); return 0;
}
int ggh_load_size_list(const char* path, int32_t* len, int cap, int* n, char* err, int errcap) {
    std::vector<int32_t> l; std::string e;
    if (!ggh::load_size_list(path, l, e)) return set_err(err, errcap, e);
    *n = (int)l.size(); if ((int)l.size() > cap) return set_err(err, errcap, "buffer too small");
    memcpy(len, l.data(), l.size() * 4); return 0;
}
int ggh_load_rates(const char* path, int kind, double* sub5, int cap5, int* rows5, double* sub3, int cap3, int* rows3, char* err, int errcap) {
    ggh::Rates a, b; std::string e; bool ok;
    if (kind == 0) ok = ggh::load_matrix_dat(path, a, b, e); else if (kind == 1) ok = ggh::load_profile(path, a, b, e); else ok = ggh::load_mapdamage(path, a, b, e);
    if (!ok) return set_err(err, errcap, e);
    *rows5 = (int)(a.size() / 12); *rows3 = (int)(b.size() / 12);
    if (*rows5 > cap5 || *rows3 > cap3) return set_err(err, errcap, "buffer too small");
    memcpy(sub5, a.data(), a.size() * 8); memcpy(sub3, b.data(), b.size() * 8); return 0;
}
int ggh_fai_load(const char* path, int from_fasta, void** handle, int* n, char* err, int errcap) {
    std::vector<ggh::FaiEntry>* v = new std::vector<ggh::FaiEntry>(); std::string e; bool ok;
    if (from_fasta) { std::vector<uint8_t> b; ok = ggh::read_file(path, b, e) && ggh::build_fai(b, *v, e); } else ok = ggh::parse_fai(path, *v, e);
    if (!ok) { delete v; return set_err(err, errcap, e); }
    *handle = v; *n = (int)v->size(); return 0;
}
int ggh_fai_get(void* handle, int i, char* name, int namecap, int64_t* len, uint64_t* offset, int32_t* blen, int32_t* llen) {
    std::vector<ggh::FaiEntry>* v = (std::vector<ggh::FaiEntry>*)handle;
    if (!v || i < 0 || i >= (int)v->size()) return -1;
    snprintf(name, namecap, "%s", (*v)[i].name.c_str()); *len = (*v)[i].len; *offset = (*v)[i].offset; *blen = (*v)[i].line_blen; *llen = (*v)[i].line_len; return 0;
}
void ggh_fai_free(void* handle) { delete (std::vector<ggh::FaiEntry>*)handle; }
int ggh_apportion(uint64_t n, double compB, double compC, double compE, int n_endo, const uint64_t* cont_len, int n_cont,
                  const double* bact_w, int n_bact, uint64_t seed, uint64_t* totals3, uint64_t* endo, uint64_t* cont, uint64_t* bact,
                  char* err, int errcap) {
    ggh::ApportionIn in; in.n_fragments = n; in.compB = compB; in.compC = compC; in.compE = compE; in.n_endo = n_endo; in.seed = seed;
    in.cont_len.assign(cont_len, cont_len + n_cont); in.bact_w.assign(bact_w, bact_w + n_bact);
    ggh::ApportionOut out; std::string e;
    if (!ggh::apportion(in, out, e)) return set_err(err, errcap, e);
    totals3[0] = out.nE; totals3[1] = out.nC; totals3[2] = out.nB;
    for (size_t i = 0; i < out.endo.size(); i++) endo[i] = out.endo[i];
    for (size_t i = 0; i < out.cont.size(); i++) cont[i] = out.cont[i];
    for (size_t i = 0; i < out.bact.size(); i++) bact[i] = out.bact[i];
    return 0;
}
uint64_t ggh_binomial(uint64_t n, double p, uint64_t seed, uint64_t stream) { return ggh::binomial_draw(n, p, seed, stream); }
int ggh_bgzf_compress(const uint8_t* in, size_t n, int threads, int level, int eof_marker, uint8_t* out, size_t cap, size_t* out_n) {
    std::vector<uint8_t> o; std::string e;
    if (!ggh::bgzf_compress(in, n, threads, level, o, eof_marker != 0, e)) return -1;
    *out_n = o.size();
    if (o.size() > cap) return -2;
    memcpy(out, o.data(), o.size()); return 0;
}
// test hook: the deflate code + header for a histogram, and a reference encoding of `in` with it (one BGZF member per 65280 bytes)
int ggh_deflate_code(const uint64_t* hist, uint32_t* sym257, uint8_t* hdr, int hdr_cap, int* hdr_nbits) {
    ggh::DeflateCode c; std::string e;
    if (!ggh::build_deflate_code(hist, c, e)) return -1;
    memcpy(sym257, c.sym, sizeof c.sym);
    if ((int)c.header.size() > hdr_cap) return -2;
    memcpy(hdr, c.header.data(), c.header.size()); *hdr_nbits = c.header_nbits; return 0;
}
// test hook: the host model of the device LZ encoder (gg_lz.h): BGZF members of `in`, and the sampled symbol counts (316) if hist != NULL
int ggh_bgzf_lz_model(const uint8_t* in, size_t n, int eof_marker, uint8_t* out, size_t cap, size_t* out_n, uint64_t* hist) {
    std::vector<uint8_t> o; std::string e;
    if (!ggh::bgzf_lz_model(in, n, o, eof_marker != 0, e, hist)) return -1;
    *out_n = o.size();
    if (o.size() > cap) return -2;
    memcpy(out, o.data(), o.size()); return 0;
}
uint32_t ggh_crc32_x8n(uint64_t n) { return ggh::crc32_x8n(n); }
uint32_t ggh_crc32_mul(uint32_t a, uint32_t b) { return ggh::crc32_mul(a, b); }
int ggh_load_dnacomp(const char* path, int dist, double* s5p, double* s5m, double* s3p, double* s3m, char* err, int errcap) {
    ggh::CompTables t; std::string e;
    if (!ggh::load_dnacomp(path, dist, t, e)) return set_err(err, errcap, e);
    const size_t n = (size_t)2 * dist * 4 * 8;
    memcpy(s5p, t.s5p.data(), n); memcpy(s5m, t.s5m.data(), n); memcpy(s3p, t.s3p.data(), n); memcpy(s3m, t.s3m.data(), n); return 0;
}
int ggh_bam_from_fasta(const uint8_t* recs, size_t n, const char* path, char* err, int errcap) {
    std::vector<uint8_t> bam; ggh::BamHeader h; ggh::bam_add_program(h, "test", "ggh_bam_from_fasta");
    ggh::bam_header_bytes(h, bam);
    size_t p = 0;
    while (p < n) {
        const uint8_t* e = (const uint8_t*)memchr(recs + p, '\n', n - p); const size_t ie = e ? (size_t)(e - recs) : n;
        const size_t ss = std::min(ie + 1, n);
        const uint8_t* e2 = (const uint8_t*)memchr(recs + ss, '\n', n - ss); const size_t se = e2 ? (size_t)(e2 - recs) : n;
        ggh::BamRecord r; r.name.assign((const char*)recs + p + (recs[p] == '>' ? 1 : 0), ie - p - (recs[p] == '>' ? 1 : 0)); r.seq.assign((const char*)recs + ss, se - ss);
        r.qual.assign(r.seq.size(), 0);
        ggh::bam_append_record(r, bam);
        p = se + 1;
    }
    std::vector<uint8_t> z; std::string e;
    if (!ggh::bgzf_compress(bam.data(), bam.size(), 1, 0, z, true, e)) { snprintf(err, errcap, "%s", e.c_str()); return 1; }
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(z.data(), 1, z.size(), f) != z.size()) { if (f) fclose(f); snprintf(err, errcap, "cannot write %s", path); return 1; }
    fclose(f);
    return 0;
}
int ggh_bam_to_fasta(const char* path, uint8_t* out, size_t cap, size_t* out_n, char* err, int errcap) {
    ggh::BamHeader h; std::vector<ggh::BamRecord> recs; std::string e;
    if (!ggh::bam_read(path, h, recs, e)) { snprintf(err, errcap, "%s", e.c_str()); return 1; }
    std::string s;
    for (size_t i = 0; i < recs.size(); i++) s += ">" + recs[i].name + "\n" + recs[i].seq + "\n";
    *out_n = s.size();
    if (s.size() > cap) { snprintf(err, errcap, "buffer too small"); return 1; }
    memcpy(out, s.data(), s.size());
    return 0;
}
int ggh_lognormal_table(double loc, double scale, int cap, int32_t* len, double* cum) {
    std::vector<int32_t> l; std::vector<double> c; ggh::lognormal_table(loc, scale, cap, l, c);
    memcpy(len, l.data(), l.size() * 4); memcpy(cum, c.data(), c.size() * 8); return (int)l.size();
}
}
