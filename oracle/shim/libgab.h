/*
 * libgab.h -- SHIM written for this repo (NOT the upstream grenaud/libgab, which is fetched by
 * `git clone` at the reference's build time, Makefile:24, and is absent here).  It provides exactly
 * the helper symbols that the UNMODIFIED reference sources src/{fragSim,deamSim,adptSim}.cpp use, with
 * semantics inferred from their call sites.  Binaries built with it are labelled
 * "reference sources + our shim".  Test/baseline infrastructure only.
 */
#ifndef SHIM_LIBGAB_H
#define SHIM_LIBGAB_H
#include <string>
#include <vector>
#include <map>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <functional>
#include <stdint.h>
#include <sys/time.h>
#include <unistd.h>
#include <limits.h>

using namespace std;

template <typename T> inline string stringify(const T& v) { ostringstream o; o << v; return o.str(); }
template <typename T> inline T destringify(const string& s) {
    istringstream i(s); T v;
    if (!(i >> v)) { cerr << "destringify: cannot convert \"" << s << "\"" << endl; exit(1); }
    return v;
}
inline string booleanAsString(bool b) { return b ? "currently turned on/used" : "not on/not used"; }
inline vector<string> allTokens(const string& s, char delim) {
    vector<string> out; string cur;
    for (size_t i = 0; i < s.size(); i++) { if (s[i] == delim) { out.push_back(cur); cur.clear(); } else cur += s[i]; }
    out.push_back(cur);
    return out;
}
inline bool strEndsWith(const string& s, const string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }
inline bool strBeginsWith(const string& s, const string& pre) { return s.size() >= pre.size() && s.compare(0, pre.size(), pre) == 0; }

/* uniform helpers: "self-seed from the clock unless the caller already seeded" (bool = !seedSpecified) */
inline void shim_seed_once(bool fromTime) {
    static bool done = false;
    if (fromTime && !done) { timeval t; gettimeofday(&t, NULL); srand((unsigned)(t.tv_sec * 1000 + t.tv_usec / 1000)); }
    done = true;
}
inline double randomProb(bool fromTime = true) { shim_seed_once(fromTime); return double(rand()) / double(RAND_MAX); }
inline int randomInt(int lo, int hi, bool fromTime = true) { shim_seed_once(fromTime); return lo + int(rand() % (unsigned)(hi - lo + 1)); }
inline bool randomBool(bool fromTime = true) { shim_seed_once(fromTime); return (rand() & 0x10000) != 0; }
inline string randomDNASeq(int n) { string s(n, 'A'); for (int i = 0; i < n; i++) s[i] = "ACGT"[(rand() >> 8) & 3]; return s; }

/* lower-case acgt count as resolved: fragSim --case (the documented --methyl mode, README "methylated and unmethylated") passes
 * lower-case bases through isResolvedDNAstring (fragSim.cpp:333-339,1472) and could not emit a single methylated base otherwise */
inline bool isResolvedDNA(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't'; }
inline int baseResolved2int(char c) {
    if (c == 'A') return 0; if (c == 'C') return 1; if (c == 'G') return 2; if (c == 'T') return 3;
    cerr << "baseResolved2int: invalid base " << c << endl; exit(1);
}
inline char complement(char c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return 'N';
    }
}
inline string reverseComplement(const string& s) { string r(s.size(), 'N'); for (size_t i = 0; i < s.size(); i++) r[i] = complement(s[s.size() - 1 - i]); return r; }

/* mapDamage2prof.cpp:337 (-h format only): the row number right-aligned in a field of `width` characters (inferred from the name) */
inline string printIntAsWhitePaddedString(int i, int width) { string t = stringify(i); while ((int)t.size() < width) t = " " + t; return t; }
template <typename T> inline string thousandSeparator(const T v) {
    string d = stringify(v), out; int n = 0;
    for (int i = int(d.size()) - 1; i >= 0; i--) { out += d[i]; if (++n % 3 == 0 && i > 0 && isdigit(d[i - 1])) out += ','; }
    reverse(out.begin(), out.end()); return out;
}
template <typename T> inline string vectorToString(const vector<T>& v, const string& sep = ",") {
    string out; for (size_t i = 0; i < v.size(); i++) { if (i) out += sep; out += stringify(v[i]); } return out;
}
template <typename T> inline string arrayToString(const T* a, int n, const string& sep = ",") {
    string out; for (int i = 0; i < n; i++) { if (i) out += sep; out += stringify(a[i]); } return out;
}
/* directory of the executable, with trailing slash (deamSim.cpp:1105 appends "matrices/...") */
inline string getCWD(const char* argv0) {
    char buf[PATH_MAX]; string p;
    ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n > 0) { buf[n] = 0; p = buf; } else p = argv0;
    size_t k = p.find_last_of('/');
    return k == string::npos ? string("./") : p.substr(0, k + 1);
}
inline string returnGitHubVersion(const string&, const string&) { return "shim"; }
#endif
