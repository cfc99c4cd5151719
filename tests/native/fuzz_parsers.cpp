// Mutates valid inputs and feeds every host parser of the drop-in executables (they face user files); built with ASan/UBSan by
// tests/test_host_fuzz.py.  usage: fuzz_parsers WORKDIR ITERS SEED (NAME SEEDFILE)...
#include "../../gargammel_b200/csrc/gg_host.h"
#include <random>
#include <unistd.h>
#include <signal.h>
#include <fstream>
#include <iostream>
#include <cstring>
using namespace ggh;
static std::vector<uint8_t> slurp(const char* p) { std::ifstream f(p, std::ios::binary); return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
static void spit(const std::string& p, const std::vector<uint8_t>& b) { std::ofstream f(p, std::ios::binary); f.write((const char*)b.data(), b.size()); }
static std::vector<uint8_t> mutate(std::vector<uint8_t> b, std::mt19937_64& r) {
    int k = 1 + r() % 6;
    while (k--) {
        if (b.empty()) { b.push_back((uint8_t)r()); continue; }
        switch (r() % 7) {
        case 0: b[r() % b.size()] = (uint8_t)r(); break;
        case 1: b.erase(b.begin() + r() % b.size()); break;
        case 2: b.insert(b.begin() + r() % b.size(), (uint8_t)"\n\t 0-9.eE+\r>"[r() % 12]); break;
        case 3: b.resize(r() % b.size()); break;
        case 4: { size_t a = r() % b.size(), n = std::min<size_t>(r() % 64, b.size() - a); b.insert(b.begin() + r() % b.size(), b.begin() + a, b.begin() + a + n); } break;
        case 5: { size_t a = r() % b.size(); const char* t[] = {"nan", "-1", "1e308", "99999999999999999999", "", "\t\t", "0"}; const char* s = t[r() % 7]; b.insert(b.begin() + a, s, s + strlen(s)); } break;
        case 6: { size_t a = r() % b.size(), n = std::min<size_t>(r() % 32, b.size() - a); b.erase(b.begin() + a, b.begin() + a + n); } break;
        }
    }
    return b;
}
int main(int argc, char** argv) {
    const std::string dir = argv[1]; const int iters = atoi(argv[2]);
    std::mt19937_64 r(atoi(argv[3]));
    struct Seed { std::string name; std::vector<uint8_t> data; };
    std::vector<Seed> seeds;
    for (int i = 4; i + 1 < argc; i += 2) { Seed s; s.name = argv[i]; s.data = slurp(argv[i + 1]); seeds.push_back(s); }
    long ok = 0, bad = 0;
    for (int it = 0; it < iters; it++) {
        const Seed& s = seeds[r() % seeds.size()];
        std::vector<uint8_t> m = mutate(s.data, r);
        std::string err; alarm(20);
        bool res = false;
        if (s.name == "fai") { spit(dir + "/x.fai", m); std::vector<FaiEntry> o; res = parse_fai(dir + "/x.fai", o, err); }
        else if (s.name == "fasta") { std::vector<FaiEntry> o; res = build_fai(m, o, err); if (res) { spit(dir + "/g.fa", m); remove((dir + "/g.fa.fai").c_str()); GenomeFile g; load_genome(dir + "/g.fa", g, err); } }
        else if (s.name == "sizelist") { spit(dir + "/x.sl", m); std::vector<int32_t> l; res = load_size_list(dir + "/x.sl", l, err); }
        else if (s.name == "sizefreq") { spit(dir + "/x.sf", m); std::vector<int32_t> l; std::vector<double> c; res = load_size_freq(dir + "/x.sf", l, c, err); }
        else if (s.name == "dnacomp") { spit(dir + "/x.dc", m); CompTables c; res = load_dnacomp(dir + "/x.dc", 1 + (int)(r() % 5), c, err); }
        else if (s.name == "dat5") { spit(dir + "/m-5.dat", m); Rates a, b; res = load_matrix_dat(dir + "/m-", a, b, err); }
        else if (s.name == "dat3") { spit(dir + "/m-3.dat", m); Rates a, b; res = load_matrix_dat(dir + "/m-", a, b, err); }
        else if (s.name == "prof5") { spit(dir + "/p_5p.prof", m); Rates a, b; res = load_profile(dir + "/p_", a, b, err); }
        else if (s.name == "mapd") { spit(dir + "/x.mis", m); Rates a, b; res = load_mapdamage(dir + "/x.mis", a, b, err); }
        else if (s.name == "list") { spit(dir + "/list", m); std::vector<std::string> n; std::vector<double> w; res = load_bact_list(dir + "/list", n, w, err); }
        else if (s.name == "bam") { spit(dir + "/x.bam", m); BamHeader h; std::vector<BamRecord> rr; res = bam_read(dir + "/x.bam", h, rr, err); }
        (res ? ok : bad)++;
    }
    alarm(0);
    printf("done: %ld accepted, %ld rejected\n", ok, bad);
    return 0;
}
