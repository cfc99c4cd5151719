// The whole-file reader (read_records) and the streaming reader (record_chunk_reader) of deamSim / adptSim normalise any FASTA record
// stream the same way (or both reject it); chunk_records tiles the result.  Built with ASan/UBSan by tests/test_host_fuzz.py.
// usage: fuzz_records ITERS SEED WORKFILE
#include "../../gargammel_b200/csrc/cli_common.h"
#include <random>
#include <fstream>
using namespace cli;
static void spit(const std::string& p, const std::string& b) { std::ofstream f(p, std::ios::binary); f.write(b.data(), b.size()); }
int main(int argc, char** argv) {
    std::mt19937_64 r(atoi(argv[2])); const int iters = atoi(argv[1]); const std::string path = argv[3]; long same = 0, both_err = 0, diff = 0;
    for (int it = 0; it < iters; it++) {
        std::string s; const int nrec = r() % 6;
        for (int k = 0; k < nrec; k++) {
            if (r() % 10) s += ">"; s += "id" + std::to_string(r() % 100);
            if (r() % 8 == 0) s += "\r";
            s += "\n";
            const int nl = r() % 4;
            for (int l = 0; l < nl; l++) { const int len = r() % 9; for (int q = 0; q < len; q++) s += "ACGTNacgt>"[r() % (r() % 20 ? 9 : 10)]; if (r() % 9 == 0) s += "\r"; if (l + 1 < nl || r() % 5) s += "\n"; }
            if (r() % 7 == 0) s += "\n";
        }
        spit(path, s);
        std::vector<uint8_t> a; uint64_t na = 0; std::string ea; const bool oka = read_records(path, a, &na, ea);
        BoundedQueue<InChunk> q(1000); std::string eb; uint64_t nb = 0;
        record_chunk_reader(path, 1 + r() % 40, &q, &eb, &nb);
        std::vector<uint8_t> b; InChunk c; uint64_t recs = 0; while (q.pop(c)) { b.insert(b.end(), c.bytes.begin(), c.bytes.end()); recs += c.recs; }
        const bool okb = eb.empty();
        if (!oka && !okb) both_err++;
        else if (oka && okb && a == b && na == nb && recs == nb) { same++; std::vector<Chunk> ch = chunk_records(a, 1 + r() % 50); size_t t = 0; uint64_t rr = 0; for (auto& x : ch) { if (x.off != t) abort(); t += x.bytes; rr += x.recs; } if (t != a.size() || rr != na) { printf("chunk_records mismatch\n"); return 1; } }
        else { diff++; if (diff < 4) printf("DIFF oka=%d okb=%d na=%llu nb=%llu\nIN<<%s>>\nA<<%.*s>>\nB<<%.*s>>\n", oka, okb, (unsigned long long)na, (unsigned long long)nb, s.c_str(), (int)a.size(), (const char*)a.data(), (int)b.size(), (const char*)b.data()); }
    }
    printf("same %ld, both rejected %ld, different %ld\n", same, both_err, diff);
    return diff ? 1 : 0;
}
