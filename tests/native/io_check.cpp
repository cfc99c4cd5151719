// Host-side I/O helpers of the drop-in executables, checked without a GPU (tests/test_host_io.py compiles and runs this):
//   io_check sink FILE   Sink::write / Sink::write_big interleaved into one plain file == the concatenation of the pieces
//   io_check gzip FILE   gzip_file(FILE): FILE.gz appears (BGZF members + EOF marker), FILE is removed
#define main gg_fused_driver_main
#include "../../gargammel_b200/csrc/cli_gargammel.cpp"
#undef main
#include <random>

static int check_sink(const char* path) {
    std::mt19937_64 rng(7); std::vector<uint8_t> all;
    cli::Sink s; if (!s.open_plain(path)) return 2;
    const size_t sizes[] = {10, 0, 1, 5u << 20, 3, 40u << 20, 1u << 20, 33u << 20, 7, 2};
    for (size_t k = 0; k < sizeof sizes / sizeof sizes[0]; k++) {
        std::vector<uint8_t> b(sizes[k]);
        for (size_t i = 0; i < b.size(); i += 8) { const uint64_t v = rng(); memcpy(&b[i], &v, std::min<size_t>(8, b.size() - i)); }
        if (!((k % 3 == 2) ? s.write(b.data(), b.size()) : s.write_big(b.data(), b.size()))) return 3;
        all.insert(all.end(), b.begin(), b.end());
    }
    if (!s.close()) return 4;
    std::vector<uint8_t> back; std::string e;
    if (!ggh::read_file(path, back, e)) return 5;
    printf("%zu bytes %s\n", all.size(), all == back ? "SAME" : "DIFFERENT");
    return all == back ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc != 3) return 64;
    if (!strcmp(argv[1], "sink")) return check_sink(argv[2]);
    if (!strcmp(argv[1], "gzip")) return gzip_file(argv[2], false) ? 0 : 1;
    return 64;
}
