// bin/fragSim -- drop-in for the reference's src/fragSim (fragSim.cpp): same flags, last argv is the FASTA, same
// record grammar on stdout / -o gz, same final stderr line.  The sampling runs on the GPU through the C-ABI.
#include "cli_common.h"
using namespace cli;

int main(int argc, char** argv) {
    int sizeFragments = 20; uint64_t nFragments = 10; int minLen = 0, maxLen = 1000;
    bool noRev = false, seedSpecified = false, specifiedLength = false, specifiedLoc = false, specifiedScale = false; uint64_t seed = 0;
    double location = -1, scale = -1;
    bool uniqTags = false, tagb = false, gcBiasB = false; double gcBias = 0;
    std::string compFile; int distFromEnd = 1;                                      // fragSim.cpp:356-359
    std::string sizeList, sizeFreq, outGz, tag;
    std::string outBAM; bool uncompressedBAM = false;                                // -b / -u, fragSim.cpp:384-385,398
    bool keepCase = false;                                                          // fragSim.cpp:387 (uppercase = true)
    bool circ = false; std::string circName;                                        // fragSim.cpp:388-389
    const std::string usage = std::string("\n This program takes a fasta file representing a chromosome and generates\n") +
        " aDNA fragments according to a certain distribution (B200 build)\n\n " + argv[0] + " [options]  [chromosome fasta] \n\n" +
        "\t\t-n\t[number]\t\t\tGenerate [number] fragments (default: 10)\n\t\t--norev\t\t\t\t\tDo not reverse complement (default: rev. comp half of seqs.)\n" +
        "\t\t--seed\t[int]\t\t\t\tUse [seed] as seed for the random number generator (default random seed each execution)\n" +
        "\t\t-o\t[fasta out]\t\t\tWrite output as a zipped fasta (default: fasta in STDOUT)\n\t\t-tag\t[tag]\t\t\t\tAppend this string to deflines\n" +
        "\t\t-m\t[length]\t\t\tMinimum fragments length (default: 0)\n\t\t-M\t[length]\t\t\tMaximum fragments length (default: 1000)\n" +
        "\t\t-l\t[length]\t\t\tGenerate fragments of fixed length  (default: 20)\n\t\t-s\t[file]\t\t\t\tOpen file with size distribution (one fragment length per line)\n" +
        "\t\t-f\t[file]\t\t\t\tOpen file with size frequency: length[TAB]freq\n" +
        "\t\t--loc\t[float] --scale\t[float]\tlog-normal fragment lengths\n" +
        "\t\t--comp\t[file] --dist\t[n]\t\tbase composition at the fragment ends (mapDamage dnacomp.txt)\n" +
        "\t\t-uniq\t\t\t\t\tMake the fragment names unique by appending a counter (host-side pass, like the reference)\n" +
        "\t\t--circ\t[REF NAME]\t\t\tAssume [REF NAME] is circular\n" +
        "\t\t-gc\t[gc bias]\t\t\tUse GC bias factor\n\t\t--case\t\t\t\t\tDo not set the sequence to upper-case (default: uppercase the seqs.)\n\t\t-b\t[bam out]\t\t\tWrite output as an unaligned BAM file\n\t\t-u\t\t\t\t\tProduce uncompressed BAM (good for UNIX pipe)\n\tNot in this build: --fq --trim5p\n";
    if (is_help(argc, argv)) { std::cout << usage << std::endl; return 1; }        // fragSim.cpp:463-469
    for (int i = 1; i < argc - 1; i++) {                                            // fragSim.cpp:471-625
        const std::string a = argv[i];
        if (a == "-n") { nFragments = (uint64_t)to_ll(argv[++i], "-n"); continue; }
        if (a == "-l") { sizeFragments = (int)to_ll(argv[++i], "-l"); specifiedLength = true; continue; }
        if (a == "-s") { sizeList = argv[++i]; continue; }
        if (a == "-f") { sizeFreq = argv[++i]; continue; }
        if (a == "-m") { minLen = (int)to_ll(argv[++i], "-m"); continue; }
        if (a == "-M") { maxLen = (int)to_ll(argv[++i], "-M"); continue; }
        if (a == "--norev") { noRev = true; continue; }
        if (a == "--seed") { seed = (uint64_t)to_ll(argv[++i], "--seed"); seedSpecified = true; continue; }
        if (a == "-o") { outGz = argv[++i]; continue; }
        if (a == "-tag") { tag = argv[++i]; tagb = true; continue; }
        if (a == "-uniq") { uniqTags = true; continue; }                              // fragSim.cpp:481-484
        if (a == "-tmp") { ++i; continue; }                                         // gz inputs are inflated in memory
        if (a == "--loc") { location = to_d(argv[++i], "--loc"); specifiedLoc = true; continue; }
        if (a == "--scale") { scale = to_d(argv[++i], "--scale"); specifiedScale = true; continue; }
        if (a == "--comp") { compFile = argv[++i]; continue; }
        if (a == "--dist") { distFromEnd = (int)to_ll(argv[++i], "--dist"); continue; }
        if (a == "-gc") { gcBias = to_d(argv[++i], "-gc"); gcBiasB = true; continue; }      // fragSim.cpp:589-594
        if (a == "--circ") { circ = true; circName = argv[++i]; continue; }             // fragSim.cpp:553-557
        if (a == "-b") { outBAM = argv[++i]; continue; }                               // fragSim.cpp:531-536
        if (a == "-u") { uncompressedBAM = true; continue; }                          // fragSim.cpp:538-541
        if (a == "--trim5p")
            return unsupported(a, "SURVEY.md 8f-3");
        if (a == "--case") { keepCase = true; continue; }                              // fragSim.cpp:548-551
        if (a == "--fq") return unsupported(a, "SURVEY.md 8f-3");
        return die("Error: unknown option " + a);                                   // fragSim.cpp:622-623
    }
    if (specifiedLength && sizeFragments > maxLen)                                   // fragSim.cpp:638-643
        return die("Error: specified size " + std::to_string(sizeFragments) + " is longer than the maximum allowed fragment size " + std::to_string(maxLen) + ", use option -M to change this.");
    if (specifiedLoc && !specifiedScale) return die("Error: cannot specify --loc without --scale");      // fragSim.cpp:645-653
    if (specifiedScale && !specifiedLoc) return die("Error: cannot specify --scale without --loc");
    if (specifiedScale && specifiedLoc) {                                                                // :655-669
        if (specifiedLength) return die("Error: cannot specify --scale and --loc with -l");
        if (!sizeList.empty()) return die("Error: cannot specify --scale and --loc with -s");
        if (!sizeFreq.empty()) return die("Error: cannot specify --scale and --loc with -f");
        if (!(scale > 0)) return die("Error: --scale must be positive");
    }
    if (!sizeList.empty() && !sizeFreq.empty()) return die("Error: cannot specify -s with -f");          // :682-685
    if (specifiedLength && !sizeList.empty()) return die("Error: cannot specify -l and -s");             // :688-691
    if (specifiedLength && !sizeFreq.empty()) return die("Error: cannot specify -l and -f");             // :693-696
    if (!outGz.empty() && !outBAM.empty()) return die("Error: cannot specify both -o and -b");           // fragSim.cpp:675-678
    if (tag.size() > 23) return die("Error: -tag longer than 23 characters");
    if (!seedSpecified) seed = time_seed();
    const std::string inFile = argv[argc - 1];                                      // fragSim.cpp:1026

    std::string err;
    std::vector<int32_t> lens; std::vector<double> cum;
    if (!sizeList.empty() && !ggh::load_size_list(sizeList, lens, err)) return die(err);
    if (!sizeFreq.empty() && !ggh::load_size_freq(sizeFreq, lens, cum, err)) return die(err);
    if (specifiedLoc) { if (maxLen > 4095) maxLen = 4095; ggh::lognormal_table(location, scale, 4095, lens, cum); }   // int(lognormal(loc,scale)), fragSim.cpp:184
    ggh::GenomeFile G;
    if (!ggh::load_genome(inFile, G, err)) return die(err);
    std::cerr << "Mapped " << inFile << " into memory" << std::endl;                // fragSim.cpp:1149

    int circChr = -1;
    if (circ) {                                                                     // fragSim.cpp:1171-1178
        for (size_t i = 0; i < G.fai.size() && circChr < 0; i++) if (G.fai[i].name == circName) circChr = (int)i;
        if (circChr < 0) return die("ERROR: Cannot find chromosome " + circName + " which is specified via --circ");
    }

    gg_ctx* ctx = NULL;
    if (open_ctx(seed, &ctx)) return 1;
    int gid = -1;
    if (keepCase) gg_tables_set_case(ctx, 1);                                       // the upload then keeps a case bit per base
    if (ggh::upload_genome(ctx, G, &gid, circChr) != GG_OK) return gg_fail(ctx, "genome upload");
    std::vector<uint8_t>().swap(G.bytes);
    const int mode = !sizeList.empty() ? GG_SIZE_LIST : (!sizeFreq.empty() || specifiedLoc) ? GG_SIZE_FREQ : GG_SIZE_FIXED;
    if (gg_tables_set_sizes(ctx, mode, lens.data(), cum.data(), (int)lens.size(), sizeFragments, minLen, maxLen) != GG_OK) return gg_fail(ctx, "size tables");
    gg_tables_set_norev(ctx, noRev ? 1 : 0);
    if (!compFile.empty()) {                                                        // --comp / --dist, fragSim.cpp:798-1021
        ggh::CompTables ct;
        if (!ggh::load_dnacomp(compFile, distFromEnd, ct, err)) return die(err);
        if (gg_tables_set_composition(ctx, 0, ct.dist, ct.s5p.data(), ct.s5m.data(), ct.s3p.data(), ct.s3m.data()) != GG_OK) return gg_fail(ctx, "--comp");
    }
    if (gcBiasB) {                                                                  // fragSim.cpp:1190-1269
        std::cerr << "Computing average GC content" << std::endl;
        double m = 0, sd = 0;
        if (gg_tables_set_gcbias(ctx, 0, gid, compFile.empty() ? 0 : distFromEnd, gcBias, &m, &sd) != GG_OK) return gg_fail(ctx, "-gc");
        std::cerr << "GC content: mean:" << m << " stdev:" << sd << std::endl;
    }
    if (nFragments) {
        gg_range r; memset(&r, 0, sizeof r); r.first_id = 0; r.count = nFragments; r.genome = gid; r.damage_slot = -1; r.comp_slot = compFile.empty() ? -1 : 0; r.gc_slot = gcBiasB ? 0 : -1;
        snprintf(r.tag, sizeof r.tag, "%s", uniqTags ? "" : tag.c_str());        // with -uniq the tag goes after the count (fragSim.cpp:1505,131-133)
        if (gg_plan_set(ctx, &r, 1) != GG_OK) return gg_fail(ctx, "plan");
    }
    if (!outBAM.empty()) {                                                         // unaligned BAM: name = defline, flag 4, phred 0 (fragSim.cpp:1589-1601)
        ggh::BamHeader bh; ggh::bam_add_program(bh, "fragSim", command_line(argc, argv));     // :1306-1315
        BamSink bam;
        if (!bam.open(outBAM, bh, uncompressedBAM)) return die("Cannot open write to BAM output file " + outBAM);
        std::vector<uint8_t> host; UniqSuffixer uniq; uniq.reset(tag, tagb); std::string renamed; bool ok = true;
        const uint64_t chunk = 4u << 20;
        for (uint64_t first = 0; first < nFragments && ok; first += chunk) {
            const uint64_t cnt = std::min<uint64_t>(chunk, nFragments - first);
            gg_out o;
            if (gg_run_fused(ctx, first, cnt, GG_EMIT_FRAG, &o) != GG_OK) return gg_fail(ctx, "fragSim kernel");
            host.resize(o.bytes); size_t n = 0;
            if (gg_out_fetch(ctx, &o, host.data(), host.size(), &n) != GG_OK) return gg_fail(ctx, "fetch");
            const uint8_t* d = host.data();
            if (uniqTags) { renamed.clear(); uniq.apply((const char*)host.data(), n, renamed); d = (const uint8_t*)renamed.data(); n = renamed.size(); }
            for_each_record(d, n, [&](const std::string& name, const std::string& seq) {
                ggh::BamRecord r; r.name = name; r.seq = seq; r.flag = 4; r.qual.assign(seq.size(), 0);
                ok = ok && bam.add(ctx, r);
            });
        }
        if (!ok || !bam.close(ctx)) return die("Error: write failed");
        gg_ctx_destroy(ctx);
        std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << nFragments << " sequences" << std::endl;
        return 0;
    }
    Sink out;
    if (!outGz.empty()) { if (!out.open_gz(outGz)) return die("Cannot write to file " + outGz); }       // fragSim.cpp:1297-1303
    else out.open_stdout();
    std::vector<uint8_t> host; UniqSuffixer uniq; uniq.reset(tag, tagb); std::string renamed;
    if (uniqTags) {                                                                 // host-side counters over the plain stream
        const uint64_t chunk = 4u << 20;
        for (uint64_t first = 0; first < nFragments; first += chunk) {
            const uint64_t cnt = std::min<uint64_t>(chunk, nFragments - first);
            gg_out o;
            if (gg_run_fused(ctx, first, cnt, GG_EMIT_FRAG, &o) != GG_OK) return gg_fail(ctx, "fragSim kernel");
            host.resize(o.bytes); size_t n = 0;
            if (gg_out_fetch(ctx, &o, host.data(), host.size(), &n) != GG_OK) return gg_fail(ctx, "fetch");
            renamed.clear(); uniq.apply((const char*)host.data(), n, renamed); if (!out.write(renamed.data(), renamed.size())) return die("Error: write failed");
        }
    } else {                                                                        // pinned staging, copies overlapped with the kernels, .gz made on the device
        const uint64_t rec_max = std::max<uint64_t>(1, gg_max_record_bytes(ctx, GG_EMIT_FRAG));
        const uint64_t chunk = std::max<uint64_t>(1u << 16, std::min<uint64_t>(4u << 20, (1ull << 30) / rec_max));
        const size_t cap = (size_t)(std::min<uint64_t>(chunk, std::max<uint64_t>(nFragments, 1)) * rec_max + 65536);
        void* pin = NULL;
        if (gg_host_alloc(ctx, cap, &pin) != GG_OK) return gg_fail(ctx, "pinned buffer");
        for (uint64_t first = 0; first < nFragments; first += chunk)
            if (!out.write_run(ctx, first, std::min<uint64_t>(chunk, nFragments - first), GG_EMIT_FRAG, pin, cap)) return die(std::string("Error: fragSim kernel / write failed: ") + gg_last_error(ctx));
        gg_host_free(ctx, pin);
    }
    if (!out.close()) return die("Error: write failed");
    gg_ctx_destroy(ctx);
    std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << nFragments << " sequences" << std::endl;   // fragSim.cpp:1785
    return 0;
}
