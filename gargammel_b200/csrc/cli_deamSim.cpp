// bin/deamSim -- drop-in for the reference's src/deamSim (deamSim.cpp): same flags, last argv is the FASTA (plain or gz),
// "ID\nSEQ\n" records on stdout / -o gz, same final stderr line.  The per-base damage runs on the GPU through the C-ABI.
#include "cli_common.h"
using namespace cli;

int main(int argc, char** argv) {
    std::string outBAM; bool uncompressedBAM = false;
    std::string outGz, matrixFile, profFile, mapdamageFile, mapdamageProtocol, matrix, matrixFileNoMeth, matrixFileWiMeth;
    bool useBriggs = false, seedSpecified = false, lastRowMatrix = true, matrixSpecified = false, putDeamInName = false;
    bool useBriggsss = false; double s5p = 0, s3p = 0, l5p = 0, l3p = 0;            // -damagess, deamSim.cpp:310-340
    double v = 0, l = 0, d = 0, s = 0; uint64_t seed = 0;
    const std::string usage = std::string("\t") + argv[0] + " [options]  [fasta file]\n\n This program reads a fasta file containing aDNA fragments and\n adds deamination according to a certain model (B200 build)\n\n" +
        "\t\t-o\t[fasta out]\t\t\tWrite fasta output as a zipped fasta\n\t\t-last\t\t\t\t\tdo not use the last matrix row beyond the matrix\n\t\t--seed\t[int]\n\n" +
        "\t\t-mapdamage  [mis.txt] [protocol]\n\t\t-profile  [prof file prefix]\n\t\t-matfile  [matrix file prefix]\n\t\t-mat      [single|double]\n\t\t-damage     [v,l,d,s]\n" +
        "\t\t-name\t\t\t\t\tPut a tag in the read name with deam bases\n\t\t-matfilenonmeth  [matrix file prefix]\tRead the matrix file of substitutions for non-methylated Cs\n\t\t-matfilemeth  [matrix file prefix]\tRead the matrix file of substitutions for methylated Cs\n\t\t-b\t[bam out]\t\t\tRead a BAM file, write a BAM file\n\t\t-u\t\t\t\t\tProduce uncompressed BAM\n";
    if (is_help(argc, argv)) { std::cout << "Usage:" << std::endl << std::endl << usage << std::endl; return 1; }   // deamSim.cpp:199-207
    for (int i = 1; i < argc - 1; i++) {                                            // deamSim.cpp:213-378
        const std::string a = argv[i];
        if (a == "-mat") {
            matrix = argv[++i];
            if (matrix != "single" && matrix != "double") return die("Specify either \"single\" or \"double\" (without quotes), you entered " + matrix);
            matrixSpecified = true; continue;
        }
        if (a == "--seed") { seed = (uint64_t)to_ll(argv[++i], "--seed"); seedSpecified = true; continue; }
        if (a == "-profile") { profFile = argv[++i]; continue; }
        if (a == "-matfile") { matrixFile = argv[++i]; continue; }
        if (a == "-mapdamage") {
            if (i + 2 >= argc) return die("Error: -mapdamage needs a file and a protocol");
            mapdamageFile = argv[++i]; mapdamageProtocol = argv[++i];
            if (mapdamageProtocol != "single" && mapdamageProtocol != "double") return die("The protocol specified should be 'single' or 'double' without the quotes, got: " + mapdamageProtocol);
            continue;
        }
        if (a == "-damage") {
            const std::string p = argv[++i]; useBriggs = true;
            std::vector<std::string> t = ggh::all_tokens(p, ',');
            if (t.size() != 4) return die("Specify 4 comma-separated values for the Briggs model, you entered:" + p);
            v = to_d(t[0].c_str(), "-damage"); l = to_d(t[1].c_str(), "-damage"); d = to_d(t[2].c_str(), "-damage"); s = to_d(t[3].c_str(), "-damage");
            if (v < 0 || v > 1) return die("Nick frequency must be between 0 and 1, entered:" + t[0]);
            if (l < 0 || l > 1) return die("Geometric parameter for overhang must be between 0 and 1, entered:" + t[1]);
            if (d < 0 || d > 1) return die("double-strand deamination prob. must be between 0 and 1, entered:" + t[2]);
            if (s < 0 || s > 1) return die("single-strand deamination prob. must be between 0 and 1, entered:" + t[3]);
            continue;
        }
        if (a == "-damagess") {                                                     // deamSim.cpp:310-340 (hidden in the reference's usage text)
            const std::string p = argv[++i]; useBriggsss = true;
            std::vector<std::string> t = ggh::all_tokens(p, ',');
            if (t.size() != 5) return die("Specify 4 comma-separated values for the Briggs model, you entered:" + p);          // (sic) :314-316
            d = to_d(t[0].c_str(), "-damagess"); s5p = to_d(t[1].c_str(), "-damagess"); s3p = to_d(t[2].c_str(), "-damagess");
            l5p = to_d(t[3].c_str(), "-damagess"); l3p = to_d(t[4].c_str(), "-damagess");
            if (l5p < 0 || l5p > 1) return die("Geometric parameter for overhang must be between 0 and 1, entered:" + t[3]);
            if (l3p < 0 || l3p > 1) return die("Geometric parameter for overhang must be between 0 and 1, entered:" + t[4]);
            if (d < 0 || d > 1) return die("double-strand deamination prob. must be between 0 and 1, entered:" + t[0]);
            if (s5p < 0 || s5p > 1) return die("single-strand deamination prob. must be between 0 and 1, entered:" + t[1]);
            if (s3p < 0 || s3p > 1) return die("single-strand deamination prob. must be between 0 and 1, entered:" + t[2]);
            continue;
        }
        if (a == "-o") { outGz = argv[++i]; continue; }
        if (a == "-last") { lastRowMatrix = false; continue; }
        if (a == "-v") return unsupported(a, "the per-read log of deamSim.cpp:1458-1465,1602-1611 is not produced by this build");
        if (a == "-name") { putDeamInName = true; continue; }                        // deamSim.cpp:342-345
        if (a == "-matfilenonmeth") { matrixFileNoMeth = argv[++i]; continue; }      // deamSim.cpp:256-261
        if (a == "-matfilemeth") { matrixFileWiMeth = argv[++i]; continue; }         // deamSim.cpp:263-268
        if (a == "-b") { outBAM = argv[++i]; continue; }                              // deamSim.cpp:347-352: BAM in, BAM out
        if (a == "-u") { uncompressedBAM = true; continue; }                         // deamSim.cpp:354-357
        return die("Error: unknown option " + a);                                   // deamSim.cpp:376-377
    }
    const bool mapd = !mapdamageFile.empty();
    if (useBriggs && (matrixSpecified || mapd)) return die("Error: cannot specify both a file and Briggs parameter model");        // :380-383
    if (useBriggsss && (matrixSpecified || mapd)) return die("Error: cannot specify both a file and Briggs parameter model");      // :385-388
    if (useBriggs && useBriggsss) return die("Error: cannot specify both a file and Briggs parameter model");                      // :390-393
    const bool noMeth = !matrixFileNoMeth.empty(), wiMeth = !matrixFileWiMeth.empty();
    if (!useBriggs && !useBriggsss && !matrixSpecified && matrixFile.empty() && !mapd && profFile.empty() && !noMeth && !wiMeth)
        return die("Please specify the matrix to use or use the Briggs parameter model");                                          // :394-405
    if (!profFile.empty() && (!matrixFile.empty() || noMeth || wiMeth)) return die("Please specify damage profile or substitution matrix");   // :407-412
    if (!matrixFile.empty() && noMeth && wiMeth) return die("Please specify matrix with/without methylation but not the overall matrix");     // :414-419
    if (noMeth != wiMeth) return die("Please specify both the matrix with/without methylation, not just one");                     // :421-430
    if (!outGz.empty() && !outBAM.empty()) return die("Error: cannot specify both -o and -b");           // deamSim.cpp:432-435
    if (!seedSpecified) seed = time_seed();
    const std::string inputFile = argv[argc - 1];                                   // deamSim.cpp:436

    gg_ctx* ctx = NULL;
    if (open_ctx(seed, &ctx)) return 1;
    print_rss("context open");
    std::string err; ggh::Rates s5, s3;
    if (useBriggsss) { if (gg_tables_set_briggs_ss(ctx, 0, d, s5p, s3p, l5p, l3p) != GG_OK) return gg_fail(ctx, "-damagess"); }   // checked first, like :1337
    else if (useBriggs) { if (gg_tables_set_briggs(ctx, 0, v, l, d, s) != GG_OK) return gg_fail(ctx, "-damage"); }
    else if (noMeth) {                                                              // deamSim.cpp:673-954: <prefix>5.dat / <prefix>3.dat of each table
        ggh::Rates m5, m3;
        std::cerr << "Files deamination with methylation" << std::endl << matrixFileWiMeth << "5.dat" << std::endl << matrixFileWiMeth << "3.dat" << std::endl;      // :683-685
        if (!ggh::load_matrix_dat(matrixFileWiMeth, m5, m3, err)) return die(err);
        std::cerr << "Files deamination without methylation" << std::endl << matrixFileNoMeth << "5.dat" << std::endl << matrixFileNoMeth << "3.dat" << std::endl;   // :824-826
        if (!ggh::load_matrix_dat(matrixFileNoMeth, s5, s3, err)) return die(err);
        if (gg_tables_set_matrix_meth(ctx, 0, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12), m5.data(), (int)(m5.size() / 12), m3.data(), (int)(m3.size() / 12), lastRowMatrix) != GG_OK)
            return gg_fail(ctx, "-matfilemeth/-matfilenonmeth");
    } else if (mapd) {                                                              // deamSim.cpp:453-668
        if (!ggh::load_mapdamage(mapdamageFile, s5, s3, err)) return die(err);
        if (gg_tables_set_singledouble(ctx, 0, mapdamageProtocol == "single" ? GG_DMG_SINGLE : GG_DMG_DOUBLE, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12)) != GG_OK) return gg_fail(ctx, "-mapdamage");
    } else if (!profFile.empty()) {                                                 // :956-1094
        if (!ggh::load_profile(profFile, s5, s3, err)) return die(err);
        if (gg_tables_set_matrix(ctx, 0, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12), lastRowMatrix) != GG_OK) return gg_fail(ctx, "-profile");
    } else if (!matrixFile.empty()) {                                               // :1098-1241
        if (!ggh::load_matrix_dat(matrixFile, s5, s3, err)) return die(err);
        if (gg_tables_set_matrix(ctx, 0, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12), lastRowMatrix) != GG_OK) return gg_fail(ctx, "-matfile");
    } else {                                                                        // -mat single|double: <exe dir>/matrices/<m>-{5,3}.dat (:1105-1106)
        if (!ggh::load_matrix_dat(exe_dir(argv[0]) + "matrices/" + matrix + "-", s5, s3, err)) return die(err);
        if (gg_tables_set_singledouble(ctx, 0, matrix == "single" ? GG_DMG_SINGLE : GG_DMG_DOUBLE, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12)) != GG_OK) return gg_fail(ctx, "-mat");
    }
    gg_tables_set_deam_naming(ctx, putDeamInName ? 1 : 0);
    if (!outBAM.empty()) {
        // BAM in, BAM out (deamSim.cpp:1260-1286,1305-1310,2016-2024): every alignment keeps its fields, QueryBases are replaced and, with
        // -name, the positions go into an XD:Z tag instead of the read name
        ggh::BamHeader bh; std::vector<ggh::BamRecord> recs;
        if (!ggh::bam_read(inputFile, bh, recs, err)) return die("Could not open input BAM file  " + inputFile);
        ggh::bam_add_program(bh, "deamSim", command_line(argc, argv));
        BamSink bam;
        if (!bam.open(outBAM, bh, uncompressedBAM)) return die("Could not open output BAM file  " + outBAM);
        std::vector<uint8_t> in, host; bool ok = true;
        for (size_t r0 = 0; r0 < recs.size() && ok; ) {
            in.clear(); size_t r1 = r0;
            while (r1 < recs.size() && in.size() < STREAM_CHUNK_BYTES) { in.push_back('>'); in.insert(in.end(), recs[r1].name.begin(), recs[r1].name.end()); in.push_back('\n'); in.insert(in.end(), recs[r1].seq.begin(), recs[r1].seq.end()); in.push_back('\n'); r1++; }
            gg_out o; size_t n = 0;
            if (gg_run_deam(ctx, in.data(), in.size(), 0, (uint64_t)r0, &o) != GG_OK) return gg_fail(ctx, "deamSim kernel");
            host.resize(o.bytes);
            if (gg_out_fetch(ctx, &o, host.data(), host.size(), &n) != GG_OK) return gg_fail(ctx, "fetch");
            size_t k = r0;
            for_each_record(host.data(), n, [&](const std::string& name, const std::string& seq) {
                if (k >= r1) { ok = false; return; }
                ggh::BamRecord& R = recs[k++];
                R.seq = seq;
                if (name.size() > R.name.size()) R.add_tag_z("XD", name.substr(R.name.size() + 6));      // "_DEAM:" + list (deamSim.cpp:2018-2022)
                ok = ok && bam.add(ctx, R);
            });
            ok = ok && k == r1;
            r0 = r1;
        }
        if (!ok || !bam.close(ctx)) return die("Error: write failed");
        gg_ctx_destroy(ctx);
        std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << recs.size() << " sequences" << std::endl;
        return 0;
    }
    Sink out;
    if (!outGz.empty()) { if (!out.open_gz(outGz)) return die("Cannot write to file " + outGz); } else out.open_stdout();
    // reader thread -> GPU (this thread) -> writer thread: bounded memory whatever the input size, file I/O overlapped with the device work
    uint64_t nrec = 0; std::string rerr; bool wfailed = false;
    BoundedQueue<InChunk> qin(2); BoundedQueue<OutChunk> qout(2);
    std::thread rd(record_chunk_reader, inputFile, STREAM_CHUNK_BYTES, &qin, &rerr, &nrec);
    std::thread wr(chunk_writer, &qout, std::vector<Sink*>(1, &out), &wfailed);
    uint64_t first = 0; InChunk c; std::string gerr;
    while (qin.pop(c)) {
        gg_out o; OutChunk oc;
        if (gg_run_deam(ctx, c.bytes.data(), c.bytes.size(), 0, first, &o) != GG_OK || !fetch_out(ctx, o, out, 0, &oc)) { gerr = std::string("deamSim kernel: ") + gg_last_error(ctx); qin.abort(); break; }
        qout.push(std::move(oc));
        first += c.recs;
        if (wfailed) { qin.abort(); break; }
    }
    qout.close(); rd.join(); wr.join();
    print_rss("stream done");
    if (!rerr.empty()) return die(rerr);
    if (!gerr.empty()) return die("Error: " + gerr);
    if (wfailed || !out.close()) return die("Error: write failed");
    gg_ctx_destroy(ctx);
    std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << nrec << " sequences" << std::endl;   // deamSim.cpp:2053
    return 0;
}
