// bin/adptSim -- drop-in for the reference's src/adptSim (adptSim.cpp): same flags, last argv is the FASTA (plain or gz);
// stdout 4-line records, -arts/-artp plain files, -fr/-rr gz files; same final stderr line.  GPU through the C-ABI.
#include "cli_common.h"
using namespace cli;

int main(int argc, char** argv) {
    std::string adF = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG";   // adptSim.cpp:28
    std::string adS = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT";         // adptSim.cpp:29
    int desiredLength = 100; std::string outArt, fwd, rev = "/dev/null"; bool arts = false, artp = false, fa = false;
    uint64_t seed = time_seed(); bool addInName = false; std::string tag;
    std::string outBAM; bool bamPaired = false, uncompressedBAM = false;
    const std::string usage = std::string("\t") + argv[0] + " [options]  [fasta file]\n\n This program reads a fasta file containing aDNA fragments and\n adds adapters / trims to the read length (B200 build)\n\n" +
        "\t\t-arts\t[out]\t\tOutput single-end reads as ART (unzipped fasta)\n\t\t-artp\t[out]\t\tOutput reads as ART (unzipped fasta) with wrap-around for paired-end mode\n" +
        "\t\t-fr\t[out fwdr]\tOutput forward read as zipped fasta\n\t\t-rr\t[out rwdr]\tOutput reverse read as zipped fasta\n" +
        "\t-f\t[seq]\t\t\tAdapter that is observed after the forward read\n\t-s\t[seq]\t\t\tAdapter that is observed after the reverse read\n\t-l\t[length]\t\tDesired read length  (Default:  100)\n" +
        "\t-name\t\t\t\tAppend _adpt/_rand to deflines if adapters / random bases are added\n\t-tag\t[tag]\t\t\tAppend this string to deflines\n\t\t-bs\t[bam out]\tRead BAM, write single-end reads as unaligned BAM\n\t\t-bp\t[bam out]\tRead BAM, write paired-end reads as unaligned BAM\n\t\t-u\t\t\tProduce uncompressed BAM\n";
    if (is_help(argc, argv)) { std::cout << "Usage:" << std::endl << std::endl << usage << std::endl; return 1; }   // adptSim.cpp:82-90
    for (int i = 1; i < argc - 1; i++) {                                            // adptSim.cpp:92-177
        const std::string a = argv[i];
        if (a == "-arts") { outArt = argv[++i]; arts = true; continue; }
        if (a == "-artp") { outArt = argv[++i]; artp = true; continue; }
        if (a == "-l") { desiredLength = (int)to_ll(argv[++i], "-l"); continue; }
        if (a == "-f") { adF = argv[++i]; continue; }
        if (a == "-s") { adS = argv[++i]; continue; }
        if (a == "-fr") { fwd = argv[++i]; fa = true; continue; }
        if (a == "-rr") { rev = argv[++i]; fa = true; continue; }
        if (a == "--seed") { seed = (uint64_t)to_ll(argv[++i], "--seed"); continue; }    // extension: the reference pads from an unseeded rand()
        if (a == "-name") { addInName = true; continue; }                           // adptSim.cpp:163-166
        if (a == "-tag") { tag = argv[++i]; continue; }                              // adptSim.cpp:168-173
        if (a == "-bs") { outBAM = argv[++i]; bamPaired = false; continue; }         // adptSim.cpp:115-121: BAM in, single-end BAM out
        if (a == "-bp") { outBAM = argv[++i]; bamPaired = true; continue; }          // adptSim.cpp:123-129: paired-end BAM out
        if (a == "-u") { uncompressedBAM = true; continue; }                         // adptSim.cpp:94-97
        return die("Error: unknown option " + a);                                   // adptSim.cpp:175-176
    }
    if (fa && fwd.empty()) return die("Error: need to specify the output forward read ");                  // :179-182
    if (fa && (arts || artp)) return die("Error: cannot produce both fastq and ART output");               // :189-192
    const std::string infileName = argv[argc - 1];                                  // adptSim.cpp:199

    gg_ctx* ctx = NULL;
    if (open_ctx(seed, &ctx)) return 1;
    if (gg_tables_set_adapters(ctx, adF.c_str(), adS.c_str(), desiredLength) != GG_OK) return gg_fail(ctx, "adapters");
    if (!outBAM.empty()) {
        // BAM in, BAM out (adptSim.cpp:217-243,273-279,328-373): fresh alignments named like the input read (no "_r"), flag 4 or 77/141,
        // phred 0; -name / -tag go into an XA:Z tag ("AD" adapter bases added, ",RD" random bases added, then the tag)
        std::string err; ggh::BamHeader bh; std::vector<ggh::BamRecord> recs;
        if (!ggh::bam_read(infileName, bh, recs, err)) return die("Could not open input BAM file  " + infileName);
        ggh::bam_add_program(bh, "adptSim", command_line(argc, argv));
        BamSink bam;
        if (!bam.open(outBAM, bh, uncompressedBAM)) return die("Could not open output BAM file  " + outBAM);
        if (gg_tables_set_adpt_naming(ctx, 0, "") != GG_OK) return gg_fail(ctx, "-tag");
        std::vector<uint8_t> in, host; bool ok = true;
        for (size_t r0 = 0; r0 < recs.size() && ok; ) {
            in.clear(); size_t r1 = r0;
            while (r1 < recs.size() && in.size() < STREAM_CHUNK_BYTES) { in.push_back('>'); in.insert(in.end(), recs[r1].name.begin(), recs[r1].name.end()); in.push_back('\n'); in.insert(in.end(), recs[r1].seq.begin(), recs[r1].seq.end()); in.push_back('\n'); r1++; }
            gg_out o; size_t n = 0;
            if (gg_run_adpt(ctx, in.data(), in.size(), bamPaired ? GG_EMIT_STDOUT4 : GG_EMIT_ARTS, (uint64_t)r0, &o) != GG_OK) return gg_fail(ctx, "adptSim kernel");
            host.resize(o.bytes);
            if (gg_out_fetch(ctx, &o, host.data(), host.size(), &n) != GG_OK) return gg_fail(ctx, "fetch");
            size_t k = 2 * r0;                                                      // output line pairs: forward read [, reverse read] per input record
            for_each_record(host.data(), n, [&](const std::string&, const std::string& seq) {
                const size_t src = bamPaired ? k / 2 : k - r0; const bool second = bamPaired && (k & 1);
                k++;
                if (src >= r1) { ok = false; return; }
                ggh::BamRecord R; R.name = recs[src].name; R.seq = seq; R.qual.assign(seq.size(), 0);
                R.flag = !bamPaired ? 4 : second ? 141 : 77;                        // adptSim.cpp:18-20,337-342
                if (addInName || !tag.empty()) {
                    std::string t; const int L0 = (int)recs[src].seq.size();
                    if (addInName) { if (L0 < desiredLength) t += "AD"; if (L0 < desiredLength && L0 + (int)adF.size() < desiredLength) t += ",RD"; }     // flags of the forward read (:294-310,344-351)
                    t += tag;
                    R.add_tag_z("XA", t);
                }
                ok = ok && bam.add(ctx, R);
            });
            ok = ok && k == (bamPaired ? 2 * r1 : r0 + r1);
            r0 = r1;
        }
        if (!ok || !bam.close(ctx)) return die("Error: write failed");
        gg_ctx_destroy(ctx);
        std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << recs.size() << " sequences" << std::endl;
        return 0;
    }
    if (gg_tables_set_adpt_naming(ctx, addInName ? 1 : 0, tag.c_str()) != GG_OK) return gg_fail(ctx, "-tag");
    // one sink per output file: -fr and -rr are two streams of the same input (adptSim.cpp:397-398)
    struct Pass { int emit; Sink sink; };
    std::vector<Pass> passes;
    if (fa) {
        Pass p1; p1.emit = GG_EMIT_FR; if (!p1.sink.open_gz(fwd)) return die("Cannot write to file " + fwd); passes.push_back(p1);
        Pass p2; p2.emit = GG_EMIT_RR; if (!p2.sink.open_gz(rev)) return die("Cannot write to file " + rev); passes.push_back(p2);
    } else if (arts || artp) {
        Pass p; p.emit = artp ? GG_EMIT_ARTP : GG_EMIT_ARTS; if (!p.sink.open_plain(outArt)) return die("Cannot write to file " + outArt); passes.push_back(p);
    } else { Pass p; p.emit = GG_EMIT_STDOUT4; p.sink.open_stdout(); passes.push_back(p); }
    // reader thread -> GPU (this thread) -> writer thread: bounded memory whatever the input size, file I/O overlapped with the device work
    uint64_t nrec = 0; std::string rerr; bool wfailed = false;
    BoundedQueue<InChunk> qin(2); BoundedQueue<OutChunk> qout(2);
    std::vector<Sink*> sinks; for (size_t k = 0; k < passes.size(); k++) sinks.push_back(&passes[k].sink);
    std::thread rd(record_chunk_reader, infileName, STREAM_CHUNK_BYTES, &qin, &rerr, &nrec);
    std::thread wr(chunk_writer, &qout, sinks, &wfailed);
    uint64_t first = 0; InChunk c; std::string gerr;
    while (qin.pop(c) && gerr.empty()) {
        for (size_t k = 0; k < passes.size(); k++) {
            gg_out o; OutChunk oc;
            if (gg_run_adpt(ctx, c.bytes.data(), c.bytes.size(), passes[k].emit, first, &o) != GG_OK || !fetch_out(ctx, o, passes[k].sink, (int)k, &oc)) { gerr = std::string("adptSim kernel: ") + gg_last_error(ctx); qin.abort(); break; }
            qout.push(std::move(oc));
        }
        first += c.recs;
        if (wfailed) { qin.abort(); break; }
    }
    qout.close(); rd.join(); wr.join();
    if (!rerr.empty()) return die(rerr);
    if (!gerr.empty()) return die("Error: " + gerr);
    if (wfailed) return die("Error: write failed");
    for (size_t k = 0; k < passes.size(); k++) if (!passes[k].sink.close()) return die("Error: write failed");
    gg_ctx_destroy(ctx);
    std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << nrec << " sequences" << std::endl;   // adptSim.cpp:437
    return 0;
}
