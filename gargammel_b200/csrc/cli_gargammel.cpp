// bin/gargammel-b200 -- fused driver for the middle of gargammel.pl: scans <dir>/{endo,cont,bact},
// apportions the fragments across the sources (gargammel.pl:1256-1283,1320-1349,1421-1431), and runs
// fragSim -> deamSim -> adptSim as ONE kernel per GPU instead of ~N process launches, 3 gzip and 2 gunzip hops.
// Writes <prefix>_a.fa (what adptSim -artp/-arts writes for ART, gargammel.pl:1838-1848); with --keep also the
// intermediate <prefix>.{e,b,c}.fa.gz and <prefix>_d.fa.gz of the reference.  ART stays on the host (north_star): when an
// art_illumina is found (--art PATH, next to this binary like gargammel.pl:232-239, or on PATH) it is run exactly as gargammel.pl:1856-1869
// does and the outputs get the reference's final names <prefix>_a.fa.gz, <prefix>_s.fq.gz / _s1.fq.gz + _s2.fq.gz (:1871-1889).
// The closing `gzip -f` steps are done in this process: <prefix>_a.fa.gz is made on the device during the run, the FASTQ files by all host threads.
#include "cli_common.h"
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>
#include <limits.h>
#include <algorithm>
#include <thread>
#include <math.h>
using namespace cli;

static bool ends_with(const std::string& s, const char* suf) { size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }
// listdirFa, gargammel.pl:156-183 (sorted here: readdir order is not portable)
static std::vector<std::string> list_fa(const std::string& dir) {
    std::vector<std::string> out; DIR* d = opendir(dir.c_str());
    if (!d) return out;
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        struct stat st; if (stat((dir + n).c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
        if (ends_with(n, ".fa") || ends_with(n, ".fa.gz") || ends_with(n, ".fna") || ends_with(n, ".fasta") || ends_with(n, ".fasta.gz")) out.push_back(dir + n);
    }
    closedir(d); std::sort(out.begin(), out.end());
    return out;
}
// listdirCdir, gargammel.pl:186-210: every subdirectory (sorted here)
static std::vector<std::string> list_dirs(const std::string& dir) {
    std::vector<std::string> out; DIR* d = opendir(dir.c_str());
    if (!d) return out;
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        struct stat st; if (stat((dir + n).c_str(), &st) == 0 && S_ISDIR(st.st_mode)) out.push_back(dir + n);
    }
    closedir(d); std::sort(out.begin(), out.end());
    return out;
}
// runcmd, gargammel.pl:71-84: echo the command, run it through bash unless --mock, die on a non-zero status
static bool run_cmd(const std::string& cmd, bool mock) {
    std::cerr << "running cmd " << cmd << std::endl;
    if (mock) return true;
    const std::string full = "bash -c '" + cmd + "'";
    const int rc = system(full.c_str());
    if (rc != 0) { std::cerr << "system  cmd " << cmd << " failed: " << rc << std::endl; return false; }
    return true;
}
// The `gzip -f FILE` steps that end gargammel.pl (:1871-1889), in this process: FILE is deflated block-wise on every host thread
// (BGZF members at gzip's default level 6: zcat / gzstream / htslib read it like any .gz) into FILE.gz and removed.  A single gzip
// process takes minutes on the FASTQ files of a large run; the command is echoed like the others so that --mock shows the step.
static bool gzip_file(const std::string& path, bool mock) {
    std::cerr << "running cmd gzip -f " << path << "   (in this process: parallel BGZF members)" << std::endl;
    if (mock) return true;
    FILE* in = fopen(path.c_str(), "rb");
    if (!in) { std::cerr << "gzip: cannot read " << path << std::endl; return false; }
    const std::string gz = path + ".gz";
    FILE* out = fopen(gz.c_str(), "wb");
    if (!out) { fclose(in); std::cerr << "gzip: cannot write " << gz << std::endl; return false; }
    const size_t B = 0xff00, piece = B * 1024;                                       // whole BGZF blocks per piece (about 64 MB)
    std::vector<uint8_t> buf(piece), z; std::string e; bool ok = true; size_t got;
    while (ok && (got = fread(buf.data(), 1, piece, in)) > 0)
        ok = ggh::bgzf_compress(buf.data(), got, 0, 6, z, false, e) && fwrite(z.data(), 1, z.size(), out) == z.size();
    ok = ok && !ferror(in) && ggh::bgzf_compress(NULL, 0, 1, 6, z, true, e) && fwrite(z.data(), 1, z.size(), out) == z.size();   // the EOF marker block
    fclose(in);
    ok = (fclose(out) == 0) && ok;
    if (!ok) { std::cerr << "gzip: failed on " << path << (e.empty() ? "" : ": " + e) << std::endl; remove(gz.c_str()); return false; }
    remove(path.c_str());
    return true;
}
static bool is_exe(const std::string& p) { struct stat st; return !p.empty() && stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode) && access(p.c_str(), X_OK) == 0; }
// art_illumina: --art PATH, else <dir of this binary>/../art_src_MountRainier/art_illumina (the reference's layout, gargammel.pl:239), else $PATH
static std::string find_art(const std::string& given, const char* argv0) {
    if (!given.empty()) return is_exe(given) ? given : std::string();
    char self[PATH_MAX]; const ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
    std::string dir = n > 0 ? std::string(self, (size_t)n) : std::string(argv0);
    const size_t sl = dir.rfind('/'); dir = sl == std::string::npos ? "." : dir.substr(0, sl);
    const std::string cand[2] = {dir + "/../art_src_MountRainier/art_illumina", dir + "/art_src_MountRainier/art_illumina"};
    for (int k = 0; k < 2; k++) if (is_exe(cand[k])) return cand[k];
    const char* path = getenv("PATH");
    for (const std::string& d : ggh::all_tokens(path ? path : "", ':')) if (is_exe(d + "/art_illumina")) return d + "/art_illumina";
    return std::string();
}

struct DamageSpec { std::string matfile, briggs, mapd_file, mapd_proto; bool any() const { return !matfile.empty() || !briggs.empty() || !mapd_file.empty(); } };

static int set_damage(gg_ctx* ctx, int slot, const DamageSpec& d) {
    std::string err; ggh::Rates s5, s3;
    if (!d.matfile.empty()) {
        if (!ggh::load_matrix_dat(d.matfile, s5, s3, err)) return die(err);
        if (gg_tables_set_matrix(ctx, slot, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12), 1) != GG_OK) return gg_fail(ctx, "-matfile");
    } else if (!d.briggs.empty()) {
        std::vector<std::string> t = ggh::all_tokens(d.briggs, ',');
        if (t.size() != 4) return die("Specify 4 comma-separated values for the Briggs model, you entered:" + d.briggs);
        if (gg_tables_set_briggs(ctx, slot, to_d(t[0].c_str(), "-damage"), to_d(t[1].c_str(), "-damage"), to_d(t[2].c_str(), "-damage"), to_d(t[3].c_str(), "-damage")) != GG_OK) return gg_fail(ctx, "-damage");
    } else if (!d.mapd_file.empty()) {
        if (d.mapd_proto != "single" && d.mapd_proto != "double") return die("The protocol specified should be 'single' or 'double' without the quotes, got: " + d.mapd_proto);
        if (!ggh::load_mapdamage(d.mapd_file, s5, s3, err)) return die(err);
        if (gg_tables_set_singledouble(ctx, slot, d.mapd_proto == "single" ? GG_DMG_SINGLE : GG_DMG_DOUBLE, s5.data(), (int)(s5.size() / 12), s3.data(), (int)(s3.size() / 12)) != GG_OK) return gg_fail(ctx, "-mapdamage");
    }
    return 0;
}

int main(int argc, char** argv) {
    double compB = 0, compC = 0, compE = 1; std::string comp = "0,0,1";
    uint64_t n = 1000; double coverage = -1; int fraglength = 35, minsize = 0, maxsize = 1000, rl = 75, gpus = 1;
    std::string sizeList, sizeFreq, prefix, adF = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG", adS = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT";   // gargammel.pl:255-257
    bool uniq = false, mock = false, no_art = false;
    std::string ss = "HS25", art_given; int qs = 0, qs2 = 0;                           // ART options of gargammel.pl:460-462 (defaults :268-270)
    bool se = false, keep = false, n_given = false, seedSpecified = false, lognorm = false; uint64_t seed = 0; double loc = 0, scale = 0;
    DamageSpec dAll, dE, dB, dC;
    bool methyl = false; std::string matNoMeth, matWiMeth;                          // --methyl -matfilenonmeth -matfilemeth (gargammel.pl:431-432,457)
    std::string misE, misB, misC; int distmis = 1;                                  // --misince/--misincb/--misincc/--distmis (gargammel.pl:300-303)
    const std::string usage = std::string("\n usage:\t") + argv[0] + " <options> [directory with fasta directories]\n\n This directory should contain endo/ cont/ bact/ (bact/ needs a `list` file when --comp B > 0);\n endo/ holds 1 (haploid) or 2 (diploid) fasta files, or one directory per cell C0 .. CN with 1 or 2 files each\n\n" +
        "\t--comp [B,C,E]\tcomposition (default 0,0,1)\n\t-o\t\toutput prefix (default: [input dir]/simadna)\n\t-n [number] | -c [coverage]\n\t-l [len] | -s [file] | -f [file]\tfragment sizes\n\t-minsize -maxsize\n" +
        "\t-matfile/-damage/-mapdamage f proto\t(endogenous + bacterial; suffix e/b/c for one source)\n\t--methyl -matfilenonmeth [prefix] -matfilemeth [prefix]\tlower-case c/g in the genomes are methylated cytosines\n\t-fa -sa -rl -se\tadapters, read length, single-end\n\t-ss [system] -qs [int] -qs2 [int]\tART sequencing system and quality shifts\n" +
        "\t--art [path]\tart_illumina to run on <prefix>_a.fa (default: ../art_src_MountRainier/art_illumina next to this binary, then $PATH)\n\t--noart\t\tstop after <prefix>_a.fa\n" +
        "\t--mock\t\tprint what would be run, run nothing\n\t--seed [int]\t--gpus [N]\t--keep\n";
    if (is_help(argc, argv)) { std::cout << usage << std::endl; return 1; }
    for (int i = 1; i < argc - 1; i++) {
        std::string a = argv[i]; if (a.size() > 2 && a[0] == '-' && a[1] == '-') a = a.substr(1);      // Getopt::Long takes -x and --x
        auto next = [&]() -> const char* { if (i + 1 >= argc - 1 + 1) { std::cerr << "Error: option " << a << " needs a value" << std::endl; exit(1); } return argv[++i]; };
        if (a == "-comp") { comp = next(); continue; }
        if (a == "-o") { prefix = next(); continue; }
        if (a == "-n") { n = (uint64_t)to_ll(next(), "-n"); n_given = true; continue; }
        if (a == "-c") { coverage = to_d(next(), "-c"); continue; }
        if (a == "-l") { fraglength = (int)to_ll(next(), "-l"); continue; }
        if (a == "-s") { sizeList = next(); continue; }
        if (a == "-f") { sizeFreq = next(); continue; }
        if (a == "-loc") { loc = to_d(next(), "-loc"); lognorm = true; continue; }
        if (a == "-scale") { scale = to_d(next(), "-scale"); lognorm = true; continue; }
        if (a == "-minsize") { minsize = (int)to_ll(next(), "-minsize"); continue; }
        if (a == "-maxsize") { maxsize = (int)to_ll(next(), "-maxsize"); continue; }
        if (a == "-matfile") { dAll.matfile = next(); continue; }
        if (a == "-damage") { dAll.briggs = next(); continue; }
        if (a == "-mapdamage") { dAll.mapd_file = next(); dAll.mapd_proto = next(); continue; }
        if (a == "-matfilee") { dE.matfile = next(); continue; }  if (a == "-damagee") { dE.briggs = next(); continue; }
        if (a == "-matfileb") { dB.matfile = next(); continue; }  if (a == "-damageb") { dB.briggs = next(); continue; }
        if (a == "-matfilec") { dC.matfile = next(); continue; }  if (a == "-damagec") { dC.briggs = next(); continue; }
        if (a == "-mapdamagee") { dE.mapd_file = next(); dE.mapd_proto = next(); continue; }
        if (a == "-mapdamageb") { dB.mapd_file = next(); dB.mapd_proto = next(); continue; }
        if (a == "-mapdamagec") { dC.mapd_file = next(); dC.mapd_proto = next(); continue; }
        if (a == "-misince") { misE = next(); continue; }  if (a == "-misincb") { misB = next(); continue; }  if (a == "-misincc") { misC = next(); continue; }
        if (a == "-distmis") { distmis = (int)to_ll(next(), "-distmis"); continue; }
        if (a == "-fa") { adF = next(); continue; }  if (a == "-sa") { adS = next(); continue; }
        if (a == "-rl") { rl = (int)to_ll(next(), "-rl"); continue; }
        if (a == "-se") { se = true; continue; }
        if (a == "-seed") { seed = (uint64_t)to_ll(next(), "--seed"); seedSpecified = true; continue; }
        if (a == "-gpus") { gpus = (int)to_ll(next(), "--gpus"); continue; }
        if (a == "-keep") { keep = true; continue; }
        if (a == "-uniq") { uniq = true; continue; }                                // fragSim -uniq for every source (gargammel.pl:1506-1507,1544,1628,1675)
        if (a == "-mock") { mock = true; continue; }                               // gargammel.pl:71-84
        if (a == "-ss") { ss = next(); continue; }
        if (a == "-qs") { qs = (int)to_ll(next(), "-qs"); continue; }
        if (a == "-qs2") { qs2 = (int)to_ll(next(), "-qs2"); continue; }
        if (a == "-art") { art_given = next(); continue; }
        if (a == "-noart") { no_art = true; continue; }
        if (a == "-methyl") { methyl = true; continue; }                           // gargammel.pl:320-322,461
        if (a == "-matfilenonmeth") { matNoMeth = next(); continue; }
        if (a == "-matfilemeth") { matWiMeth = next(); continue; }
        return die("Unknown option: " + a);
    }
    if (coverage >= 0 && n_given) return die("Must use the -c or -n but not both");                    // gargammel.pl:753-756
    std::vector<std::string> c3 = ggh::all_tokens(comp, ',');
    if (c3.size() != 3) return die("the --comp option must have 3 comma-delimited fields, found " + std::to_string(c3.size()));   // :748-751
    compB = to_d(c3[0].c_str(), "--comp"); compC = to_d(c3[1].c_str(), "--comp"); compE = to_d(c3[2].c_str(), "--comp");
    if (coverage >= 0 && compE == 0) return die("Cannot use the endogenous coverage -c and not specify some level of endogenous content.");   // :778-780
    if ((!matNoMeth.empty() || !matWiMeth.empty()) && !methyl) return die("Must define --methyl for the fragment selection");            // gargammel.pl:650-660
    if (matNoMeth.empty() != matWiMeth.empty()) return die("Must defined matrix file for both methylated and non-methylated");           // :662-668 (sic)
    if (!matNoMeth.empty() && (dE.any() || dB.any() || dC.any() || dAll.any()))
        return die("Cannot specify as of yet, different deamination profiles for the endogenous, bacterial, contaminant and methylation deamination");   // :670-678
    if (methyl && matNoMeth.empty()) return die("Must define --matfilenonmeth if methylation is used");                                  // :691-695
    if (methyl && matWiMeth.empty()) return die("Must define --matfilemeth if methylation is used");                                     // :697-701
    if (dAll.any()) {                                                              // -matfile/-damage/-mapdamage => endogenous + bacterial only (gargammel.pl:619-648)
        if (dE.any() || dB.any() || dC.any()) return die("Specify either patterns for endogenous and bacterial using either -matfile or -damage but do not specify other parameters");
        dE = dAll; dB = dAll;
    }
    for (size_t i = 0; i < adF.size(); i++) adF[i] = (char)toupper(adF[i]);         // uc(), gargammel.pl:469-475
    for (size_t i = 0; i < adS.size(); i++) adS[i] = (char)toupper(adS[i]);
    if (!seedSpecified) seed = time_seed();
    std::string dir = argv[argc - 1]; if (dir.empty() || dir[dir.size() - 1] != '/') dir += "/";
    if (prefix.empty()) prefix = dir + "simadna";                                   // gargammel.pl usage
    std::string err;

    // ---- scan the three directories (gargammel.pl:817-1181)
    std::vector<std::string> fE = list_fa(dir + "endo/"), fC = list_fa(dir + "cont/"), fB = list_fa(dir + "bact/");
    // multi-cell mode (gargammel.pl:1017-1087): endo/ holds directories C0 .. CN (one cell each, all haploid or all diploid) and no fasta file
    std::vector<std::vector<std::string> > cells;
    if (compE > 0) {
        const std::vector<std::string> cd = list_dirs(dir + "endo/");
        if (!cd.empty()) {
            if (!fE.empty()) return die("The endogenous directory " + dir + "endo/ must directories called CX where X is 0...N without any other fasta files, found: " + std::to_string(fE.size()) + " fasta files");
            long maxc = 0; bool c0 = false;
            for (size_t k = 0; k < cd.size(); k++) {
                const size_t sl = cd[k].rfind('/'); const std::string nm = cd[k].substr(sl + 1);
                bool okn = nm.size() >= 2 && nm[0] == 'C'; for (size_t q = 1; q < nm.size() && okn; q++) okn = nm[q] >= '0' && nm[q] <= '9';
                if (!okn) return die("Directories " + dir + "endo/ does not correspond to the pattern CX where X is 0...N, found: " + cd[k]);
                const long v = strtol(nm.c_str() + 1, NULL, 10); if (v > maxc) maxc = v; if (nm == "C0") c0 = true;
            }
            if (!c0) return die("Directories " + dir + "endo/ does not correspond to the pattern CX where X is 0...N, must start with C0");
            for (long i = 0; i <= maxc; i++) {
                const std::string cdir = dir + "endo/C" + std::to_string(i);
                if (std::find(cd.begin(), cd.end(), cdir) == cd.end()) return die("Missing directory " + cdir);
                cells.push_back(list_fa(cdir + "/"));
                if (i == 0) { if (cells[0].size() != 1 && cells[0].size() != 2) return die("The endogenous directory must have 1 (haploid) or 2 (diploid) files, directory in question: " + cdir + ", found " + std::to_string(cells[0].size())); }
                else if (cells[(size_t)i].size() != cells[0].size()) return die("The endogenous directory must have all have 1 (haploid) or 2 (diploid) files, directory in question: " + cdir + ", found: " + std::to_string(cells[(size_t)i].size()));
            }
            fE = cells[0];                                                          // ploidy and the genome size come from cell 0 (gargammel.pl:1058-1061,1107-1126)
            std::cerr << "Found " << cd.size() << " directories: C[0..C" << maxc << "]" << std::endl;
        }
    }
    if (compE > 0 && fE.size() != 1 && fE.size() != 2) return die("The endogenous directory must have 1 (haploid) or 2 (diploid) files, found: " + std::to_string(fE.size()));
    if (compC > 0 && fC.empty()) return die("If you want contamination, please have at least one file in the cont/ directory");
    std::vector<double> bw;
    if (compB > 0) {
        if (fB.empty()) return die("If you want bacterial contamination, please have at least one file in the bact/ directory");
        std::vector<std::string> names; std::vector<double> w;
        if (!ggh::load_bact_list(dir + "bact/list", names, w, err)) return die(err);
        bw.assign(fB.size(), 0.0); std::vector<int> seen(names.size(), 0);
        for (size_t j = 0; j < fB.size(); j++)
            for (size_t k = 0; k < names.size(); k++) if (dir + "bact/" + names[k] == fB[j]) { bw[j] = w[k]; seen[k]++; }   // counts go to the file NAMED in the list (DESIGN.md deviation from gargammel.pl:1343)
        for (size_t k = 0; k < names.size(); k++) {
            if (seen[k] == 0) return die("Fasta  " + names[k] + " was not found in the directory but it was found in the list");
            if (seen[k] > 1) return die("Fasta  " + names[k] + " was found twice in the list");
        }
    }
    // ---- load genomes (only the sources that get fragments)
    struct Src { std::string path; ggh::GenomeFile g; char kind; int idx; int cell; };
    std::vector<Src> srcs;
    int cell_now = 0;
    auto add = [&](const std::vector<std::string>& files, char kind, bool wanted) -> bool {
        for (size_t i = 0; i < files.size(); i++) {
            Src s; s.path = files[i]; s.kind = kind; s.idx = (int)i; s.cell = cell_now;
            if (wanted) { if (!ggh::load_genome(files[i], s.g, err)) return false; std::cerr << "Found " << (kind == 'e' ? "endogenous" : kind == 'c' ? "contaminant" : "bacterial") << " file " << files[i] << " (" << s.g.total_len << " bp)" << std::endl; }
            else s.g.total_len = 0;
            srcs.push_back(s);
        }
        return true;
    };
    if (cells.empty()) { if (!add(fE, 'e', compE > 0)) return die(err); }
    else for (size_t i = 0; i < cells.size(); i++) { cell_now = (int)i; if (!add(cells[i], 'e', compE > 0)) return die(err); }   // cells in order, each e1 then e2 (gargammel.pl:1464-1470)
    cell_now = 0;
    if (!add(fB, 'b', compB > 0) || !add(fC, 'c', compC > 0)) return die(err);
    uint64_t sumE = 0; std::vector<uint64_t> contLen;
    for (size_t i = 0; i < srcs.size(); i++) { if (srcs[i].kind == 'e' && srcs[i].cell == 0) sumE += srcs[i].g.total_len; if (srcs[i].kind == 'c') contLen.push_back(srcs[i].g.total_len); }

    // ---- size model + average size (gargammel.pl:1185-1229)
    std::vector<int32_t> lens; std::vector<double> cum; double avg = fraglength;
    if (!sizeList.empty()) { if (!ggh::load_size_list(sizeList, lens, err)) return die(err); double t = 0; for (size_t i = 0; i < lens.size(); i++) t += lens[i]; avg = t / lens.size(); }
    else if (!sizeFreq.empty()) { if (!ggh::load_size_freq(sizeFreq, lens, cum, err)) return die(err); double t = 0, prev = 0; for (size_t i = 0; i < lens.size(); i++) { t += lens[i] * (cum[i] - prev); prev = cum[i]; } avg = t / cum.back(); }
    else if (lognorm) { if (!(scale > 0)) return die("-scale must be positive"); if (maxsize > 4095) maxsize = 4095; ggh::lognormal_table(loc, scale, 4095, lens, cum); avg = exp(loc + scale * scale / 2); }   // gargammel.pl:1221-1222
    // ---- number of fragments (gargammel.pl:1236-1266)
    ggh::ApportionIn in; in.compB = compB; in.compC = compC; in.compE = compE; in.n_endo = (int)std::max<size_t>(fE.size(), 1); in.cont_len = contLen; in.bact_w = bw; in.seed = seed; in.n_cells = (int)cells.size();
    if (coverage >= 0) {
        double use = avg; if (se) { if (rl < avg) use = rl; } else if (2 * rl < avg) use = 2 * rl;
        const double sumEeff = fE.size() == 2 ? sumE / 2.0 : (double)sumE;
        const uint64_t nE = (uint64_t)(coverage * (sumEeff / use));
        n = (uint64_t)((double)nE / compE);
        in.has_nE = true; in.nE = nE;                                              // the endogenous count stays what the coverage asks for (gargammel.pl:1255)
    }
    in.n_fragments = n;
    ggh::ApportionOut ap;
    if (!ggh::apportion(in, ap, err)) return die(err);
    std::cerr << "We will generate: \n\n" << ap.nE + ap.nB + ap.nC << "\ttotal fragments\n" << ap.nE << "\tendogenous fragments\n" << ap.nC << "\tcontaminant fragments\n" << ap.nB << "\tbacterial fragments\n";
    if (!cells.empty()) {                                                           // gargammel.pl:1402; cell 0 carries the script's forced extra fragment(s)
        uint64_t sum = 0;
        for (size_t i = 0; i < cells.size(); i++) { std::cerr << "Will extract " << ap.cell_e1[i] + ap.cell_e2[i] << " from cell: " << dir << "endo/C" << i << "/\tendogenous fragments\n"; sum += ap.cell_e1[i] + ap.cell_e2[i]; }
        ap.nE = sum;                                                                // what the script's fragSim calls add up to: nE+1 (haploid) / nE+2 (diploid)
    }
    const uint64_t total = ap.nE + ap.nB + ap.nC;

    // ---- plan in the reference's output order e, b, c with its tag grammar (gargammel.pl:1476,1521,1554,1600,1648; 1698-1830)
    struct PlanItem { size_t src; uint64_t count; std::string tag; int slot; int comp; };
    std::vector<PlanItem> items;
    for (size_t i = 0; i < srcs.size(); i++) {
        const Src& s = srcs[i]; PlanItem it; it.src = i; it.count = 0; it.slot = -1; it.comp = -1;
        if (s.kind == 'e' && !cells.empty()) {                                      // tags e1_<cell> / e2_<cell>, or e for every haploid cell (gargammel.pl:1476,1521,1554)
            it.count = s.idx == 0 ? ap.cell_e1[(size_t)s.cell] : ap.cell_e2[(size_t)s.cell];
            it.tag = fE.size() == 2 ? std::string(s.idx == 0 ? "e1_" : "e2_") + std::to_string(s.cell) : std::string("e"); it.slot = dE.any() ? 0 : -1; it.comp = misE.empty() ? -1 : 0;
        } else
        if (s.kind == 'e') { it.count = (size_t)s.idx < ap.endo.size() ? ap.endo[s.idx] : 0; it.tag = fE.size() == 2 ? (s.idx == 0 ? "e1_0" : "e2_0") : "e"; it.slot = dE.any() ? 0 : -1; it.comp = misE.empty() ? -1 : 0; }
        if (s.kind == 'b') { it.count = (size_t)s.idx < ap.bact.size() ? ap.bact[s.idx] : 0; it.tag = "b" + std::to_string(s.idx + 1); it.slot = dB.any() ? 1 : -1; it.comp = misB.empty() ? -1 : 1; }
        if (s.kind == 'c') { it.count = (size_t)s.idx < ap.cont.size() ? ap.cont[s.idx] : 0; it.tag = "c" + std::to_string(s.idx + 1); it.slot = dC.any() ? 2 : -1; it.comp = misC.empty() ? -1 : 2; }
        if (methyl) it.slot = 0;                                                    // every source goes through deamSim with the two tables (gargammel.pl:1698-1811)
        if (it.count) items.push_back(it);
    }
    if (gpus < 1) gpus = 1;
    const int emit = se ? GG_EMIT_ARTS : GG_EMIT_ARTP;
    // ---- what follows the kernel: art_illumina and the final names, exactly as gargammel.pl:1856-1889 composes them
    if (!art_given.empty() && !is_exe(art_given)) return die("Cannot find file " + art_given);                     // fileExists, gargammel.pl:212-218
    const std::string art = no_art ? std::string() : find_art(art_given, argv[0]);
    const std::string cmd_art = art + " -ss " + ss + " -amp -na " + (se ? "    " : " -p ") + " -i " + prefix + "_a.fa  -l " + std::to_string(rl) + "  -c 1  -qs  " + std::to_string(qs) +
                                "  -qs2 " + std::to_string(qs2) + "  -o " + prefix + "_s";
    bool a_gz_on_device = false;                                    // set by the streaming path once <prefix>_a.fa.gz is complete
    auto post = [&]() -> int {
        if (art.empty()) {
            std::cerr << (no_art ? "--noart: " : "art_illumina not found (--art PATH): ") << "stopping after " << prefix << "_a.fa" << std::endl;
            return 0;
        }
        if (!run_cmd(cmd_art, mock)) return 1;
        if (a_gz_on_device) {                                      // <prefix>_a.fa.gz is already there (members made on the GPU next to <prefix>_a.fa)
            std::cerr << "running cmd gzip -f " << prefix << "_a.fa   (done on the device: removing the plain file)" << std::endl;
            remove((prefix + "_a.fa").c_str());
        } else if (!gzip_file(prefix + "_a.fa", mock)) return 1;
        if (se) { if (!gzip_file(prefix + "_s.fq", mock)) return 1; std::cerr << "Single-end reads available here: " << prefix << "_s.fq.gz" << std::endl; }
        else {
            if (!gzip_file(prefix + "_s1.fq", mock) || !gzip_file(prefix + "_s2.fq", mock)) return 1;
            std::cerr << "Paired-end forward reads available here: " << prefix << "_s1.fq.gz" << std::endl << "Paired-end reverse reads available here: " << prefix << "_s2.fq.gz" << std::endl;
        }
        std::cerr << "Program finished successfully" << std::endl;
        return 0;
    };
    if (mock) {                                                    // print the plan instead of the reference's ~N fragSim/deamSim/adptSim command lines
        uint64_t first = 0;
        for (size_t k = 0; k < items.size(); k++) {
            std::cerr << "fused range ids [" << first << "," << first + items[k].count << ") tag " << items[k].tag << " file " << srcs[items[k].src].path
                      << " damage slot " << items[k].slot << " comp slot " << items[k].comp << std::endl;
            first += items[k].count;
        }
        std::cerr << "running fused kernel fragSim->deamSim->adptSim " << (se ? "-arts " : "-artp ") << prefix << "_a.fa on " << gpus << " GPU(s)" << std::endl;
        return post();
    }
    // ---- one worker per GPU: rank r takes the r-th contiguous slice of the id space; outputs are concatenated in rank order
    std::vector<std::string> shard(gpus), shard_err(gpus); std::vector<int> rcs(gpus, 0);
    std::vector<std::vector<std::string> > shard_keep(gpus, std::vector<std::string>(4));
    // tables + plan of one rank's context; returns false (with shard_err[r]) on failure
    auto setup = [&](int r, gg_ctx* ctx) -> bool {
        auto fail = [&](const char* what) -> bool { shard_err[r] = std::string(what) + ": " + gg_last_error(ctx); return false; };
        std::vector<gg_range> plan; uint64_t first = 0;
        if (methyl) gg_tables_set_case(ctx, 1);                                     // fragSim --case for every source (gargammel.pl:1486,1526,1561,1605,1653)
        for (size_t k = 0; k < items.size(); k++) {
            int gid = -1;
            if (ggh::upload_genome(ctx, srcs[items[k].src].g, &gid) != GG_OK) return fail("genome upload");
            gg_range rg; memset(&rg, 0, sizeof rg); rg.first_id = first; rg.count = items[k].count; rg.genome = gid; rg.damage_slot = items[k].slot; rg.comp_slot = items[k].comp; rg.gc_slot = -1;
            snprintf(rg.tag, sizeof rg.tag, "%s", uniq ? "" : items[k].tag.c_str()); plan.push_back(rg); first += items[k].count;   // -uniq: "_<count>_<tag>" is appended on the host
        }
        const int mode = !sizeList.empty() ? GG_SIZE_LIST : (!sizeFreq.empty() || lognorm) ? GG_SIZE_FREQ : GG_SIZE_FIXED;
        if (gg_tables_set_sizes(ctx, mode, lens.data(), cum.data(), (int)lens.size(), fraglength, minsize, maxsize) != GG_OK) return fail("size tables");
        if ((dE.any() && set_damage(ctx, 0, dE)) || (dB.any() && set_damage(ctx, 1, dB)) || (dC.any() && set_damage(ctx, 2, dC))) return false;
        if (methyl) {
            ggh::Rates n5, n3, w5, w3; std::string e2;
            if (!ggh::load_matrix_dat(matNoMeth, n5, n3, e2) || !ggh::load_matrix_dat(matWiMeth, w5, w3, e2)) { shard_err[r] = e2; return false; }
            if (gg_tables_set_matrix_meth(ctx, 0, n5.data(), (int)(n5.size() / 12), n3.data(), (int)(n3.size() / 12), w5.data(), (int)(w5.size() / 12), w3.data(), (int)(w3.size() / 12), 1) != GG_OK)
                return fail("-matfilemeth/-matfilenonmeth");
        }
        const std::string* mis[3] = {&misE, &misB, &misC};
        for (int m = 0; m < 3; m++) if (!mis[m]->empty()) {                        // fragSim --comp F --dist D per source (gargammel.pl:1492-1495,1615-1618,1664-1667)
            ggh::CompTables ct; std::string e2;
            if (!ggh::load_dnacomp(*mis[m], distmis, ct, e2)) { shard_err[r] = e2; return false; }
            if (gg_tables_set_composition(ctx, m, ct.dist, ct.s5p.data(), ct.s5m.data(), ct.s3p.data(), ct.s3m.data()) != GG_OK) return fail("--misinc");
        }
        if (gg_tables_set_adapters(ctx, adF.c_str(), adS.c_str(), rl) != GG_OK) return fail("adapters");
        if (!plan.empty() && gg_plan_set(ctx, plan.data(), (int)plan.size()) != GG_OK) return fail("plan");
        return true;
    };
    const int emits[3] = {emit, GG_EMIT_FRAG, GG_EMIT_DEAM};
    // Streaming path (everything but -uniq): every rank writes its slice straight to disk through a pinned buffer, the D2H copies
    // overlapped with the kernels and the .gz outputs compressed on the device; nothing is held in host memory.  With several GPUs
    // the ranks write <file>.part<r>, concatenated in rank order afterwards (BGZF members are independent).
    const bool streaming = !uniq;
    const bool a_gz = streaming && !art.empty() && !getenv("GG_HOST_GZIP_A");      // ART will run: <prefix>_a.fa.gz comes from the device too
    const char* gz_names[5] = {"_d.fa.gz", ".e.fa.gz", ".b.fa.gz", ".c.fa.gz", "_a.fa.gz"};
    auto part_name = [&](const std::string& final_name, int r) { return gpus == 1 ? final_name : final_name + ".part" + std::to_string(r); };
    auto stream_worker = [&](int r) {
        gg_ctx* ctx = NULL;
        if (gg_ctx_create(r, seed, &ctx) != GG_OK) { shard_err[r] = "cannot open CUDA device " + std::to_string(r) + " (no CPU fallback)"; rcs[r] = 1; return; }
        void* pin = NULL;
        auto bail = [&](const std::string& m) { if (!m.empty()) shard_err[r] = m; rcs[r] = 1; if (pin) gg_host_free(ctx, pin); gg_ctx_destroy(ctx); };
        if (!setup(r, ctx)) return bail("");
        const uint64_t lo = total * (uint64_t)r / (uint64_t)gpus, hi = total * (uint64_t)(r + 1) / (uint64_t)gpus;
        uint64_t rec_max = 1; for (int p = 0; p < (keep ? 3 : 1); p++) rec_max = std::max<uint64_t>(rec_max, gg_max_record_bytes(ctx, emits[p]));
        const uint64_t chunk = std::max<uint64_t>(1u << 16, std::min<uint64_t>(4u << 20, (1ull << 30) / rec_max));   // <= 1 GiB of pinned memory
        const size_t cap = (size_t)(std::min<uint64_t>(chunk, std::max<uint64_t>(hi - lo, 1)) * rec_max + 65536);
        if (gg_host_alloc(ctx, cap, &pin) != GG_OK) return bail(std::string("pinned buffer: ") + gg_last_error(ctx));
        Sink a; if (!a.open_plain(part_name(prefix + "_a.fa", r))) return bail("Cannot write to file " + prefix + "_a.fa");
        for (uint64_t x = lo; x < hi; x += chunk)
            if (!a.write_run(ctx, x, std::min<uint64_t>(chunk, hi - x), emit, pin, cap)) return bail(std::string("fused kernel / write: ") + gg_last_error(ctx));
        if (!a.close()) return bail("Error: write failed");
        if (a_gz) {                                                // what `gzip -f <prefix>_a.fa` leaves behind after ART (gargammel.pl:1871): made here, on the device
            Sink z; if (!z.open_gz(part_name(prefix + "_a.fa.gz", r))) return bail("Cannot write to file " + prefix + "_a.fa.gz");
            for (uint64_t x = lo; x < hi; x += chunk)
                if (!z.write_run(ctx, x, std::min<uint64_t>(chunk, hi - x), emit, pin, cap)) return bail(std::string("fused kernel / write: ") + gg_last_error(ctx));
            if (!z.close(gpus == 1)) return bail("Error: write failed");
        }
        if (keep) {
            // <prefix>_d.fa.gz = deamSim output in e,b,c order; <prefix>.{e,b,c}.fa.gz = the fragSim outputs per source class
            Sink z[4];
            for (int k = 0; k < 4; k++) if (!z[k].open_gz(part_name(prefix + gz_names[k], r))) return bail("Cannot write intermediate files for prefix " + prefix);
            for (uint64_t x = lo; x < hi; x += chunk)
                if (!z[0].write_run(ctx, x, std::min<uint64_t>(chunk, hi - x), GG_EMIT_DEAM, pin, cap)) return bail(std::string("fused kernel / write: ") + gg_last_error(ctx));
            const uint64_t edge[4] = {0, ap.nE, ap.nE + ap.nB, total};             // ids are in e, b, c order
            for (int k = 0; k < 3; k++) {
                const uint64_t c0 = std::max(lo, edge[k]), c1 = std::min(hi, edge[k + 1]);
                for (uint64_t x = c0; x < c1; x += chunk)
                    if (!z[1 + k].write_run(ctx, x, std::min<uint64_t>(chunk, c1 - x), GG_EMIT_FRAG, pin, cap)) return bail(std::string("fused kernel / write: ") + gg_last_error(ctx));
            }
            for (int k = 0; k < 4; k++) if (!z[k].close(gpus == 1)) return bail("Error: write failed");
        }
        gg_host_free(ctx, pin); gg_ctx_destroy(ctx);
    };
    auto worker = [&](int r) {
        if (streaming) return stream_worker(r);
        gg_ctx* ctx = NULL;
        if (gg_ctx_create(r, seed, &ctx) != GG_OK) { shard_err[r] = "cannot open CUDA device " + std::to_string(r) + " (no CPU fallback)"; rcs[r] = 1; return; }
        auto fail = [&](const char* what) { shard_err[r] = std::string(what) + ": " + gg_last_error(ctx); rcs[r] = 1; gg_ctx_destroy(ctx); };
        if (!setup(r, ctx)) { rcs[r] = 1; gg_ctx_destroy(ctx); return; }
        const uint64_t lo = total * (uint64_t)r / (uint64_t)gpus, hi = total * (uint64_t)(r + 1) / (uint64_t)gpus;
        const uint64_t chunk = 8u << 20;
        for (int pass = 0; pass < (keep ? 3 : 1); pass++)
            for (uint64_t a = lo; a < hi; a += chunk) {
                const uint64_t cnt = std::min<uint64_t>(chunk, hi - a);
                gg_out o;
                if (gg_run_fused(ctx, a, cnt, emits[pass], &o) != GG_OK) return fail("fused kernel");
                std::string& dst = pass == 0 ? shard[r] : shard_keep[r][pass];
                const size_t old = dst.size(); dst.resize(old + o.bytes); size_t got = 0;
                if (gg_out_fetch(ctx, &o, &dst[old], o.bytes, &got) != GG_OK) return fail("fetch");
            }
        gg_ctx_destroy(ctx);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < gpus; r++) th.push_back(std::thread(worker, r));
    for (int r = 0; r < gpus; r++) th[r].join();
    for (int r = 0; r < gpus; r++) if (rcs[r]) return die("Error: " + shard_err[r]);
    if (streaming) {
        if (gpus > 1) {                                            // concatenate the parts in rank order
            std::vector<char> buf(64u << 20);
            for (int k = -1; k < 5; k++) {
                if (k >= 0 && (k == 4 ? !a_gz : !keep)) continue;
                const std::string fin = prefix + (k < 0 ? "_a.fa" : gz_names[k]);
                FILE* out = fopen(fin.c_str(), "wb"); if (!out) return die("Cannot write to file " + fin);
                for (int r = 0; r < gpus; r++) {
                    const std::string pn = part_name(fin, r);
                    FILE* in2 = fopen(pn.c_str(), "rb"); if (!in2) return die("Cannot read " + pn);
                    size_t got; while ((got = fread(buf.data(), 1, buf.size(), in2)) > 0) if (fwrite(buf.data(), 1, got, out) != got) return die("Error: write failed");
                    fclose(in2); remove(pn.c_str());
                }
                if (k >= 0) { static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0}; fwrite(eof, 1, 28, out); }
                if (fclose(out) != 0) return die("Error: write failed");
            }
        }
        std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << total << " sequences to " << prefix << "_a.fa" << std::endl;
        a_gz_on_device = a_gz;
        return post();
    }
    // -uniq: one defline counter per source file, exactly one fragSim invocation each in the reference
    auto uniq_pass = [&](std::vector<std::string>& parts) {
        std::string all; for (size_t r = 0; r < parts.size(); r++) { all += parts[r]; std::string().swap(parts[r]); }
        std::string out2; out2.reserve(all.size() + all.size() / 16);
        UniqSuffixer u; size_t p = 0;
        for (size_t k = 0; k < items.size(); k++) {
            u.reset(items[k].tag, true);
            size_t q = p; uint64_t lines = 0;
            while (q < all.size() && lines < 2 * items[k].count) { const char* e = (const char*)memchr(all.data() + q, '\n', all.size() - q); q = e ? (size_t)(e - all.data()) + 1 : all.size(); lines++; }
            u.apply(all.data() + p, q - p, out2); p = q;
        }
        parts.assign(1, out2);
    };
    if (uniq) {
        uniq_pass(shard);
        if (keep) for (int pass = 1; pass < 3; pass++) { std::vector<std::string> v; for (int r = 0; r < gpus; r++) v.push_back(shard_keep[r][pass]); uniq_pass(v); for (int r = 0; r < gpus; r++) shard_keep[r][pass] = r == 0 ? v[0] : std::string(); }
    }
    Sink out;
    if (!out.open_plain(prefix + "_a.fa")) return die("Cannot write to file " + prefix + "_a.fa");
    for (size_t r = 0; r < shard.size(); r++) if (!out.write(shard[r].data(), shard[r].size())) return die("Error: write failed");
    if (!out.close()) return die("Error: write failed");
    if (keep) {
        // <prefix>_d.fa.gz = deamSim output in e,b,c order; <prefix>.{e,b,c}.fa.gz = the fragSim outputs per source class
        Sink d; if (!d.open_gz(prefix + "_d.fa.gz")) return die("Cannot write to file " + prefix + "_d.fa.gz");
        for (int r = 0; r < gpus; r++) d.write(shard_keep[r][2].data(), shard_keep[r][2].size());
        d.close();
        std::string all; for (int r = 0; r < gpus; r++) all += shard_keep[r][1];
        Sink se_, sb_, sc_;
        if (!se_.open_gz(prefix + ".e.fa.gz") || !sb_.open_gz(prefix + ".b.fa.gz") || !sc_.open_gz(prefix + ".c.fa.gz")) return die("Cannot write intermediate files for prefix " + prefix);
        size_t p = 0; uint64_t rec = 0;
        while (p < all.size()) {                                   // records are in e,b,c order: split by count
            size_t e1 = all.find('\n', p); size_t e2 = e1 == std::string::npos ? e1 : all.find('\n', e1 + 1);
            if (e2 == std::string::npos) break;
            Sink& s = rec < ap.nE ? se_ : rec < ap.nE + ap.nB ? sb_ : sc_;
            s.write(all.data() + p, e2 + 1 - p); p = e2 + 1; rec++;
        }
        se_.close(); sb_.close(); sc_.close();
    }
    std::cerr << "Program " << argv[0] << " terminated succesfully, wrote " << total << " sequences to " << prefix << "_a.fa" << std::endl;
    return post();
}
