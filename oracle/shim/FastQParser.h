/* FastQParser.h -- SHIM written for this repo (upstream class lives in un-vendored libgab).
 * FastQParser(file, isFasta): reads (optionally gzipped) FASTA/FASTQ; getID() keeps the leading '>' / '@'
 * (deamSim.cpp:2038 prints it raw).  FASTA sequences may span several lines. */
#ifndef SHIM_FASTQPARSER_H
#define SHIM_FASTQPARSER_H
#include <zlib.h>
#include <string>
#include <iostream>
#include <cstdlib>
class FastQObj {
    std::string id, seq, qual;
public:
    friend class FastQParser;
    std::string* getID() { return &id; }
    std::string* getSeq() { return &seq; }
    std::string* getQual() { return &qual; }
};
class FastQParser {
    gzFile f; bool fasta; FastQObj cur; std::string pending; bool havePending, eof;
    bool readLine(std::string& line) {
        line.clear(); char buf[1 << 16];
        while (true) {
            if (!gzgets(f, buf, sizeof buf)) return !line.empty();
            line += buf;
            if (!line.empty() && line[line.size() - 1] == '\n') { line.erase(line.size() - 1); if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1); return true; }
        }
    }
public:
    FastQParser(const std::string& name, bool isFasta) : fasta(isFasta), havePending(false), eof(false) {
        f = gzopen(name.c_str(), "rb");
        if (!f) { std::cerr << "FastQParser: cannot open " << name << std::endl; exit(1); }
        gzbuffer(f, 1 << 20);
    }
    ~FastQParser() { if (f) gzclose(f); }
    bool hasData() {
        std::string line;
        if (havePending) { line = pending; havePending = false; }
        else { do { if (!readLine(line)) return false; } while (line.empty()); }
        cur.id = line; cur.seq.clear(); cur.qual.clear();
        if (fasta) {
            while (readLine(line)) { if (!line.empty() && line[0] == '>') { pending = line; havePending = true; break; } cur.seq += line; }
        } else {
            std::string plus;
            if (!readLine(cur.seq) || !readLine(plus) || !readLine(cur.qual)) { std::cerr << "FastQParser: truncated fastq record " << cur.id << std::endl; exit(1); }
        }
        return true;
    }
    FastQObj* getData() { return &cur; }
};
#endif
