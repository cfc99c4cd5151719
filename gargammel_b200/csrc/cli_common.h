// cli_common.h -- shared plumbing of the drop-in executables (bin/fragSim, bin/deamSim, bin/adptSim, bin/gargammel-b200).
// Host side in C++11 above the C-ABI (include/gargammel_b200.h); mirrors the reference's CLI conventions:
// options over argv[1..argc-2], last argv is the input, errors on stderr + exit 1 (fragSim.cpp:463-471,622-623).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <sys/time.h>
#include <unistd.h>
#include <sys/types.h>
#include <limits.h>
#include <string>
#include <algorithm>
#include <vector>
#include <iostream>
#include <unordered_map>
#include <deque>
#include <thread>
#include <mutex>
#include <condition_variable>
#include "gg_host.h"

namespace cli {

inline bool is_help(int argc, char** argv) {
    return argc == 1 || (argc == 2 && (!strcmp(argv[1], "-h") || !strcmp(argv[1], "-help") || !strcmp(argv[1], "--help")));
}
inline int die(const std::string& msg) { std::cerr << msg << std::endl; return 1; }
inline int unsupported(const std::string& opt, const char* why) {
    std::cerr << "Error: option " << opt << " is not supported by the B200 build (" << why << ")" << std::endl;
    return 1;
}
// destringify<T>(argv) of the reference: exit(1) on garbage
inline long long to_ll(const char* s, const char* opt) {
    char* e; long long v = strtoll(s, &e, 10);
    if (e == s) { std::cerr << "Error: cannot convert \"" << s << "\" for option " << opt << std::endl; exit(1); }
    return v;
}
inline double to_d(const char* s, const char* opt) {
    char* e; double v = strtod(s, &e);
    if (e == s) { std::cerr << "Error: cannot convert \"" << s << "\" for option " << opt << std::endl; exit(1); }
    return v;
}
inline uint64_t time_seed() {       // the reference seeds from the clock when --seed is absent (fragSim.cpp:713-715)
    timeval t; gettimeofday(&t, NULL);
    return (uint64_t)t.tv_sec * 1000000ull + (uint64_t)t.tv_usec;
}
inline std::string exe_dir(const char* argv0) {     // getCWD(argv[0]) of libgab: directory of the executable, trailing '/'
    // the path the program was called by wins (a src/deamSim symlink in a gargammel tree finds src/matrices/ next to it); a bare name
    // found through $PATH resolves through /proc/self/exe
    char buf[PATH_MAX]; std::string p = argv0 ? argv0 : "";
    if (p.find('/') == std::string::npos) {
        ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
        if (n > 0) { buf[n] = 0; p = buf; }
    }
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string("./") : p.substr(0, k + 1);
}

// output sink: stdout, a plain file, or a gzip file (ogzstream of the reference).  gz output is written as BGZF
// (blocked gzip: any gzip reader inflates it, htslib can index it), deflated on all host threads.
struct Sink {
    FILE* f; bool own, bgzf; std::vector<uint8_t> pend;
    Sink() : f(NULL), own(false), bgzf(false) {}
    bool open_stdout() { f = stdout; own = false; setvbuf(stdout, NULL, _IOFBF, 1 << 22); return true; }
    bool open_plain(const std::string& p) { f = fopen(p.c_str(), "wb"); own = true; if (f) setvbuf(f, NULL, _IOFBF, 1 << 22); return f != NULL; }
    bool open_gz(const std::string& p, bool append = false) { f = fopen(p.c_str(), append ? "ab" : "wb"); own = true; bgzf = true; return f != NULL; }
    bool flush_blocks(bool final, bool eof_marker = true) {
        const size_t B = 0xff00, whole = final ? pend.size() : pend.size() / B * B;
        if (whole) {
            std::vector<uint8_t> z; std::string e;
            if (!ggh::bgzf_compress(pend.data(), whole, 0, 1, z, false, e)) return false;
            if (fwrite(z.data(), 1, z.size(), f) != z.size()) return false;
            pend.erase(pend.begin(), pend.begin() + whole);
        }
        if (final && eof_marker) {
            std::vector<uint8_t> z; std::string e;
            if (!ggh::bgzf_compress(NULL, 0, 1, 1, z, true, e)) return false;     // EOF marker block
            if (fwrite(z.data(), 1, z.size(), f) != z.size()) return false;
        }
        return true;
    }
    // A large piece for a plain file of our own: disjoint slices written with pwrite by a few threads.  One thread copying into the page
    // cache moves 1-2 GB/s -- the fused driver's kernel produces a 0.73 GB piece in 60 ms, so behind a fast file system (tmpfs, striped
    // NVMe) the single copy would be the whole wall clock.  Pipes and stdout (not seekable), small pieces and .gz sinks keep fwrite.
    bool write_big(const void* d, size_t n) {
        const char* env = getenv("GG_PWRITE_MIN");
        const size_t min_bytes = env ? (size_t)strtoull(env, NULL, 10) : ((size_t)32 << 20);
        if (bgzf || !own || !f || n < min_bytes || n < 2) return write(d, n);
        if (fflush(f) != 0) return false;
        const off_t pos = ftello(f);
        if (pos < 0) return write(d, n);
        const int fd = fileno(f);
        const size_t nt = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency())), n / std::max<size_t>(1, min_bytes / 2) + 1));
        const size_t per = (n + nt - 1) / nt;
        std::vector<int> ok(nt, 1); std::vector<std::thread> th;
        auto work = [&](size_t t) {
            size_t a = std::min(n, t * per); const size_t b = std::min(n, a + per);
            while (a < b) {
                const ssize_t w = pwrite(fd, (const char*)d + a, std::min<size_t>(b - a, (size_t)64 << 20), pos + (off_t)a);
                if (w <= 0) { ok[t] = 0; return; }
                a += (size_t)w;
            }
        };
        for (size_t t = 1; t < nt; t++) th.push_back(std::thread(work, t));
        work(0);
        for (size_t t = 0; t < th.size(); t++) th[t].join();
        for (size_t t = 0; t < nt; t++) if (!ok[t]) return false;
        return fseeko(f, pos + (off_t)n, SEEK_SET) == 0;
    }
    bool write(const void* d, size_t n) {
        if (!bgzf) return fwrite(d, 1, n, f) == n;
        pend.insert(pend.end(), (const uint8_t*)d, (const uint8_t*)d + n);
        return pend.size() >= (8u << 20) ? flush_blocks(false) : true;
    }
    // a stream still on the device (the gg_out of a gg_run_*): plain sinks fetch and write it; gz sinks compress it ON THE DEVICE
    // (gg_bgzf_dev) and fetch only the BGZF members.  GG_HOST_BGZF=1 keeps the host zlib path (for comparison).
    bool write_out(gg_ctx* ctx, const gg_out& o, std::vector<uint8_t>& host) {
        size_t n = 0;
        if (!bgzf || getenv("GG_HOST_BGZF")) {
            host.resize(o.bytes);
            if (gg_out_fetch(ctx, &o, host.data(), host.size(), &n) != GG_OK) return false;
            return write(host.data(), n);
        }
        if (!pend.empty()) {                                               // members are independent: flush what the host path buffered
            std::vector<uint8_t> z; std::string e;
            if (!ggh::bgzf_compress(pend.data(), pend.size(), 0, 1, z, false, e) || fwrite(z.data(), 1, z.size(), f) != z.size()) return false;
            pend.clear();
        }
        gg_out z;
        if (gg_bgzf_dev(ctx, &o, &z) != GG_OK) return false;
        host.resize(z.bytes);
        if (gg_out_fetch(ctx, &z, host.data(), host.size(), &n) != GG_OK) return false;
        return fwrite(host.data(), 1, n, f) == n;
    }
    // eof_marker = false: a part that will be concatenated with others (BGZF members are independent; one marker ends the file)
    bool close(bool eof_marker = true) {
        bool ok = true;
        if (f) { if (bgzf) ok = flush_blocks(true, eof_marker); ok = (fflush(f) == 0) && ok; if (own) fclose(f); f = NULL; }
        return ok;
    }
    // device-resident run -> this sink through a pinned buffer, chunk by chunk with the copies overlapped (gg_run_fused[_bgzf]_to_host)
    bool write_run(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, void* pinned, size_t cap) {
        gg_out o;
        if (bgzf && !getenv("GG_HOST_BGZF")) {
            if (!pend.empty() && !flush_blocks(true, false)) return false;
            if (gg_run_fused_bgzf_to_host(ctx, first, count, emit, pinned, cap, 1u << 19, &o) != GG_OK) return false;
            return fwrite(pinned, 1, o.bytes, f) == o.bytes;
        }
        if (gg_run_fused_to_host(ctx, first, count, emit, pinned, cap, 1u << 20, &o) != GG_OK) return false;
        return write_big(pinned, o.bytes);
    }
};

// Read a FASTA record stream as the reference's FastQParser(file, true) does (gz or plain) and normalise it to
// 2-line records "ID\nSEQ\n" (multi-line sequences are joined).  Returns false on error.
inline bool read_records(const std::string& path, std::vector<uint8_t>& out, uint64_t* n_rec, std::string& err) {
    std::vector<uint8_t> raw;
    if (!ggh::read_file(path, raw, err)) return false;
    *n_rec = 0;
    // fast path: already strictly 2-line
    bool two_line = true; size_t p = 0, lines = 0;
    while (p < raw.size()) {
        const uint8_t* e = (const uint8_t*)memchr(raw.data() + p, '\n', raw.size() - p);
        size_t le = e ? (size_t)(e - raw.data()) : raw.size();
        const bool hdr = raw[p] == '>';
        if ((lines & 1) == 0 ? !hdr : hdr) { two_line = false; break; }
        if (le > p && raw[le - 1] == '\r') { two_line = false; break; }
        lines++; p = le + 1;
    }
    if (two_line && (lines & 1) == 0) {
        if (!raw.empty() && raw.back() != '\n') raw.push_back('\n');
        out.swap(raw); *n_rec = lines / 2; return true;
    }
    out.clear(); out.reserve(raw.size() + 1);
    p = 0; bool have = false, in_seq = false;
    while (p < raw.size()) {
        const uint8_t* e = (const uint8_t*)memchr(raw.data() + p, '\n', raw.size() - p);
        size_t le = e ? (size_t)(e - raw.data()) : raw.size();
        size_t l1 = le; if (l1 > p && raw[l1 - 1] == '\r') l1--;
        if (l1 > p && raw[p] == '>') {
            if (have) { if (!in_seq) {} out.push_back('\n'); }
            out.insert(out.end(), raw.begin() + p, raw.begin() + l1); out.push_back('\n');
            have = true; in_seq = true; (*n_rec)++;
        } else if (l1 > p) {
            if (!have) { err = "input does not start with a '>' record: " + path; return false; }
            out.insert(out.end(), raw.begin() + p, raw.begin() + l1);
        }
        p = le + 1;
    }
    if (have) out.push_back('\n');
    return true;
}

// split a normalised 2-line stream into chunks of at most `max_bytes` on record boundaries; returns (offset, bytes, records)
struct Chunk { size_t off, bytes; uint64_t recs; };
inline std::vector<Chunk> chunk_records(const std::vector<uint8_t>& s, size_t max_bytes) {
    std::vector<Chunk> out; size_t start = 0, p = 0; uint64_t recs = 0, lines = 0;
    while (p < s.size()) {
        const uint8_t* e = (const uint8_t*)memchr(s.data() + p, '\n', s.size() - p);
        p = e ? (size_t)(e - s.data()) + 1 : s.size();
        if ((++lines & 1) == 0) {
            recs++;
            if (p - start >= max_bytes) { Chunk c = {start, p - start, recs}; out.push_back(c); start = p; recs = 0; }
        }
    }
    if (p > start) { Chunk c = {start, p - start, recs}; out.push_back(c); }
    return out;
}

// ---- streaming: deamSim / adptSim read their input chunk by chunk (bounded memory, whatever the file size) while the GPU works on the
// previous chunk and a writer thread drains the one before: reader thread -> [chunks] -> main thread (H2D, kernel, D2H) -> [outputs] -> writer
template <class T> struct BoundedQueue {
    std::mutex m; std::condition_variable cv; std::deque<T> q; size_t cap; bool closed;
    explicit BoundedQueue(size_t c) : cap(c), closed(false) {}
    void push(T&& v) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return q.size() < cap || closed; }); if (!closed) q.push_back(std::move(v)); cv.notify_all(); }
    bool pop(T& v) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !q.empty() || closed; }); if (q.empty()) return false; v = std::move(q.front()); q.pop_front(); cv.notify_all(); return true; }
    void close() { std::unique_lock<std::mutex> l(m); closed = true; cv.notify_all(); }
    // close and drop what is queued (consumer gave up)
    void abort() { std::unique_lock<std::mutex> l(m); closed = true; q.clear(); cv.notify_all(); }
};
struct InChunk { std::vector<uint8_t> bytes; uint64_t recs; InChunk() : recs(0) {} };
struct OutChunk { std::vector<uint8_t> bytes; int sink; bool raw; OutChunk() : sink(0), raw(false) {} };   // raw: BGZF members made on the device, written as they are

// Reads a FASTA record stream as the reference's FastQParser(file, true) does (gz or plain), normalises it to 2-line records
// "ID\nSEQ\n" (multi-line sequences are joined, '\r' dropped) and hands it on in chunks of about `target` bytes cut on record boundaries.
inline void record_chunk_reader(const std::string path, size_t target, BoundedQueue<InChunk>* q, std::string* err, uint64_t* n_rec) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) { *err = "Unable to open file " + path; q->close(); return; }
    gzbuffer(f, 1 << 20);
    std::vector<uint8_t> raw(8u << 20); std::vector<uint8_t> carry;            // carry: the unfinished last line of the previous block
    InChunk cur; cur.bytes.reserve(target + (1u << 20));
    bool have = false; *n_rec = 0;
    auto line = [&](const uint8_t* p, size_t len) -> bool {
        if (len && p[len - 1] == '\r') len--;
        if (len && p[0] == '>') {
            if (have) cur.bytes.push_back('\n');                                 // ends the previous record's (joined) sequence line
            if (have && cur.bytes.size() >= target) { q->push(std::move(cur)); cur = InChunk(); cur.bytes.reserve(target + (1u << 20)); }
            cur.bytes.insert(cur.bytes.end(), p, p + len); cur.bytes.push_back('\n');
            have = true; cur.recs++; (*n_rec)++;
        } else if (len) {
            if (!have) { *err = "input does not start with a '>' record: " + path; return false; }
            cur.bytes.insert(cur.bytes.end(), p, p + len);
        }
        return true;
    };
    bool ok = true;
    while (ok) {
        const int n = gzread(f, raw.data(), (unsigned)raw.size());
        if (n < 0) { *err = "Read error in " + path; ok = false; break; }
        if (n == 0) break;
        size_t p = 0;
        while (p < (size_t)n) {
            const uint8_t* e = (const uint8_t*)memchr(raw.data() + p, '\n', (size_t)n - p);
            if (!e) { carry.insert(carry.end(), raw.begin() + p, raw.begin() + n); break; }
            const size_t le = (size_t)(e - raw.data());
            if (!carry.empty()) { carry.insert(carry.end(), raw.begin() + p, raw.begin() + le); ok = line(carry.data(), carry.size()); carry.clear(); }
            else ok = line(raw.data() + p, le - p);
            if (!ok) break;
            p = le + 1;
        }
    }
    if (ok && !carry.empty()) ok = line(carry.data(), carry.size());
    gzclose(f);
    if (ok) { if (have) cur.bytes.push_back('\n'); if (!cur.bytes.empty()) q->push(std::move(cur)); }
    q->close();
}
// the writer thread: drains the outputs in order into their sinks
inline void chunk_writer(BoundedQueue<OutChunk>* q, std::vector<Sink*> sinks, bool* failed) {
    OutChunk c;
    while (q->pop(c)) {
        Sink* s = sinks[c.sink];
        const bool ok = c.raw ? fwrite(c.bytes.data(), 1, c.bytes.size(), s->f) == c.bytes.size() : s->write(c.bytes.data(), c.bytes.size());
        if (!ok) { *failed = true; q->abort(); return; }
    }
}
// a device-resident result -> an OutChunk: plain sinks get the bytes, gz sinks BGZF members compressed on the device
inline bool fetch_out(gg_ctx* ctx, const gg_out& o, const Sink& sink, int sink_index, OutChunk* oc) {
    size_t n = 0; oc->sink = sink_index;
    if (!sink.bgzf || getenv("GG_HOST_BGZF")) {
        oc->raw = false; oc->bytes.resize(o.bytes);
        return gg_out_fetch(ctx, &o, oc->bytes.data(), oc->bytes.size(), &n) == GG_OK;
    }
    gg_out z;
    if (gg_bgzf_dev(ctx, &o, &z) != GG_OK) return false;
    oc->raw = true; oc->bytes.resize(z.bytes);
    return gg_out_fetch(ctx, &z, oc->bytes.data(), oc->bytes.size(), &n) == GG_OK;
}
constexpr size_t STREAM_CHUNK_BYTES = 64u << 20;

// fragSim -uniq (fragSim.cpp:119-135,1579-1584): every defline gets "_<k>" where k counts how often that exact defline has
// been produced so far, then "_<tag>" if a tag was given.  Like the reference this is a host-side map over all deflines
// ("note: this might not be practical for large datasets", fragSim.cpp:429-430); it runs over the byte stream the GPU wrote
// (2-line records whose deflines carry no tag).
struct UniqSuffixer {
    std::unordered_map<std::string, unsigned int> seen; std::string tag; bool tagb;
    UniqSuffixer() : tagb(false) {}
    void reset(const std::string& t, bool has_tag) { seen.clear(); tag = t; tagb = has_tag; }
    // rewrites `in` (whole records) into out; header lines are those at even line index
    void apply(const char* in, size_t n, std::string& out) {
        size_t p = 0; bool hdr = true;
        while (p < n) {
            const char* e = (const char*)memchr(in + p, '\n', n - p);
            const size_t le = e ? (size_t)(e - in) : n;
            if (hdr) {
                std::string key(in + p + 1, le - p - 1);                 // defline without '>'
                unsigned int& c = seen[key]; c++;
                out.append(in + p, le - p); out += '_'; out += std::to_string(c);
                if (tagb) { out += '_'; out += tag; }
                out += '\n';
            } else { out.append(in + p, le - p); out += '\n'; }
            hdr = !hdr; p = le + 1;
        }
    }
};

// Unaligned BAM out (fragSim -b, deamSim -b, adptSim -bs/-bp; -u = members stored instead of deflated): records are serialised on
// the host, the BGZF members are made by the device encoder (gg_bgzf_host; GG_HOST_BGZF=1: zlib on host threads)
struct BamSink {
    FILE* f; bool stored; std::vector<uint8_t> pend, host; uint64_t n_rec;
    BamSink() : f(NULL), stored(false), n_rec(0) {}
    bool open(const std::string& path, const ggh::BamHeader& h, bool uncompressed) {
        f = fopen(path.c_str(), "wb"); stored = uncompressed;
        if (f) ggh::bam_header_bytes(h, pend);
        return f != NULL;
    }
    bool flush(gg_ctx* ctx, bool final) {
        const size_t B = 0xff00, whole = final ? pend.size() : pend.size() / B * B;
        if (whole) {
            if (stored || getenv("GG_HOST_BGZF")) {
                std::vector<uint8_t> z; std::string e;
                if (!ggh::bgzf_compress(pend.data(), whole, 0, stored ? 0 : 1, z, false, e) || fwrite(z.data(), 1, z.size(), f) != z.size()) return false;
            } else {
                gg_out o; size_t n = 0;
                if (gg_bgzf_host(ctx, pend.data(), whole, &o) != GG_OK) return false;
                host.resize(o.bytes);
                if (gg_out_fetch(ctx, &o, host.data(), host.size(), &n) != GG_OK || fwrite(host.data(), 1, n, f) != n) return false;
            }
            pend.erase(pend.begin(), pend.begin() + whole);
        }
        if (final) {
            std::vector<uint8_t> z; std::string e;
            if (!ggh::bgzf_compress(NULL, 0, 1, 1, z, true, e) || fwrite(z.data(), 1, z.size(), f) != z.size()) return false;     // EOF marker block
        }
        return true;
    }
    bool add(gg_ctx* ctx, const ggh::BamRecord& r) { ggh::bam_append_record(r, pend); n_rec++; return pend.size() >= (64u << 20) ? flush(ctx, false) : true; }
    bool close(gg_ctx* ctx) { bool ok = f && flush(ctx, true); if (f) { ok = (fclose(f) == 0) && ok; f = NULL; } return ok; }
};
// the records of a 2-line stream, one callback per record: (id line without its first byte, sequence)
template <class F> inline void for_each_record(const uint8_t* d, size_t n, F fn) {
    size_t p = 0;
    while (p < n) {
        const uint8_t* e = (const uint8_t*)memchr(d + p, '\n', n - p); const size_t ie = e ? (size_t)(e - d) : n;
        const size_t ss = std::min(ie + 1, n);
        const uint8_t* e2 = (const uint8_t*)memchr(d + ss, '\n', n - ss); const size_t se = e2 ? (size_t)(e2 - d) : n;
        fn(std::string((const char*)d + p + 1, ie - p - 1), std::string((const char*)d + ss, se - ss));
        p = se + 1;
    }
}
inline std::string command_line(int argc, char** argv) { std::string c; for (int i = 0; i < argc; i++) { c += argv[i]; c += " "; } return c; }

// GG_PRINT_RSS=1: peak / current resident set of this process on stderr (the streaming CLIs promise a bounded one)
inline void print_rss(const char* where) {
    if (!getenv("GG_PRINT_RSS")) return;
    FILE* f = fopen("/proc/self/status", "r"); if (!f) return;
    char line[256];
    while (fgets(line, sizeof line, f)) if (!strncmp(line, "VmHWM", 5) || !strncmp(line, "VmRSS", 5) || !strncmp(line, "RssAnon", 7) || !strncmp(line, "RssFile", 7) || !strncmp(line, "RssShmem", 8)) std::cerr << "[" << where << "] " << line;
    fclose(f);
}
inline int gg_fail(gg_ctx* ctx, const char* what) {
    std::cerr << "Error: " << what << ": " << gg_last_error(ctx) << std::endl; return 1;
}
inline int open_ctx(uint64_t seed, gg_ctx** ctx) {
    int dev = 0; const char* d = getenv("GG_DEVICE"); if (d) dev = atoi(d);
    if (gg_ctx_create(dev, seed, ctx) != GG_OK) {
        std::cerr << "Error: cannot open CUDA device " << dev << " (this build has no CPU fallback)" << std::endl; return 1;
    }
    return 0;
}

} // namespace cli
