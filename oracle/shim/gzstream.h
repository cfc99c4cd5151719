/* gzstream.h -- SHIM over zlib written for this repo (upstream gzstream is not vendored in the reference). */
#ifndef SHIM_GZSTREAM_H
#define SHIM_GZSTREAM_H
#include <zlib.h>
#include <iostream>
#include <streambuf>
#include <cstring>
class gzstreambuf : public std::streambuf {
    static const int BUFSZ = 1 << 16;
    gzFile f; char buf[BUFSZ]; bool opened, writing;
public:
    gzstreambuf() : f(NULL), opened(false), writing(false) { setp(buf, buf + BUFSZ - 1); setg(buf, buf, buf); }
    ~gzstreambuf() { close(); }
    bool is_open() const { return opened; }
    gzstreambuf* open(const char* name, std::ios::openmode mode) {
        if (opened) return NULL;
        writing = (mode & std::ios::out) != 0;
        f = gzopen(name, writing ? ((mode & std::ios::app) ? "ab" : "wb") : "rb");
        if (!f) return NULL;
        opened = true; setp(buf, buf + BUFSZ - 1); setg(buf, buf, buf); return this;
    }
    gzstreambuf* close() { if (opened) { sync(); opened = false; if (gzclose(f) == Z_OK) return this; } return NULL; }
    int flush_buffer() { int w = int(pptr() - pbase()); if (w && gzwrite(f, pbase(), w) != w) return EOF; pbump(-w); return w; }
    virtual int overflow(int c = EOF) {
        if (!opened || !writing) return EOF;
        if (c != EOF) { *pptr() = char(c); pbump(1); }
        return flush_buffer() == EOF ? EOF : c;
    }
    virtual int underflow() {
        if (gptr() && gptr() < egptr()) return *reinterpret_cast<unsigned char*>(gptr());
        if (!opened || writing) return EOF;
        int n = gzread(f, buf, BUFSZ);
        if (n <= 0) return EOF;
        setg(buf, buf, buf + n);
        return *reinterpret_cast<unsigned char*>(gptr());
    }
    virtual int sync() { if (opened && writing && pptr() && pptr() > pbase()) { if (flush_buffer() == EOF) return -1; } return 0; }
};
class gzstreambase : virtual public std::ios {
protected: gzstreambuf buf;
public:
    gzstreambase() { init(&buf); }
    void open(const char* name, std::ios::openmode mode) { if (!buf.open(name, mode)) clear(rdstate() | std::ios::badbit); else clear(); }
    void close() { if (buf.is_open()) if (!buf.close()) clear(rdstate() | std::ios::badbit); }
    gzstreambuf* rdbuf() { return &buf; }
};
class igzstream : public gzstreambase, public std::istream {
public:
    igzstream() : std::istream(&buf) {}
    void open(const char* name, std::ios::openmode mode = std::ios::in) { gzstreambase::open(name, mode); }
};
class ogzstream : public gzstreambase, public std::ostream {
public:
    ogzstream() : std::ostream(&buf) {}
    void open(const char* name, std::ios::openmode mode = std::ios::out) { gzstreambase::open(name, mode); }
};
#endif
