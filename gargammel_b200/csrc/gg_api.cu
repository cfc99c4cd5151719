// gg_api.cu -- implementation of the C-ABI in include/gargammel_b200.h.
// Host side of the device path: owns the CUDA context state (stream, events, HBM-resident genomes and tables),
// turns the reference's floating-point tables into exact integer thresholds, sizes shared memory, launches the
// kernels in gg_kernels.cu.  There is no CPU fallback: without a CUDA device every gg_run_* fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/gargammel_b200.h"
#include "gg_host.h"
#include "gg_kernels.cuh"

using namespace ggk;

namespace {

struct DBuf {                       // grow-only device buffer
    void* p; size_t cap;
    DBuf() : p(nullptr), cap(0) {}
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = n + (n >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; cudaGetLastError(); e = cudaMalloc(&p, n); want = n; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return (T*)p; }
};

struct HostContig { std::string name; uint64_t len, offset; int32_t blen, llen; uint64_t gbase, cum_end; int fai_index; };
struct HostGenome {
    std::vector<HostContig> contigs;      // name-sorted (std::map order, fragSim.cpp:215,1152)
    std::vector<int> fai2sorted;
    uint64_t G, n_units;
    uint32_t* d_seq2; uint32_t* d_nmask; size_t seq2_bytes, nmask_bytes;
    uint32_t* d_cmask; size_t cmask_bytes;                       // fragSim --case: 1 bit per base, 1 = lower-case (null without --case)
    int first_contig;
    int circ_sorted; uint32_t wrap_s0; uint64_t wrap_gbase;     // --circ: index of the circular contig in `contigs` (-1 none) + its wrap image
};
struct HostDamage {
    int mode, rows5, rows3, clamp_last;
    int n5t, n3t;                                  // matrix: terminal rows drawn per base (draw specification v2)
    std::vector<uint4> term, acc; std::vector<uint32_t> sd5, sd3, geoL, geoV, geoA, geoB; std::vector<uint8_t> guideA, guideB; uint32_t Ks, Kd, K3;
    double pL, pV, pA, pB;                         // parameters of the geometric tables geoL/geoV and geoA (matrix: majorant M; Briggs: s) / geoB (Briggs: d)
    DBuf d_term, d_acc, d_sd5, d_sd3, d_geoL, d_geoV, d_geoA, d_geoB, d_guideA, d_guideB;
    HostDamage() : mode(GG_DMG_NONE), rows5(0), rows3(0), clamp_last(1), n5t(0), n3t(0), Ks(0), Kd(0), K3(0), pL(0), pV(0), pA(0), pB(0) {}
};
const int NT_MAX = 8;                // at most this many terminal rows per end are drawn per base
const double NT_KAPPA = 1.25;        // a row is terminal when its total rate exceeds NT_KAPPA x the largest rate of the rows >= NT_MAX

const int GEO_TABLE = 4096;          // geometric inverse-CDF table length (overhangs / nick position)
const int MAX_FRAG_LEN = 4095;

// p = (u31 + 0.5) / 2^31.   number of u31 with p <= c   (cum[a] <= p <= cum[a+1] picks, deamSim.cpp:1850; p <= cProb, fragSim.cpp:165)
uint32_t thr_le(double c) {
    if (!(c == c)) return 0;
    const double y = c * 2147483648.0 - 0.5;
    if (y < 0.0) return 0;
    const double k = floor(y) + 1.0;
    return k >= 2147483648.0 ? 2147483648u : (uint32_t)k;
}
// number of u31 with p < c   (randomProb() < rate, deamSim.cpp:1488-1587,1948-1993)
uint32_t thr_lt(double c) {
    if (!(c == c)) return 0;
    const double y = c * 2147483648.0 - 0.5;
    if (y <= 0.0) return 0;
    const double k = ceil(y);
    return k >= 2147483648.0 ? 2147483648u : (uint32_t)k;
}

// inverse-CDF table of geometric_distribution<int>(p): CDF(k) = 1-(1-p)^(k+1) as a running product (deamSim.cpp:297,1478-1479,1503-1507)
void geo_table(double p, int n, std::vector<uint32_t>& t) {
    t.resize(n);
    double q = 1.0; uint32_t prev = 0;
    for (int k = 0; k < n; k++) { q *= (1.0 - p); uint32_t v = thr_le(1.0 - q); if (v < prev) v = prev; t[k] = v; prev = v; }
}
// 512-entry guide of such a table: guide[c] = first k with (c << 22) < t[k], capped at 255 (the device scans upwards from there)
void geo_guide(const std::vector<uint32_t>& t, std::vector<uint8_t>& g) {
    g.resize(512);
    size_t k = 0;
    for (uint32_t c = 0; c < 512; c++) {
        const uint32_t lo = c << 22;
        while (k < t.size() && lo >= t[k]) k++;
        g[c] = (uint8_t)std::min<size_t>(k, 255);
    }
}

} // namespace

struct gg_ctx {
    int device; cudaStream_t st; cudaEvent_t ev0, ev1;
    cudaStream_t st_copy; cudaEvent_t ev_k[2], ev_c[2];     // egress pipeline of gg_run_fused_to_host
    int out_slab;                                          // which of d_out/d_out2 the next fused run writes
    uint64_t seed; std::string err;
    int sm_count; size_t smem_optin;
    // genomes
    std::vector<HostGenome> genomes;
    DBuf d_cum_end, d_gbase, d_clen, d_name_off, d_name_len, d_names, d_genomes;
    bool ctab_dirty;
    int n_contigs_total, names_bytes;
    // sizes
    int size_mode, fixed_len, minlen, maxlen, norev;
    int keep_case;                                         // fragSim --case
    std::vector<int32_t> size_len; std::vector<uint32_t> size_thr;
    DBuf d_size_len, d_size_thr;
    bool sizes_set;
    // damage
    HostDamage dmg[GG_MAX_DAMAGE_SLOTS];
    // base composition (fragSim --comp)
    struct HostComp { int dist; double maxP, maxM; DBuf d_tab; } comp[GG_MAX_COMP_SLOTS];
    // GC bias (fragSim -gc)
    struct HostGc { bool on; int dist, genome, wmax; double bias, mean, sd; DBuf d_thr; } gc[GG_MAX_GC_SLOTS];
    // adapters
    std::string adF, adS; int read_len;
    DBuf d_adF2, d_adS2, d_rcS2, d_adF, d_adS;
    bool ad_dirty, ad_acgt;                            // ad_acgt: both adapters are upper-case ACGT (what the 2-bit fused path can carry)
    AdptNaming naming; int deam_name;
    // plan
    std::vector<gg_range> plan; DBuf d_ranges; bool plan_dirty;
    // run state
    std::vector<std::pair<void*, size_t> > pool;     // device blocks of cleared genomes, reused by the next uploads
    DBuf d_tmp_fa, d_tmp_desc;                       // upload staging (raw FASTA bytes, contig descriptors)
    DBuf d_out, d_out2, d_tile_state, d_ticket_stats, d_in, d_nl, d_scan, d_sizes, d_flush;
    DBuf d_case, d_cmask_ptrs;                       // fragSim --case: per-fragment descriptors of a run, case-mask pointer per genome
    DBuf d_chain_a, d_chain_b; DBuf* fused_dst;      // methylation plans run as chained stage kernels: intermediate streams; where gg_run_fused writes (null = the slabs)
    DBuf d_bgzf, d_bgzf2, d_bgzf_tab, d_bgzf_state;  // device BGZF: output slabs (ping-pong), code/CRC tables, look-back state + histogram
    int bgzf_slab, bgzf_reuse_code, bgzf_hdr_nbits;  // pipeline state of gg_run_fused_bgzf_to_host
    std::vector<uint32_t> bgzf_static;               // CRC tables and shift polynomials (built once)
    void* h_pinned; size_t h_pinned_cap;
    unsigned long long* h_pub; unsigned long long* d_pub;   // 512 u64 of mapped pinned memory: small results published by k_publish
    double last_ms;

    int fail(int code, const std::string& msg) { err = msg; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e); cudaGetLastError(); return GG_ERR_CUDA;
    }
};

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ctx->cuda_fail(e__, #call); } while (0)

namespace {

// n u64 words of device memory -> host, through mapped pinned memory (see k_publish); syncs the compute stream
int fetch_small(gg_ctx* ctx, const void* dsrc, unsigned long long* host, int n, cudaEvent_t after = nullptr) {
    if (n > 512) return ctx->fail(GG_ERR_ARG, "fetch_small: too many words");
    if (!ctx->h_pub) {
        CK(cudaHostAlloc((void**)&ctx->h_pub, 512 * 8, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&ctx->d_pub, ctx->h_pub, 0));
    }
    CK(launch_publish((const unsigned long long*)dsrc, ctx->d_pub, n, ctx->st));
    if (after) CK(cudaEventRecord(after, ctx->st));                         // the publish kernel belongs to the timed run
    CK(cudaStreamSynchronize(ctx->st));
    memcpy(host, ctx->h_pub, (size_t)n * 8);
    return GG_OK;
}

template <typename T> int upload_vec(gg_ctx* ctx, DBuf& b, const std::vector<T>& v, size_t slack_elems = 0) {
    const size_t bytes = (v.size() + slack_elems) * sizeof(T);
    CK(b.ensure(bytes ? bytes : sizeof(T)));
    if (slack_elems) CK(cudaMemsetAsync(b.p, 0, bytes, ctx->st));
    if (!v.empty()) CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));     // v may be a temporary
    return GG_OK;
}

// 2-bit packing with one guard word in front (device pointers are taken at word 1)
std::vector<uint32_t> pack2(const std::string& s) {
    std::vector<uint32_t> w(1 + (s.size() + 15) / 16, 0u);
    for (size_t i = 0; i < s.size(); i++) {
        uint32_t c = s[i] == 'A' ? 0u : s[i] == 'C' ? 1u : s[i] == 'G' ? 2u : 3u;
        w[1 + (i >> 4)] |= c << (2 * (i & 15));
    }
    return w;
}
std::string revcomp(const std::string& s) {
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); i++) {
        char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    return r;
}

int sync_contig_tables(gg_ctx* ctx) {
    if (!ctx->ctab_dirty) return GG_OK;
    std::vector<uint64_t> cum_end, gbase; std::vector<uint32_t> clen, name_off; std::vector<uint16_t> name_len; std::vector<char> names;
    std::vector<GenomeDev> gd;
    for (size_t g = 0; g < ctx->genomes.size(); g++) {
        HostGenome& H = ctx->genomes[g];
        H.first_contig = (int)cum_end.size();
        for (size_t i = 0; i < H.contigs.size(); i++) {
            const HostContig& c = H.contigs[i];
            cum_end.push_back(c.cum_end); gbase.push_back(c.gbase); clen.push_back((uint32_t)c.len);
            name_off.push_back((uint32_t)names.size()); name_len.push_back((uint16_t)c.name.size());
            names.insert(names.end(), c.name.begin(), c.name.end());
        }
        GenomeDev d; d.seq2 = H.d_seq2; d.nmask = H.d_nmask; d.G = H.G; d.first_contig = H.first_contig; d.n_contigs = (int)H.contigs.size();
        d.circ_contig = H.circ_sorted >= 0 ? H.first_contig + H.circ_sorted : -1; d.wrap_s0 = H.wrap_s0; d.wrap_gbase = H.wrap_gbase;
        gd.push_back(d);
    }
    ctx->n_contigs_total = (int)cum_end.size(); ctx->names_bytes = (int)names.size();
    int rc;
    if ((rc = upload_vec(ctx, ctx->d_cum_end, cum_end))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_gbase, gbase))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_clen, clen))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_name_off, name_off))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_name_len, name_len))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_names, names, 16))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_genomes, gd))) return rc;
    ctx->ctab_dirty = false;
    return GG_OK;
}

int sync_adapters(gg_ctx* ctx) {
    if (!ctx->ad_dirty) return GG_OK;
    int rc;
    if ((rc = upload_vec(ctx, ctx->d_adF2, pack2(ctx->adF), 3))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_adS2, pack2(ctx->adS), 3))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_rcS2, pack2(revcomp(ctx->adS)), 3))) return rc;
    std::vector<uint8_t> f(ctx->adF.begin(), ctx->adF.end()), s(ctx->adS.begin(), ctx->adS.end());
    if ((rc = upload_vec(ctx, ctx->d_adF, f, 16))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_adS, s, 16))) return rc;
    ctx->ad_dirty = false;
    return GG_OK;
}

int sync_plan(gg_ctx* ctx) {
    if (!ctx->plan_dirty) return GG_OK;
    std::vector<RangeDev> r(ctx->plan.size());
    for (size_t i = 0; i < r.size(); i++) {
        memset(&r[i], 0, sizeof(RangeDev));
        r[i].first_id = ctx->plan[i].first_id; r[i].count = ctx->plan[i].count;
        r[i].genome = ctx->plan[i].genome; r[i].slot = ctx->plan[i].damage_slot;
        r[i].gc = (ctx->plan[i].gc_slot >= 0 && ctx->gc[ctx->plan[i].gc_slot].on) ? ctx->plan[i].gc_slot : -1;
        r[i].comp = (ctx->plan[i].comp_slot >= 0 && ctx->comp[ctx->plan[i].comp_slot].dist > 0) ? ctx->plan[i].comp_slot : -1;
        r[i].tag_len = (int)strnlen(ctx->plan[i].tag, sizeof(ctx->plan[i].tag) - 1);
        memcpy(r[i].tag, ctx->plan[i].tag, r[i].tag_len);
    }
    int rc = upload_vec(ctx, ctx->d_ranges, r);
    if (rc) return rc;
    ctx->plan_dirty = false;
    return GG_OK;
}

void fill_damage_dev(const HostDamage& H, ggd::DamageDev* D) {
    memset(D, 0, sizeof *D);
    D->mode = H.mode; D->rows5 = H.rows5; D->rows3 = H.rows3; D->clamp_last = H.clamp_last;
    D->term = H.d_term.as<uint4>(); D->acc = H.d_acc.as<uint4>(); D->n5t = H.n5t; D->n3t = H.n3t;
    D->sd5 = H.d_sd5.as<uint32_t>(); D->sd3 = H.d_sd3.as<uint32_t>();
    D->Ks = H.Ks; D->Kd = H.Kd; D->geoL = H.d_geoL.as<uint32_t>(); D->geoV = H.d_geoV.as<uint32_t>(); D->n_geo = GEO_TABLE;
    D->geoM = H.d_geoA.as<uint32_t>(); D->geoS = H.d_geoA.as<uint32_t>(); D->geoD = H.d_geoB.as<uint32_t>();
    D->guideA = H.d_guideA.as<uint8_t>(); D->guideB = H.d_guideB.as<uint8_t>();
    // -damagess keeps its 3' overhang table where plain Briggs keeps the nick table
    D->K3 = H.mode == GG_DMG_BRIGGS_SS ? H.K3 : H.Ks; D->geoL3 = H.mode == GG_DMG_BRIGGS_SS ? D->geoV : D->geoL;
    auto inv_log = [](double p) { return p <= 0.0 ? -INFINITY : p >= 1.0 ? -0.0f : (float)(1.0 / log(1.0 - p)); };   // p = 0: never a success (k = n); p = 1: always k = 0
    D->inv_logL = inv_log(H.pL); D->inv_logV = inv_log(H.pV);
    D->inv_logL3 = H.mode == GG_DMG_BRIGGS_SS ? D->inv_logV : D->inv_logL;
}

int digits10(uint64_t v) { int n = 1; while (v >= 10) { v /= 10; n++; } return n; }

// largest fragment length the size model can emit
int max_frag_len(const gg_ctx* ctx) {
    if (ctx->size_mode == GG_SIZE_FIXED) return (ctx->fixed_len >= ctx->minlen && ctx->fixed_len <= ctx->maxlen) ? ctx->fixed_len : -1;
    int m = -1;                                  // -1: no admissible length at all
    for (size_t i = 0; i < ctx->size_len.size(); i++)
        if (ctx->size_len[i] >= 0 && ctx->size_len[i] >= ctx->minlen && ctx->size_len[i] <= ctx->maxlen) m = std::max(m, ctx->size_len[i]);
    return m;
}
int emit_lines(int emit) { return emit == GG_EMIT_STDOUT4 ? 2 : 1; }
int emit_seqlen(int emit, int L, int RL) { return emit <= GG_EMIT_DEAM ? L : (emit == GG_EMIT_ARTP ? 2 * RL : RL); }

// upper bounds over the current plan: header bytes, record bytes
void plan_bounds(const gg_ctx* ctx, int emit, int Lmax, int* hdr_max, uint64_t* rec_max) {
    size_t name = 1, tag = 0; uint64_t clen = 1, cend = 1;
    for (size_t i = 0; i < ctx->plan.size(); i++) {
        const gg_range& r = ctx->plan[i];
        tag = std::max(tag, strnlen(r.tag, sizeof(r.tag) - 1));
        if (r.genome >= 0 && r.genome < (int)ctx->genomes.size()) {
            const HostGenome& HG = ctx->genomes[r.genome];
            for (size_t c = 0; c < HG.contigs.size(); c++) {
                name = std::max(name, HG.contigs[c].name.size());
                clen = std::max(clen, HG.contigs[c].len);
                // the defline's end is start+length, un-wrapped on a --circ contig (fragSim.cpp:1489): it can pass the contig length by a fragment
                cend = std::max(cend, HG.contigs[c].len + ((int)c == HG.circ_sorted ? (uint64_t)std::max(Lmax, 1) + 2 * GG_MAX_COMP_DIST + 1 : 0));
            }
        }
    }
    const int h = 1 + (int)name + 3 + digits10(clen) + 1 + digits10(cend) + 1 + digits10((uint64_t)std::max(Lmax, 1)) + (int)tag;
    *hdr_max = h;
    const int S = emit_seqlen(emit, Lmax, ctx->read_len);
    uint64_t b = 0;
    for (int ln = 0; ln < emit_lines(emit); ln++) b += (uint64_t)h + 2 + 1 + (uint64_t)S + 1;   // +2: "_r"
    *rec_max = b;
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

// best-fit block from the context's pool of cleared genome storage, else cudaMalloc (cudaFree/cudaMalloc cost tens of ms)
cudaError_t pool_alloc(gg_ctx* ctx, size_t bytes, void** p, size_t* got) {
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); i++)
        if (ctx->pool[i].second >= bytes && (best < 0 || ctx->pool[i].second < ctx->pool[best].second)) best = (int)i;
    if (best >= 0 && ctx->pool[best].second <= 2 * bytes + (1u << 20)) {
        *p = ctx->pool[best].first; *got = ctx->pool[best].second;
        ctx->pool.erase(ctx->pool.begin() + best);
        return cudaSuccess;
    }
    *got = bytes;
    return cudaMalloc(p, bytes);
}

// shared-memory carve-up: one region per warp holding a tile of F (<= 32) fragments, plus the CTA-shared tables
SmemLayout make_layout(int F, uint64_t rec_max, int max_chunks, int e_frags, int line_words, int warps, int thr_bytes, int flags /* 1: staging aliases the fragment words, 2: per-fragment Briggs header words (-damagess) */,
                       int tab_ranges = 0, int tab_genomes = 0, int tab_contigs = 0, int tab_names = 0, bool tab_on = false,
                       int tab_sizes = 0, int adF_words = 0, int adS_words = 0) {
    SmemLayout s; int o = align_up(fused_desc_bytes(), 16);   // fixed-layout per-fragment descriptors (WarpDesc, gg_kernels.cu)
    o += 16;                                   // readable slack before staging / guard word in front of the damaged-fragment words
    // the damaged-fragment words sit at a FIXED offset behind the descriptors (an immediate in the kernel's addressing).  adptSim
    // dialects: staging aliases them (the words are dead once the line images exist, before staging is written) and is made large
    // enough to hold them; fragSim / deamSim records: staging follows them
    s.dwords = o;
    const bool alias_dwords = (flags & 1) != 0;
    s.stage_bytes = (int)((uint64_t)std::min(F, e_frags) * rec_max);      // one emission part (half a tile by default)
    const int dw_bytes = 4 * (F * max_chunks + 4);
    if (alias_dwords) { s.stage = o; s.stage_bytes = std::max(s.stage_bytes, dw_bytes); }
    else { o += align_up(dw_bytes, 16) + 16; s.stage = o; }
    o += align_up(s.stage_bytes, 16) + 48;     // (+ misalignment pad (<16) + read slack)
    s.bh = o; if (flags & 2) o += 8 * F;
    s.lwords = o; o += 4 * (F * line_words + 4);
    s.cmap = o; o += F * max_chunks;
    s.gmap = o;
    s.warp_bytes = align_up(o, 16);
    s.thr = 0; s.thr_bytes = thr_bytes; s.warp0 = align_up(thr_bytes, 16);
    s.tab = -1; s.tab_bytes = 0; s.n_contigs = tab_contigs; s.names_bytes = tab_names;
    s.tab_ranges = s.tab_genomes = s.tab_cum = s.tab_gbase = s.tab_name8 = s.tab_len = s.tab_noff = s.tab_nlen = s.tab_names = 0;
    s.tab_sthr = s.tab_slen = s.n_sizes_sm = 0; for (int k = 0; k < 3; k++) { s.tab_ad[k] = 0; s.ad_words[k] = 0; }
    if (tab_on) {                              // small plan / genome / contig tables, CTA-shared
        int t = 0;
        s.tab_cum = t; t += 8 * tab_contigs; s.tab_gbase = t; t += 8 * tab_contigs; s.tab_name8 = t; t += 8 * tab_contigs;
        s.tab_ranges = t; t += (int)sizeof(RangeDev) * tab_ranges; s.tab_genomes = t; t += (int)sizeof(GenomeDev) * tab_genomes;
        s.tab_len = t; t += 4 * tab_contigs; s.tab_noff = t; t += 4 * tab_contigs; s.tab_nlen = t; t += align_up(2 * tab_contigs, 4);
        s.tab_names = t; t += tab_names; t = align_up(t, 4);
        s.tab_sthr = t; t += 4 * tab_sizes; s.tab_slen = t; t += 4 * tab_sizes; s.n_sizes_sm = tab_sizes;
        s.ad_words[0] = adF_words; s.ad_words[1] = s.ad_words[2] = adS_words;
        for (int k = 0; k < 3; k++) { s.tab_ad[k] = t; t += 4 * s.ad_words[k]; }
        s.tab = s.warp0; s.tab_bytes = align_up(t, 16); s.warp0 += s.tab_bytes;
    }
    for (int i = 0; i < 4; i++) s.thr_slot[i] = s.term_slot[i] = -1;
    s.total = s.warp0 + warps * s.warp_bytes;
    return s;
}

} // namespace

extern "C" {

int gg_abi_version(void) { return GG_ABI_VERSION; }

int gg_ctx_create(int device, uint64_t seed, gg_ctx** out) {
    if (!out) return GG_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) { cudaGetLastError(); return GG_ERR_CUDA; }   // no GPU: fail loudly, never fall back
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return GG_ERR_CUDA; }
    gg_ctx* ctx = new gg_ctx();
    ctx->device = device; ctx->seed = seed; ctx->st = nullptr; ctx->ev0 = ctx->ev1 = nullptr; ctx->st_copy = nullptr; ctx->out_slab = 0;
    for (int i = 0; i < 2; i++) ctx->ev_k[i] = ctx->ev_c[i] = nullptr;
    ctx->ctab_dirty = true; ctx->n_contigs_total = 0;
    ctx->size_mode = GG_SIZE_FIXED; ctx->fixed_len = 20; ctx->minlen = 0; ctx->maxlen = 1000; ctx->norev = 0; ctx->sizes_set = false; ctx->keep_case = 0; ctx->fused_dst = nullptr;
    ctx->adF = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG";   // adptSim.cpp:28
    ctx->adS = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT";         // adptSim.cpp:29
    ctx->read_len = 100; ctx->ad_dirty = true; ctx->ad_acgt = true; ctx->plan_dirty = true; memset(&ctx->naming, 0, sizeof ctx->naming); ctx->deam_name = 0;                  // adptSim.cpp:31
    ctx->h_pinned = nullptr; ctx->h_pinned_cap = 0; ctx->last_ms = 0.0; ctx->h_pub = ctx->d_pub = nullptr;
    ctx->bgzf_slab = 0; ctx->bgzf_reuse_code = 0; ctx->bgzf_hdr_nbits = 0;
    for (int i = 0; i < GG_MAX_GC_SLOTS; i++) { ctx->gc[i].on = false; ctx->gc[i].wmax = -1; }
    for (int i = 0; i < GG_MAX_COMP_SLOTS; i++) { ctx->comp[i].dist = 0; ctx->comp[i].maxP = ctx->comp[i].maxM = 1.0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_k[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_k[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_c[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_c[1], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); delete ctx; return GG_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount; ctx->smem_optin = prop.sharedMemPerBlockOptin;
    *out = ctx;
    return GG_OK;
}

int gg_ctx_destroy(gg_ctx* ctx) {
    if (!ctx) return GG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    for (size_t g = 0; g < ctx->genomes.size(); g++) { cudaFree(ctx->genomes[g].d_seq2); cudaFree(ctx->genomes[g].d_nmask); if (ctx->genomes[g].d_cmask) cudaFree(ctx->genomes[g].d_cmask); }
    DBuf* bufs[] = {&ctx->d_cum_end, &ctx->d_gbase, &ctx->d_clen, &ctx->d_name_off, &ctx->d_name_len, &ctx->d_names, &ctx->d_genomes,
                    &ctx->d_size_len, &ctx->d_size_thr, &ctx->d_adF2, &ctx->d_adS2, &ctx->d_rcS2, &ctx->d_adF, &ctx->d_adS, &ctx->d_ranges,
                    &ctx->d_tmp_fa, &ctx->d_tmp_desc, &ctx->d_out, &ctx->d_out2, &ctx->d_tile_state, &ctx->d_ticket_stats, &ctx->d_in, &ctx->d_nl, &ctx->d_scan, &ctx->d_sizes, &ctx->d_flush,
                    &ctx->d_case, &ctx->d_cmask_ptrs, &ctx->d_chain_a, &ctx->d_chain_b,
                    &ctx->d_bgzf, &ctx->d_bgzf2, &ctx->d_bgzf_tab, &ctx->d_bgzf_state};
    for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); i++) bufs[i]->release();
    for (int s = 0; s < GG_MAX_COMP_SLOTS; s++) ctx->comp[s].d_tab.release();
    for (int s = 0; s < GG_MAX_GC_SLOTS; s++) ctx->gc[s].d_thr.release();
    for (int s = 0; s < GG_MAX_DAMAGE_SLOTS; s++) {
        ctx->dmg[s].d_term.release(); ctx->dmg[s].d_acc.release(); ctx->dmg[s].d_sd5.release(); ctx->dmg[s].d_sd3.release(); ctx->dmg[s].d_geoL.release(); ctx->dmg[s].d_geoV.release(); ctx->dmg[s].d_geoA.release(); ctx->dmg[s].d_geoB.release(); ctx->dmg[s].d_guideA.release(); ctx->dmg[s].d_guideB.release();
    }
    for (size_t i = 0; i < ctx->pool.size(); i++) cudaFree(ctx->pool[i].first);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_pub) cudaFreeHost(ctx->h_pub);
    cudaStreamSynchronize(ctx->st_copy);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_k[i]); cudaEventDestroy(ctx->ev_c[i]); }
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaStreamDestroy(ctx->st); cudaStreamDestroy(ctx->st_copy);
    delete ctx;
    return GG_OK;
}

int gg_ctx_set_seed(gg_ctx* ctx, uint64_t seed) { if (!ctx) return GG_ERR_ARG; ctx->seed = seed; return GG_OK; }
const char* gg_last_error(gg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* gg_ctx_stream(gg_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }
double gg_last_device_ms(gg_ctx* ctx) { return ctx ? ctx->last_ms : 0.0; }

// ------------------------------------------------------------------------------------------------ genome
int gg_genome_upload(gg_ctx* ctx, const uint8_t* fasta, size_t n, const gg_fai_entry* fai, int n_chr, int* genome_id_out) {
    return gg_genome_upload_circ(ctx, fasta, n, fai, n_chr, -1, genome_id_out);
}
int gg_genome_upload_circ(gg_ctx* ctx, const uint8_t* fasta, size_t n, const gg_fai_entry* fai, int n_chr, int circ_chr, int* genome_id_out) {
    if (!ctx || !fasta || !fai || n_chr <= 0 || circ_chr >= n_chr) return ctx ? ctx->fail(GG_ERR_ARG, "gg_genome_upload: bad arguments") : GG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    HostGenome H; H.d_seq2 = H.d_nmask = nullptr; H.d_cmask = nullptr; H.cmask_bytes = 0; H.first_contig = 0; H.circ_sorted = -1; H.wrap_s0 = 0; H.wrap_gbase = 0;
    std::vector<HostContig> cs;
    for (int i = 0; i < n_chr; i++) {
        HostContig c; c.name = fai[i].name ? fai[i].name : "";
        if (fai[i].len <= 0 || fai[i].line_blen <= 0 || fai[i].line_len < fai[i].line_blen) return ctx->fail(GG_ERR_ARG, "gg_genome_upload: bad .fai entry for " + c.name);
        if ((uint64_t)fai[i].len > 0xFFFFFFFFull) return ctx->fail(GG_ERR_ARG, "gg_genome_upload: contig longer than 2^32-1 bases: " + c.name);
        if (c.name.size() > 4096) return ctx->fail(GG_ERR_ARG, "gg_genome_upload: contig name too long");
        c.len = (uint64_t)fai[i].len; c.offset = fai[i].offset; c.blen = fai[i].line_blen; c.llen = fai[i].line_len; c.fai_index = i;
        c.gbase = c.cum_end = 0;
        cs.push_back(c);
    }
    // std::map<string,...>::insert keeps the first of duplicate names and iterates in lexicographic order (fragSim.cpp:215,266,1152)
    std::stable_sort(cs.begin(), cs.end(), [](const HostContig& a, const HostContig& b) { return a.name < b.name; });
    H.fai2sorted.assign(n_chr, -1);
    uint64_t G = 0, gb = 64;                 // 64 guard bases so that reverse-strand windows never index below 0
    for (size_t i = 0; i < cs.size(); i++) {
        if (i && cs[i].name == cs[i - 1].name) continue;
        HostContig c = cs[i];
        G += c.len; c.cum_end = G; c.gbase = gb; gb += (c.len + 63) / 64 * 64;
        H.fai2sorted[c.fai_index] = (int)H.contigs.size();
        H.contigs.push_back(c);
    }
    // --circ (fragSim.cpp:1171-1178): the named contig (a duplicate name resolves to the map's entry, i.e. the first) gets a
    // wrap image behind the last contig: K bases before its last base, then its first K+1 bases, K = longest window
    int wrap_desc = -1; uint32_t wrap_keff = 0, wrap_srclen = 0, wrap_len = 0;
    if (circ_chr >= 0) {
        for (size_t i = 0; i < H.contigs.size(); i++) if (H.contigs[i].name == (fai[circ_chr].name ? fai[circ_chr].name : "")) H.circ_sorted = (int)i;
        if (H.circ_sorted < 0) return ctx->fail(GG_ERR_ARG, "gg_genome_upload_circ: circular contig not found");
        const HostContig& cc = H.contigs[H.circ_sorted];
        const uint32_t K = (uint32_t)(MAX_FRAG_LEN + 2 * GG_MAX_COMP_DIST + 1);
        wrap_srclen = (uint32_t)cc.len; wrap_keff = std::min<uint32_t>(K, wrap_srclen - 1); wrap_len = wrap_keff + K + 1;
        H.wrap_s0 = wrap_srclen - 1 - wrap_keff; H.wrap_gbase = gb; gb += ((uint64_t)wrap_len + 63) / 64 * 64;
    }
    H.G = G; H.n_units = gb / 32 + 2;
    // device: raw FASTA (temporary), contig descriptors (temporary), packed arrays (resident)
    const size_t nc = H.contigs.size() + (circ_chr >= 0 ? 1 : 0);
    std::vector<uint64_t> off(nc), gbase(nc); std::vector<uint32_t> len(nc); std::vector<int32_t> blen(nc), llen(nc);
    for (size_t i = 0; i < H.contigs.size(); i++) { off[i] = H.contigs[i].offset; gbase[i] = H.contigs[i].gbase; len[i] = (uint32_t)H.contigs[i].len; blen[i] = H.contigs[i].blen; llen[i] = H.contigs[i].llen; }
    if (circ_chr >= 0) {
        const HostContig& cc = H.contigs[H.circ_sorted];
        wrap_desc = (int)nc - 1; off[nc - 1] = cc.offset; gbase[nc - 1] = H.wrap_gbase; len[nc - 1] = wrap_len; blen[nc - 1] = cc.blen; llen[nc - 1] = cc.llen;
    }
    // staging (reused across uploads): raw FASTA bytes + contig descriptors
    const size_t desc_bytes = nc * (8 + 8 + 4 + 4 + 4) + 64;
    CK(ctx->d_tmp_fa.ensure(n + 16));
    CK(ctx->d_tmp_desc.ensure(desc_bytes));
    uint8_t* d_fa = ctx->d_tmp_fa.as<uint8_t>();
    uint64_t* d_off = ctx->d_tmp_desc.as<uint64_t>(); uint64_t* d_gb = d_off + nc;
    uint32_t* d_len = (uint32_t*)(d_gb + nc); int32_t* d_bl = (int32_t*)(d_len + nc); int32_t* d_ll = d_bl + nc;
    void* p1 = nullptr; void* p2 = nullptr; void* p3 = nullptr;
    cudaError_t e = pool_alloc(ctx, (H.n_units * 2 + 4) * 4, &p1, &H.seq2_bytes);
    if (e == cudaSuccess) e = pool_alloc(ctx, (H.n_units + 8) * 4, &p2, &H.nmask_bytes);   // slack: pass A reads the mask in aligned 16-byte pieces
    if (e == cudaSuccess && ctx->keep_case) e = pool_alloc(ctx, (H.n_units + 4) * 4, &p3, &H.cmask_bytes);   // --case: one more bit per base
    H.d_seq2 = (uint32_t*)p1; H.d_nmask = (uint32_t*)p2; H.d_cmask = (uint32_t*)p3;
#define TRY(x) if (e == cudaSuccess) e = (x)
    TRY(cudaMemcpyAsync(d_fa, fasta, n, cudaMemcpyHostToDevice, ctx->st));
    TRY(cudaMemcpyAsync(d_off, off.data(), nc * 8, cudaMemcpyHostToDevice, ctx->st));
    TRY(cudaMemcpyAsync(d_gb, gbase.data(), nc * 8, cudaMemcpyHostToDevice, ctx->st));
    TRY(cudaMemcpyAsync(d_len, len.data(), nc * 4, cudaMemcpyHostToDevice, ctx->st));
    TRY(cudaMemcpyAsync(d_bl, blen.data(), nc * 4, cudaMemcpyHostToDevice, ctx->st));
    TRY(cudaMemcpyAsync(d_ll, llen.data(), nc * 4, cudaMemcpyHostToDevice, ctx->st));
    // every unit (incl. the guard units in front and the slack behind) is written by the pack kernel
    TRY(launch_pack_genome(d_fa, n, d_off, d_len, d_bl, d_ll, d_gb, (int)nc, wrap_desc, wrap_keff, wrap_srclen, H.n_units, H.d_seq2, H.d_nmask, H.d_cmask, ctx->st));
    TRY(cudaMemsetAsync(H.d_seq2 + H.n_units * 2, 0, 16, ctx->st));
    TRY(cudaMemsetAsync(H.d_nmask + H.n_units, 0xFF, 32, ctx->st));
    if (H.d_cmask) TRY(cudaMemsetAsync(H.d_cmask + H.n_units, 0, 16, ctx->st));
    TRY(cudaStreamSynchronize(ctx->st));
#undef TRY
    if (e != cudaSuccess) { if (p1) cudaFree(p1); if (p2) cudaFree(p2); if (p3) cudaFree(p3); return ctx->cuda_fail(e, "gg_genome_upload"); }
    ctx->genomes.push_back(H);
    ctx->ctab_dirty = true;
    if (genome_id_out) *genome_id_out = (int)ctx->genomes.size() - 1;
    return GG_OK;
}
int gg_genome_count(gg_ctx* ctx) { return ctx ? (int)ctx->genomes.size() : 0; }
int gg_genome_clear(gg_ctx* ctx) {
    if (!ctx) return GG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st));
    for (size_t g = 0; g < ctx->genomes.size(); g++) {
        ctx->pool.push_back(std::make_pair((void*)ctx->genomes[g].d_seq2, ctx->genomes[g].seq2_bytes));
        ctx->pool.push_back(std::make_pair((void*)ctx->genomes[g].d_nmask, ctx->genomes[g].nmask_bytes));
        if (ctx->genomes[g].d_cmask) ctx->pool.push_back(std::make_pair((void*)ctx->genomes[g].d_cmask, ctx->genomes[g].cmask_bytes));
    }
    ctx->genomes.clear(); ctx->plan.clear(); ctx->ctab_dirty = true; ctx->plan_dirty = true;
    return GG_OK;
}

int gg_genome_fetch(gg_ctx* ctx, int genome_id, int chr, int64_t start, int len, char* dst) {
    if (!ctx || genome_id < 0 || genome_id >= (int)ctx->genomes.size() || !dst || len < 0) return ctx ? ctx->fail(GG_ERR_ARG, "gg_genome_fetch: bad arguments") : GG_ERR_ARG;
    const HostGenome& H = ctx->genomes[genome_id];
    if (chr < 0 || chr >= (int)H.fai2sorted.size() || H.fai2sorted[chr] < 0) return ctx->fail(GG_ERR_ARG, "gg_genome_fetch: bad contig");
    const HostContig& c = H.contigs[H.fai2sorted[chr]];
    if (start < 0 || (uint64_t)start + (uint64_t)len > c.len) return ctx->fail(GG_ERR_ARG, "gg_genome_fetch: range outside contig");
    if (len == 0) return GG_OK;
    CK(cudaSetDevice(ctx->device));
    const uint64_t b0 = c.gbase + (uint64_t)start, b1 = b0 + (uint64_t)len - 1;
    const uint64_t w0 = b0 >> 4, w1 = b1 >> 4, m0 = b0 >> 5, m1 = b1 >> 5;
    std::vector<uint32_t> w(w1 - w0 + 1), m(m1 - m0 + 1);
    CK(cudaMemcpyAsync(w.data(), H.d_seq2 + w0, w.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(m.data(), H.d_nmask + m0, m.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (int i = 0; i < len; i++) {
        const uint64_t b = b0 + (uint64_t)i;
        const bool bad = (m[(b >> 5) - m0] >> (b & 31)) & 1u;
        dst[i] = bad ? 'N' : "ACGT"[(w[(b >> 4) - w0] >> (2 * (b & 15))) & 3u];
    }
    return GG_OK;
}

// ------------------------------------------------------------------------------------------------ tables
int gg_tables_set_sizes(gg_ctx* ctx, int mode, const int32_t* len, const double* cum, int n, int fixed_len, int minlen, int maxlen) {
    if (!ctx) return GG_ERR_ARG;
    if (mode != GG_SIZE_FIXED && mode != GG_SIZE_LIST && mode != GG_SIZE_FREQ) return ctx->fail(GG_ERR_ARG, "gg_tables_set_sizes: unknown mode");
    if (mode != GG_SIZE_FIXED && (n <= 0 || !len || (mode == GG_SIZE_FREQ && !cum))) return ctx->fail(GG_ERR_ARG, "gg_tables_set_sizes: empty size table");
    CK(cudaSetDevice(ctx->device));
    ctx->size_mode = mode; ctx->fixed_len = fixed_len; ctx->minlen = minlen; ctx->maxlen = maxlen;
    ctx->size_len.clear(); ctx->size_thr.clear();
    if (mode != GG_SIZE_FIXED) {
        ctx->size_len.assign(len, len + n);
        if (mode == GG_SIZE_FREQ) {
            ctx->size_thr.resize(n);
            uint32_t prev = 0;
            for (int i = 0; i < n; i++) {      // p <= cProb[i], fragSim.cpp:163-165; a non-increasing cum entry can never be the first match
                uint32_t t = thr_le(cum[i]); if (t < prev) t = prev; ctx->size_thr[i] = t; prev = t;
            }
        }
    }
    if (max_frag_len(ctx) > MAX_FRAG_LEN) return ctx->fail(GG_ERR_ARG, "gg_tables_set_sizes: fragment lengths above 4095 are not supported");
    int rc;
    if ((rc = upload_vec(ctx, ctx->d_size_len, ctx->size_len))) return rc;
    if ((rc = upload_vec(ctx, ctx->d_size_thr, ctx->size_thr))) return rc;
    ctx->sizes_set = true;
    return GG_OK;
}
int gg_tables_set_norev(gg_ctx* ctx, int norev) { if (!ctx) return GG_ERR_ARG; ctx->norev = norev ? 1 : 0; return GG_OK; }
int gg_tables_set_case(gg_ctx* ctx, int keep_case) { if (!ctx) return GG_ERR_ARG; ctx->keep_case = keep_case ? 1 : 0; return GG_OK; }

static int check_rates(gg_ctx* ctx, const double* sub, int rows, const char* what) {
    for (int r = 0; r < rows; r++)
        for (int b = 0; b < 4; b++) {
            double sum = 0.0;
            for (int a = 0; a < 3; a++) {
                const double v = sub[(size_t)r * 12 + b * 3 + a];
                if (!(v >= 0.0)) return ctx->fail(GG_ERR_ARG, std::string(what) + ": negative or NaN substitution rate");
                sum += v;
            }
            if (sum > 1.0) return ctx->fail(GG_ERR_ARG, std::string(what) + ": the sum of the substitution probabilities exceeds 1");   // deamSim.cpp:1154-1157
        }
    return GG_OK;
}

// total substitution rate of base b in a row of 12, summed in the reference's order (a ascending, deamSim.cpp:1820-1830)
static double row_rate(const double* row, int b) {
    double sum = 0.0;
    for (int a = 0; a < 4; a++) { if (a == b) continue; sum += row[a < b ? b * 3 + a : b * 3 + (a - 1)]; }
    return sum;
}
static double row_max(const double* row) { double m = 0.0; for (int b = 0; b < 4; b++) m = std::max(m, row_rate(row, b)); return m; }

// thresholds of the reference's categorical pick for base b of a row (deamSim.cpp:1797-1863): K[a] = number of u31 with p <= sumProb[a+1]
static bool categorical_thresholds(const double* row, int b, uint32_t k[4]) {
    double sumProb[5], _sumProb[5]; sumProb[0] = 0.0; _sumProb[0] = 0.0;
    double sum = 0.0;
    for (int a = 0; a < 4; a++) {
        if (a == b) continue;
        const int idx = a < b ? b * 3 + a : b * 3 + (a - 1);
        sum += row[idx]; _sumProb[a + 1] = row[idx];
    }
    _sumProb[b + 1] = 1.0 - sum;
    double sacc = 0;
    for (int a = 0; a < 5; a++) { sacc += _sumProb[a]; sumProb[a] = sacc; }
    uint32_t prev = 0;
    for (int a = 0; a < 4; a++) { uint32_t v = thr_le(sumProb[a + 1]); if (v < prev) v = prev; k[a] = v; prev = v; }
    return k[3] == 2147483648u;          // the device picks (u>=K1)+(u>=K2)+(u>=K3): exact iff the last cumulative value covers every u31
}

int gg_tables_set_matrix(gg_ctx* ctx, int slot, const double* sub5, int rows5, const double* sub3, int rows3, int clamp_last) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS || !sub5 || !sub3 || rows5 <= 0 || rows3 <= 0) return ctx->fail(GG_ERR_ARG, "gg_tables_set_matrix: bad arguments");
    int rc;
    if ((rc = check_rates(ctx, sub5, rows5, "gg_tables_set_matrix"))) return rc;
    if ((rc = check_rates(ctx, sub3, rows3, "gg_tables_set_matrix"))) return rc;
    CK(cudaSetDevice(ctx->device));
    HostDamage& H = ctx->dmg[slot];
    H.mode = GG_DMG_MATRIX; H.rows5 = rows5; H.rows3 = rows3; H.clamp_last = clamp_last ? 1 : 0;
    // row that applies at distance `dist` from end e (0 = 5', 1 = 3'); nullptr = left untouched (-last, deamSim.cpp:1811-1817)
    auto eff_row = [&](int e, int dist) -> const double* {
        const double* sub = e ? sub3 : sub5; const int n = e ? rows3 : rows5;
        if (dist >= n) { if (!H.clamp_last) return nullptr; dist = n - 1; }
        return sub + (size_t)dist * 12;
    };
    // draw specification v2: the rows of each end whose total rate stands out are drawn per base, every other position by
    // thinning under the majorant M = the largest total rate among them
    double tail = 0.0;
    for (int e = 0; e < 2; e++) {
        const double* sub = e ? sub3 : sub5; const int n = e ? rows3 : rows5;
        for (int r = NT_MAX; r < n; r++) tail = std::max(tail, row_max(sub + (size_t)r * 12));
        if (H.clamp_last) tail = std::max(tail, row_max(sub + (size_t)(n - 1) * 12));
    }
    double M = tail; int nt[2];
    for (int e = 0; e < 2; e++) {
        nt[e] = 0;
        for (int dist = 0; dist < NT_MAX; dist++) { const double* row = eff_row(e, dist); if (row && row_max(row) > NT_KAPPA * tail) nt[e] = dist + 1; }
        for (int dist = nt[e]; dist < NT_MAX; dist++) { const double* row = eff_row(e, dist); if (row) M = std::max(M, row_max(row)); }
    }
    H.n5t = nt[0]; H.n3t = nt[1]; H.pA = M;
    // terminal rows: the reference's categorical pick; K[a] = number of u31 with p <= sumProb[a+1]
    H.term.assign((size_t)(H.n5t + H.n3t) * 4 + 4, make_uint4(0, 0, 0, 0));
    for (int e = 0; e < 2; e++)
        for (int t = 0; t < nt[e]; t++) {
            const double* row = eff_row(e, t);
            for (int b = 0; b < 4; b++) {
                uint32_t k[4];
                if (!row) {                                  // identity: u >= K for every a < b, u < K for a >= b
                    for (int a = 0; a < 4; a++) k[a] = a < b ? 0u : 2147483648u;
                } else {
                    // cumulative table exactly as deamSim.cpp:1797-1837 builds it (same operation order, fp64)
                    if (!categorical_thresholds(row, b, k)) return ctx->fail(GG_ERR_ARG, "gg_tables_set_matrix: cumulative probabilities of a row do not reach 1");
                }
                H.term[(size_t)((e ? H.n5t : 0) + t) * 4 + b] = make_uint4(k[0], k[1], k[2], k[3]);
            }
        }
    // acceptance rows of the thinned positions: rows5 real rows + sentinel, then sentinel + rows3 real rows in reverse order (distance
    // rows3-1 .. 0); the sentinel is the last row when clamping (deamSim.cpp:1811-1813) or "unchanged" with -last (:1814-1816).
    // K_k = number of u31 with p < (r_1 + .. + r_k) / M over the other bases in ascending order.
    H.acc.assign((size_t)(rows5 + rows3 + 2) * 4, make_uint4(0, 0, 0, 0));
    for (int e = 0; e < 2; e++) {
        const int rows = e == 0 ? rows5 : rows3; const size_t base = e == 0 ? 0 : (size_t)rows5 + 1;
        for (int r = 0; r <= rows; r++) {
            const double* row = eff_row(e, r);
            for (int b = 0; b < 4; b++) {
                uint32_t k[3] = {0u, 0u, 0u};
                if (row && M > 0.0) {
                    double cum = 0.0; int n = 0; uint32_t prev = 0;
                    for (int a = 0; a < 4; a++) {
                        if (a == b) continue;
                        cum += row[a < b ? b * 3 + a : b * 3 + (a - 1)];
                        uint32_t v = thr_lt(cum / M); if (v < prev) v = prev; k[n++] = v; prev = v;
                    }
                }
                const size_t slot_r = e == 0 ? (size_t)r : base + (size_t)(rows - r);
                H.acc[slot_r * 4 + b] = make_uint4(k[0], k[1], k[2], 0u);
            }
        }
    }
    geo_table(M, GEO_TABLE, H.geoA); geo_guide(H.geoA, H.guideA);
    if ((rc = upload_vec(ctx, H.d_term, H.term))) return rc;
    if ((rc = upload_vec(ctx, H.d_acc, H.acc))) return rc;
    if ((rc = upload_vec(ctx, H.d_guideA, H.guideA))) return rc;
    return upload_vec(ctx, H.d_geoA, H.geoA);
}

int gg_tables_set_matrix_meth(gg_ctx* ctx, int slot, const double* no5, int rows_no5, const double* no3, int rows_no3,
                              const double* wi5, int rows_wi5, const double* wi3, int rows_wi3, int clamp_last) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS || !no5 || !no3 || !wi5 || !wi3 || rows_no5 <= 0 || rows_no3 <= 0 || rows_wi5 <= 0 || rows_wi3 <= 0)
        return ctx->fail(GG_ERR_ARG, "gg_tables_set_matrix_meth: bad arguments");
    int rc;
    if ((rc = check_rates(ctx, no5, rows_no5, "gg_tables_set_matrix_meth"))) return rc;
    if ((rc = check_rates(ctx, no3, rows_no3, "gg_tables_set_matrix_meth"))) return rc;
    if ((rc = check_rates(ctx, wi5, rows_wi5, "gg_tables_set_matrix_meth"))) return rc;
    if ((rc = check_rates(ctx, wi3, rows_wi3, "gg_tables_set_matrix_meth"))) return rc;
    CK(cudaSetDevice(ctx->device));
    HostDamage& H = ctx->dmg[slot];
    H.mode = GG_DMG_METH; H.clamp_last = clamp_last ? 1 : 0; H.n5t = H.n3t = 0;
    H.rows5 = std::max(rows_no5, rows_wi5); H.rows3 = std::max(rows_no3, rows_wi3);
    // term[(row)*8 + m*4 + b]: rows5 5' rows, then rows3 3' rows; m = 1 methylated.  A table with fewer rows than the other of its end
    // repeats its last row (deamSim.cpp:1665-1668) or, with -last, leaves the position alone (w = 0; :1669-1671)
    H.term.assign((size_t)(H.rows5 + H.rows3) * 8, make_uint4(0, 0, 0, 0));
    for (int e = 0; e < 2; e++)
        for (int m = 0; m < 2; m++) {
            const double* sub = e ? (m ? wi3 : no3) : (m ? wi5 : no5);
            const int n = e ? (m ? rows_wi3 : rows_no3) : (m ? rows_wi5 : rows_no5), R = e ? H.rows3 : H.rows5;
            for (int r = 0; r < R; r++)
                for (int b = 0; b < 4; b++) {
                    uint4& T = H.term[(size_t)((e ? H.rows5 : 0) + r) * 8 + (size_t)m * 4 + b];
                    if (r >= n && !H.clamp_last) { T = make_uint4(0, 0, 0, 0); continue; }
                    uint32_t k[4];
                    if (!categorical_thresholds(sub + (size_t)std::min(r, n - 1) * 12, b, k)) return ctx->fail(GG_ERR_ARG, "gg_tables_set_matrix_meth: cumulative probabilities of a row do not reach 1");
                    T = make_uint4(k[0], k[1], k[2], k[3]);
                }
        }
    return upload_vec(ctx, H.d_term, H.term);
}

int gg_tables_set_singledouble(gg_ctx* ctx, int slot, int protocol, const double* sub5, int rows5, const double* sub3, int rows3) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS || !sub5 || !sub3 || rows5 <= 0 || rows3 <= 0 ||
        (protocol != GG_DMG_SINGLE && protocol != GG_DMG_DOUBLE)) return ctx->fail(GG_ERR_ARG, "gg_tables_set_singledouble: bad arguments");
    CK(cudaSetDevice(ctx->device));
    HostDamage& H = ctx->dmg[slot];
    H.mode = protocol; H.rows5 = rows5; H.rows3 = rows3; H.clamp_last = 1;
    H.sd5.resize(rows5); H.sd3.resize(rows3);
    for (int r = 0; r < rows5; r++) H.sd5[r] = thr_lt(sub5[(size_t)r * 12 + 5]);                                      // C->T, deamSim.cpp:1948
    for (int r = 0; r < rows3; r++) H.sd3[r] = thr_lt(sub3[(size_t)r * 12 + (protocol == GG_DMG_SINGLE ? 5 : 6)]);    // deamSim.cpp:1961 / :1993
    int rc;
    if ((rc = upload_vec(ctx, H.d_sd5, H.sd5))) return rc;
    return upload_vec(ctx, H.d_sd3, H.sd3);
}

int gg_tables_set_briggs(gg_ctx* ctx, int slot, double v, double l, double d, double s) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS) return ctx->fail(GG_ERR_ARG, "gg_tables_set_briggs: bad slot");
    if (!(v >= 0 && v <= 1)) return ctx->fail(GG_ERR_ARG, "Nick frequency must be between 0 and 1");                          // deamSim.cpp:298-301
    if (!(l >= 0 && l <= 1)) return ctx->fail(GG_ERR_ARG, "Geometric parameter for overhang must be between 0 and 1");
    if (!(d >= 0 && d <= 1)) return ctx->fail(GG_ERR_ARG, "double-strand deamination prob. must be between 0 and 1");
    if (!(s >= 0 && s <= 1)) return ctx->fail(GG_ERR_ARG, "single-strand deamination prob. must be between 0 and 1");
    CK(cudaSetDevice(ctx->device));
    HostDamage& H = ctx->dmg[slot];
    H.mode = GG_DMG_BRIGGS; H.rows5 = H.rows3 = 0; H.clamp_last = 1;
    H.Ks = thr_lt(s); H.Kd = thr_lt(d); H.pL = l; H.pV = v; H.pA = s; H.pB = d;
    // overhang lengths (l), nick position (v), and the geometric skips of the event draws over the single-stranded (s) and
    // double-stranded (d) positions (draw specification v2)
    geo_table(l, GEO_TABLE, H.geoL); geo_table(v, GEO_TABLE, H.geoV); geo_table(s, GEO_TABLE, H.geoA); geo_table(d, GEO_TABLE, H.geoB);
    geo_guide(H.geoA, H.guideA); geo_guide(H.geoB, H.guideB);
    int rc;
    if ((rc = upload_vec(ctx, H.d_guideA, H.guideA))) return rc;
    if ((rc = upload_vec(ctx, H.d_guideB, H.guideB))) return rc;
    if ((rc = upload_vec(ctx, H.d_geoL, H.geoL))) return rc;
    if ((rc = upload_vec(ctx, H.d_geoV, H.geoV))) return rc;
    if ((rc = upload_vec(ctx, H.d_geoA, H.geoA))) return rc;
    return upload_vec(ctx, H.d_geoB, H.geoB);
}
int gg_tables_set_briggs_ss(gg_ctx* ctx, int slot, double d, double s5, double s3, double l5, double l3) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS) return ctx->fail(GG_ERR_ARG, "gg_tables_set_briggs_ss: bad slot");
    if (!(l5 >= 0 && l5 <= 1) || !(l3 >= 0 && l3 <= 1)) return ctx->fail(GG_ERR_ARG, "Geometric parameter for overhang must be between 0 and 1");   // deamSim.cpp:329-334
    if (!(d >= 0 && d <= 1)) return ctx->fail(GG_ERR_ARG, "double-strand deamination prob. must be between 0 and 1");
    if (!(s5 >= 0 && s5 <= 1) || !(s3 >= 0 && s3 <= 1)) return ctx->fail(GG_ERR_ARG, "single-strand deamination prob. must be between 0 and 1");
    CK(cudaSetDevice(ctx->device));
    HostDamage& H = ctx->dmg[slot];
    H.mode = GG_DMG_BRIGGS_SS; H.rows5 = H.rows3 = 0; H.clamp_last = 1;
    H.Ks = thr_lt(s5); H.K3 = thr_lt(s3); H.Kd = thr_lt(d); H.pL = l5; H.pV = l3;
    geo_table(l5, GEO_TABLE, H.geoL); geo_table(l3, GEO_TABLE, H.geoV);      // geometric_distribution<int>(l5) / (l3), deamSim.cpp:325-326,1342-1343
    int rc;
    if ((rc = upload_vec(ctx, H.d_geoL, H.geoL))) return rc;
    return upload_vec(ctx, H.d_geoV, H.geoV);
}
int gg_tables_set_composition(gg_ctx* ctx, int slot, int dist, const double* s5p, const double* s5m, const double* s3p, const double* s3m) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_COMP_SLOTS || dist < 1 || dist > GG_MAX_COMP_DIST || !s5p || !s5m || !s3p || !s3m) return ctx->fail(GG_ERR_ARG, "gg_tables_set_composition: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)2 * dist * 4;
    std::vector<double> t; t.insert(t.end(), s5p, s5p + n); t.insert(t.end(), s5m, s5m + n); t.insert(t.end(), s3p, s3p + n); t.insert(t.end(), s3m, s3m + n);
    // maxProb exactly as fragSim.cpp:1030-1047 multiplies it (5p+ / 5p- first, then 3p+ / 3p-)
    double maxP = 1.0, maxM = 1.0;
    for (int i = 0; i < 2 * dist; i++) maxP = maxP * *std::max_element(s5p + 4 * i, s5p + 4 * i + 4);
    for (int i = 0; i < 2 * dist; i++) maxM = maxM * *std::max_element(s5m + 4 * i, s5m + 4 * i + 4);
    for (int i = 0; i < 2 * dist; i++) maxP = maxP * *std::max_element(s3p + 4 * i, s3p + 4 * i + 4);
    for (int i = 0; i < 2 * dist; i++) maxM = maxM * *std::max_element(s3m + 4 * i, s3m + 4 * i + 4);
    if (!(maxP > 0.0) || !(maxM > 0.0)) return ctx->fail(GG_ERR_ARG, "gg_tables_set_composition: non-positive composition table");
    int rc = upload_vec(ctx, ctx->comp[slot].d_tab, t);
    if (rc) return rc;
    ctx->comp[slot].dist = dist; ctx->comp[slot].maxP = maxP; ctx->comp[slot].maxM = maxM;
    ctx->plan_dirty = true;
    return GG_OK;
}
int gg_tables_clear_composition(gg_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= GG_MAX_COMP_SLOTS) return GG_ERR_ARG;
    ctx->comp[slot].dist = 0; ctx->plan_dirty = true; return GG_OK;
}
// survival thresholds for every window length W <= wmax and G/C count g <= W: survive iff rP <= pSurv (fragSim.cpp:1524-1529)
static int build_gc_table(gg_ctx* ctx, int slot, int wmax) {
    gg_ctx::HostGc& H = ctx->gc[slot];
    std::vector<uint32_t> thr((size_t)(wmax + 1) * (size_t)(wmax + 2) / 2);
    static const double inv_sqrt_2pi = 0.3989422804014327;                                // pdfNorm, fragSim.cpp:96-102
    const double maxNorm = inv_sqrt_2pi / H.sd * exp(-0.5 * ((H.mean - H.mean) / H.sd) * ((H.mean - H.mean) / H.sd));
    for (int W = 0; W <= wmax; W++)
        for (int g = 0; g <= W; g++) {
            uint32_t t = 2147483648u;                                                     // W == 0: gcContent is NaN, rP > NaN is false: never killed
            if (W > 0) {
                const double gcContent = double(g) / double(W);
                double pSurv = 1 / (1 + exp(H.bias * (gcContent - H.mean)));
                const double a = (gcContent - H.mean) / H.sd;
                pSurv = pSurv * ((inv_sqrt_2pi / H.sd * exp(-0.5 * a * a)) / maxNorm);
                t = pSurv == pSurv ? thr_le(pSurv) : 2147483648u;                        // NaN (a genome whose probes all have the same GC content: sd = 0): "rP > NaN" is false, never killed
            }
            thr[(size_t)W * (W + 1) / 2 + g] = t;
        }
    int rc = upload_vec(ctx, H.d_thr, thr);
    if (rc) return rc;
    H.wmax = wmax;
    return GG_OK;
}
int gg_tables_set_gcbias(gg_ctx* ctx, int slot, int genome_id, int comp_dist, double gc_bias, double* mean_out, double* sd_out) {
    if (!ctx) return GG_ERR_ARG;
    if (slot < 0 || slot >= GG_MAX_GC_SLOTS || genome_id < 0 || genome_id >= (int)ctx->genomes.size() || comp_dist < 0 || comp_dist > GG_MAX_COMP_DIST)
        return ctx->fail(GG_ERR_ARG, "gg_tables_set_gcbias: bad arguments");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sync_contig_tables(ctx))) return rc;
    CK(ctx->d_ticket_stats.ensure(256));
    unsigned long long* d_sums = (unsigned long long*)(ctx->d_ticket_stats.as<uint8_t>() + 128);
    CK(cudaMemsetAsync(d_sums, 0, 24, ctx->st));
    const HostGenome& HG = ctx->genomes[genome_id];
    GenomeDev G; G.seq2 = HG.d_seq2; G.nmask = HG.d_nmask; G.G = HG.G; G.first_contig = HG.first_contig; G.n_contigs = (int)HG.contigs.size();
    G.circ_contig = -1; G.wrap_s0 = 0; G.wrap_gbase = 0;       // the probe loop has no --circ branch (fragSim.cpp:1214-1225)
    ggd::Key key; key.k0 = (uint32_t)ctx->seed; key.k1 = (uint32_t)(ctx->seed >> 32);
    CK(launch_gc_probe(G, ctx->d_cum_end.as<uint64_t>(), ctx->d_gbase.as<uint64_t>(), ctx->d_clen.as<uint32_t>(), comp_dist, key, d_sums, ctx->st));
    unsigned long long sums[3];
    CK(cudaMemcpyAsync(sums, d_sums, sizeof sums, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (sums[2]) return ctx->fail(GG_ERR_GIVEUP, "gg_tables_set_gcbias: a GC probe exhausted its attempts (genome without a resolved 50-bp window?)");
    gg_ctx::HostGc& H = ctx->gc[slot];
    const double n = (double)GG_GC_PROBES, w = (double)(GG_GC_PROBE_LEN + 2 * comp_dist);
    H.mean = (double)sums[0] / (n * w);
    const double var = (double)sums[1] / (n * w * w) - H.mean * H.mean;
    H.sd = sqrt(var > 0 ? var : 0);
    H.bias = gc_bias; H.dist = comp_dist; H.genome = genome_id; H.on = true; H.wmax = -1;
    if (mean_out) *mean_out = H.mean;
    if (sd_out) *sd_out = H.sd;
    ctx->plan_dirty = true;
    return GG_OK;
}
int gg_tables_clear_gcbias(gg_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= GG_MAX_GC_SLOTS) return GG_ERR_ARG;
    ctx->gc[slot].on = false; ctx->plan_dirty = true; return GG_OK;
}
int gg_tables_clear_damage(gg_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= GG_MAX_DAMAGE_SLOTS) return GG_ERR_ARG;
    ctx->dmg[slot].mode = GG_DMG_NONE; return GG_OK;
}
int gg_tables_set_adapters(gg_ctx* ctx, const char* fwd, const char* rev, int read_len) {
    if (!ctx || !fwd || !rev) return ctx ? ctx->fail(GG_ERR_ARG, "gg_tables_set_adapters: bad arguments") : GG_ERR_ARG;
    if (read_len < 1 || read_len > MAX_FRAG_LEN) return ctx->fail(GG_ERR_ARG, "gg_tables_set_adapters: read length must be in [1,4095]");
    std::string f(fwd), s(rev);
    // the reference appends the adapter strings as they are (adptSim.cpp:298-315): lower case or index placeholders (N) pass through the
    // stand-alone adptSim stage, which works on ASCII; the fused kernel carries sequences in 2 bits and needs upper-case ACGT
    bool acgt = true;
    for (size_t i = 0; i < f.size(); i++) { if ((unsigned char)f[i] < 0x21 || f[i] == '>') return ctx->fail(GG_ERR_ARG, "gg_tables_set_adapters: adapters must be printable"); if (!strchr("ACGT", f[i])) acgt = false; }
    for (size_t i = 0; i < s.size(); i++) { if ((unsigned char)s[i] < 0x21 || s[i] == '>') return ctx->fail(GG_ERR_ARG, "gg_tables_set_adapters: adapters must be printable"); if (!strchr("ACGT", s[i])) acgt = false; }
    ctx->adF = f; ctx->adS = s; ctx->read_len = read_len; ctx->ad_dirty = true; ctx->ad_acgt = acgt;
    return GG_OK;
}

int gg_tables_set_deam_naming(gg_ctx* ctx, int on) { if (!ctx) return GG_ERR_ARG; ctx->deam_name = on ? 1 : 0; return GG_OK; }
int gg_tables_set_adpt_naming(gg_ctx* ctx, int add_in_name, const char* tag) {
    if (!ctx) return GG_ERR_ARG;
    const size_t n = tag ? strlen(tag) : 0;
    if (n > sizeof(ctx->naming.tag) - 1) return ctx->fail(GG_ERR_ARG, "gg_tables_set_adpt_naming: -tag longer than 23 characters");
    memset(&ctx->naming, 0, sizeof ctx->naming);
    ctx->naming.add_in_name = add_in_name ? 1 : 0; ctx->naming.tag_len = (int)n;
    if (n) memcpy(ctx->naming.tag, tag, n);
    return GG_OK;
}

int gg_plan_set(gg_ctx* ctx, const gg_range* ranges, int n) {
    if (!ctx || n < 0 || (n > 0 && !ranges)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_plan_set: bad arguments") : GG_ERR_ARG;
    if (n > 65535) return ctx->fail(GG_ERR_ARG, "gg_plan_set: at most 65535 ranges");
    uint64_t next = n ? ranges[0].first_id : 0;
    for (int i = 0; i < n; i++) {
        if (ranges[i].first_id != next) return ctx->fail(GG_ERR_ARG, "gg_plan_set: ranges must be contiguous and ascending in fragment id");
        if (ranges[i].genome < 0 || ranges[i].genome >= (int)ctx->genomes.size()) return ctx->fail(GG_ERR_ARG, "gg_plan_set: unknown genome id");
        if (ranges[i].damage_slot < -1 || ranges[i].damage_slot >= GG_MAX_DAMAGE_SLOTS) return ctx->fail(GG_ERR_ARG, "gg_plan_set: bad damage slot");
        if (ranges[i].comp_slot < -1 || ranges[i].comp_slot >= GG_MAX_COMP_SLOTS) return ctx->fail(GG_ERR_ARG, "gg_plan_set: bad composition slot");
        if (ranges[i].gc_slot < -1 || ranges[i].gc_slot >= GG_MAX_GC_SLOTS) return ctx->fail(GG_ERR_ARG, "gg_plan_set: bad GC-bias slot");
        if (ranges[i].count == 0) return ctx->fail(GG_ERR_ARG, "gg_plan_set: empty range");
        next += ranges[i].count;
    }
    ctx->plan.assign(ranges, ranges + n);
    ctx->plan_dirty = true;
    return GG_OK;
}

uint64_t gg_max_record_bytes(gg_ctx* ctx, int emit) {
    if (!ctx || emit < GG_EMIT_FRAG || emit > GG_EMIT_RR) return 0;
    int h; uint64_t r; plan_bounds(ctx, emit, max_frag_len(ctx), &h, &r); return r;
}

// ------------------------------------------------------------------------------------------------ fused run
static int run_meth_chain(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, int slot, gg_out* out);

int gg_run_fused(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, gg_out* out) {
    if (!ctx || !out) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_fused: bad arguments") : GG_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (emit > GG_EMIT_FRAG && emit <= GG_EMIT_RR && !ctx->plan.empty()) {
        // methylation (-matfilemeth/-matfilenonmeth) needs the case of every base: such a plan runs as the three stage kernels chained on
        // the device (fragSim --case -> deamSim -> adptSim), like gargammel.pl --methyl, which hands every source the same two tables (:1698-1811)
        int meth = -1; bool uniform = true;
        for (size_t i = 0; i < ctx->plan.size(); i++) {
            const int sl = ctx->plan[i].damage_slot;
            if (sl >= 0 && sl < GG_MAX_DAMAGE_SLOTS && ctx->dmg[sl].mode == GG_DMG_METH) meth = sl;
        }
        if (meth >= 0) {
            for (size_t i = 0; i < ctx->plan.size(); i++) if (ctx->plan[i].damage_slot != meth) uniform = false;
            if (!uniform) return ctx->fail(GG_ERR_ARG, "gg_run_fused: a plan that uses a methylation slot must use it for every range (gargammel.pl:670-677)");
            return run_meth_chain(ctx, first, count, emit, meth, out);
        }
    }
    if (emit < GG_EMIT_FRAG || emit > GG_EMIT_RR) return ctx->fail(GG_ERR_ARG, "gg_run_fused: unknown emit mode");
    if (emit >= GG_EMIT_STDOUT4 && !ctx->ad_acgt) return ctx->fail(GG_ERR_ARG, "gg_run_fused: adapters with characters other than upper-case ACGT are only supported by the stand-alone adptSim stage (gg_run_adpt)");
    if (ctx->plan.empty() || ctx->genomes.empty()) return ctx->fail(GG_ERR_STATE, "gg_run_fused: no plan / genome");
    const uint64_t p0 = ctx->plan.front().first_id, p1 = ctx->plan.back().first_id + ctx->plan.back().count;
    if (first < p0 || first + count > p1) return ctx->fail(GG_ERR_ARG, "gg_run_fused: fragment ids outside the plan");
    for (size_t i = 0; i < ctx->plan.size(); i++) {
        const int s = ctx->plan[i].damage_slot;
        if (emit != GG_EMIT_FRAG && s >= 0 && ctx->dmg[s].mode == GG_DMG_NONE) return ctx->fail(GG_ERR_STATE, "gg_run_fused: plan references an unset damage slot");
        if (ctx->genomes[ctx->plan[i].genome].G < 2) return ctx->fail(GG_ERR_ARG, "gg_run_fused: genome too small");
    }
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sync_contig_tables(ctx)) || (rc = sync_adapters(ctx)) || (rc = sync_plan(ctx))) return rc;
    if (count == 0) { out->dev_ptr = ctx->d_out.p; ctx->last_ms = 0; return GG_OK; }
    const bool with_case = ctx->keep_case && emit == GG_EMIT_FRAG;      // later stages upper-case what they read (deamSim.cpp:1319, adptSim.cpp:288)
    if (with_case)
        for (size_t i = 0; i < ctx->plan.size(); i++)
            if (!ctx->genomes[ctx->plan[i].genome].d_cmask) return ctx->fail(GG_ERR_STATE, "gg_run_fused: --case was set after a genome upload (call gg_tables_set_case first)");

    const int Lmax = max_frag_len(ctx);
    if (Lmax < 0 || Lmax > MAX_FRAG_LEN) return ctx->fail(GG_ERR_ARG, "gg_run_fused: no admissible fragment length (check -l/-s/-f against -m/-M)");
    int hdr_max; uint64_t rec_max; plan_bounds(ctx, emit, Lmax, &hdr_max, &rec_max);
    const int Smax = emit_seqlen(emit, Lmax, ctx->read_len);
    const int max_chunks = (Lmax + 15) / 16;
    // packed line images (adptSim dialects); an odd word stride per fragment: the owner lanes' accesses then fall into 32 different banks
    const int line_words = emit >= GG_EMIT_STDOUT4 ? ((emit_lines(emit) * ((Smax + 15) / 16 + 2)) | 1) : 0;
    // whole-group stores need >= 13 header bytes between two sequence lines: name + tag + 9 >= 13
    size_t min_name_tag = 1000;
    for (size_t i = 0; i < ctx->plan.size(); i++) {
        size_t nm = 1000;
        const HostGenome& HG = ctx->genomes[ctx->plan[i].genome];
        for (size_t c = 0; c < HG.contigs.size(); c++) nm = std::min(nm, HG.contigs[c].name.size());
        min_name_tag = std::min(min_name_tag, nm + strnlen(ctx->plan[i].tag, sizeof(ctx->plan[i].tag) - 1));
    }
    const int fast_boundary = (min_name_tag >= 4 && !getenv("GG_FUSED_NOFAST")) ? 1 : 0;

    // One warp per tile of F = 32 fragments (fewer only if 32 maximal records do not fit one warp's shared memory);
    // warps per CTA chosen to fill the SM: limited by shared memory, registers and the kernel's launch bound.
    const size_t sm_budget = std::min<size_t>(227 * 1024, ctx->smem_optin) - 2560;   // static: length guide table + two-digit table (2256 B)
    // specialised kernels: 1 = every used damage slot is a matrix, 2 = every used slot is Briggs, 0 = any mix
    bool has_matrix = false, has_briggs = false, has_other = false;
    for (size_t i = 0; i < ctx->plan.size(); i++) {
        const int sl = ctx->plan[i].damage_slot; if (sl < 0) continue;
        const int m = ctx->dmg[sl].mode;
        if (m == GG_DMG_MATRIX) has_matrix = true; else if (m == GG_DMG_BRIGGS) has_briggs = true; else if (m != GG_DMG_NONE) has_other = true;
    }
    int dmode = has_other || (has_matrix && has_briggs) || !fast_boundary ? 0 : has_briggs ? 2 : 1;   // (short headers: only the general variant is built)
    const int emitc = emit <= GG_EMIT_DEAM ? 0 : emit == GG_EMIT_STDOUT4 ? 2 : 1;
    // matrix-only plans keep their threshold tables in shared memory (L1 is only a few KB at this carve-out)
    int thr_bytes = 0, thr_slot[GG_MAX_DAMAGE_SLOTS] = {-1, -1, -1, -1}, term_slot[GG_MAX_DAMAGE_SLOTS] = {-1, -1, -1, -1};
    if (dmode == 1 && emit != GG_EMIT_FRAG) {
        bool used[GG_MAX_DAMAGE_SLOTS] = {false, false, false, false};
        for (size_t i = 0; i < ctx->plan.size(); i++) if (ctx->plan[i].damage_slot >= 0) used[ctx->plan[i].damage_slot] = true;
        for (int sl = 0; sl < GG_MAX_DAMAGE_SLOTS; sl++)
            if (used[sl] && ctx->dmg[sl].mode == GG_DMG_MATRIX) {
                thr_slot[sl] = thr_bytes / 16; thr_bytes += 64 * (ctx->dmg[sl].rows5 + ctx->dmg[sl].rows3 + 2);
                term_slot[sl] = thr_bytes / 16; thr_bytes += 64 * (ctx->dmg[sl].n5t + ctx->dmg[sl].n3t);
            }
        if (thr_bytes > 48 * 1024) { dmode = 0; thr_bytes = 0; for (int sl = 0; sl < GG_MAX_DAMAGE_SLOTS; sl++) thr_slot[sl] = term_slot[sl] = -1; }   // huge matrices: read them from global
    }
    // ... and so do the skip tables of the event draws (matrix- or Briggs-only plans): the first Lmax+1 thresholds (a skip never exceeds
    // the record) and the 512-byte guide, once per slot (matrix) or twice (Briggs: rate s, rate d)
    int geo_n = 0, geo_slot[GG_MAX_DAMAGE_SLOTS] = {-1, -1, -1, -1};
    if ((dmode == 1 || dmode == 2) && emit != GG_EMIT_FRAG) {
        geo_n = (std::min(GEO_TABLE, Lmax + 1) + 3) & ~3;
        int tb = thr_bytes;
        for (size_t i = 0; i < ctx->plan.size(); i++) {
            const int sl = ctx->plan[i].damage_slot;
            if (sl < 0 || geo_slot[sl] >= 0 || ctx->dmg[sl].mode != (dmode == 1 ? GG_DMG_MATRIX : GG_DMG_BRIGGS)) continue;
            geo_slot[sl] = tb; tb += (dmode == 1 ? 1 : 2) * (4 * geo_n + 512);
        }
        if (tb <= 56 * 1024) thr_bytes = tb;
        else { dmode = 0; thr_bytes = 0; geo_n = 0; for (int sl = 0; sl < GG_MAX_DAMAGE_SLOTS; sl++) geo_slot[sl] = thr_slot[sl] = term_slot[sl] = -1; }   // (records of thousands of bases: the general variant reads its tables from global memory)
    }
    int regs = 64, bps0 = 0; CK(fused_prepare(fast_boundary, dmode, emitc, 32, 4096, &bps0, &regs));
    const char* envW = getenv("GG_FUSED_WARPS"); const char* envC = getenv("GG_FUSED_CTAS");
    int F = 32;
    // fragments per emission part (the staging buffer holds that many records): half a tile unless GG_FUSED_E says otherwise (tuning knob)
    const char* envE = getenv("GG_FUSED_E");
    auto e_frags = [&](int f) { const int e = envE ? std::max(1, atoi(envE)) : (f + 1) / 2; return std::min(f, e); };
    const int lflags = (emit >= GG_EMIT_STDOUT4 ? 1 : 0) | (dmode == 0 ? 2 : 0);
    { const char* envF = getenv("GG_FUSED_F"); if (envF) F = std::max(1, std::min(32, atoi(envF))); }   // tuning knob, like GG_FUSED_WARPS
    while (F > 1 && (size_t)make_layout(F, rec_max, max_chunks, e_frags(F), line_words, 1, thr_bytes, lflags).total > sm_budget) F >>= 1;
    if ((size_t)make_layout(F, rec_max, max_chunks, e_frags(F), line_words, 1, thr_bytes, lflags).total > sm_budget) return ctx->fail(GG_ERR_ARG, "gg_run_fused: a single record does not fit shared memory");
    const int ctas = envC ? std::max(1, atoi(envC)) : 1;
    const size_t per_warp = (size_t)make_layout(F, rec_max, max_chunks, e_frags(F), line_words, 1, thr_bytes, lflags).warp_bytes;
    int warps = (int)((sm_budget / (size_t)ctas - (size_t)thr_bytes - 16) / per_warp);
    warps = std::min(warps, 65536 / (regs * 32) / ctas);
    warps = std::min(warps, GG_FUSED_MAXT / 32);
    if (envW) warps = std::min(warps, std::max(1, atoi(envW)));
    warps = std::max(warps, 1);
    SmemLayout sm = make_layout(F, rec_max, max_chunks, e_frags(F), line_words, warps, thr_bytes, lflags);
    // small plan / genome / contig tables ride along in shared memory when that costs no resident warp (pass A walks them with
    // dependent loads; L1 is only a few KB at this carve-out)
    if (!getenv("GG_FUSED_NOTAB") && ctx->plan.size() <= 16 && ctx->genomes.size() <= 16 && ctx->n_contigs_total <= 96 && ctx->names_bytes <= 1024) {
        static_assert(sizeof(RangeDev) % 8 == 0 && sizeof(GenomeDev) % 8 == 0, "the shared-memory table copy moves whole words");
        SmemLayout st = make_layout(F, rec_max, max_chunks, e_frags(F), line_words, warps, thr_bytes, lflags,
                                    (int)ctx->plan.size(), (int)ctx->genomes.size(), ctx->n_contigs_total, ctx->names_bytes, true,
                                    (ctx->size_mode == GG_SIZE_FREQ && ctx->size_len.size() <= 512) ? (int)ctx->size_len.size() : 0,
                                    (ctx->adF.size() <= 512 && ctx->adS.size() <= 512 && emit >= GG_EMIT_STDOUT4) ? (int)(1 + (ctx->adF.size() + 15) / 16 + 3) : 0,
                                    (ctx->adF.size() <= 512 && ctx->adS.size() <= 512 && emit >= GG_EMIT_STDOUT4) ? (int)(1 + (ctx->adS.size() + 15) / 16 + 3) : 0);
        if ((size_t)st.total <= sm_budget) sm = st;
    }
    for (int i = 0; i < GG_MAX_DAMAGE_SLOTS; i++) { sm.thr_slot[i] = thr_slot[i]; sm.term_slot[i] = term_slot[i]; sm.geo_slot[i] = geo_slot[i]; }
    sm.geo_n = geo_n;
    const uint64_t n_tiles64 = (count + (uint64_t)F - 1) / (uint64_t)F;
    if (n_tiles64 > 0x7FFFFFF0ull) return ctx->fail(GG_ERR_ARG, "gg_run_fused: too many fragments for one run; split the id range");
    const int n_tiles = (int)n_tiles64;

    const uint64_t out_cap = count * rec_max + 64;
    DBuf& oslab = ctx->fused_dst ? *ctx->fused_dst : ctx->out_slab ? ctx->d_out2 : ctx->d_out;
    CK(oslab.ensure(out_cap));
    if (with_case) {
        CK(ctx->d_case.ensure((size_t)count * 20 + 64));
        std::vector<const uint32_t*> cm(ctx->genomes.size());
        for (size_t g = 0; g < cm.size(); g++) cm[g] = ctx->genomes[g].d_cmask;
        int rcc = upload_vec(ctx, ctx->d_cmask_ptrs, cm); if (rcc) return rcc;
    }
    CK(ctx->d_tile_state.ensure((size_t)n_tiles * 8));
    CK(ctx->d_ticket_stats.ensure(256));

    FusedParams P; memset(&P, 0, sizeof P);
    P.first = first; P.count = count; P.F = F; P.n_tiles = n_tiles; P.emit = emit; P.norev = ctx->norev; P.fast_boundary = fast_boundary; P.dmode = dmode;
    if (emitc) {                                     // adptSim dialects: the sequence-line geometry is the same for every fragment
        P.S_ad = Smax; P.LWL = (Smax + 15) / 16 + 2; P.lw_stride = line_words;
        P.gpl_ad = (uint32_t)((Smax + 15) / 16 + 1); P.gpf_ad = (uint32_t)emit_lines(emit) * P.gpl_ad;
        P.gmagic = (uint32_t)((0x100000000ull + P.gpf_ad - 1) / P.gpf_ad);
        P.sfx0 = emit == GG_EMIT_RR ? 2 : 0;
    }
    P.E = e_frags(F); P.lut = 0x54474341u;
    P.key.k0 = (uint32_t)ctx->seed; P.key.k1 = (uint32_t)(ctx->seed >> 32); P.ks = ggd::make_key_sched(P.key.k0, P.key.k1);
    P.ranges = ctx->d_ranges.as<RangeDev>(); P.n_ranges = (int)ctx->plan.size();
    P.genomes = ctx->d_genomes.as<GenomeDev>(); P.n_genomes = (int)ctx->genomes.size();
    P.ctab.cum_end = ctx->d_cum_end.as<uint64_t>(); P.ctab.gbase = ctx->d_gbase.as<uint64_t>(); P.ctab.len = ctx->d_clen.as<uint32_t>();
    P.ctab.name_off = ctx->d_name_off.as<uint32_t>(); P.ctab.name_len = ctx->d_name_len.as<uint16_t>(); P.ctab.names = ctx->d_names.as<char>();
    P.size_mode = ctx->size_mode; P.fixed_len = ctx->fixed_len; P.minlen = ctx->minlen; P.maxlen = ctx->maxlen; P.n_sizes = (int)ctx->size_len.size();
    P.size_thr = ctx->d_size_thr.as<uint32_t>(); P.size_len = ctx->d_size_len.as<int32_t>();
    for (int s = 0; s < GG_MAX_DAMAGE_SLOTS; s++) fill_damage_dev(ctx->dmg[s], &P.dmg[s]);
    for (int s = 0; s < GG_MAX_GC_SLOTS; s++) {
        if (ctx->gc[s].on && ctx->gc[s].wmax < Lmax + 2 * ctx->gc[s].dist + 1) { int rc2 = build_gc_table(ctx, s, Lmax + 2 * ctx->gc[s].dist + 1); if (rc2) return rc2; }   // +1: the --circ corner window
        P.gc[s].thr = ctx->gc[s].d_thr.as<uint32_t>(); P.gc[s].wmax = ctx->gc[s].wmax; P.gc[s].on = ctx->gc[s].on ? 1 : 0;
    }
    for (int s = 0; s < GG_MAX_COMP_SLOTS; s++) { P.comp[s].dist = ctx->comp[s].dist; P.comp[s].tab = ctx->comp[s].d_tab.as<double>(); P.comp[s].maxP = ctx->comp[s].maxP; P.comp[s].maxM = ctx->comp[s].maxM; }
    P.ad.adF2 = ctx->d_adF2.as<uint32_t>() + 1; P.ad.adS2 = ctx->d_adS2.as<uint32_t>() + 1; P.ad.rcS2 = ctx->d_rcS2.as<uint32_t>() + 1;
    P.ad.lenF = (int)ctx->adF.size(); P.ad.lenS = (int)ctx->adS.size(); P.ad.read_len = ctx->read_len;
    P.out = oslab.as<uint8_t>(); P.out_cap = out_cap;
    P.tile_state = ctx->d_tile_state.as<unsigned long long>();
    P.ticket = ctx->d_ticket_stats.as<unsigned int>();
    P.stats = (unsigned long long*)(ctx->d_ticket_stats.as<uint8_t>() + 64);
    P.sm = sm;
    if (with_case) {
        P.case_off = ctx->d_case.as<unsigned long long>(); P.case_gpos = P.case_off + count; P.case_meta = (uint32_t*)(P.case_gpos + count);
    }

    int bps = 0;
    CK(fused_prepare(fast_boundary, dmode, emitc, warps * 32, sm.total, &bps, &regs));
    if (bps < 1) return ctx->fail(GG_ERR_CUDA, "gg_run_fused: kernel does not fit on an SM");
    bps = std::min(bps, ctas);
    const uint64_t want_ctas = (n_tiles64 + (uint64_t)warps - 1) / (uint64_t)warps;
    const int grid = (int)std::min<uint64_t>(want_ctas, (uint64_t)bps * (uint64_t)ctx->sm_count);

    if (getenv("GG_FUSED_DEBUG"))
        fprintf(stderr, "[gg_run_fused] variant fast=%d dmode=%d emitc=%d regs=%d | F=%d E=%d warps=%d ctas/SM=%d grid=%d | smem/warp=%d B (stage %d) tables=%d B total=%d B\n",
                fast_boundary, dmode, emitc, regs, F, P.E, warps, bps, grid, sm.warp_bytes, sm.stage_bytes, sm.warp0, sm.total);
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    CK(cudaMemsetAsync(ctx->d_tile_state.p, 0, (size_t)n_tiles * 8, ctx->st));
    CK(cudaMemsetAsync(ctx->d_ticket_stats.p, 0, 256, ctx->st));
    CK(launch_fused(P, emitc, grid, warps * 32, ctx->st));
    if (with_case) CK(launch_case_apply(count, P.case_off, P.case_gpos, P.case_meta, ctx->d_cmask_ptrs.as<const uint32_t*>(), P.out, ctx->st));
    unsigned long long stats[ST_N];
    { int rcp = fetch_small(ctx, P.stats, stats, ST_N, ctx->ev1); if (rcp) return rcp; }
    float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    out->dev_ptr = oslab.p; out->bytes = stats[ST_TOTAL]; out->records = count;
    out->attempts = stats[ST_ATTEMPTS]; out->gather_bytes = stats[ST_GATHER]; out->device_ms = ms; out->launches = with_case ? 3 : 2;   // k_fused (+ k_case_apply) + k_publish
    if (stats[ST_ERR] & ERRF_GIVEUP) return ctx->fail(GG_ERR_GIVEUP, "gg_run_fused: a fragment exhausted its sampling attempts (genome without an admissible window?)");
    if (stats[ST_ERR] & ERRF_OVERFLOW) return ctx->fail(GG_ERR_CAPACITY, "gg_run_fused: internal staging/output overflow");
    return GG_OK;
}

namespace {
// the chunked egress pipelines steer gg_run_fused / gg_bgzf_dev through these context fields; whatever way the pipeline leaves
// (CUDA error, capacity error, success) they go back to their defaults
struct PipelineReset {
    gg_ctx* c;
    explicit PipelineReset(gg_ctx* ctx) : c(ctx) {}
    ~PipelineReset() { c->out_slab = 0; c->bgzf_slab = 0; c->bgzf_reuse_code = 0; }
};
}

// ------------------------------------------------------------------------------------------------ stand-alone stages
namespace {
// stand-alone stages may be chained on the device (input = a previous run's output): write to the other slab
DBuf& pick_out(gg_ctx* ctx, const void* input) {
    const uint8_t* in = (const uint8_t*)input; const uint8_t* o = ctx->d_out.as<uint8_t>();
    return (o && in >= o && in < o + ctx->d_out.cap) ? ctx->d_out2 : ctx->d_out;
}
// newline index of a device-resident record stream; returns n_rec and the RecStream
int index_records(gg_ctx* ctx, const uint8_t* d_in, size_t n, RecStream* R, int* launches) {
    R->in = d_in; R->n = n; R->nl = nullptr; R->n_nl = 0; R->n_rec = 0;
    if (n == 0) return GG_OK;
    const int nb = (int)((n + NL_BLOCK_BYTES - 1) / NL_BLOCK_BYTES);
    CK(ctx->d_scan.ensure(((size_t)nb + (nb + 2047) / 2048 + 16) * 8));
    unsigned long long* counts = ctx->d_scan.as<unsigned long long>();
    unsigned long long* scratch = counts + nb;
    unsigned long long* total = scratch + (nb + 2047) / 2048 + 1;
    CK(launch_count_newlines(d_in, n, counts, nb, ctx->st));
    CK(launch_exclusive_scan(counts, (uint64_t)nb, scratch, total, ctx->st));
    unsigned long long h_total = 0; uint8_t last = 0;
    CK(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&last, d_in + n - 1, 1, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(ctx->d_nl.ensure((size_t)(h_total + 2) * 8));
    CK(launch_fill_newlines(d_in, n, counts, ctx->d_nl.as<uint64_t>(), nb, ctx->st));
    *launches += 5;
    const uint64_t n_lines = h_total + (last != '\n' ? 1 : 0);
    if (n_lines & 1) return ctx->fail(GG_ERR_PARSE, "record stream has an odd number of lines (expected ID line + sequence line per record)");
    R->nl = ctx->d_nl.as<uint64_t>(); R->n_nl = h_total; R->n_rec = n_lines / 2;
    return GG_OK;
}
}

static int run_deam_into(gg_ctx* ctx, const void* recs_dev, size_t n, int slot, uint64_t first_id, DBuf& ob, gg_out* out);
static int run_adpt_into(gg_ctx* ctx, const void* recs_dev, size_t n, int emit, uint64_t first_id, DBuf& ob, gg_out* out);

int gg_run_deam_dev(gg_ctx* ctx, const void* recs_dev, size_t n, int slot, uint64_t first_id, gg_out* out) {
    if (!ctx || !out || (n && !recs_dev)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_deam: bad arguments") : GG_ERR_ARG;
    return run_deam_into(ctx, recs_dev, n, slot, first_id, pick_out(ctx, recs_dev), out);
}
static int run_deam_into(gg_ctx* ctx, const void* recs_dev, size_t n, int slot, uint64_t first_id, DBuf& ob, gg_out* out) {
    memset(out, 0, sizeof *out);
    if (slot < -1 || slot >= GG_MAX_DAMAGE_SLOTS) return ctx->fail(GG_ERR_ARG, "gg_run_deam: bad damage slot");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_ticket_stats.ensure(256));
    unsigned long long* d_stats = (unsigned long long*)(ctx->d_ticket_stats.as<uint8_t>() + 64);
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    CK(cudaMemsetAsync(ctx->d_ticket_stats.p, 0, 256, ctx->st));
    ggd::DamageDev D; memset(&D, 0, sizeof D);
    const int has = (slot >= 0 && ctx->dmg[slot].mode != GG_DMG_NONE) ? 1 : 0;
    if (has) fill_damage_dev(ctx->dmg[slot], &D);
    ggd::Key key; key.k0 = (uint32_t)ctx->seed; key.k1 = (uint32_t)(ctx->seed >> 32);
    // ---- single pass (k_deam_stream): no record index, no host round trip.  It covers the event-drawn models without -name (record
    // sizes unchanged) on a 16-byte aligned stream; ERRF_FALLBACK (very short lines, records longer than its look-ahead) -> general path
    if (n && !getenv("GG_DEAM_GENERAL") && ((uintptr_t)recs_dev & 15) == 0 && (!has || ((D.mode == GG_DMG_MATRIX || D.mode == GG_DMG_BRIGGS) && !ctx->deam_name))) {
        const uint64_t n_tiles64 = (n + (uint64_t)deam_stream_tile_bytes() - 1) / (uint64_t)deam_stream_tile_bytes();
        if (n_tiles64 < 0xFFFFFFF0ull) {
            CK(ob.ensure(n + 1 + 64));
            CK(ctx->d_tile_state.ensure((size_t)n_tiles64 * 8));
            CK(cudaMemsetAsync(ctx->d_tile_state.p, 0, (size_t)n_tiles64 * 8, ctx->st));
            CK(launch_deam_stream((const uint8_t*)recs_dev, n, ob.as<uint8_t>(), D, has, key, first_id, ctx->d_tile_state.as<unsigned long long>(),
                                  ctx->d_ticket_stats.as<unsigned int>(), d_stats, (uint32_t)n_tiles64, ctx->sm_count, ctx->st));
            unsigned long long stats[ST_N];
            { int rcp = fetch_small(ctx, d_stats, stats, ST_N, ctx->ev1); if (rcp) return rcp; }
            if (!(stats[ST_ERR] & ERRF_FALLBACK)) {
                float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
                ctx->last_ms = ms;
                const uint64_t lines = stats[ST_TOTAL];
                out->dev_ptr = ob.p; out->bytes = n + stats[ST_ATTEMPTS]; out->records = lines / 2; out->in_bytes = n; out->device_ms = ms; out->launches = 2;   // k_deam_stream + k_publish
                if (lines & 1) return ctx->fail(GG_ERR_PARSE, "record stream has an odd number of lines (expected ID line + sequence line per record)");
                if (stats[ST_ERR] & ERRF_PARSE) return ctx->fail(GG_ERR_PARSE, "gg_run_deam: a record does not start with '>'");
                return GG_OK;
            }
            CK(cudaMemsetAsync(ctx->d_ticket_stats.p, 0, 256, ctx->st));         // general path below (the event times then cover both attempts)
        }
    }
    RecStream R; int launches = 0;
    int rc = index_records(ctx, (const uint8_t*)recs_dev, n, &R, &launches);
    if (rc) return rc;
    uint64_t out_bytes = n ? (R.n_nl == 2 * R.n_rec ? n : n + 1) : 0;
    if (ctx->deam_name && has && R.n_rec) {
        // -name: record sizes depend on the damage -> sizes, scan, write (the draws are counter-based, so both passes agree)
        const uint64_t nbk = (R.n_rec + 1 + 2047) / 2048;
        CK(ctx->d_sizes.ensure((size_t)(R.n_rec + 1 + nbk + 16) * 8));
        unsigned long long* sizes = ctx->d_sizes.as<unsigned long long>();
        unsigned long long* scratch = sizes + R.n_rec + 1;
        unsigned long long* total = scratch + nbk + 1;
        CK(cudaMemsetAsync(sizes + R.n_rec, 0, 8, ctx->st));                        // sentinel so that off[r+1] exists for the last record
        CK(launch_deam_name_sizes(R, D, key, first_id, sizes, d_stats, ctx->st));
        CK(launch_exclusive_scan(sizes, R.n_rec + 1, scratch, total, ctx->st));
        unsigned long long h_total = 0;
        CK(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        out_bytes = h_total;
        CK(ob.ensure(out_bytes + 64));
        CK(launch_deam_name_records(R, sizes, ob.as<uint8_t>(), D, key, first_id, ctx->st));
        launches += 5;
    } else {
        CK(ob.ensure(out_bytes + 64));
        CK(launch_deam_records(R, ob.as<uint8_t>(), D, has, key, first_id, d_stats, ctx->st));
        launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->st));
    unsigned long long stats[ST_N];
    CK(cudaMemcpyAsync(stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    out->dev_ptr = ob.p; out->bytes = out_bytes; out->records = R.n_rec; out->in_bytes = n; out->device_ms = ms; out->launches = launches;
    if (stats[ST_ERR] & ERRF_PARSE) return ctx->fail(GG_ERR_PARSE, "gg_run_deam: a record does not start with '>'");
    return GG_OK;
}

int gg_run_adpt_dev(gg_ctx* ctx, const void* recs_dev, size_t n, int emit, uint64_t first_id, gg_out* out) {
    if (!ctx || !out || (n && !recs_dev)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_adpt: bad arguments") : GG_ERR_ARG;
    return run_adpt_into(ctx, recs_dev, n, emit, first_id, pick_out(ctx, recs_dev), out);
}
static int run_adpt_into(gg_ctx* ctx, const void* recs_dev, size_t n, int emit, uint64_t first_id, DBuf& ob, gg_out* out) {
    memset(out, 0, sizeof *out);
    if (emit < GG_EMIT_STDOUT4 || emit > GG_EMIT_RR) return ctx->fail(GG_ERR_ARG, "gg_run_adpt: emit mode must be an adptSim dialect");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sync_adapters(ctx))) return rc;
    CK(ctx->d_ticket_stats.ensure(256));
    unsigned long long* d_stats = (unsigned long long*)(ctx->d_ticket_stats.as<uint8_t>() + 64);
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    CK(cudaMemsetAsync(ctx->d_ticket_stats.p, 0, 256, ctx->st));
    RecStream R; int launches = 0;
    if ((rc = index_records(ctx, (const uint8_t*)recs_dev, n, &R, &launches))) return rc;
    unsigned long long h_total = 0;
    if (R.n_rec) {
        const uint64_t nb = (R.n_rec + 2047) / 2048;
        CK(ctx->d_sizes.ensure((size_t)(R.n_rec + nb + 16) * 8));
        unsigned long long* sizes = ctx->d_sizes.as<unsigned long long>();
        unsigned long long* scratch = sizes + R.n_rec;
        unsigned long long* total = scratch + nb + 1;
        Adapters ad; memset(&ad, 0, sizeof ad);
        ad.lenF = (int)ctx->adF.size(); ad.lenS = (int)ctx->adS.size(); ad.read_len = ctx->read_len;
        CK(launch_adpt_sizes(R, emit, ad, ctx->naming, sizes, d_stats, ctx->st));
        CK(launch_exclusive_scan(sizes, R.n_rec, scratch, total, ctx->st));
        CK(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        CK(ob.ensure(h_total + 64));
        ggd::Key key; key.k0 = (uint32_t)ctx->seed; key.k1 = (uint32_t)(ctx->seed >> 32);
        CK(launch_adpt_records(R, sizes, ob.as<uint8_t>(), emit, ad, ctx->d_adF.as<uint8_t>(), ctx->d_adS.as<uint8_t>(), ctx->naming, key, first_id, ctx->st));
        launches += 5;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->st));
    unsigned long long stats[ST_N];
    CK(cudaMemcpyAsync(stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    out->dev_ptr = ob.p; out->bytes = h_total; out->records = R.n_rec; out->in_bytes = n; out->device_ms = ms; out->launches = launches;
    if (stats[ST_ERR] & ERRF_PARSE) return ctx->fail(GG_ERR_PARSE, "gg_run_adpt: a record does not start with '>'");
    return GG_OK;
}

// a methylation plan: fragSim (--case) -> deamSim -> [adptSim] as stage kernels over device-resident streams; the last stage writes
// where gg_run_fused would have written (so the egress pipelines work unchanged), the intermediate streams live in their own buffers
static int run_meth_chain(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, int slot, gg_out* out) {
    if (!ctx->keep_case) return ctx->fail(GG_ERR_STATE, "gg_run_fused: methylation tables need the case of the genome (gg_tables_set_case before gg_genome_upload)");
    struct DstReset { gg_ctx* c; ~DstReset() { c->fused_dst = nullptr; } } reset = {ctx};
    DBuf& final_dst = ctx->out_slab ? ctx->d_out2 : ctx->d_out;
    gg_out o1, o2, o3;
    ctx->fused_dst = &ctx->d_chain_a;
    int rc = gg_run_fused(ctx, first, count, GG_EMIT_FRAG, &o1);
    ctx->fused_dst = nullptr;
    if (rc) return rc;
    rc = run_deam_into(ctx, o1.dev_ptr, (size_t)o1.bytes, slot, first, emit == GG_EMIT_DEAM ? final_dst : ctx->d_chain_b, &o2);
    if (rc) return rc;
    *out = o2;
    if (emit != GG_EMIT_DEAM) {
        rc = run_adpt_into(ctx, o2.dev_ptr, (size_t)o2.bytes, emit, first, final_dst, &o3);
        if (rc) return rc;
        *out = o3; out->device_ms += o2.device_ms; out->launches += o2.launches;
    }
    out->records = count; out->attempts = o1.attempts; out->gather_bytes = o1.gather_bytes; out->in_bytes = 0;
    out->device_ms += o1.device_ms; out->launches += o1.launches;
    ctx->last_ms = out->device_ms;
    return GG_OK;
}

static int upload_input(gg_ctx* ctx, const uint8_t* recs, size_t n) {
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_in.ensure(n + 64));
    if (n) CK(cudaMemcpyAsync(ctx->d_in.p, recs, n, cudaMemcpyHostToDevice, ctx->st));
    return GG_OK;
}
int gg_run_deam(gg_ctx* ctx, const uint8_t* recs, size_t n, int slot, uint64_t first_id, gg_out* out) {
    if (!ctx || (n && !recs)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_deam: bad arguments") : GG_ERR_ARG;
    int rc = upload_input(ctx, recs, n); if (rc) return rc;
    return gg_run_deam_dev(ctx, ctx->d_in.p, n, slot, first_id, out);
}
int gg_run_adpt(gg_ctx* ctx, const uint8_t* recs, size_t n, int emit, uint64_t first_id, gg_out* out) {
    if (!ctx || (n && !recs)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_adpt: bad arguments") : GG_ERR_ARG;
    int rc = upload_input(ctx, recs, n); if (rc) return rc;
    return gg_run_adpt_dev(ctx, ctx->d_in.p, n, emit, first_id, out);
}

int gg_run_fused_to_host(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, void* host_dst, size_t cap, uint64_t chunk, gg_out* out) {
    if (!ctx || !out || (!host_dst && count)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_fused_to_host: bad arguments") : GG_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (chunk == 0) chunk = 1u << 20;
    CK(cudaSetDevice(ctx->device));
    PipelineReset reset(ctx);
    uint8_t* dst = (uint8_t*)host_dst; size_t used = 0; int rc = GG_OK;
    for (uint64_t a = 0, i = 0; a < count; a += chunk, i++) {
        const int slab = (int)(i & 1);
        if (i >= 2) CK(cudaEventSynchronize(ctx->ev_c[slab]));          // the copy that last read this slab has finished
        ctx->out_slab = slab;
        gg_out o;
        rc = gg_run_fused(ctx, first + a, std::min<uint64_t>(chunk, count - a), emit, &o);   // syncs the compute stream, returns the byte count
        if (rc != GG_OK) break;
        if (used + o.bytes > cap) { rc = ctx->fail(GG_ERR_CAPACITY, "gg_run_fused_to_host: destination too small"); break; }
        CK(cudaMemcpyAsync(dst + used, o.dev_ptr, o.bytes, cudaMemcpyDeviceToHost, ctx->st_copy));   // overlaps the next chunk's kernel
        CK(cudaEventRecord(ctx->ev_c[slab], ctx->st_copy));
        used += o.bytes; out->bytes += o.bytes; out->records += o.records; out->attempts += o.attempts; out->gather_bytes += o.gather_bytes;
        out->device_ms += o.device_ms; out->launches += o.launches;
    }
    ctx->out_slab = 0;
    cudaError_t e = cudaStreamSynchronize(ctx->st_copy);
    if (rc != GG_OK) return rc;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gg_run_fused_to_host: copy");
    out->dev_ptr = nullptr;
    ctx->last_ms = out->device_ms;
    return GG_OK;
}

// ------------------------------------------------------------------------------------------------ device BGZF
int gg_bgzf_dev(gg_ctx* ctx, const gg_out* in, gg_out* out) {
    if (!ctx || !in || !out) return ctx ? ctx->fail(GG_ERR_ARG, "gg_bgzf_dev: bad arguments") : GG_ERR_ARG;
    memset(out, 0, sizeof *out);
    out->records = in->records; out->in_bytes = in->bytes;
    if (in->bytes == 0) return GG_OK;
    if (!in->dev_ptr || ((uintptr_t)in->dev_ptr & 15)) return ctx->fail(GG_ERR_ARG, "gg_bgzf_dev: input must be a 16-byte aligned device stream");
    CK(cudaSetDevice(ctx->device));
    const uint64_t n = in->bytes;
    const uint64_t n_blocks = (n + BGZF_BLOCK - 1) / BGZF_BLOCK;
    if (n_blocks > 0xFFFFFFF0ull) return ctx->fail(GG_ERR_ARG, "gg_bgzf_dev: stream too long");
    const uint64_t out_cap = n + n_blocks * (18 + 5 + 8) + 64;            // every member at worst stored
    // the encoder with LZ77 matches (k_bgzf_lz) unless GG_BGZF_MODE=huff asks for the Huffman-only one (k_bgzf_deflate)
    const char* mode_env = getenv("GG_BGZF_MODE");
    const bool use_lz = !(mode_env && strcmp(mode_env, "huff") == 0);
    const int HN = use_lz ? 320 : 256;                                      // histogram counters
    DBuf& oslab = ctx->bgzf_slab ? ctx->d_bgzf2 : ctx->d_bgzf;
    CK(oslab.ensure(out_cap));
    CK(ctx->d_bgzf_tab.ensure((size_t)std::max(BGZF_TAB_WORDS, BGZL_TAB_WORDS) * 4));
    CK(ctx->d_bgzf_state.ensure((size_t)n_blocks * 8 + (size_t)HN * 8 + 64));
    unsigned long long* d_state = ctx->d_bgzf_state.as<unsigned long long>();
    unsigned long long* d_hist = d_state + n_blocks;                         // HN counters
    unsigned long long* d_total = d_hist + HN;                               // [0] bytes, [1] error flag
    unsigned int* d_ticket = (unsigned int*)(d_total + 2);                   // [0] encoder, [1] counting pass
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    CK(cudaMemsetAsync(d_state, 0, (size_t)n_blocks * 8 + (size_t)HN * 8 + 32, ctx->st));
    int launches = 2;                                                       // the encoder + k_publish
    if (!(ctx->bgzf_reuse_code && ctx->bgzf_hdr_nbits > 0)) {      // (the chunked pipeline keeps the code of its first chunk)
        if (ctx->bgzf_static.empty()) {                                     // CRC tables and shift polynomials, in the Huffman-only layout
            ctx->bgzf_static.assign(BGZF_TAB_WORDS, 0u);
            uint32_t t[4][256]; ggh::crc32_tables(t);
            memcpy(&ctx->bgzf_static[BGZF_TAB_CRC], t, sizeof t);
            for (int k = 0; k < BGZF_THREADS; k++) ctx->bgzf_static[BGZF_TAB_SHIFT + k] = ggh::crc32_x8n((uint64_t)BGZF_RUN * (uint64_t)(BGZF_THREADS - 1 - k));
            for (int k = 0; k < 17; k++) ctx->bgzf_static[BGZF_TAB_X8 + k] = ggh::crc32_x8n(1ull << k);
        }
        std::vector<uint32_t> tab;
        if (use_lz) {
            // 1. symbol counts of every stride-th member under the encoder's own parse -> one literal/length + distance code for the stream
            BgzfLzParams C; memset(&C, 0, sizeof C);
            C.in = (const uint8_t*)in->dev_ptr; C.n = n; C.stride = (uint32_t)std::max<uint64_t>(1, (n_blocks + gglz::SAMPLE_MEMBERS - 1) / gglz::SAMPLE_MEMBERS);
            C.n_blocks = (uint32_t)((n_blocks + C.stride - 1) / C.stride); C.ticket = d_ticket + 1; C.hist = d_hist;
            CK(launch_bgzf_lz(C, true, (int)std::min<uint64_t>(C.n_blocks, (uint64_t)ctx->sm_count), ctx->st));
            unsigned long long hist[gglz::N_SYM];
            { int rcp = fetch_small(ctx, d_hist, hist, gglz::N_SYM); if (rcp) return rcp; }
            ggh::LzCode code; std::string e;
            uint64_t h64[gglz::N_SYM]; for (int i = 0; i < gglz::N_SYM; i++) h64[i] = hist[i];
            if (!ggh::build_lz_code(h64, code, e)) return ctx->fail(GG_ERR_STATE, e);
            if (code.header_nbits > BGZL_HDR_WORDS * 32) return ctx->fail(GG_ERR_STATE, "gg_bgzf_dev: code description too long");
            tab.assign(BGZL_TAB_WORDS, 0u);
            memcpy(&tab[BGZL_TAB_CRC], &ctx->bgzf_static[BGZF_TAB_CRC], 1024 * 4);
            memcpy(&tab[BGZL_TAB_SHIFT], &ctx->bgzf_static[BGZF_TAB_SHIFT], BGZF_THREADS * 4);
            memcpy(&tab[BGZL_TAB_X8], &ctx->bgzf_static[BGZF_TAB_X8], 17 * 4);
            memcpy(&tab[BGZL_TAB_SYM], code.sym, sizeof code.sym);
            for (size_t i = 0; i < code.header.size(); i++) tab[BGZL_TAB_HDR + (i >> 2)] |= (uint32_t)code.header[i] << (8 * (i & 3));
            ctx->bgzf_hdr_nbits = code.header_nbits;
        } else {
            // 1. sampled byte histogram (<= ~32 MB of the stream) -> one dynamic Huffman code for the whole stream
            const uint64_t stride = std::max<uint64_t>(1, (n >> 4) / (2u << 20));
            CK(launch_byte_hist((const uint8_t*)in->dev_ptr, n, stride, d_hist, ctx->st));
            unsigned long long hist[256];
            { int rcp = fetch_small(ctx, d_hist, hist, 256); if (rcp) return rcp; }
            ggh::DeflateCode code; std::string e;
            uint64_t h64[256]; for (int i = 0; i < 256; i++) h64[i] = hist[i];
            if (!ggh::build_deflate_code(h64, code, e)) return ctx->fail(GG_ERR_STATE, e);
            if (code.header_nbits > BGZF_HDR_WORDS * 32) return ctx->fail(GG_ERR_STATE, "gg_bgzf_dev: code description too long");
            tab = ctx->bgzf_static;
            memcpy(&tab[BGZF_TAB_SYM], code.sym, sizeof code.sym);
            for (size_t i = 0; i < code.header.size(); i++) tab[BGZF_TAB_HDR + (i >> 2)] |= (uint32_t)code.header[i] << (8 * (i & 3));
            ctx->bgzf_hdr_nbits = code.header_nbits;
        }
        CK(cudaMemcpyAsync(ctx->d_bgzf_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));                                 // tab is a local
        launches = 4;                                                       // + the counting / histogram pass + its k_publish
    }
    // 2. one CTA per 65280-byte member, persistent over the SMs
    if (use_lz) {
        BgzfLzParams P; memset(&P, 0, sizeof P);
        P.in = (const uint8_t*)in->dev_ptr; P.n = n; P.out = oslab.as<uint8_t>(); P.out_cap = out_cap;
        P.tab = ctx->d_bgzf_tab.as<uint32_t>(); P.hdr_nbits = ctx->bgzf_hdr_nbits; P.n_blocks = (uint32_t)n_blocks; P.stride = 1;
        P.state = d_state; P.ticket = d_ticket; P.total = d_total; P.hist = d_hist;
        CK(launch_bgzf_lz(P, false, (int)std::min<uint64_t>(n_blocks, (uint64_t)ctx->sm_count), ctx->st));
    } else {
        BgzfParams P;
        P.in = (const uint8_t*)in->dev_ptr; P.n = n; P.out = oslab.as<uint8_t>(); P.out_cap = out_cap;
        P.tab = ctx->d_bgzf_tab.as<uint32_t>(); P.hdr_nbits = ctx->bgzf_hdr_nbits; P.n_blocks = (uint32_t)n_blocks;
        P.state = d_state; P.ticket = d_ticket; P.total = d_total;
        CK(launch_bgzf_deflate(P, (int)std::min<uint64_t>(n_blocks, (uint64_t)ctx->sm_count), ctx->st));
    }
    unsigned long long tot[2];
    { int rcp = fetch_small(ctx, d_total, tot, 2, ctx->ev1); if (rcp) return rcp; }
    float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (tot[1]) return ctx->fail(GG_ERR_CAPACITY, "gg_bgzf_dev: output overflow");
    out->dev_ptr = oslab.p; out->bytes = tot[0]; out->device_ms = ms; out->launches = launches;
    return GG_OK;
}

int gg_run_fused_bgzf_to_host(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, void* host_dst, size_t cap, uint64_t chunk, gg_out* out) {
    if (!ctx || !out || (!host_dst && count)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_run_fused_bgzf_to_host: bad arguments") : GG_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (chunk == 0) chunk = 1u << 20;
    CK(cudaSetDevice(ctx->device));
    PipelineReset reset(ctx);
    uint8_t* dst = (uint8_t*)host_dst; size_t used = 0; int rc = GG_OK;
    ctx->bgzf_hdr_nbits = 0;                                                // the first chunk builds the code, the others reuse it
    for (uint64_t a = 0, i = 0; a < count; a += chunk, i++) {
        const int slab = (int)(i & 1);
        if (i >= 2) CK(cudaEventSynchronize(ctx->ev_c[slab]));              // the copy that last read this BGZF slab has finished
        gg_out o, z;
        rc = gg_run_fused(ctx, first + a, std::min<uint64_t>(chunk, count - a), emit, &o);
        if (rc != GG_OK) break;
        ctx->bgzf_slab = slab; ctx->bgzf_reuse_code = 1;
        rc = gg_bgzf_dev(ctx, &o, &z);                                      // syncs the compute stream: the members are complete
        if (rc != GG_OK) break;
        if (used + z.bytes > cap) { rc = ctx->fail(GG_ERR_CAPACITY, "gg_run_fused_bgzf_to_host: destination too small"); break; }
        CK(cudaMemcpyAsync(dst + used, z.dev_ptr, z.bytes, cudaMemcpyDeviceToHost, ctx->st_copy));   // overlaps the next chunk's kernels
        CK(cudaEventRecord(ctx->ev_c[slab], ctx->st_copy));
        used += z.bytes; out->bytes += z.bytes; out->in_bytes += o.bytes; out->records += o.records; out->attempts += o.attempts;
        out->gather_bytes += o.gather_bytes; out->device_ms += o.device_ms + z.device_ms; out->launches += o.launches + z.launches;
    }
    ctx->bgzf_slab = 0; ctx->bgzf_reuse_code = 0;
    cudaError_t e = cudaStreamSynchronize(ctx->st_copy);
    if (rc != GG_OK) return rc;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gg_run_fused_bgzf_to_host: copy");
    out->dev_ptr = nullptr;
    ctx->last_ms = out->device_ms;
    return GG_OK;
}

int gg_bgzf_host(gg_ctx* ctx, const void* src, size_t n, gg_out* out) {
    if (!ctx || !out || (!src && n)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_bgzf_host: bad arguments") : GG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_in.ensure(n + 64));
    if (n) CK(cudaMemcpyAsync(ctx->d_in.p, src, n, cudaMemcpyHostToDevice, ctx->st));
    gg_out in; memset(&in, 0, sizeof in);
    in.dev_ptr = ctx->d_in.p; in.bytes = n;
    return gg_bgzf_dev(ctx, &in, out);
}

int gg_out_fetch(gg_ctx* ctx, const gg_out* out, void* host_dst, size_t cap, size_t* n) {
    if (!ctx || !out || (!host_dst && out->bytes)) return ctx ? ctx->fail(GG_ERR_ARG, "gg_out_fetch: bad arguments") : GG_ERR_ARG;
    if (n) *n = (size_t)out->bytes;
    if (out->bytes > cap) return ctx->fail(GG_ERR_CAPACITY, "gg_out_fetch: destination too small");
    if (out->bytes == 0) return GG_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host_dst, out->dev_ptr, out->bytes, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GG_OK;
}

int gg_host_alloc(gg_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return ctx ? ctx->fail(GG_ERR_ARG, "gg_host_alloc: bad arguments") : GG_ERR_ARG;
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return ctx->fail(GG_ERR_NOMEM, "gg_host_alloc: cannot pin host memory"); }
    return GG_OK;
}
int gg_host_free(gg_ctx* ctx, void* p) {
    if (!ctx) return GG_ERR_ARG;
    if (p) CK(cudaFreeHost(p));
    return GG_OK;
}

int gg_flush_l2(gg_ctx* ctx) {
    if (!ctx) return GG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = 256ull << 20;          // > 126 MB L2
    CK(ctx->d_flush.ensure(bytes));
    CK(launch_l2_flush(ctx->d_flush.as<uint4>(), bytes / 16, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return GG_OK;
}

} // extern "C"
