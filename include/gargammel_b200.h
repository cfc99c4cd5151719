/*
 * gargammel_b200.h -- C-ABI of the B200-native gargammel hot path
 * (fragSim -> deamSim -> adptSim per-fragment pipeline).
 *
 * The reference (grenaud/gargammel) has no library / FFI interface: the boundary
 * of this path is the process CLI of three executables called by gargammel.pl
 * (gargammel.pl:71-84,232-248).  This header is the thin C boundary between the
 * C++11 host side (bin/fragSim, bin/deamSim, bin/adptSim, bin/gargammel-b200,
 * the Python harness) and the hand-written sm_100a kernels.  Every entry point
 * cites the reference code it replaces.  Plain pointers and sizes only; no C++
 * or torch types cross this boundary; no exception crosses it.
 *
 * All functions return GG_OK (0) or a negative gg_status.  gg_last_error(ctx)
 * returns a human-readable message for the last failure on that context.
 * A context is bound to one CUDA device and one stream and is not thread-safe.
 * There is NO CPU fallback: every gg_run_* launches CUDA kernels or fails.
 */
#ifndef GARGAMMEL_B200_H
#define GARGAMMEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_ABI_VERSION 1

typedef enum gg_status {
    GG_OK            =  0,
    GG_ERR_ARG       = -1,  /* bad argument / inconsistent tables            */
    GG_ERR_CUDA      = -2,  /* CUDA runtime error (message has the detail)    */
    GG_ERR_NOMEM     = -3,  /* host or device allocation failed               */
    GG_ERR_STATE     = -4,  /* call order: tables/plan/genome missing         */
    GG_ERR_PARSE     = -5,  /* malformed FASTA record stream                  */
    GG_ERR_GIVEUP    = -6,  /* a fragment exhausted GG_MAX_ATTEMPTS attempts  */
    GG_ERR_CAPACITY  = -7   /* caller buffer too small                        */
} gg_status;

/* Rejection loop bound per fragment (fragSim.cpp:1361 loops forever on an all-N genome). */
#define GG_MAX_ATTEMPTS 65536u

/* ---- fragment length model: selectFragmentLength, fragSim.cpp:137-187 ---- */
typedef enum gg_size_mode {
    GG_SIZE_FIXED = 0,   /* -l L              fragSim.cpp:181      */
    GG_SIZE_LIST  = 1,   /* -s file           fragSim.cpp:149-152  */
    GG_SIZE_FREQ  = 2    /* -f file           fragSim.cpp:156-178  */
} gg_size_mode;

/* ---- deamination model per damage slot ---- */
typedef enum gg_damage_mode {
    GG_DMG_NONE   = 0,
    GG_DMG_MATRIX = 1,   /* -matfile / -profile    deamSim.cpp:1794-1930 */
    GG_DMG_BRIGGS = 2,   /* -damage v,l,d,s        deamSim.cpp:1477-1613 */
    GG_DMG_SINGLE = 3,   /* -mat single / -mapdamage f single  deamSim.cpp:1941-1971 */
    GG_DMG_DOUBLE = 4,   /* -mat double / -mapdamage f double  deamSim.cpp:1973-2002 */
    GG_DMG_BRIGGS_SS = 5,/* -damagess d,s5,s3,l5,l3 (single-strand libraries)  deamSim.cpp:310-340,1337-1468 */
    GG_DMG_METH   = 6    /* -matfilemeth + -matfilenonmeth (lower-case = methylated)  deamSim.cpp:673-954,1630-1787 */
} gg_damage_mode;
#define GG_MAX_DAMAGE_SLOTS 4
#define GG_MAX_COMP_SLOTS 4
#define GG_MAX_COMP_DIST 32
#define GG_MAX_GC_SLOTS 4
#define GG_GC_PROBES 1000000u   /* fragSim.cpp:1198 */
#define GG_GC_PROBE_LEN 50       /* fragSim.cpp:1199 */

/* ---- what the last stage writes ---- */
typedef enum gg_emit_mode {
    GG_EMIT_FRAG    = 0, /* fragSim record  ">def\nFRAG\n"             fragSim.cpp:1603      */
    GG_EMIT_DEAM    = 1, /* deamSim record  "ID\nSEQ\n"                deamSim.cpp:2038      */
    GG_EMIT_STDOUT4 = 2, /* adptSim default "ID\nF\nID_r\nR\n"         adptSim.cpp:408       */
    GG_EMIT_ARTS    = 3, /* adptSim -arts   "ID\nF\n"                  adptSim.cpp:402       */
    GG_EMIT_ARTP    = 4, /* adptSim -artp   "ID\nF revcomp(R)\n"       adptSim.cpp:403-406   */
    GG_EMIT_FR      = 5, /* adptSim -fr     "ID\nF\n"                  adptSim.cpp:397       */
    GG_EMIT_RR      = 6  /* adptSim -rr     "ID_r\nR\n"                adptSim.cpp:398       */
} gg_emit_mode;

/* One line of a samtools .fai index (faidx1_t, fragSim.cpp:61-65,249-259). */
typedef struct gg_fai_entry {
    const char* name;        /* NUL-terminated contig name                    */
    int64_t     len;         /* bases                                         */
    uint64_t    offset;      /* byte offset of the first base in the FASTA    */
    int32_t     line_blen;   /* bases per line                                */
    int32_t     line_len;    /* bytes per line incl. newline                  */
} gg_fai_entry;

/* One contiguous run of fragment ids drawn from one genome file: replaces one
 * `fragSim -tag T -n N file` invocation of gargammel.pl:1476-1683 and the choice
 * of deamination for that source (gargammel.pl:640-648,1698-1830). */
typedef struct gg_range {
    uint64_t first_id;       /* global fragment id of the first fragment      */
    uint64_t count;          /* number of fragments                           */
    int32_t  genome;         /* id returned by gg_genome_upload               */
    int32_t  damage_slot;    /* 0..GG_MAX_DAMAGE_SLOTS-1, or -1 = no damage   */
    int32_t  comp_slot;      /* base-composition table (fragSim --comp), 0..GG_MAX_COMP_SLOTS-1, or -1 = none */
    int32_t  gc_slot;        /* GC-bias table (fragSim -gc), 0..GG_MAX_GC_SLOTS-1, or -1 = none */
    char     tag[24];        /* NUL-terminated -tag string appended to defline (fragSim.cpp:1505-1507) */
} gg_range;

/* Result of a gg_run_*: a device-resident byte stream. */
typedef struct gg_out {
    void*    dev_ptr;        /* device pointer to the first output byte (owned by ctx) */
    uint64_t bytes;          /* bytes produced                                */
    uint64_t records;        /* fragments / records processed                 */
    uint64_t attempts;       /* sampling attempts incl. rejected (fused/frag) */
    uint64_t gather_bytes;   /* algorithmic genome-gather bytes: sum ceil(L/4)+ceil(L/8) */
    uint64_t in_bytes;       /* input record bytes (stand-alone deam/adpt)    */
    double   device_ms;      /* CUDA-event time of the kernels of this run    */
    int32_t  launches;       /* kernels launched by this run                  */
    int32_t  reserved;
} gg_out;

typedef struct gg_ctx gg_ctx;

/* ---- lifetime ---- */
int  gg_abi_version(void);
int  gg_ctx_create(int device, uint64_t seed, gg_ctx** out);
int  gg_ctx_destroy(gg_ctx* ctx);
int  gg_ctx_set_seed(gg_ctx* ctx, uint64_t seed);
const char* gg_last_error(gg_ctx* ctx);
/* the CUDA stream the kernels run on (cudaStream_t as void*), for external event timing */
void* gg_ctx_stream(gg_ctx* ctx);

/* ---- genome: replaces IndexedGenome (mmap + fetchSeq), fragSim.cpp:192-331,1148-1166 ----
 * fasta/n: host bytes of ONE plain FASTA file (caller-owned, may be freed after return);
 * fai: its index.  Bases are upper-cased (fragSim.cpp:315), 2-bit packed + 1-bit non-ACGT mask,
 * and stay resident in HBM.  Contigs are concatenated in lexicographic name order like the
 * reference's std::map (fragSim.cpp:215,1152-1166). */
int  gg_genome_upload(gg_ctx* ctx, const uint8_t* fasta, size_t n,
                      const gg_fai_entry* fai, int n_chr, int* genome_id_out);
/* fragSim --circ NAME (fragSim.cpp:553-556,1171-1178,1407-1426): contig fai[circ_chr] is circular.  A window that runs off its end
 * continues at its first base; like the reference, the wrapped window skips the contig's LAST base (temp1 is one base short,
 * fragSim.cpp:1413-1415) and the defline keeps idx and idx+len un-wrapped.  circ_chr = -1: same as gg_genome_upload. */
int  gg_genome_upload_circ(gg_ctx* ctx, const uint8_t* fasta, size_t n,
                           const gg_fai_entry* fai, int n_chr, int circ_chr, int* genome_id_out);
int  gg_genome_count(gg_ctx* ctx);
/* drop every uploaded genome and the plan (tables stay) */
int  gg_genome_clear(gg_ctx* ctx);
/* debug/parity helper: unpack `len` bases of contig `chr` (index in the caller's fai order) from `start` into ASCII (N for masked) */
int  gg_genome_fetch(gg_ctx* ctx, int genome_id, int chr, int64_t start, int len, char* dst);

/* ---- tables ---- */
/* fragSim -l / -s / -f, -m, -M (fragSim.cpp:137-187,719-784,1375-1381).
 * FREQ: len[i], cum[i] = running sum exactly as fragSim.cpp:765-768 accumulates it. */
int  gg_tables_set_sizes(gg_ctx* ctx, int mode, const int32_t* len, const double* cum, int n,
                         int fixed_len, int minlen, int maxlen);
int  gg_tables_set_norev(gg_ctx* ctx, int norev);   /* --norev, fragSim.cpp:1463-1467 */
/* fragSim --case (fragSim.cpp:548-551,318-326; gargammel.pl --methyl :1486,1526,1561,1605,1653): fragments keep the case of the genome
 * file (lower-case c / g = methylated cytosine on the + / - strand, README "methylated and unmethylated").  Set it BEFORE
 * gg_genome_upload: the upload then also keeps a 1-bit case mask per base; reverse-strand fragments complement within the case. */
int  gg_tables_set_case(gg_ctx* ctx, int keep_case);
/* fragSim --comp/--dist (fragSim.cpp:798-1047,1609-1759): base-composition acceptance.  Each table is [2*dist][4] (A,C,G,T) and
 * already holds "frequency - background + 0.25" (fragSim.cpp:948-1012): s5p = 5' end on the + strand, s5m on the - strand, s3p/s3m the
 * 3' end; rows 0..dist-1 are the positions outside the fragment on the 5' side / inside on the 3' side (reference order). */
int  gg_tables_set_composition(gg_ctx* ctx, int slot, int dist, const double* s5p, const double* s5m, const double* s3p, const double* s3m);
int  gg_tables_clear_composition(gg_ctx* ctx, int slot);
/* fragSim -gc bias (fragSim.cpp:1190-1269,1509-1576): GC content of 1,000,000 random 50-bp windows of `genome_id` gives mean/stdev; every
 * attempt then survives with probability 1/(1+exp(bias*(gc-mean))) * N(gc; mean, sd)/N(mean; mean, sd) (drawn twice without --comp, as the
 * reference does).  comp_dist = the --dist of the composition slot used together with it (0 without --comp).  Call after
 * gg_tables_set_sizes and gg_genome_upload.  mean_out/sd_out may be NULL. */
int  gg_tables_set_gcbias(gg_ctx* ctx, int slot, int genome_id, int comp_dist, double gc_bias, double* mean_out, double* sd_out);
int  gg_tables_clear_gcbias(gg_ctx* ctx, int slot);
/* deamSim -matfile/-profile (deamSim.cpp:956-1241): sub5[r][12], sub3[r][12] with index 0 = terminal base
 * (i.e. AFTER the 3' reversal of deamSim.cpp:1236); clamp_last = !(-last) (deamSim.cpp:1811-1817). */
int  gg_tables_set_matrix(gg_ctx* ctx, int slot, const double* sub5, int rows5,
                          const double* sub3, int rows3, int clamp_last);
/* deamSim -matfilenonmeth P -matfilemeth Q (deamSim.cpp:673-954,1630-1787): a lower-case base of the record is methylated and takes the
 * "wi" tables, an upper-case base the "no" tables (same layout as gg_tables_set_matrix; the four tables may have different row counts).
 * The record is not upper-cased first (deamSim.cpp:1319); every resolved base that has a row comes out in upper case, unresolved bases
 * and positions beyond a -last table keep their bytes.  Drawn per base.  The fused entry points run a plan whose every range uses such
 * a slot as fragSim (--case) -> deamSim -> adptSim stage kernels chained on the device. */
int  gg_tables_set_matrix_meth(gg_ctx* ctx, int slot, const double* no5, int rows_no5, const double* no3, int rows_no3,
                               const double* wi5, int rows_wi5, const double* wi3, int rows_wi3, int clamp_last);
/* deamSim -mat single|double / -mapdamage (deamSim.cpp:1941-2002); protocol = GG_DMG_SINGLE or GG_DMG_DOUBLE */
int  gg_tables_set_singledouble(gg_ctx* ctx, int slot, int protocol, const double* sub5, int rows5,
                                const double* sub3, int rows3);
/* deamSim -damage v,l,d,s (deamSim.cpp:285-306,1477-1613) */
int  gg_tables_set_briggs(gg_ctx* ctx, int slot, double v, double l, double d, double s);
/* deamSim -damagess d,s5,s3,l5,l3 (deamSim.cpp:310-340,1337-1468): C->T only; 5'/3' overhangs ~ geometric(l5)/(l3) deaminate at s5/s3,
 * the double-stranded middle at d; a base inside both overhangs takes the nearer end */
int  gg_tables_set_briggs_ss(gg_ctx* ctx, int slot, double d, double s5, double s3, double l5, double l3);
int  gg_tables_clear_damage(gg_ctx* ctx, int slot);
/* deamSim -name (deamSim.cpp:342-345,2029-2037): append "_DEAM:p1,p2,..." to the ID of damaged records; stand-alone deamSim stage only */
int  gg_tables_set_deam_naming(gg_ctx* ctx, int put_deam_in_name);
/* adptSim -f -s -l (adptSim.cpp:28-31,131-147) */
int  gg_tables_set_adapters(gg_ctx* ctx, const char* fwd, const char* rev, int read_len);
/* adptSim -name / -tag (adptSim.cpp:163-173,375-393), stand-alone adptSim stage only; tag may be NULL */
int  gg_tables_set_adpt_naming(gg_ctx* ctx, int add_in_name, const char* tag);

/* ---- plan: which fragment id comes from which genome with which tag/damage ---- */
int  gg_plan_set(gg_ctx* ctx, const gg_range* ranges, int n);

/* ---- runs ---- */
/* fragment ids [first, first+count) of the plan; fused frag->deam->adpt in ONE kernel; emit selects the
 * last stage's dialect (GG_EMIT_FRAG = fragSim only; GG_EMIT_DEAM = fragSim|deamSim; others add adptSim). */
int  gg_run_fused(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, gg_out* out);
/* stand-alone stages over a HOST byte stream of 2-line FASTA records (what the previous stage wrote):
 * deamSim main loop deamSim.cpp:1300-2042; adptSim main loop adptSim.cpp:263-420.
 * first_id = fragment id of the first record (0 for the reference CLI). */
int  gg_run_deam(gg_ctx* ctx, const uint8_t* recs, size_t n, int slot, uint64_t first_id, gg_out* out);
int  gg_run_adpt(gg_ctx* ctx, const uint8_t* recs, size_t n, int emit, uint64_t first_id, gg_out* out);
/* same, with the record stream already resident on the device (used for device-timed benchmarks) */
int  gg_run_deam_dev(gg_ctx* ctx, const void* recs_dev, size_t n, int slot, uint64_t first_id, gg_out* out);
int  gg_run_adpt_dev(gg_ctx* ctx, const void* recs_dev, size_t n, int emit, uint64_t first_id, gg_out* out);

/* gg_run_fused + egress in one call: the id range is processed in chunks of `chunk` fragments (0 = default) through two
 * device slabs, and each chunk's bytes are copied to host_dst while the next chunk is being generated (second stream).
 * host_dst should be pinned memory for the copies to overlap.  out->device_ms = sum of the kernels' event times. */
int  gg_run_fused_to_host(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, void* host_dst, size_t cap,
                          uint64_t chunk, gg_out* out);
/* ---- device BGZF: replaces the `| gzip >` hops of gargammel.pl:1510,1729,1873-1884 and ogzstream (-o) ----
 * Compresses a device-resident stream (the gg_out of a gg_run_*) into BGZF members (blocked gzip: zcat/gzstream/htslib read it)
 * on the GPU: deflate with LZ77 matches (greedy parse over warp-private hash tables; files of zlib -1's size) and one dynamic literal/length +
 * distance code per stream, one member per 65280 input bytes, CRC-32 computed on the device.  The members are a pure function of the input
 * bytes.  Environment GG_BGZF_MODE=huff selects the Huffman-only encoder (no matches: six times the kernel speed, files 30-40 % larger).
 * out->dev_ptr/bytes = the members back to back (fetch with gg_out_fetch); the 28-byte BGZF EOF marker is NOT appended
 * (the writer of the file appends it once).  out->in_bytes = uncompressed bytes. */
int  gg_bgzf_dev(gg_ctx* ctx, const gg_out* in, gg_out* out);
/* gg_run_fused + gg_bgzf_dev in chunks of `chunk` fragments (0 = 1 Mi), the D2H copy of chunk k's members overlapping the kernels of
 * chunk k+1; one code (built from the first chunk) for the whole stream.  host_dst should be pinned.  out->bytes = BGZF bytes
 * written to host_dst, out->in_bytes = plain bytes they inflate to. */
int  gg_run_fused_bgzf_to_host(gg_ctx* ctx, uint64_t first, uint64_t count, int emit, void* host_dst, size_t cap,
                               uint64_t chunk, gg_out* out);
/* the same for a stream in host memory (H2D copy, then gg_bgzf_dev) */
int  gg_bgzf_host(gg_ctx* ctx, const void* src, size_t n, gg_out* out);

/* page-locked host memory for the *_to_host entry points (plain malloc'ed memory works too, at a fraction of the PCIe rate) */
int  gg_host_alloc(gg_ctx* ctx, size_t bytes, void** out);
int  gg_host_free(gg_ctx* ctx, void* p);

/* copy a run's bytes to the host (pinned staging inside; blocking) */
int  gg_out_fetch(gg_ctx* ctx, const gg_out* out, void* host_dst, size_t cap, size_t* n);
/* CUDA-event time (ms) of the kernels of the last gg_run_* on this context */
double gg_last_device_ms(gg_ctx* ctx);
/* upper bound of output bytes per fragment for the current tables/plan and emit mode (0 on error) */
uint64_t gg_max_record_bytes(gg_ctx* ctx, int emit);
/* write 1<<log2_bytes bytes of scratch to evict L2 between timed iterations */
int  gg_flush_l2(gg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GARGAMMEL_B200_H */
