"""ctypes binding of include/gargammel_b200.h (product) -- also reused, with the orc_ prefix, by oracle/oracle.py.

The same `Context` class drives either library so that the parity tests run the CUDA path and the CPU
oracle through identical call sequences."""
import ctypes as C
import os
from collections import namedtuple

EMIT_FRAG, EMIT_DEAM, EMIT_STDOUT4, EMIT_ARTS, EMIT_ARTP, EMIT_FR, EMIT_RR = range(7)
SIZE_FIXED, SIZE_LIST, SIZE_FREQ = range(3)
BGZF_EOF = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
DMG_NONE, DMG_MATRIX, DMG_BRIGGS, DMG_SINGLE, DMG_DOUBLE, DMG_BRIGGS_SS, DMG_METH = range(7)
DEFAULT_ADAPTER_F = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG"   # adptSim.cpp:28
DEFAULT_ADAPTER_S = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT"         # adptSim.cpp:29

FaiEntry = namedtuple("FaiEntry", "name len offset line_blen line_len")
Range = namedtuple("Range", "first_id count genome damage_slot tag comp_slot gc_slot", defaults=(-1, -1))
RunInfo = namedtuple("RunInfo", "bytes records attempts gather_bytes in_bytes device_ms launches")


class GGError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gargammel_b200 error %d: %s" % (code, msg))
        self.code = code


class _Fai(C.Structure):
    _fields_ = [("name", C.c_char_p), ("len", C.c_int64), ("offset", C.c_uint64),
                ("line_blen", C.c_int32), ("line_len", C.c_int32)]


class _Range(C.Structure):
    _fields_ = [("first_id", C.c_uint64), ("count", C.c_uint64), ("genome", C.c_int32),
                ("damage_slot", C.c_int32), ("comp_slot", C.c_int32), ("gc_slot", C.c_int32), ("tag", C.c_char * 24)]


class _Out(C.Structure):
    _fields_ = [("dev_ptr", C.c_void_p), ("bytes", C.c_uint64), ("records", C.c_uint64), ("attempts", C.c_uint64),
                ("gather_bytes", C.c_uint64), ("in_bytes", C.c_uint64), ("device_ms", C.c_double),
                ("launches", C.c_int32), ("reserved", C.c_int32)]


def lib_path():
    # GG_LIB selects an experimental build of the same library (tuning only)
    return os.environ.get("GG_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgargammel_b200.so")


_LIB = None


def load_library():
    """Load the CUDA library.  Raises if it has not been built: there is no CPU fallback."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise ImportError("gargammel_b200: %s is missing -- run `make` (or __graft_entry__.build()); "
                              "there is no CPU fallback" % p)
        _LIB = C.CDLL(p)
    return _LIB


def _info(o):
    return RunInfo(o.bytes, o.records, o.attempts, o.gather_bytes, o.in_bytes, o.device_ms, o.launches)


def _darr(a):
    import numpy as np
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One (process, device) context.  `lib`/`prefix` select the product ('gg_') or the oracle ('orc_')."""

    def __init__(self, device=0, seed=0, lib=None, prefix="gg_"):
        self._lib = lib if lib is not None else load_library()
        self._p = prefix
        self._oracle = prefix != "gg_"
        self._h = C.c_void_p()
        if self._oracle:
            rc = self._f("ctx_create")(C.c_uint64(seed), C.byref(self._h))
        else:
            rc = self._f("ctx_create")(C.c_int(device), C.c_uint64(seed), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise GGError(rc, "cannot create context (no CUDA device?)")
        self._keep = []

    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def _ck(self, rc):
        if rc != 0:
            f = self._f("last_error")
            f.restype = C.c_char_p
            raise GGError(rc, (f(self._h) or b"").decode("utf-8", "replace"))

    def close(self):
        if self._h:
            self._f("ctx_destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state
    def set_seed(self, seed):
        self._ck(self._f("ctx_set_seed")(self._h, C.c_uint64(seed)))

    def genome_clear(self):
        self._ck(self._f("genome_clear")(self._h))

    def upload_genome(self, fasta, fai, circ=-1):
        """fasta: bytes, or any object with data_ptr()/numel() (a pinned torch uint8 tensor), or a numpy uint8 array.
        circ: index into `fai` of the contig fragSim --circ names (-1 = none)."""
        arr = (_Fai * len(fai))()
        for i, e in enumerate(fai):
            arr[i].name = e.name.encode()
            arr[i].len, arr[i].offset, arr[i].line_blen, arr[i].line_len = e.len, e.offset, e.line_blen, e.line_len
        gid = C.c_int(-1)
        if hasattr(fasta, "data_ptr"):
            ptr, n = C.c_void_p(fasta.data_ptr()), int(fasta.numel())
        elif hasattr(fasta, "ctypes"):
            ptr, n = C.c_void_p(fasta.ctypes.data), int(fasta.size)
        else:
            ptr, n = C.cast(C.c_char_p(fasta), C.c_void_p), len(fasta)
        if circ < 0:
            self._ck(self._f("genome_upload")(self._h, ptr, C.c_size_t(n), arr, C.c_int(len(fai)), C.byref(gid)))
        else:
            self._ck(self._f("genome_upload_circ")(self._h, ptr, C.c_size_t(n), arr, C.c_int(len(fai)), C.c_int(circ), C.byref(gid)))
        return gid.value

    def genome_fetch(self, genome, chr_index, start, length):
        out = C.create_string_buffer(max(length, 1))
        self._ck(self._f("genome_fetch")(self._h, C.c_int(genome), C.c_int(chr_index), C.c_int64(start), C.c_int(length), out))
        return out.raw[:length]

    def set_sizes(self, mode, lens=None, cum=None, fixed_len=20, minlen=0, maxlen=1000):
        import numpy as np
        n = 0 if lens is None else len(lens)
        la = np.ascontiguousarray(lens if lens is not None else [0], dtype=np.int32)
        ca = np.ascontiguousarray(cum if cum is not None else [0.0] * max(n, 1), dtype=np.float64)
        self._ck(self._f("tables_set_sizes")(self._h, C.c_int(mode), la.ctypes.data_as(C.POINTER(C.c_int32)),
                                             ca.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(n), C.c_int(fixed_len),
                                             C.c_int(minlen), C.c_int(maxlen)))

    def set_norev(self, norev):
        self._ck(self._f("tables_set_norev")(self._h, C.c_int(1 if norev else 0)))

    def set_matrix(self, slot, sub5, sub3, clamp_last=True):
        a5, p5 = _darr(sub5)
        a3, p3 = _darr(sub3)
        self._ck(self._f("tables_set_matrix")(self._h, C.c_int(slot), p5, C.c_int(a5.size // 12), p3, C.c_int(a3.size // 12),
                                              C.c_int(1 if clamp_last else 0)))

    def set_case(self, keep_case=True):
        """fragSim --case; call before upload_genome"""
        self._ck(self._f("tables_set_case")(self._h, C.c_int(1 if keep_case else 0)))

    def set_matrix_meth(self, slot, no5, no3, wi5, wi3, clamp_last=True):
        """deamSim -matfilenonmeth (no5, no3) -matfilemeth (wi5, wi3)"""
        a = [_darr(x) for x in (no5, no3, wi5, wi3)]
        self._ck(self._f("tables_set_matrix_meth")(self._h, C.c_int(slot), a[0][1], C.c_int(a[0][0].size // 12), a[1][1], C.c_int(a[1][0].size // 12),
                                                   a[2][1], C.c_int(a[2][0].size // 12), a[3][1], C.c_int(a[3][0].size // 12), C.c_int(1 if clamp_last else 0)))

    def set_singledouble(self, slot, protocol, sub5, sub3):
        a5, p5 = _darr(sub5)
        a3, p3 = _darr(sub3)
        self._ck(self._f("tables_set_singledouble")(self._h, C.c_int(slot), C.c_int(protocol), p5, C.c_int(a5.size // 12),
                                                    p3, C.c_int(a3.size // 12)))

    def set_composition(self, slot, dist, s5p, s5m, s3p, s3m):
        a = [_darr(x) for x in (s5p, s5m, s3p, s3m)]
        self._ck(self._f("tables_set_composition")(self._h, C.c_int(slot), C.c_int(dist), a[0][1], a[1][1], a[2][1], a[3][1]))

    def clear_composition(self, slot):
        self._ck(self._f("tables_clear_composition")(self._h, C.c_int(slot)))

    def set_gcbias(self, slot, genome, comp_dist, gc_bias):
        m, sd = C.c_double(), C.c_double()
        self._ck(self._f("tables_set_gcbias")(self._h, C.c_int(slot), C.c_int(genome), C.c_int(comp_dist), C.c_double(gc_bias), C.byref(m), C.byref(sd)))
        return m.value, sd.value

    def clear_gcbias(self, slot):
        self._ck(self._f("tables_clear_gcbias")(self._h, C.c_int(slot)))

    def set_briggs(self, slot, v, l, d, s):
        self._ck(self._f("tables_set_briggs")(self._h, C.c_int(slot), C.c_double(v), C.c_double(l), C.c_double(d), C.c_double(s)))

    def set_briggs_ss(self, slot, d, s5, s3, l5, l3):
        self._ck(self._f("tables_set_briggs_ss")(self._h, C.c_int(slot), C.c_double(d), C.c_double(s5), C.c_double(s3), C.c_double(l5), C.c_double(l3)))

    def clear_damage(self, slot):
        self._ck(self._f("tables_clear_damage")(self._h, C.c_int(slot)))

    def set_adapters(self, fwd=DEFAULT_ADAPTER_F, rev=DEFAULT_ADAPTER_S, read_len=100):
        self._ck(self._f("tables_set_adapters")(self._h, fwd.encode(), rev.encode(), C.c_int(read_len)))

    def set_deam_naming(self, on=True):
        self._ck(self._f("tables_set_deam_naming")(self._h, C.c_int(1 if on else 0)))

    def set_adpt_naming(self, add_in_name=False, tag=""):
        self._ck(self._f("tables_set_adpt_naming")(self._h, C.c_int(1 if add_in_name else 0), (tag or "").encode()))

    def set_plan(self, ranges):
        arr = (_Range * max(len(ranges), 1))()
        for i, r in enumerate(ranges):
            arr[i].first_id, arr[i].count, arr[i].genome, arr[i].damage_slot = r.first_id, r.count, r.genome, r.damage_slot
            arr[i].comp_slot = r.comp_slot
            arr[i].gc_slot = r.gc_slot
            arr[i].tag = (r.tag or "").encode()
        self._ck(self._f("plan_set")(self._h, arr, C.c_int(len(ranges))))

    # ---- runs: return (bytes, RunInfo)
    def run_fused(self, first, count, emit, fetch=True):
        o = _Out()
        if self._oracle:
            cap = 4096 + count * 4400
            buf = C.create_string_buffer(cap)
            n = C.c_size_t(0)
            self._ck(self._f("run_fused")(self._h, C.c_uint64(first), C.c_uint64(count), C.c_int(emit), buf, C.c_size_t(cap), C.byref(n), C.byref(o)))
            return buf.raw[:n.value], _info(o)
        self._ck(self._f("run_fused")(self._h, C.c_uint64(first), C.c_uint64(count), C.c_int(emit), C.byref(o)))
        return (self._fetch(o) if fetch else None), _info(o)

    def run_fused_into(self, first, count, emit, host_ptr, cap):
        """run + D2H into caller memory (e.g. a pinned torch tensor's data_ptr()); returns RunInfo."""
        o = _Out()
        self._ck(self._f("run_fused")(self._h, C.c_uint64(first), C.c_uint64(count), C.c_int(emit), C.byref(o)))
        n = C.c_size_t(0)
        self._ck(self._f("out_fetch")(self._h, C.byref(o), C.c_void_p(host_ptr), C.c_size_t(cap), C.byref(n)))
        return _info(o)

    def run_fused_to_host(self, first, count, emit, host_ptr, cap, chunk=0):
        """chunked run with the D2H copies overlapped (gg_run_fused_to_host); host_ptr should be pinned memory."""
        o = _Out()
        self._ck(self._f("run_fused_to_host")(self._h, C.c_uint64(first), C.c_uint64(count), C.c_int(emit), C.c_void_p(host_ptr),
                                              C.c_size_t(cap), C.c_uint64(chunk), C.byref(o)))
        return _info(o)

    def run_fused_bgzf(self, first, count, emit, host_ptr=None, cap=0):
        """gg_run_fused + gg_bgzf_dev (device BGZF) + D2H of the compressed members.  With host_ptr: copies into caller memory and
        returns (RunInfo of the run, RunInfo of the compression); without: returns (bgzf bytes, run info, bgzf info).
        The 28-byte BGZF EOF marker (api.BGZF_EOF) is not included."""
        o, z = _Out(), _Out()
        self._ck(self._f("run_fused")(self._h, C.c_uint64(first), C.c_uint64(count), C.c_int(emit), C.byref(o)))
        self._ck(self._f("bgzf_dev")(self._h, C.byref(o), C.byref(z)))
        if host_ptr is None:
            return self._fetch(z), _info(o), _info(z)
        n = C.c_size_t(0)
        self._ck(self._f("out_fetch")(self._h, C.byref(z), C.c_void_p(host_ptr), C.c_size_t(cap), C.byref(n)))
        return _info(o), _info(z)

    def run_fused_bgzf_to_host(self, first, count, emit, host_ptr, cap, chunk=0):
        """gg_run_fused_bgzf_to_host: chunked fused run + device BGZF, D2H of the members overlapped with the next chunk."""
        o = _Out()
        self._ck(self._f("run_fused_bgzf_to_host")(self._h, C.c_uint64(first), C.c_uint64(count), C.c_int(emit), C.c_void_p(host_ptr),
                                                   C.c_size_t(cap), C.c_uint64(chunk), C.byref(o)))
        return _info(o)

    def bgzf(self, data: bytes):
        """gg_bgzf_host: BGZF members of a host byte string, compressed on the device (no EOF marker)."""
        z = _Out()
        src = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data) if len(data) else (C.c_uint8 * 1)()
        self._ck(self._f("bgzf_host")(self._h, src, C.c_size_t(len(data)), C.byref(z)))
        return self._fetch(z), _info(z)

    def _fetch(self, o):
        buf = C.create_string_buffer(max(int(o.bytes), 1))
        n = C.c_size_t(0)
        self._ck(self._f("out_fetch")(self._h, C.byref(o), buf, C.c_size_t(int(o.bytes)), C.byref(n)))
        return buf.raw[:n.value]

    def _run_stage(self, name, data, arg, first_id):
        o = _Out()
        src = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data) if len(data) else (C.c_uint8 * 1)()
        if self._oracle:
            cap = 4096 + 4 * len(data) + 4 * 4200 * (data.count(b"\n") // 2 + 1)
            buf = C.create_string_buffer(cap)
            n = C.c_size_t(0)
            self._ck(self._f(name)(self._h, src, C.c_size_t(len(data)), C.c_int(arg), C.c_uint64(first_id), buf, C.c_size_t(cap), C.byref(n), C.byref(o)))
            return buf.raw[:n.value], _info(o)
        self._ck(self._f(name)(self._h, src, C.c_size_t(len(data)), C.c_int(arg), C.c_uint64(first_id), C.byref(o)))
        return self._fetch(o), _info(o)

    def run_deam(self, data: bytes, slot, first_id=0):
        return self._run_stage("run_deam", data, slot, first_id)

    def run_adpt(self, data: bytes, emit, first_id=0):
        return self._run_stage("run_adpt", data, emit, first_id)

    def last_device_ms(self):
        f = self._f("last_device_ms")
        f.restype = C.c_double
        return f(self._h)

    def max_record_bytes(self, emit):
        f = self._f("max_record_bytes")
        f.restype = C.c_uint64
        return f(self._h, C.c_int(emit))

    def flush_l2(self):
        self._ck(self._f("flush_l2")(self._h))
