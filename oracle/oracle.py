"""Loader of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "gg_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def load():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def OracleContext(seed=0):
    from gargammel_b200.api import Context
    return Context(seed=seed, lib=load(), prefix="orc_")


def philox(seed, frag_id, block, stream):
    out = (C.c_uint32 * 4)()
    load().orc_philox(C.c_uint64(seed), C.c_uint64(frag_id), C.c_uint32(block), C.c_uint32(stream), out)
    return list(out)


def philox_raw(ctr, key):
    out = (C.c_uint32 * 4)()
    load().orc_philox_raw((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def ref_bin(name):
    """Path of the reference's own filter built from its sources + our shim (oracle/Makefile), or None."""
    p = os.path.join(HERE, "_ref", "src", name)
    return p if os.path.exists(p) else None
