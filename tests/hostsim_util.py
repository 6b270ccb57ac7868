"""TEST CODE — builds and drives tests/hostsim (the engine's pass-1 core compiled as plain C++)."""
import ctypes as C
import os
import subprocess

import numpy as np

from pbsim_b200 import capi
from tests import expand

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "hostsim", "hostsim.cpp"), os.path.join(ROOT, "pbsim_b200", "csrc", "host_model.cpp")]
DEPS = SRC + [os.path.join(ROOT, "pbsim_b200", "csrc", f) for f in ("sim_core.cuh", "model_image.hpp", "sample_plan.hpp")] + [
    os.path.join(ROOT, "include", "pbsim_cuda.h")]
LIB = os.path.join(HERE, "hostsim", "libhostsim.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", LIB] + SRC)
        L = C.CDLL(LIB)
        capi.declare_host(L)
        L.hostsim_run.restype = C.c_long
        L.hostsim_run.argtypes = [C.POINTER(capi.Model), C.c_void_p, C.c_void_p, C.c_long, C.c_int,
                                  C.POINTER(C.c_double), C.c_int, C.c_uint32, C.c_void_p, C.c_long,
                                  C.c_longlong, C.c_long]
        L.hostsim_run_sample.restype = C.c_long
        L.hostsim_run_sample.argtypes = [C.POINTER(capi.Model), C.c_void_p, C.c_void_p, C.c_long, C.c_int,
                                         C.POINTER(C.c_double), C.c_int, C.c_uint32, C.c_void_p, C.c_long,
                                         C.c_longlong, C.c_long, C.c_char_p, C.c_void_p, C.c_long]
        L.hostsim_info.restype = C.POINTER(C.c_int64)
        L.hostsim_counts.restype = C.POINTER(C.c_uint32)
        L.hostsim_accuracy.restype = C.POINTER(C.c_double)
        L.hostsim_events.restype = C.POINTER(C.c_uint8)
        L.hostsim_events.argtypes = [C.POINTER(C.c_long)]
        _lib = L
    return _lib


def run(model, genome_upper, hp, seq_num, bias, rng_mode, seed, draws, len_quota, max_reads=0, pool=None,
        batch_reads=0):
    """Returns list of per-subread dicts with the event stream attached.  pool: --method sample, the quality strings"""
    L = lib()
    g = np.frombuffer(genome_upper, dtype=np.uint8)
    hp = np.ascontiguousarray(hp, dtype=np.int16)
    b = (C.c_double * 12)(*bias)
    if draws is None:
        draws = np.zeros(1, dtype=np.int32)
    draws = np.ascontiguousarray(draws, dtype=np.int32)
    if pool is not None:
        quals = b"".join(pool)
        qstart = np.zeros(len(pool) + 1, dtype=np.int64)
        qstart[1:] = np.cumsum([len(x) for x in pool])
        n = L.hostsim_run_sample(model.ptr, g.ctypes.data, hp.ctypes.data, len(g), seq_num, b, rng_mode, seed,
                                 draws.ctypes.data, len(draws), int(len_quota), len(pool), quals, qstart.ctypes.data,
                                 batch_reads)
    else:
        n = L.hostsim_run(model.ptr, g.ctypes.data, hp.ctypes.data, len(g), seq_num, b, rng_mode, seed,
                          draws.ctypes.data, len(draws), int(len_quota), max_reads)
    if n < 0:
        raise RuntimeError("hostsim_run failed: %d" % n)
    info = np.ctypeslib.as_array(L.hostsim_info(), shape=(n, 12)).copy()
    counts = np.ctypeslib.as_array(L.hostsim_counts(), shape=(n, 3)).copy()
    acc = np.ctypeslib.as_array(L.hostsim_accuracy(), shape=(n,)).copy()
    nb = C.c_long()
    evp = L.hostsim_events(C.byref(nb))
    ev = np.ctypeslib.as_array(evp, shape=(nb.value,)).copy() if nb.value else np.zeros(0, np.uint8)
    qs = model.view.method != capi.METHOD_ERRHMM
    out = []
    for i in range(n):
        rid, pas, a, off, wlen, rlen, ncol, minus, nent, evoff, dstart, ovf = [int(x) for x in info[i]]
        if qs:
            e = ev[evoff:evoff + 2 * nent].view(np.uint16)
        else:
            e = ev[evoff:evoff + nent]
        out.append(dict(read_id=rid, pas=pas, acc=a, offset=off, wlen=wlen, rlen=rlen, ncol=ncol, minus=minus,
                        events=e, draw_start=dstart, overflow=ovf, nsub=int(counts[i, 0]), nins=int(counts[i, 1]),
                        ndel=int(counts[i, 2]), accuracy=float(acc[i])))
    return out


def records_from_events(model, subreads, genome_upper, seq_num):
    """Expand + format every subread; returns (reads_bytes, maf_bytes)."""
    v = model.view
    qs = v.method != capi.METHOD_ERRHMM
    sample = v.method == capi.METHOD_SAMPLE
    reads, maf = [], []
    for s in subreads:
        if sample:  # the window may outlast the read: the MAF line reports what was consumed (:1847)
            seq, qual, mref, mread = expand.expand_qshmm(s["events"], genome_upper, s["offset"], s["wlen"], s["minus"],
                                                         partial=True)
            shown = s["ncol"] - s["nins"]
        else:
            f = expand.expand_qshmm if qs else expand.expand_errhmm
            seq, qual, mref, mread = f(s["events"], genome_upper, s["offset"], s["wlen"], s["minus"])
            shown = s["wlen"]
        assert len(seq) == s["rlen"] and len(mref) == s["ncol"]
        r, m = expand.format_records(v.pass_num, v.id_prefix.decode(), seq_num, s["read_id"], s["pas"], s["offset"],
                                     shown, len(genome_upper), s["minus"], seq, qual, mref, mread,
                                     v.accuracy_mean)
        reads.append(r)
        maf.append(m)
    return b"".join(reads), b"".join(maf)
