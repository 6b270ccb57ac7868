"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  Nothing under pbsim_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

METHOD_QS = 1
METHOD_ERR = 2


def build(force=False):
    src_newer = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH)
        for f in ("pbsim_oracle.c", "pbsim_oracle.h", "glibc_rand.c", "philox.h")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class Stats(C.Structure):
    _fields_ = [
        ("res_num", C.c_int64),
        ("res_pass_num", C.c_int64),
        ("res_len_total", C.c_int64),
        ("res_len_min", C.c_int64),
        ("res_len_max", C.c_int64),
        ("res_sub_num", C.c_int64),
        ("res_ins_num", C.c_int64),
        ("res_del_num", C.c_int64),
        ("res_depth", C.c_double),
        ("res_len_mean", C.c_double),
        ("res_len_sd", C.c_double),
        ("res_accuracy_mean", C.c_double),
        ("res_accuracy_sd", C.c_double),
        ("res_sub_rate", C.c_double),
        ("res_ins_rate", C.c_double),
        ("res_del_rate", C.c_double),
        ("accuracy_total", C.c_double),
    ]


READINFO_DTYPE = np.dtype(
    [
        ("read_id", "<i8"),
        ("pass", "<i4"),
        ("acc", "<i4"),
        ("offset", "<i8"),
        ("wlen", "<i8"),
        ("rlen", "<i8"),
        ("ncol", "<i8"),
        ("strand", "<i4"),
        ("nsub", "<i4"),
        ("nins", "<i4"),
        ("ndel", "<i4"),
        ("draw_start", "<i8"),
        ("accuracy", "<f8"),
    ],
    align=True,
)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_new.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_char_p
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_long, C.c_long, C.c_double,
                                     C.c_double, C.c_long, C.c_long, C.c_long, C.c_double, C.c_char_p]
        L.orc_load_model.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_build_tables.argtypes = [C.c_void_p]
        L.orc_prepass_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.orc_finish_bias.argtypes = [C.c_void_p]
        L.orc_set_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int]
        L.orc_get_bias.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_get_hp.restype = C.POINTER(C.c_int16)
        L.orc_get_hp.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_get_seq.restype = C.POINTER(C.c_char)
        L.orc_get_seq.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_rng_glibc.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_rng_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_rng_philox.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_simulate_wgs.argtypes = [C.c_void_p, C.c_double]
        L.orc_reset_outputs.argtypes = [C.c_void_p]
        L.orc_simulate_sample.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_char_p, C.c_void_p]
        L.orc_simulate_set.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_char_p, C.c_void_p]
        L.orc_get_ssp.restype = C.c_int64
        L.orc_get_ssp.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        for name in ("orc_out_reads", "orc_out_maf"):
            getattr(L, name).restype = C.POINTER(C.c_char)
            getattr(L, name).argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.orc_get_readinfo.restype = C.c_void_p
        L.orc_get_readinfo.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_get_draw_log.restype = C.POINTER(C.c_int32)
        L.orc_get_draw_log.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_draws_consumed.restype = C.c_int64
        L.orc_draws_consumed.argtypes = [C.c_void_p]
        for name in ("orc_get_freq_len", "orc_get_freq_accuracy"):
            getattr(L, name).restype = C.POINTER(C.c_int64)
            getattr(L, name).argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_get_table.restype = C.c_int64
        L.orc_get_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.orc_get_emis2del.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_get_thresholds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_model_exists.argtypes = [C.c_void_p, C.c_int]
        L.orc_model_range.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.orc_glibc_rand_fill.argtypes = [C.c_uint32, C.c_int64, C.c_void_p]
        L.orc_philox_block.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def glibc_rand(seed, n):
    out = np.empty(n, dtype=np.int32)
    lib().orc_glibc_rand_fill(seed, n, out.ctypes.data)
    return out


def philox_block(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.empty(4, dtype=np.uint32)
    lib().orc_philox_block(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def truncate_accuracy(x):
    """set_sim_param: int(x*100)*0.01 (pbsim.cpp:1660)."""
    return int(x * 100) * 0.01


class Oracle:
    """One reference run: parameters + model -> tables; then per sequence simulate_wgs()."""

    def __init__(self, method, model_path, pass_num=1, accuracy_mean=0.85, accuracy_mean_set=False,
                 len_min=100, len_max=1000000, len_mean=9000.0, len_sd=7000.0,
                 ratio=(6, 55, 39), hp_del_bias=1.0, id_prefix="S"):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_new())
        self.method = 3 if method == "sample" else (METHOD_QS if method in ("qshmm", METHOD_QS) else METHOD_ERR)
        self.pass_num = pass_num
        self.hp_del_bias = hp_del_bias
        if accuracy_mean_set:
            accuracy_mean = truncate_accuracy(accuracy_mean)
        self.accuracy_mean = accuracy_mean
        self._keep = []
        self._chk(self.L.orc_set_params(self.h, self.method, pass_num, accuracy_mean, len_min, len_max, len_mean,
                                        len_sd, ratio[0], ratio[1], ratio[2], hp_del_bias,
                                        id_prefix.encode()))
        if method != "sample":  # --method sample has no model: thresholds only (set_mut)
            self._chk(self.L.orc_load_model(self.h, model_path.encode()))
            self._chk(self.L.orc_build_tables(self.h))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.orc_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.orc_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- draw sources
    def rng_glibc(self, seed):
        self.L.orc_rng_glibc(self.h, seed)

    def rng_replay(self, log):
        log = np.ascontiguousarray(log, dtype=np.int32)
        self._keep.append(log)
        self.L.orc_rng_replay(self.h, log.ctypes.data, log.size)

    def rng_philox(self, seed):
        self.L.orc_rng_philox(self.h, seed)

    # --- genome
    def hp_bias_prepass(self, seqs):
        """main(): the --hp-del-bias != 1 prepass over all sequences (pbsim.cpp:678-697)."""
        for s in seqs:
            self.L.orc_prepass_sequence(self.h, s, len(s))
        self.L.orc_finish_bias(self.h)

    def set_sequence(self, seq_bytes, seq_num):
        self.L.orc_set_sequence(self.h, seq_bytes, len(seq_bytes), seq_num)

    def bias(self):
        b = (C.c_double * 12)()
        self.L.orc_get_bias(self.h, b)
        return np.array(b[:], dtype=np.float64)

    def hp(self):
        n = C.c_int64()
        p = self.L.orc_get_hp(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def seq_upper(self):
        n = C.c_int64()
        p = self.L.orc_get_seq(self.h, C.byref(n))
        return C.string_at(p, n.value)

    # --- run
    def simulate_wgs(self, depth, reset=True):
        if reset:
            self.L.orc_reset_outputs(self.h)
        self._chk(self.L.orc_simulate_wgs(self.h, depth))
        n = C.c_int64()
        p = self.L.orc_out_reads(self.h, C.byref(n))
        reads = C.string_at(p, n.value)
        p = self.L.orc_out_maf(self.h, C.byref(n))
        maf = C.string_at(p, n.value)
        st = Stats()
        self.L.orc_get_stats(self.h, C.byref(st))
        return reads, maf, st

    def simulate_set(self, strategy, seqset, reset=True):
        """strategy 'trans' | 'templ'; seqset = list of (name, plus, minus, bases) (plus/minus ignored for templ)"""
        if reset:
            self.L.orc_reset_outputs(self.h)
        bases, start, plus, minus, ids, id_start = pack_set(seqset)
        self._chk(self.L.orc_simulate_set(self.h, 1 if strategy == "trans" else 2, len(seqset), bases,
                                          start.ctypes.data, plus.ctypes.data, minus.ctypes.data, ids,
                                          id_start.ctypes.data))
        n = C.c_int64()
        p = self.L.orc_out_reads(self.h, C.byref(n))
        reads = C.string_at(p, n.value)
        p = self.L.orc_out_maf(self.h, C.byref(n))
        maf = C.string_at(p, n.value)
        st = Stats()
        self.L.orc_get_stats(self.h, C.byref(st))
        return reads, maf, st

    def simulate_sample(self, depth, pool, reset=True):
        """pool: the quality strings get_sample_inf kept (sample_pool()), in file order"""
        if reset:
            self.L.orc_reset_outputs(self.h)
        quals = b"".join(pool)
        qstart = np.zeros(len(pool) + 1, dtype=np.int64)
        qstart[1:] = np.cumsum([len(x) for x in pool])
        self._chk(self.L.orc_simulate_sample(self.h, depth, len(pool), quals, qstart.ctypes.data))
        n = C.c_int64()
        p = self.L.orc_out_reads(self.h, C.byref(n))
        reads = C.string_at(p, n.value)
        p = self.L.orc_out_maf(self.h, C.byref(n))
        maf = C.string_at(p, n.value)
        st = Stats()
        self.L.orc_get_stats(self.h, C.byref(st))
        return reads, maf, st

    def readinfo(self):
        n = C.c_int64()
        p = self.L.orc_get_readinfo(self.h, C.byref(n))
        if n.value == 0:
            return np.empty(0, dtype=READINFO_DTYPE)
        buf = C.string_at(p, n.value * READINFO_DTYPE.itemsize)
        return np.frombuffer(buf, dtype=READINFO_DTYPE).copy()

    def draw_log(self):
        n = C.c_int64()
        p = self.L.orc_get_draw_log(self.h, C.byref(n))
        if n.value == 0:
            return np.empty(0, dtype=np.int32)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def draws_consumed(self):
        return self.L.orc_draws_consumed(self.h)

    def freq_len(self):
        n = C.c_int64()
        p = self.L.orc_get_freq_len(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def freq_accuracy(self):
        n = C.c_int64()
        p = self.L.orc_get_freq_accuracy(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    # --- tables
    def table(self, which, acc=0, state=0, cap=100001):
        out = np.zeros(cap, dtype=np.int32)
        n = self.L.orc_get_table(self.h, which, acc, state, out.ctypes.data, cap)
        return out[: max(0, min(n, cap))].copy(), n

    def emis2del(self, acc, state):
        return self.L.orc_get_emis2del(self.h, acc, state)

    def thresholds(self):
        s = np.zeros(94, dtype=np.int64)
        i = np.zeros(94, dtype=np.int64)
        d = np.zeros(94, dtype=np.int64)
        self.L.orc_get_thresholds(self.h, s.ctypes.data, i.ctypes.data, d.ctypes.data)
        return s, i, d

    def model_exists(self, acc):
        return bool(self.L.orc_model_exists(self.h, acc))

    def model_range(self):
        v = [C.c_int() for _ in range(4)]
        self.L.orc_model_range(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)


def sample_pool(fastq, len_min=100, len_max=1000000, accuracy_min=0.75, accuracy_max=1.0):
    """get_sample_inf's filter (pbsim.cpp:1214-1275): the quality lines of the 4-line FASTQ records whose length
    lies in [len_min, len_max] and whose accuracy 1 - mean(10^(-q/10)) lies in [accuracy_min, accuracy_max]."""
    prob = [pow(10, i / -10) for i in range(94)]  # qc[i].prob (:549)
    pool = []
    lines = fastq.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    for k in range(3, len(lines), 4):
        q = lines[k]
        n = len(q)
        if n < len_min or n > len_max:
            continue
        total = 0.0
        for ch in q:
            total += prob[ch - 33]
        acc = 1.0 - (total / n)
        if accuracy_min <= acc <= accuracy_max:
            pool.append(q)
    return pool


def pack_set(seqset):
    """[(name, plus, minus, bases)] -> (bases, start[n+1], plus[n], minus[n], ids, id_start[n+1])"""
    bases = b"".join(x[3] for x in seqset)
    start = np.zeros(len(seqset) + 1, dtype=np.int64)
    start[1:] = np.cumsum([len(x[3]) for x in seqset])
    plus = np.array([x[1] for x in seqset], dtype=np.int32)
    minus = np.array([x[2] for x in seqset], dtype=np.int32)
    names = [x[0].encode() if isinstance(x[0], str) else x[0] for x in seqset]
    ids = b"".join(names)
    id_start = np.zeros(len(seqset) + 1, dtype=np.int32)
    id_start[1:] = np.cumsum([len(x) for x in names])
    return bases, start, plus, minus, ids, id_start


def ssp_table(rank_max):
    ends = np.full((rank_max + 1) * 21, -1, dtype=np.int32)
    mod = np.zeros(rank_max + 1, dtype=np.int32)
    lib().orc_get_ssp(rank_max, ends.ctypes.data, mod.ctypes.data)
    return ends.reshape(rank_max + 1, 21), mod


def format_stats_set(st):
    """print_simulation_stats for the transcript / template strategies (pbsim.cpp:5547-5564)."""
    return (
        ":::: Simulation stats ::::\n\n"
        + "read num. : %d\n" % st.res_num
        + "read length mean (SD) : %f (%f)\n" % (st.res_len_mean, st.res_len_sd)
        + "read length min : %d\n" % st.res_len_min
        + "read length max : %d\n" % st.res_len_max
        + "read accuracy mean (SD) : %f (%f)\n" % (st.res_accuracy_mean, st.res_accuracy_sd)
        + "substitution rate. : %f\n" % st.res_sub_rate
        + "insertion rate. : %f\n" % st.res_ins_rate
        + "deletion rate. : %f\n" % st.res_del_rate
        + "\n"
    )


def format_stats(st, seq_num):
    """print_simulation_stats for WGS (pbsim.cpp:5541-5564), as text."""
    return (
        ":::: Simulation stats (ref.%d) ::::\n\n" % seq_num
        + "read num. : %d\n" % st.res_num
        + "depth : %f\n" % st.res_depth
        + "read length mean (SD) : %f (%f)\n" % (st.res_len_mean, st.res_len_sd)
        + "read length min : %d\n" % st.res_len_min
        + "read length max : %d\n" % st.res_len_max
        + "read accuracy mean (SD) : %f (%f)\n" % (st.res_accuracy_mean, st.res_accuracy_sd)
        + "substitution rate. : %f\n" % st.res_sub_rate
        + "insertion rate. : %f\n" % st.res_ins_rate
        + "deletion rate. : %f\n" % st.res_del_rate
        + "\n"
    )
