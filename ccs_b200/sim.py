"""Synthetic Sequel-II-shape ZMWs (ctypes over libccssim.so, include/ccssim.h -- not the product library)."""
import ctypes as C
import numpy as np
from ._lib import simlib as lib


class SimConfig(C.Structure):
    _fields_ = [("insert_mean", C.c_int32), ("insert_sd", C.c_int32), ("passes_min", C.c_int32),
                ("passes_max", C.c_int32), ("partials", C.c_int32), ("snr_mean", C.c_double * 4),
                ("snr_sd", C.c_double), ("frac_low_snr", C.c_double), ("frac_few_passes", C.c_double),
                ("seed", C.c_uint64)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def synthetic_model():
    L = lib()
    buf = np.zeros(L.ccs_model_sizeof(), dtype=np.uint8)
    L.ccs_model_synthetic(buf.ctypes.data_as(C.c_void_p))
    return buf


def get_config(config_id, **overrides):
    cfg = SimConfig()
    lib().ccs_sim_get_config(int(config_id), C.byref(cfg))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


class Zmw:
    """One simulated ZMW: truth template + reads (emission codes) + truth mapping."""
    __slots__ = ("hole", "snr", "tpl", "codes", "read_off", "cx", "strand", "tstart", "tend")

    @property
    def n_reads(self):
        return len(self.cx)

    def read(self, k):
        return self.codes[self.read_off[k]:self.read_off[k + 1]]


def simulate_zmw(model, cfg, index, max_len=70000, max_reads=64):
    L = lib()
    snr = np.zeros(4, np.float32)
    tpl = np.zeros(max_len, np.uint8)
    tlen = C.c_int32()
    cap = int(max_len * 1.3) * max_reads
    codes = np.zeros(cap, np.uint8)
    nr = C.c_int32()
    off = np.zeros(max_reads + 1, np.int64)
    cx = np.zeros(max_reads, np.uint8)
    strand = np.zeros(max_reads, np.uint8)
    ts = np.zeros(max_reads, np.int32)
    te = np.zeros(max_reads, np.int32)
    rc = L.ccs_sim_zmw(model.ctypes.data_as(C.c_void_p), C.byref(cfg), C.c_int64(index), _p(snr, C.c_float),
                       _p(tpl, C.c_uint8), C.c_int32(max_len), C.byref(tlen), _p(codes, C.c_uint8), C.c_int64(cap),
                       C.c_int32(max_reads), C.byref(nr), _p(off, C.c_int64), _p(cx, C.c_uint8), _p(strand, C.c_uint8),
                       _p(ts, C.c_int32), _p(te, C.c_int32))
    if rc != 0:
        raise RuntimeError(f"ccs_sim_zmw failed: {rc}")
    z = Zmw()
    n = nr.value
    z.hole = int(index)
    z.snr = snr
    z.tpl = tpl[:tlen.value].copy()
    z.read_off = off[:n + 1].copy()
    z.codes = codes[:off[n]].copy()
    z.cx = cx[:n].copy()
    z.strand = strand[:n].copy()
    z.tstart = ts[:n].copy()
    z.tend = te[:n].copy()
    return z


def corrupt(tpl, rate, seed):
    L = lib()
    out = np.zeros(len(tpl) * 2 + 16, np.uint8)
    olen = C.c_int32()
    mp = np.zeros(len(tpl) + 1, np.int32)
    tpl = np.ascontiguousarray(tpl, np.uint8)
    rc = L.ccs_sim_corrupt(_p(tpl, C.c_uint8), C.c_int32(len(tpl)), C.c_double(rate), C.c_uint64(seed),
                           _p(out, C.c_uint8), C.c_int32(len(out)), C.byref(olen), _p(mp, C.c_int32))
    if rc != 0:
        raise RuntimeError(f"ccs_sim_corrupt failed: {rc}")
    return out[:olen.value].copy(), mp


def simulate_batch(model, cfg, first_index, n_zmws, draft_error_rate=-1.0, threads=8):
    """ZMWs [first, first+n) as SoA numpy arrays (multi-threaded C++).  Returns a dict with the
    batch arrays, the truth templates/spans and (if draft_error_rate >= 0) corrupted drafts with
    the spans mapped onto them."""
    L = lib()
    L.ccs_sim_batch_create.restype = C.c_void_p
    h = L.ccs_sim_batch_create(model.ctypes.data_as(C.c_void_p), C.byref(cfg), C.c_int64(first_index),
                               C.c_int32(n_zmws), C.c_double(draft_error_rate), C.c_int32(threads))
    try:
        sizes = np.zeros(4, np.int64)
        L.ccs_sim_batch_sizes(C.c_void_p(h), _p(sizes, C.c_int64))
        nr, nc, nt, nd = [int(x) for x in sizes]
        out = dict(zmw_read_off=np.zeros(n_zmws + 1, np.int32), read_off=np.zeros(nr + 1, np.int64),
                   codes=np.zeros(nc, np.uint8), snr=np.zeros(4 * n_zmws, np.float32), cx=np.zeros(nr, np.uint8),
                   hole=np.zeros(n_zmws, np.int32), truth_off=np.zeros(n_zmws + 1, np.int64),
                   truth=np.zeros(nt, np.uint8), strand=np.zeros(nr, np.uint8), tstart=np.zeros(nr, np.int32),
                   tend=np.zeros(nr, np.int32), draft_off=np.zeros(n_zmws + 1, np.int64),
                   draft=np.zeros(max(nd, 1), np.uint8), dstart=np.zeros(nr, np.int32), dend=np.zeros(nr, np.int32))
        L.ccs_sim_batch_copy(C.c_void_p(h), _p(out["zmw_read_off"], C.c_int32), _p(out["read_off"], C.c_int64),
                             _p(out["codes"], C.c_uint8), _p(out["snr"], C.c_float), _p(out["cx"], C.c_uint8),
                             _p(out["hole"], C.c_int32), _p(out["truth_off"], C.c_int64), _p(out["truth"], C.c_uint8),
                             _p(out["strand"], C.c_uint8), _p(out["tstart"], C.c_int32), _p(out["tend"], C.c_int32),
                             _p(out["draft_off"], C.c_int64), _p(out["draft"], C.c_uint8), _p(out["dstart"], C.c_int32),
                             _p(out["dend"], C.c_int32))
        out["draft"] = out["draft"][:nd]
        return out
    finally:
        L.ccs_sim_batch_free(C.c_void_p(h))
