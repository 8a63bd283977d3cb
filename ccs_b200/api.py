"""ctypes mirror of the Draft/Polish stage C ABI (include/ccsgpu.h).

Everything here runs on the GPU through libccsgpu.so; there is no CPU path.
"""
import ctypes as C
import numpy as np
from ._lib import lib

CCS_OK = 0
ZMW_STATUS = ["POOR_SNR", "NO_SUBREADS", "TOO_FEW_PASSES", "LOW_PASS_SHORTCUT", "HETERODUPLEXES", "COVERAGE_DROPS",
              "INSUFFICIENT_SPANS", "TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT", "DRAFT_FAILURE", "TOO_LONG", "TOO_SHORT",
              "TOO_MANY_UNUSABLE", "EMPTY_WINDOW_DURING_POLISHING", "NON_CONVERGENT", "POOR_QUALITY",
              "EXCEPTION_THROWN", "SUCCESS"]
ZMW_SUCCESS = 16


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class CBatch(C.Structure):
    _fields_ = [("n_zmws", C.c_int32), ("n_reads", C.c_int32), ("zmw_read_off", C.POINTER(C.c_int32)),
                ("read_off", C.POINTER(C.c_int64)), ("codes", C.POINTER(C.c_uint8)), ("snr", C.POINTER(C.c_float)),
                ("cx", C.POINTER(C.c_uint8)), ("hole", C.POINTER(C.c_int32))]


class CDrafts(C.Structure):
    _fields_ = [("tpl_off", C.POINTER(C.c_int64)), ("tpl", C.POINTER(C.c_uint8)), ("strand", C.POINTER(C.c_uint8)),
                ("tstart", C.POINTER(C.c_int32)), ("tend", C.POINTER(C.c_int32)), ("rstart", C.POINTER(C.c_int32)),
                ("rend", C.POINTER(C.c_int32))]


class CDraftCfg(C.Structure):
    _fields_ = [("min_snr", C.c_double), ("min_passes", C.c_int32), ("top_passes", C.c_int32),
                ("max_poa_reads", C.c_int32), ("min_length", C.c_int32), ("max_length", C.c_int32)]


class CDraftsOut(C.Structure):
    _fields_ = [("tpl_cap", C.c_int64), ("tpl_off", C.POINTER(C.c_int64)), ("tpl", C.POINTER(C.c_uint8)),
                ("strand", C.POINTER(C.c_uint8)), ("tstart", C.POINTER(C.c_int32)), ("tend", C.POINTER(C.c_int32)),
                ("rstart", C.POINTER(C.c_int32)), ("rend", C.POINTER(C.c_int32)), ("status", C.POINTER(C.c_int32))]


class CPolishCfg(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("separation", C.c_int32), ("neighborhood", C.c_int32),
                ("min_length", C.c_int32), ("max_length", C.c_int32), ("min_rq", C.c_double),
                ("ab_mismatch_tol", C.c_double), ("min_active_fraction", C.c_double), ("min_zscore", C.c_double),
                ("window_size", C.c_int32), ("window_overlap", C.c_int32)]


class CResults(C.Structure):
    _fields_ = [("seq_cap", C.c_int64), ("seq_off", C.POINTER(C.c_int64)), ("seq", C.POINTER(C.c_uint8)),
                ("qv", C.POINTER(C.c_uint8)), ("rq", C.POINTER(C.c_float)), ("status", C.POINTER(C.c_int32)),
                ("n_passes", C.POINTER(C.c_int32)), ("iterations", C.POINTER(C.c_int32)),
                ("n_applied", C.POINTER(C.c_int32)), ("n_tested", C.POINTER(C.c_int64)),
                ("read_ll", C.POINTER(C.c_double)), ("read_status", C.POINTER(C.c_int32))]


class CStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_fill_alpha", "ms_fill_beta", "ms_score", "ms_pick", "ms_qv", "ms_h2d",
                                          "ms_draft")] + \
               [(n, C.c_int64) for n in ("launches_fill_alpha", "launches_fill_beta", "launches_score", "launches_pick",
                                         "launches_qv", "launches_draft", "bytes_fill_alpha", "bytes_fill_beta",
                                         "cells_fill", "score_items", "rounds", "h2d_bytes", "d2h_bytes")] + \
               [("ms_poa_align", C.c_double)] + \
               [(n, C.c_int64) for n in ("launches_poa", "poa_tasks", "poa_rows", "bytes_poa_align")] + \
               [("ms_resident", C.c_double), ("ms_e2e", C.c_double), ("n_zmws", C.c_int64),
                ("top_fill_alpha_bytes", C.c_int64), ("top_fill_alpha_ms", C.c_double),
                ("ms_poa_map", C.c_double), ("ms_poa_graph", C.c_double), ("launches_poa_graph", C.c_int64),
                ("bytes_poa_map", C.c_int64), ("bytes_score", C.c_int64), ("launches_pack", C.c_int64),
                ("top_fill_beta_bytes", C.c_int64), ("top_fill_beta_ms", C.c_double),
                ("top_score_bytes", C.c_int64), ("top_score_ms", C.c_double),
                ("top_poa_align_bytes", C.c_int64), ("top_poa_align_ms", C.c_double),
                ("top_poa_map_bytes", C.c_int64), ("top_poa_map_ms", C.c_double)]


class Batch:
    """Host-side SoA batch of ZMWs (+ optional drafts), numpy arrays kept alive for the C call."""

    def __init__(self, zmws, drafts=None):
        """zmws: list of sim.Zmw-like objects (snr, codes, read_off, cx, hole).
        drafts: optional list of (tpl, strand[], tstart[], tend[]) per ZMW."""
        self.n_zmws = len(zmws)
        nr = [z.n_reads for z in zmws]
        self.zmw_read_off = np.zeros(self.n_zmws + 1, np.int32)
        self.zmw_read_off[1:] = np.cumsum(nr)
        self.n_reads = int(self.zmw_read_off[-1])
        lens = np.concatenate([np.diff(z.read_off) for z in zmws]) if zmws else np.zeros(0, np.int64)
        self.read_off = np.zeros(self.n_reads + 1, np.int64)
        self.read_off[1:] = np.cumsum(lens)
        self.codes = np.ascontiguousarray(np.concatenate([z.codes for z in zmws])) if zmws else np.zeros(0, np.uint8)
        self.snr = np.ascontiguousarray(np.concatenate([z.snr for z in zmws]).astype(np.float32))
        self.cx = np.ascontiguousarray(np.concatenate([z.cx for z in zmws]).astype(np.uint8))
        self.hole = np.array([z.hole for z in zmws], np.int32)
        self.c = CBatch(self.n_zmws, self.n_reads, _p(self.zmw_read_off, C.c_int32), _p(self.read_off, C.c_int64),
                        _p(self.codes, C.c_uint8), _p(self.snr, C.c_float), _p(self.cx, C.c_uint8),
                        _p(self.hole, C.c_int32))
        self.d = None
        if drafts is not None:
            self.set_drafts(drafts)

    @classmethod
    def from_arrays(cls, zmw_read_off, read_off, codes, snr, cx, hole, tpl_off=None, tpl=None, strand=None,
                    tstart=None, tend=None):
        b = cls.__new__(cls)
        b.n_zmws = len(zmw_read_off) - 1
        b.n_reads = int(zmw_read_off[-1])
        b.zmw_read_off = np.ascontiguousarray(zmw_read_off, np.int32)
        b.read_off = np.ascontiguousarray(read_off, np.int64)
        b.codes = np.ascontiguousarray(codes, np.uint8)
        b.snr = np.ascontiguousarray(snr, np.float32)
        b.cx = np.ascontiguousarray(cx, np.uint8)
        b.hole = np.ascontiguousarray(hole, np.int32)
        b.c = CBatch(b.n_zmws, b.n_reads, _p(b.zmw_read_off, C.c_int32), _p(b.read_off, C.c_int64),
                     _p(b.codes, C.c_uint8), _p(b.snr, C.c_float), _p(b.cx, C.c_uint8), _p(b.hole, C.c_int32))
        b.d = None
        if tpl_off is not None:
            b.tpl_off = np.ascontiguousarray(tpl_off, np.int64)
            b.tpl = np.ascontiguousarray(tpl, np.uint8)
            b.strand = np.ascontiguousarray(strand, np.uint8)
            b.tstart = np.ascontiguousarray(tstart, np.int32)
            b.tend = np.ascontiguousarray(tend, np.int32)
            b.d = CDrafts(_p(b.tpl_off, C.c_int64), _p(b.tpl, C.c_uint8), _p(b.strand, C.c_uint8),
                          _p(b.tstart, C.c_int32), _p(b.tend, C.c_int32), None, None)
        return b

    def zmw_reads(self, z):
        """list of code arrays of ZMW z"""
        r0, r1 = self.zmw_read_off[z], self.zmw_read_off[z + 1]
        return [self.codes[self.read_off[r]:self.read_off[r + 1]] for r in range(r0, r1)]

    def set_drafts(self, drafts):
        self.tpl_off = np.zeros(self.n_zmws + 1, np.int64)
        self.tpl_off[1:] = np.cumsum([len(d[0]) for d in drafts])
        self.tpl = np.ascontiguousarray(np.concatenate([np.asarray(d[0], np.uint8) for d in drafts]))
        self.strand = np.ascontiguousarray(np.concatenate([np.asarray(d[1], np.uint8) for d in drafts]))
        self.tstart = np.ascontiguousarray(np.concatenate([np.asarray(d[2], np.int32) for d in drafts]))
        self.tend = np.ascontiguousarray(np.concatenate([np.asarray(d[3], np.int32) for d in drafts]))
        self.d = CDrafts(_p(self.tpl_off, C.c_int64), _p(self.tpl, C.c_uint8), _p(self.strand, C.c_uint8),
                         _p(self.tstart, C.c_int32), _p(self.tend, C.c_int32), None, None)


class CcsGpuError(RuntimeError):
    pass


class Context:
    """One GPU context (ccsgpu_create).  Raises if there is no CUDA device: no CPU fallback."""

    def __init__(self, model, device=0, budget_bytes=0):
        L = lib()
        L.ccsgpu_create.restype = C.c_void_p
        L.ccsgpu_last_error.restype = C.c_char_p
        err = C.c_int()
        self._L = L
        self.model = model
        self._h = L.ccsgpu_create(C.c_int(device), model.ctypes.data_as(C.c_void_p), C.c_size_t(budget_bytes),
                                  C.byref(err))
        if not self._h:
            raise CcsGpuError(f"ccsgpu_create failed ({err.value}): {L.ccsgpu_last_error(None).decode()}")

    def close(self):
        if self._h:
            self._L.ccsgpu_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise CcsGpuError(f"{what} failed ({rc}): {self._L.ccsgpu_last_error(C.c_void_p(self._h)).decode()}")

    def fill_alpha_beta(self, pairs, dump_pair=-1):
        """pairs: list of (snr[4], tpl, codes).  Returns dict(ll_alpha, ll_beta, status[, dumps])."""
        n = len(pairs)
        tpl_off = np.zeros(n + 1, np.int64); read_off = np.zeros(n + 1, np.int64)
        tpl_off[1:] = np.cumsum([len(p[1]) for p in pairs]); read_off[1:] = np.cumsum([len(p[2]) for p in pairs])
        tpl = np.ascontiguousarray(np.concatenate([np.asarray(p[1], np.uint8) for p in pairs]))
        codes = np.ascontiguousarray(np.concatenate([np.asarray(p[2], np.uint8) for p in pairs]))
        snr = np.ascontiguousarray(np.concatenate([np.asarray(p[0], np.float32) for p in pairs]))
        la = np.zeros(n); lb = np.zeros(n); st = np.zeros(n, np.int32)
        a = b = s = ae = be = None
        if dump_pair >= 0:
            J = len(pairs[dump_pair][1])
            a = np.zeros((J, 32), np.float32); b = np.zeros((J, 32), np.float32)
            s = np.zeros(J, np.int32); ae = np.zeros(J, np.int32); be = np.zeros(J, np.int32)
        rc = self._L.ccsgpu_fill_alpha_beta(C.c_void_p(self._h), n, _p(tpl_off, C.c_int64), _p(tpl, C.c_uint8),
                                            _p(read_off, C.c_int64), _p(codes, C.c_uint8), _p(snr, C.c_float),
                                            _p(la, C.c_double), _p(lb, C.c_double), _p(st, C.c_int32), dump_pair,
                                            _p(a, C.c_float), _p(b, C.c_float), _p(s, C.c_int32), _p(ae, C.c_int32),
                                            _p(be, C.c_int32))
        self._check(rc, "ccsgpu_fill_alpha_beta")
        out = dict(ll_alpha=la, ll_beta=lb, status=st)
        if dump_pair >= 0:
            out.update(alpha=a, beta=b, start=s, aexp=ae, bexp=be)
        return out

    def score_all(self, batch):
        delta = np.zeros((int(batch.tpl_off[-1]), 9))
        rll = np.zeros(batch.n_reads); rst = np.zeros(batch.n_reads, np.int32)
        rc = self._L.ccsgpu_score_all(C.c_void_p(self._h), C.byref(batch.c), C.byref(batch.d), _p(delta, C.c_double),
                                      _p(rll, C.c_double), _p(rst, C.c_int32))
        self._check(rc, "ccsgpu_score_all")
        return delta, rll, rst

    def default_polish_cfg(self):
        cfg = CPolishCfg()
        self._L.ccs_polish_cfg_default(C.byref(cfg))
        return cfg

    def polish(self, batch, cfg=None):
        if cfg is None:
            cfg = self.default_polish_cfg()
        nz, nr = batch.n_zmws, batch.n_reads
        cap = int(batch.tpl_off[-1]) + 1024 * nz + 1024
        while True:
            seq_off = np.zeros(nz + 1, np.int64); seq = np.zeros(cap, np.uint8); qv = np.zeros(cap, np.uint8)
            rq = np.zeros(nz, np.float32); status = np.zeros(nz, np.int32); npass = np.zeros(nz, np.int32)
            its = np.zeros(nz, np.int32); napp = np.zeros(nz, np.int32); ntest = np.zeros(nz, np.int64)
            rll = np.zeros(nr); rst = np.zeros(nr, np.int32)
            res = CResults(cap, _p(seq_off, C.c_int64), _p(seq, C.c_uint8), _p(qv, C.c_uint8), _p(rq, C.c_float),
                           _p(status, C.c_int32), _p(npass, C.c_int32), _p(its, C.c_int32), _p(napp, C.c_int32),
                           _p(ntest, C.c_int64), _p(rll, C.c_double), _p(rst, C.c_int32))
            rc = self._L.ccsgpu_polish(C.c_void_p(self._h), C.byref(batch.c), C.byref(batch.d), C.byref(cfg),
                                       C.byref(res))
            if rc == -1:
                cap = int(res.seq_cap) + 1024
                continue
            self._check(rc, "ccsgpu_polish")
            break
        return dict(seq_off=seq_off, seq=seq, qv=qv, rq=rq, status=status, n_passes=npass, iterations=its,
                    n_applied=napp, n_tested=ntest, read_ll=rll, read_status=rst)

    def default_draft_cfg(self):
        cfg = CDraftCfg()
        self._L.ccs_draft_cfg_default(C.byref(cfg))
        return cfg

    def draft(self, batch, cfg=None):
        """Draft Stage: returns dict(tpl_off, tpl, strand, tstart, tend, rstart, rend, status)."""
        if cfg is None:
            cfg = self.default_draft_cfg()
        nz, nr = batch.n_zmws, batch.n_reads
        cap = int(np.diff(batch.read_off).max(initial=1)) * nz * 2 + 1024
        while True:
            tpl_off = np.zeros(nz + 1, np.int64); tpl = np.zeros(cap, np.uint8); strand = np.zeros(nr, np.uint8)
            ts = np.zeros(nr, np.int32); te = np.zeros(nr, np.int32); rs = np.zeros(nr, np.int32)
            re = np.zeros(nr, np.int32); status = np.zeros(nz, np.int32)
            out = CDraftsOut(cap, _p(tpl_off, C.c_int64), _p(tpl, C.c_uint8), _p(strand, C.c_uint8), _p(ts, C.c_int32),
                             _p(te, C.c_int32), _p(rs, C.c_int32), _p(re, C.c_int32), _p(status, C.c_int32))
            rc = self._L.ccsgpu_draft(C.c_void_p(self._h), C.byref(batch.c), C.byref(cfg), C.byref(out))
            if rc == -1:
                cap = int(out.tpl_cap) + 1024
                continue
            self._check(rc, "ccsgpu_draft")
            break
        return dict(tpl_off=tpl_off, tpl=tpl[:tpl_off[-1]], strand=strand, tstart=ts, tend=te, rstart=rs, rend=re,
                    status=status)

    def ccs(self, batch, dcfg=None, pcfg=None):
        """Whole per-ZMW hot path (draft + polish + gates) for a batch."""
        if dcfg is None:
            dcfg = self.default_draft_cfg()
        if pcfg is None:
            pcfg = self.default_polish_cfg()
        nz, nr = batch.n_zmws, batch.n_reads
        cap = int(np.diff(batch.read_off).max(initial=1)) * nz * 2 + 1024
        while True:
            seq_off = np.zeros(nz + 1, np.int64); seq = np.zeros(cap, np.uint8); qv = np.zeros(cap, np.uint8)
            rq = np.zeros(nz, np.float32); status = np.zeros(nz, np.int32); npass = np.zeros(nz, np.int32)
            its = np.zeros(nz, np.int32); napp = np.zeros(nz, np.int32); ntest = np.zeros(nz, np.int64)
            rll = np.zeros(nr); rst = np.zeros(nr, np.int32)
            res = CResults(cap, _p(seq_off, C.c_int64), _p(seq, C.c_uint8), _p(qv, C.c_uint8), _p(rq, C.c_float),
                           _p(status, C.c_int32), _p(npass, C.c_int32), _p(its, C.c_int32), _p(napp, C.c_int32),
                           _p(ntest, C.c_int64), _p(rll, C.c_double), _p(rst, C.c_int32))
            rc = self._L.ccsgpu_ccs(C.c_void_p(self._h), C.byref(batch.c), C.byref(dcfg), C.byref(pcfg), C.byref(res))
            if rc == -1:
                cap = int(res.seq_cap) + 1024
                continue
            self._check(rc, "ccsgpu_ccs")
            break
        return dict(seq_off=seq_off, seq=seq, qv=qv, rq=rq, status=status, n_passes=npass, iterations=its,
                    n_applied=napp, n_tested=ntest, read_ll=rll, read_status=rst)

    def set_lanes(self, n):
        self._check(self._L.ccsgpu_set_lanes(C.c_void_p(self._h), int(n)), "ccsgpu_set_lanes")

    def stats(self, reset=False):
        s = CStats()
        self._check(self._L.ccsgpu_get_stats(C.c_void_p(self._h), C.byref(s), int(reset)), "ccsgpu_get_stats")
        return {n: getattr(s, n) for n, _ in CStats._fields_}
