"""ctypes access to the CPU oracle (oracle/liboracle.so) -- the checker, never the product."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = None

MUT_SUB, MUT_INS, MUT_DEL = 0, 1, 2


def olib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _lib = C.CDLL(_PATH)
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def synthetic_model():
    L = olib()
    buf = np.zeros(L.oracle_model_sizeof(), dtype=np.uint8)
    L.oracle_synthetic_model(_vp(buf))
    return buf


def tables(model, snr):
    snr = np.ascontiguousarray(snr, np.float32)
    em_match = np.zeros((36, 16)); em_ins = np.zeros((17, 16)); tr = np.zeros((36, 4)); lcw = C.c_double()
    olib().oracle_get_tables(_vp(model), _p(snr, C.c_float), _p(em_match, C.c_double), _p(em_ins, C.c_double),
                             _p(tr, C.c_double), C.byref(lcw))
    return em_match, em_ins, tr, lcw.value


def folded(model, snr):
    """The factors the recursion multiplies by: mm = fl32(em_match * match), gg = fl32(em_ins * (branch | stick))."""
    snr = np.ascontiguousarray(snr, np.float32)
    mm = np.zeros((36, 16)); gg = np.zeros((17, 16))
    olib().oracle_get_folded(_vp(model), _p(snr, C.c_float), _p(mm, C.c_double), _p(gg, C.c_double))
    return mm, gg


def fill(model, snr, tpl, codes, W=32, precision=0, dump=False):
    snr = np.ascontiguousarray(snr, np.float32)
    tpl = np.ascontiguousarray(tpl, np.uint8); codes = np.ascontiguousarray(codes, np.uint8)
    J, I = len(tpl), len(codes)
    la, lb, cells = C.c_double(), C.c_double(), C.c_int64()
    a = b = st = ae = be = None
    if dump:
        a = np.zeros((J, W), np.float32); b = np.zeros((J, W), np.float32)
        st = np.zeros(J, np.int32); ae = np.zeros(J, np.int64); be = np.zeros(J, np.int64)
    status = olib().oracle_fill(_vp(model), _p(snr, C.c_float), _p(tpl, C.c_uint8), J, _p(codes, C.c_uint8), I, W,
                                precision, C.byref(la), C.byref(lb), C.byref(cells), _p(a, C.c_float), _p(b, C.c_float),
                                _p(st, C.c_int32), _p(ae, C.c_int64), _p(be, C.c_int64))
    out = dict(status=status, ll_alpha=la.value, ll_beta=lb.value, cells=cells.value)
    if dump:
        out.update(alpha=a, beta=b, start=st, aexp=ae, bexp=be)
    return out


def score(model, snr, tpl, codes, muts, W=32, precision=0, full=True):
    snr = np.ascontiguousarray(snr, np.float32)
    tpl = np.ascontiguousarray(tpl, np.uint8); codes = np.ascontiguousarray(codes, np.uint8)
    n = len(muts)
    ty = np.array([m[0] for m in muts], np.int32); po = np.array([m[1] for m in muts], np.int32)
    ba = np.array([m[2] for m in muts], np.int32)
    inc = np.zeros(n); ful = np.zeros(n) if full else None
    rc = olib().oracle_score(_vp(model), _p(snr, C.c_float), _p(tpl, C.c_uint8), len(tpl), _p(codes, C.c_uint8),
                             len(codes), W, precision, n, _p(ty, C.c_int32), _p(po, C.c_int32), _p(ba, C.c_int32),
                             _p(inc, C.c_double), _p(ful, C.c_double))
    if rc != 0:
        raise RuntimeError(f"oracle_score status {rc}")
    return inc, ful


def _pack_reads(reads):
    off = np.zeros(len(reads) + 1, np.int64)
    for k, r in enumerate(reads):
        off[k + 1] = off[k] + len(r)
    codes = np.concatenate([np.asarray(r, np.uint8) for r in reads]) if reads else np.zeros(0, np.uint8)
    return np.ascontiguousarray(codes), off


def polish(model, snr, draft, reads, strand, tstart, tend, W=32, max_iter=-1, precision=0):
    snr = np.ascontiguousarray(snr, np.float32)
    draft = np.ascontiguousarray(draft, np.uint8)
    codes, off = _pack_reads(reads)
    n = len(reads)
    strand = np.ascontiguousarray(strand, np.int32); tstart = np.ascontiguousarray(tstart, np.int32)
    tend = np.ascontiguousarray(tend, np.int32)
    cap = len(draft) * 2 + 64
    cons = np.zeros(cap, np.uint8); qv = np.zeros(cap, np.uint8); clen = C.c_int32(); rq = C.c_double()
    stats = np.zeros(8, np.int64); rll = np.zeros(n); rst = np.zeros(n, np.int32)
    rc = olib().oracle_polish(_vp(model), _p(snr, C.c_float), _p(draft, C.c_uint8), len(draft), n, _p(codes, C.c_uint8),
                              _p(off, C.c_int64), _p(strand, C.c_int32), _p(tstart, C.c_int32), _p(tend, C.c_int32), W,
                              max_iter, precision, _p(cons, C.c_uint8), cap, C.byref(clen), _p(qv, C.c_uint8),
                              C.byref(rq), _p(stats, C.c_int64), _p(rll, C.c_double), _p(rst, C.c_int32))
    if rc != 0:
        raise RuntimeError(f"oracle_polish rc {rc}")
    L = clen.value
    return dict(consensus=cons[:L].copy(), qv=qv[:L].copy(), rq=rq.value, converged=bool(stats[0]),
                iterations=int(stats[1]), n_tested=int(stats[2]), n_applied=int(stats[3]), n_active=int(stats[4]),
                cells=int(stats[5]), read_ll=rll, read_status=rst)


def score_all(model, snr, draft, reads, strand, tstart, tend, W=32):
    snr = np.ascontiguousarray(snr, np.float32)
    draft = np.ascontiguousarray(draft, np.uint8)
    codes, off = _pack_reads(reads)
    n = len(reads)
    strand = np.ascontiguousarray(strand, np.int32); tstart = np.ascontiguousarray(tstart, np.int32)
    tend = np.ascontiguousarray(tend, np.int32)
    out = np.zeros((len(draft) + 1, 9)); rll = np.zeros(n)
    olib().oracle_score_all(_vp(model), _p(snr, C.c_float), _p(draft, C.c_uint8), len(draft), n, _p(codes, C.c_uint8),
                            _p(off, C.c_int64), _p(strand, C.c_int32), _p(tstart, C.c_int32), _p(tend, C.c_int32), W,
                            _p(out, C.c_double), _p(rll, C.c_double))
    return out, rll


def _cfg_arrays(min_passes=3, top_passes=60, max_poa_reads=5, min_length=10, max_length=50000, max_iterations=-1,
                min_snr=2.5, min_rq=0.99, min_active_fraction=0.5, min_zscore=-3.4, window_size=-1, window_overlap=-1):
    ci = np.array([min_passes, top_passes, max_poa_reads, min_length, max_length, max_iterations, window_size,
                   window_overlap], np.int32)
    cd = np.array([min_snr, min_rq, min_active_fraction, min_zscore])
    return ci, cd


def draft_zmw(snr, reads, cx, **cfg):
    ci, cd = _cfg_arrays(**cfg)
    snr = np.ascontiguousarray(snr, np.float32)
    codes, off = _pack_reads(reads)
    cx = np.ascontiguousarray(cx, np.uint8)
    n = len(reads)
    cap = int(max([len(r) for r in reads] + [1]) * 2 + 64)
    draft = np.zeros(cap, np.uint8); dlen = C.c_int32(); maps = np.zeros((n, 6), np.int32); st = C.c_int32()
    rc = olib().oracle_draft_zmw(_p(ci, C.c_int32), _p(cd, C.c_double), _p(snr, C.c_float), n, _p(codes, C.c_uint8),
                                 _p(off, C.c_int64), _p(cx, C.c_uint8), _p(draft, C.c_uint8), cap, C.byref(dlen),
                                 _p(maps, C.c_int32), C.byref(st))
    if rc != 0:
        raise RuntimeError("oracle_draft_zmw capacity")
    return dict(status=st.value, draft=draft[:dlen.value].copy(), maps=maps)


def ccs_zmw(model, snr, reads, cx, **cfg):
    ci, cd = _cfg_arrays(**cfg)
    snr = np.ascontiguousarray(snr, np.float32)
    codes, off = _pack_reads(reads)
    cx = np.ascontiguousarray(cx, np.uint8)
    n = len(reads)
    cap = int(max([len(r) for r in reads] + [1]) * 2 + 64)
    seq = np.zeros(cap, np.uint8); qv = np.zeros(cap, np.uint8); slen = C.c_int32(); rq = C.c_double()
    stats = np.zeros(8, np.int64); rll = np.zeros(n); rst = np.zeros(n, np.int32)
    rc = olib().oracle_ccs_zmw(_vp(model), _p(ci, C.c_int32), _p(cd, C.c_double), _p(snr, C.c_float), n,
                               _p(codes, C.c_uint8), _p(off, C.c_int64), _p(cx, C.c_uint8), _p(seq, C.c_uint8), cap,
                               C.byref(slen), _p(qv, C.c_uint8), C.byref(rq), _p(stats, C.c_int64), _p(rll, C.c_double),
                               _p(rst, C.c_int32))
    if rc != 0:
        raise RuntimeError("oracle_ccs_zmw capacity")
    L = slen.value
    return dict(status=int(stats[0]), np=int(stats[1]), converged=bool(stats[2]), iterations=int(stats[3]),
                n_tested=int(stats[4]), n_applied=int(stats[5]), draft_len=int(stats[6]), seq=seq[:L].copy(),
                qv=qv[:L].copy(), rq=rq.value, read_ll=rll, read_status=rst)
