"""GPU parity at bench scale and on the ugly paths.

 * >= 256 config-2, >= 64 config-3 and >= 64 config-5 ZMWs through ccsgpu_ccs against the oracle on all three
   north_star tolerances: identical status / consensus / iteration counts, QV within 1, per-read LL within 1e-4.
 * adversarial inputs compared with the oracle: unrelated and chimeric reads forced into the mapping (READ_DEAD,
   ALPHA_BETA_MISMATCH, POOR_ZSCORE, TOO_MANY_UNUSABLE), a POA vertex with more than 8 predecessors, a template that
   outgrows its capacity, an iteration cap.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_lib as O
from ccs_b200 import sim, api

pytestmark = pytest.mark.gpu
MODEL = sim.synthetic_model()
CORES = os.cpu_count() or 8


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(MODEL)
    yield c
    c.close()


def _oracle_many(zs):
    O.olib()
    with ThreadPoolExecutor(max_workers=CORES) as ex:
        return list(ex.map(lambda z: O.ccs_zmw(MODEL, z["snr"], z["reads"], z["cx"]), zs))


SCALE = max(1, int(os.environ.get("CCS_PARITY_SCALE", "1")))      # CCS_PARITY_SCALE=2 doubles the batches (profiles/)


@pytest.mark.parametrize("cfg_id,n", [(2, 256), (3, 64), (5, 64)])
def test_bench_scale_parity_with_oracle(ctx, cfg_id, n):
    n *= SCALE
    cfg = sim.get_config(cfg_id)
    a = sim.simulate_batch(MODEL, cfg, 20_000, n, -1.0, CORES)
    b = api.Batch.from_arrays(a["zmw_read_off"], a["read_off"], a["codes"], a["snr"], a["cx"], a["hole"])
    res = ctx.ccs(b)
    zs = []
    for z in range(n):
        r0, r1 = a["zmw_read_off"][z], a["zmw_read_off"][z + 1]
        zs.append(dict(snr=a["snr"][4 * z:4 * z + 4], cx=a["cx"][r0:r1],
                       reads=[a["codes"][a["read_off"][r]:a["read_off"][r + 1]] for r in range(r0, r1)]))
    ora = _oracle_many(zs)
    n_cons = 0
    worst_ll = 0.0
    for z, o in enumerate(ora):
        assert res["status"][z] == o["status"], (z, res["status"][z], o["status"])
        s0, s1 = res["seq_off"][z], res["seq_off"][z + 1]
        if o["status"] in (16, 14, 13):
            n_cons += 1
            assert np.array_equal(res["seq"][s0:s1], o["seq"]), z                          # bit-identical consensus
            assert np.max(np.abs(res["qv"][s0:s1].astype(int) - o["qv"].astype(int))) <= 1, z   # QV +-1
            assert res["iterations"][z] == o["iterations"] and res["n_applied"][z] == o["n_applied"], z
            assert res["n_passes"][z] == o["np"], z
            assert abs(res["rq"][z] - o["rq"]) < 2e-3, z
            r0, r1 = b.zmw_read_off[z], b.zmw_read_off[z + 1]
            both = ~np.isnan(o["read_ll"])
            assert np.array_equal(~np.isnan(res["read_ll"][r0:r1]), both), z
            assert np.array_equal(res["read_status"][r0:r1], o["read_status"]), z
            if both.any():
                worst_ll = max(worst_ll, float(np.max(np.abs(res["read_ll"][r0:r1][both] - o["read_ll"][both]))))
        else:
            assert s1 == s0
    assert worst_ll < 1e-4                                                                  # per-read LL within 1e-4
    assert n_cons >= 0.75 * n


def _polish_case(ctx, tpl, snr, reads, strand, ts, te, **kw):
    """one ZMW through ccsgpu_polish and through the oracle's polish with the same forced mapping"""
    z = sim.Zmw()
    z.hole = 1; z.snr = np.asarray(snr, np.float32)
    z.read_off = np.zeros(len(reads) + 1, np.int64)
    z.read_off[1:] = np.cumsum([len(r) for r in reads])
    z.codes = np.concatenate(reads).astype(np.uint8)
    z.cx = np.full(len(reads), 3, np.uint8)
    z.strand = np.asarray(strand, np.uint8); z.tstart = np.asarray(ts, np.int32); z.tend = np.asarray(te, np.int32)
    batch = api.Batch([z], [(tpl, z.strand, z.tstart, z.tend)])
    pc = ctx.default_polish_cfg()
    for k, v in kw.items():
        setattr(pc, k, v)
    res = ctx.polish(batch, pc)
    o = O.polish(MODEL, z.snr, tpl, reads, z.strand.astype(np.int32), z.tstart, z.tend,
                 max_iter=kw.get("max_iterations", -1))
    return res, o


def _good_zmw(insert, idx=3, n_reads=10):
    cfg = sim.get_config(1, insert_mean=insert)
    z = sim.simulate_zmw(MODEL, cfg, idx)
    full = [k for k in range(z.n_reads) if z.cx[k] == 3][:n_reads]
    return z, full


def test_unusable_reads_match_oracle(ctx):
    """Unrelated, chimeric and scrambled reads forced onto the template: whatever each one turns into (dead band,
    alpha/beta mismatch, poor z-score), the GPU agrees with the oracle read by read; the rest of the ZMW polishes to
    the same consensus; with too many of them the ZMW fails with TOO_MANY_UNUSABLE on both sides."""
    z, full = _good_zmw(1500)
    other = sim.simulate_zmw(MODEL, sim.get_config(1, insert_mean=1500), 99)
    rng = np.random.default_rng(5)
    J = len(z.tpl)
    draft, mp = sim.corrupt(z.tpl, 0.02, seed=11)
    def base_reads():
        return [z.read(k).copy() for k in full], [int(z.strand[k]) for k in full], [0] * len(full), [len(draft)] * len(full)
    # (a) one unrelated read, one chimera (first half right, second half from another molecule), one with a 150-base
    #     random block in the middle, one truncated read stretched over the whole template
    reads, strand, ts, te = base_reads()
    bad = [other.read(1).copy()]
    chim = z.read(full[0]).copy(); o2 = other.read(2); chim[len(chim) // 2:] = o2[:len(chim) - len(chim) // 2]
    bad.append(chim)
    blk = z.read(full[1]).copy(); blk[600:750] = rng.integers(0, 12, 150).astype(np.uint8)
    bad.append(blk)
    bad.append(z.read(full[2])[:len(z.read(full[2])) // 2].copy())
    bstr = [0, int(z.strand[full[0]]), int(z.strand[full[1]]), int(z.strand[full[2]])]
    reads += bad; strand += bstr; ts += [0] * 4; te += [len(draft)] * 4
    res, o = _polish_case(ctx, draft, z.snr, reads, strand, ts, te)
    assert np.array_equal(res["read_status"], o["read_status"]), (res["read_status"], o["read_status"])
    assert set(int(s) for s in o["read_status"][-4:]) - {0}, "the adversarial reads were all accepted"
    assert res["status"][0] in (16, 14)
    n = res["seq_off"][1]
    assert np.array_equal(res["seq"][:n], o["consensus"])
    assert np.max(np.abs(res["qv"][:n].astype(int) - o["qv"].astype(int))) <= 1
    both = ~np.isnan(o["read_ll"])
    assert np.array_equal(~np.isnan(res["read_ll"]), both)
    assert np.max(np.abs(res["read_ll"][both] - o["read_ll"][both])) < 1e-4
    # (b) more than half of the mapped reads unusable -> TOO_MANY_UNUSABLE (status 11), no consensus
    reads, strand, ts, te = base_reads()
    reads = reads[:3]; strand = strand[:3]; ts = ts[:3]; te = te[:3]
    for k in range(1, 6):
        reads.append(other.read(k).copy()); strand.append(0); ts.append(0); te.append(len(draft))
    res, o = _polish_case(ctx, draft, z.snr, reads, strand, ts, te)
    assert np.array_equal(res["read_status"], o["read_status"])
    assert o["n_active"] < 0.5 * len(reads)
    assert res["status"][0] == 11 and res["seq_off"][1] == 0


def test_zscore_filter_matches_oracle(ctx):
    """POOR_ZSCORE: a read whose LL against the draft is far below expectation is dropped when it is added; a lenient
    threshold keeps it.  Read statuses agree with the oracle under both settings."""
    z, full = _good_zmw(1200)
    rng = np.random.default_rng(9)
    draft, mp = sim.corrupt(z.tpl, 0.01, seed=4)
    reads = [z.read(k).copy() for k in full]
    strand = [int(z.strand[k]) for k in full]
    noisy = reads[0].copy()                      # every 4th base call replaced: alive, but improbable
    idx = np.arange(2, len(noisy) - 2, 4)
    noisy[idx] = rng.integers(0, 12, len(idx)).astype(np.uint8)
    reads.append(noisy); strand.append(strand[0])
    ts = [0] * len(reads); te = [len(draft)] * len(reads)
    res, o = _polish_case(ctx, draft, z.snr, reads, strand, ts, te)
    assert np.array_equal(res["read_status"], o["read_status"])
    assert res["read_status"][-1] == 5 and np.isnan(res["read_ll"][-1])          # CCS_READ_POOR_ZSCORE
    assert np.all(res["read_status"][:-1] == 0)
    n = res["seq_off"][1]
    assert np.array_equal(res["seq"][:n], o["consensus"])


def test_iteration_cap_and_junk_draft_match_oracle(ctx):
    """NON_CONVERGENT by the iteration cap; a junk draft loses every read (TOO_MANY_UNUSABLE).  (The template growth
    cap max(512, J/8) is unreachable with the z-score filter on: reads that would drive such growth are dropped when
    they are added.)"""
    z, full = _good_zmw(900)
    draft, mp = sim.corrupt(z.tpl, 0.05, seed=2)
    reads = [z.read(k).copy() for k in full]
    strand = [int(z.strand[k]) for k in full]
    res, o = _polish_case(ctx, draft, z.snr, reads, strand, [0] * len(reads), [len(draft)] * len(reads), max_iterations=2)
    assert not o["converged"] and res["status"][0] == 13 and res["iterations"][0] == o["iterations"] == 2
    n = res["seq_off"][1]
    assert np.array_equal(res["seq"][:n], o["consensus"])
    # junk: a 120-base draft (a prefix of the truth) under reads of the whole 900-base molecule -- every read dies or
    # is dropped on both sides, the ZMW fails with TOO_MANY_UNUSABLE and nothing is emitted
    short = z.tpl[:120].copy()
    res, o = _polish_case(ctx, short, z.snr, reads, strand, [0] * len(reads), [len(short)] * len(reads))
    assert np.array_equal(res["read_status"], o["read_status"])
    assert o["n_active"] == 0 and res["status"][0] == 11 and res["seq_off"][1] == 0


def test_poa_vertex_with_more_than_8_predecessors(ctx):
    """Eleven reads that each delete a different number of bases in front of the same position give that vertex eleven
    candidate predecessors; a vertex keeps the first 8 (spec), on the GPU exactly as in the oracle."""
    rng = np.random.default_rng(21)
    T = rng.integers(0, 4, 700).astype(np.uint8)
    p = 400
    reads = [T.copy()]
    for k in range(1, 12):
        reads.append(np.concatenate([T[:p - k], T[p:]]))
    codes = [(4 * rng.integers(0, 3, len(r)) + r).astype(np.uint8) for r in reads]
    z = sim.Zmw()
    z.hole = 1; z.snr = np.array([9, 16, 8.5, 13], np.float32)
    z.read_off = np.zeros(len(codes) + 1, np.int64); z.read_off[1:] = np.cumsum([len(c) for c in codes])
    z.codes = np.concatenate(codes); z.cx = np.full(len(codes), 3, np.uint8)
    z.strand = np.zeros(len(codes), np.uint8); z.tstart = np.zeros(len(codes), np.int32); z.tend = z.tstart
    batch = api.Batch([z])
    dc = ctx.default_draft_cfg(); dc.max_poa_reads = 12
    d = ctx.draft(batch, dc)
    o = O.draft_zmw(z.snr, codes, z.cx, max_poa_reads=12)
    assert d["status"][0] == o["status"]
    assert np.array_equal(d["tpl"][d["tpl_off"][0]:d["tpl_off"][1]], o["draft"])
    for k in range(len(codes)):
        mapped, strand, ts, te, rs, re = o["maps"][k]
        if mapped:
            assert (d["tstart"][k], d["tend"][k], d["rstart"][k], d["rend"][k]) == (ts, te, rs, re)


def test_draft_cascade_matches_oracle(ctx):
    """Draft cascade (docs/faq/accuracy-vs-passes.md:41-46): when the first full-length read is junk, generator 0 seeds the
    POA with it, nothing threads, too few subreads map back to that "draft"; generator 1 (seed = the full-length read closest
    to the median length, more reads) recovers the ZMW.  A ZMW made of unrelated reads only stays failed after both.
    Statuses, drafts and mappings agree with the oracle bit for bit."""
    cfg = sim.get_config(1, insert_mean=1200)
    zs = []
    for i in range(4):
        z = sim.simulate_zmw(MODEL, cfg, 40 + i)
        other = sim.simulate_zmw(MODEL, cfg, 900 + i)
        reads = [z.read(k).copy() for k in range(z.n_reads)]
        first_full = int(np.flatnonzero(z.cx == 3)[0])
        junk = other.read(int(np.flatnonzero(other.cx == 3)[0])).copy()
        reads[first_full] = junk                      # a full-length pass of a different molecule in the seed's place
        if i == 3:                                    # hopeless: every full-length pass comes from a different molecule
            for k in np.flatnonzero(z.cx == 3):
                o2 = sim.simulate_zmw(MODEL, cfg, 2000 + 17 * int(k))
                reads[int(k)] = o2.read(int(np.flatnonzero(o2.cx == 3)[0])).copy()
        zz = sim.Zmw()
        zz.hole = z.hole; zz.snr = z.snr; zz.cx = z.cx
        zz.read_off = np.zeros(len(reads) + 1, np.int64); zz.read_off[1:] = np.cumsum([len(r) for r in reads])
        zz.codes = np.concatenate(reads); zz.strand = z.strand; zz.tstart = z.tstart; zz.tend = z.tend
        zs.append(zz)
    batch = api.Batch(zs)
    d = ctx.draft(batch)
    res = ctx.ccs(batch)
    n_ok = 0
    for zi, z in enumerate(zs):
        reads = [z.read(k) for k in range(z.n_reads)]
        o = O.draft_zmw(z.snr, reads, z.cx)
        assert d["status"][zi] == o["status"], (zi, d["status"][zi], o["status"])
        assert np.array_equal(d["tpl"][d["tpl_off"][zi]:d["tpl_off"][zi + 1]], o["draft"]), zi
        r0 = batch.zmw_read_off[zi]
        if o["status"] == 16:
            n_ok += 1
            for k in range(z.n_reads):
                mapped, strand, ts, te, rs, re = o["maps"][k]
                if mapped:
                    assert (d["strand"][r0 + k], d["tstart"][r0 + k], d["tend"][r0 + k], d["rstart"][r0 + k],
                            d["rend"][r0 + k]) == (strand, ts, te, rs, re), (zi, k)
                else:
                    assert d["tend"][r0 + k] == 0
        oc = O.ccs_zmw(MODEL, z.snr, reads, z.cx)
        assert res["status"][zi] == oc["status"], (zi, res["status"][zi], oc["status"])
        if oc["status"] in (16, 14, 13):
            s0, s1 = res["seq_off"][zi], res["seq_off"][zi + 1]
            assert np.array_equal(res["seq"][s0:s1], oc["seq"])
    assert n_ok == 3 and d["status"][3] in (7, 8)     # the cascade rescued the three; the hopeless one stays failed


@pytest.mark.parametrize("wsize,wover", [(512, 64), (256, 128), (0, 64)])
def test_windowing_parameters_match_oracle(ctx, wsize, wover):
    """Windowing (docs/how-does-ccs-work.md:57-61,108-110) with non-default window sizes, and switched off: the same cuts,
    read slices, per-window polish and stitching as the oracle -- identical status, consensus, iteration and mutation
    counts and read statuses, QV within 1, summed per-read LL within 1e-4 -- including a ZMW with a partial pass that ends
    inside a window and one that is too short to be split."""
    cfg = sim.get_config(3, insert_mean=2600)
    zs = [sim.simulate_zmw(MODEL, cfg, 500 + i) for i in range(5)]
    zs.append(sim.simulate_zmw(MODEL, sim.get_config(2, insert_mean=700), 77))
    batch = api.Batch(zs)
    pc = ctx.default_polish_cfg()
    pc.window_size = wsize; pc.window_overlap = wover
    res = ctx.ccs(batch, None, pc)
    n_multi = 0
    for zi, z in enumerate(zs):
        reads = [z.read(k) for k in range(z.n_reads)]
        o = O.ccs_zmw(MODEL, z.snr, reads, z.cx, window_size=wsize, window_overlap=wover)
        assert res["status"][zi] == o["status"], (zi, res["status"][zi], o["status"])
        s0, s1 = res["seq_off"][zi], res["seq_off"][zi + 1]
        if o["status"] not in (16, 14, 13):
            assert s1 == s0
            continue
        assert np.array_equal(res["seq"][s0:s1], o["seq"]), zi
        assert np.max(np.abs(res["qv"][s0:s1].astype(int) - o["qv"].astype(int))) <= 1, zi
        assert res["iterations"][zi] == o["iterations"] and res["n_applied"][zi] == o["n_applied"], zi
        assert res["n_tested"][zi] == o["n_tested"], zi
        r0, r1 = batch.zmw_read_off[zi], batch.zmw_read_off[zi + 1]
        assert np.array_equal(res["read_status"][r0:r1], o["read_status"]), zi
        both = ~np.isnan(o["read_ll"])
        assert np.array_equal(~np.isnan(res["read_ll"][r0:r1]), both), zi
        assert np.max(np.abs(res["read_ll"][r0:r1][both] - o["read_ll"][both])) < 1e-4, zi
        if wsize and len(o["seq"]) >= 2 * wsize:
            n_multi += 1
    assert (n_multi >= 4) == (wsize > 0)
