"""GPU parity tests: CUDA Arrow path (through the C ABI) vs the CPU oracle on the same inputs.

Tolerances are BASELINE.json's: consensus identical, QV within +-1, per-read LL within 1e-4.
"""
import numpy as np
import pytest

import oracle_lib as O
from ccs_b200 import sim, api

pytestmark = pytest.mark.gpu

MODEL = sim.synthetic_model()
LL_TOL = 1e-4   # north-star tolerance on per-subread Arrow log-likelihood


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(MODEL)
    yield c
    c.close()


def _pairs(n_zmw, insert, cfg_id=2, first=0, **kw):
    cfg = sim.get_config(cfg_id, insert_mean=insert, frac_low_snr=0.0, frac_few_passes=0.0, **kw)
    out = []
    for zi in range(first, first + n_zmw):
        z = sim.simulate_zmw(MODEL, cfg, zi)
        for k in range(z.n_reads):
            t = z.tpl[z.tstart[k]:z.tend[k]]
            t = t if z.strand[k] == 0 else (3 - t[::-1])
            out.append((z.snr.copy(), np.ascontiguousarray(t), z.read(k).copy()))
    return out


def test_fill_cells_match_oracle(ctx):
    pairs = _pairs(1, 700, insert_sd=0)[:5]
    for k in range(len(pairs)):
        g = ctx.fill_alpha_beta(pairs, dump_pair=k)
        snr, t, r = pairs[k]
        o = O.fill(MODEL, snr, t, r, W=32, dump=True)
        assert g["status"][k] == 0 and o["status"] == 0
        assert np.array_equal(g["start"], o["start"])          # identical band
        assert np.array_equal(g["aexp"], o["aexp"])            # identical power-of-two scaling
        assert np.array_equal(g["bexp"][1:], o["bexp"][1:])
        assert np.allclose(g["alpha"], o["alpha"], rtol=2e-4, atol=1e-30)
        # row 0 of beta is never read by anything: compare rows >= 1 only
        J = len(t)
        rows = o["start"][:, None] + ((np.arange(32)[None, :] - o["start"][:, None]) % 32)
        mask = rows >= 1
        mask[0, :] = False
        assert np.allclose(g["beta"][mask], o["beta"][mask], rtol=2e-4, atol=1e-30)
        assert abs(g["ll_alpha"][k] - o["ll_alpha"]) < LL_TOL
        assert abs(g["ll_beta"][k] - o["ll_beta"]) < LL_TOL


@pytest.mark.parametrize("insert,n_zmw", [(300, 3), (2500, 2), (10000, 1), (25000, 1)])
def test_fill_ll_matches_oracle(ctx, insert, n_zmw):
    pairs = _pairs(n_zmw, insert, snr_sd=1.5)
    g = ctx.fill_alpha_beta(pairs)
    for k, (snr, t, r) in enumerate(pairs):
        o = O.fill(MODEL, snr, t, r, W=32)
        assert g["status"][k] == o["status"] == 0
        assert abs(g["ll_alpha"][k] - o["ll_alpha"]) < LL_TOL, (k, g["ll_alpha"][k], o["ll_alpha"])
        assert abs(g["ll_beta"][k] - o["ll_beta"]) < LL_TOL


def test_fill_edge_cases(ctx):
    rng = np.random.default_rng(5)
    pairs = []
    for J, I in [(2, 2), (2, 5), (3, 2), (5, 3), (40, 31), (33, 64), (64, 33), (1, 4), (4, 1)]:
        pairs.append((np.array([9, 16, 8.5, 13], np.float32), rng.integers(0, 4, J).astype(np.uint8),
                      rng.integers(0, 12, I).astype(np.uint8)))
    g = ctx.fill_alpha_beta(pairs)
    for k, (snr, t, r) in enumerate(pairs):
        o = O.fill(MODEL, snr, t, r, W=32)
        assert g["status"][k] == o["status"], (k, len(t), len(r))
        if o["status"] == 0:
            assert abs(g["ll_alpha"][k] - o["ll_alpha"]) < 1e-5
            assert abs(g["ll_beta"][k] - o["ll_beta"]) < 1e-5


def _zmw_batch(zs, rate=0.03):
    drafts = []
    for z in zs:
        d, mp = sim.corrupt(z.tpl, rate, seed=z.hole + 17)
        drafts.append((d, z.strand, mp[z.tstart], mp[z.tend]))
    return api.Batch(zs, drafts), drafts


def _oracle_inputs(z, draft):
    d, strand, ts, te = draft
    return d, [z.read(k) for k in range(z.n_reads)], np.asarray(strand, np.int32), ts, te


def test_score_all_matches_oracle(ctx):
    cfg = sim.get_config(1, insert_mean=400)
    zs = [sim.simulate_zmw(MODEL, cfg, i) for i in range(3)]
    batch, drafts = _zmw_batch(zs)
    delta, rll, rst = ctx.score_all(batch)
    for zi, z in enumerate(zs):
        d, reads, strand, ts, te = _oracle_inputs(z, drafts[zi])
        want, oll = O.score_all(MODEL, z.snr, d, reads, strand, ts, te)
        J = len(d)
        got = delta[batch.tpl_off[zi]:batch.tpl_off[zi + 1]]
        r0, r1 = batch.zmw_read_off[zi], batch.zmw_read_off[zi + 1]
        assert np.all(rst[r0:r1] == 0)
        assert np.max(np.abs(rll[r0:r1] - oll)) < LL_TOL
        w = want[:J]
        fin = np.isfinite(w)
        assert np.array_equal(fin, np.isfinite(got))
        err = np.abs(got[fin] - w[fin])
        assert err.max() < 2e-4, err.max()           # 12 reads x per-read 1e-5-ish
        # the decisions Polish() takes from these numbers are the same
        assert np.array_equal(got[fin] > 0, w[fin] > 0)


@pytest.mark.parametrize("insert,n_zmw,cfg_id", [(600, 6, 1), (3000, 3, 2)])
def test_polish_matches_oracle(ctx, insert, n_zmw, cfg_id):
    cfg = sim.get_config(cfg_id, insert_mean=insert, frac_low_snr=0.0, frac_few_passes=0.0)
    zs = [sim.simulate_zmw(MODEL, cfg, i) for i in range(n_zmw)]
    batch, drafts = _zmw_batch(zs)
    res = ctx.polish(batch)
    for zi, z in enumerate(zs):
        d, reads, strand, ts, te = _oracle_inputs(z, drafts[zi])
        o = O.polish(MODEL, z.snr, d, reads, strand, ts, te)
        s0, s1 = res["seq_off"][zi], res["seq_off"][zi + 1]
        assert np.array_equal(res["seq"][s0:s1], o["consensus"])              # bit-identical consensus
        assert np.max(np.abs(res["qv"][s0:s1].astype(int) - o["qv"].astype(int))) <= 1   # QV +-1
        r0, r1 = batch.zmw_read_off[zi], batch.zmw_read_off[zi + 1]
        assert np.array_equal(res["read_status"][r0:r1], o["read_status"])
        assert np.nanmax(np.abs(res["read_ll"][r0:r1] - o["read_ll"])) < LL_TOL
        assert res["iterations"][zi] == o["iterations"]
        assert res["n_applied"][zi] == o["n_applied"]
        assert res["n_tested"][zi] == o["n_tested"]
        assert abs(res["rq"][zi] - o["rq"]) < 2e-3
        assert (res["status"][zi] == api.ZMW_SUCCESS) == (o["converged"] and o["rq"] >= 0.99)


def test_polish_many_passes_matches_oracle(ctx):
    """55 passes per ZMW: more reads than one fill group (16) and than the scoring kernel's shared-memory read cache
    (48) -- both overflow paths must give the oracle's answer."""
    cfg = sim.get_config(1, insert_mean=300, passes_min=55, passes_max=55, frac_low_snr=0.0, frac_few_passes=0.0)
    zs = [sim.simulate_zmw(MODEL, cfg, i) for i in range(2)]
    assert min(z.n_reads for z in zs) > 48
    batch, drafts = _zmw_batch(zs)
    res = ctx.polish(batch)
    for zi, z in enumerate(zs):
        d, reads, strand, ts, te = _oracle_inputs(z, drafts[zi])
        o = O.polish(MODEL, z.snr, d, reads, strand, ts, te)
        s0, s1 = res["seq_off"][zi], res["seq_off"][zi + 1]
        assert np.array_equal(res["seq"][s0:s1], o["consensus"])
        assert np.max(np.abs(res["qv"][s0:s1].astype(int) - o["qv"].astype(int))) <= 1
        r0, r1 = batch.zmw_read_off[zi], batch.zmw_read_off[zi + 1]
        assert np.array_equal(res["read_status"][r0:r1], o["read_status"])
        assert np.nanmax(np.abs(res["read_ll"][r0:r1] - o["read_ll"])) < LL_TOL
        assert res["n_applied"][zi] == o["n_applied"]


@pytest.mark.timeout(600)
def test_score_kernel_variants_agree(ctx):
    """Compile-time variants of arrow_score_kernel (occupancy targets; the staged variant that brings each read's band
    window into shared memory with cp.async.bulk + mbarrier) and the unfactored generic kernel give the same delta-LLs:
    the variants bit for bit (same arithmetic, different operand paths), the generic one within 2e-4."""
    import os
    cfg = sim.get_config(2, insert_mean=1500, frac_low_snr=0.0, frac_few_passes=0.0)
    zs = [sim.simulate_zmw(MODEL, cfg, 60 + i) for i in range(4)]
    drafts = []
    for z in zs:
        d, mp = sim.corrupt(z.tpl, 0.02, seed=z.hole + 3)
        drafts.append((d, z.strand, mp[z.tstart], mp[z.tend]))
    batch = api.Batch(zs, drafts)
    base, rll0, _ = ctx.score_all(batch)
    for var in ("1", "2", "3"):
        os.environ["CCS_B200_SCORE_VARIANT"] = var
        try:
            c = api.Context(MODEL)
            d, rll, _ = c.score_all(batch)
            c.close()
        finally:
            os.environ.pop("CCS_B200_SCORE_VARIANT", None)
        fin = np.isfinite(base)
        assert np.array_equal(fin, np.isfinite(d)), var
        assert np.array_equal(base[fin], d[fin]), (var, float(np.max(np.abs(base[fin] - d[fin]))))
        assert np.array_equal(rll0, rll)
