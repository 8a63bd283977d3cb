"""GPU parity tests of the Draft Stage and of the whole per-ZMW path against the CPU oracle.

The draft stage is integer work (max-plus DP, graph bookkeeping): the bar is bit-exact -- identical
draft strings, identical strands / spans / read clips, identical statuses.
"""
import numpy as np
import pytest

import oracle_lib as O
from ccs_b200 import sim, api

pytestmark = pytest.mark.gpu
MODEL = sim.synthetic_model()


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(MODEL)
    yield c
    c.close()


def _zmws(cfg_id, n, **kw):
    cfg = sim.get_config(cfg_id, **kw)
    return [sim.simulate_zmw(MODEL, cfg, i) for i in range(n)]


def _check_draft(ctx, zs):
    batch = api.Batch(zs)
    d = ctx.draft(batch)
    for zi, z in enumerate(zs):
        reads = [z.read(k) for k in range(z.n_reads)]
        o = O.draft_zmw(z.snr, reads, z.cx)
        assert d["status"][zi] == o["status"], (zi, d["status"][zi], o["status"])
        got = d["tpl"][d["tpl_off"][zi]:d["tpl_off"][zi + 1]]
        assert np.array_equal(got, o["draft"]), (zi, len(got), len(o["draft"]))
        r0 = batch.zmw_read_off[zi]
        for k in range(z.n_reads):
            mapped, strand, ts, te, rs, re = o["maps"][k]
            if o["status"] != 16 and o["status"] != 7:
                continue
            if mapped:
                assert (d["strand"][r0 + k], d["tstart"][r0 + k], d["tend"][r0 + k], d["rstart"][r0 + k],
                        d["rend"][r0 + k]) == (strand, ts, te, rs, re), (zi, k)
            else:
                assert d["tend"][r0 + k] == 0
    return d


def test_draft_matches_oracle_small(ctx):
    _check_draft(ctx, _zmws(1, 6, insert_mean=800))


def test_draft_matches_oracle_mixed_statuses(ctx):
    # config 2 mixes low-SNR and too-few-pass ZMWs into the batch
    zs = _zmws(2, 40, insert_mean=1500, insert_sd=100, frac_low_snr=0.15, frac_few_passes=0.15)
    d = _check_draft(ctx, zs)
    st = set(int(s) for s in d["status"])
    assert 0 in st and 2 in st and 16 in st      # POOR_SNR, TOO_FEW_PASSES, passed


def test_draft_matches_oracle_10kb(ctx):
    _check_draft(ctx, _zmws(1, 2, insert_mean=10000))


def test_draft_edge_cases(ctx):
    # no reads / one read / tiny reads never crash and agree on the status
    z0 = _zmws(1, 1, insert_mean=300)[0]
    zs = []
    for keep in (0, 1, 2, 3):
        z = sim.Zmw()
        z.hole = keep; z.snr = z0.snr
        sel = [1 + k for k in range(keep)]
        z.read_off = np.zeros(keep + 1, np.int64)
        parts = []
        for j, k in enumerate(sel):
            parts.append(z0.read(k)); z.read_off[j + 1] = z.read_off[j] + len(parts[-1])
        z.codes = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
        z.cx = z0.cx[sel]; z.strand = z0.strand[sel]; z.tstart = z0.tstart[sel]; z.tend = z0.tend[sel]
        zs.append(z)
    _check_draft(ctx, zs)


@pytest.mark.parametrize("insert,n,cfg_id", [(700, 8, 2), (3000, 2, 1)])
def test_ccs_pipeline_matches_oracle(ctx, insert, n, cfg_id):
    zs = _zmws(cfg_id, n, insert_mean=insert, insert_sd=0)
    batch = api.Batch(zs)
    res = ctx.ccs(batch)
    for zi, z in enumerate(zs):
        reads = [z.read(k) for k in range(z.n_reads)]
        o = O.ccs_zmw(MODEL, z.snr, reads, z.cx)
        assert res["status"][zi] == o["status"], (zi, res["status"][zi], o["status"])
        if o["status"] in (16, 14, 13):   # SUCCESS / POOR_QUALITY / NON_CONVERGENT carry a consensus
            s0, s1 = res["seq_off"][zi], res["seq_off"][zi + 1]
            assert np.array_equal(res["seq"][s0:s1], o["seq"])
            assert np.max(np.abs(res["qv"][s0:s1].astype(int) - o["qv"].astype(int))) <= 1
            assert res["n_passes"][zi] == o["np"]
            assert abs(res["rq"][zi] - o["rq"]) < 2e-3
            r0, r1 = batch.zmw_read_off[zi], batch.zmw_read_off[zi + 1]
            both = ~np.isnan(o["read_ll"])
            assert np.array_equal(~np.isnan(res["read_ll"][r0:r1]), both)
            assert np.max(np.abs(res["read_ll"][r0:r1][both] - o["read_ll"][both])) < 1e-4
