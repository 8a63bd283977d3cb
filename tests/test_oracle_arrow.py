"""First-principles pins of the CPU oracle (the reference has no golden vectors for this
path -- SURVEY.md 8c -- so the oracle is pinned by what can be derived independently):

 1. brute-force enumeration of every alignment path (pure Python, no DP) == FillAlpha LL
 2. log alpha(I,J) == log beta(0,0)         (the reference's own ALPHA_BETA_MISMATCH check)
 3. incremental Evaluator::LL(mutation) == refill on the mutated template
 4. the W=32 leading-edge band loses nothing against a 256-row band
 5. strand symmetry of the Integrator
 6. Polish recovers the true template from a corrupted draft; QVs are high where it is right
"""
import math
import numpy as np
import pytest

import oracle_lib as O
from ccs_b200 import sim

MODEL = O.synthetic_model()
SNR = np.array([9.0, 16.0, 8.5, 13.0], np.float32)


def model_fields(buf):
    off = 64
    f = np.frombuffer(buf.tobytes(), dtype=np.float64, offset=off)
    snr_lo, snr_hi = f[0:4], f[4:8]
    trans = f[8:8 + 192].reshape(16, 3, 4)
    em = f[200:200 + 576].reshape(3, 16, 12)
    cw = f[776]
    return snr_lo, snr_hi, trans, em, cw


def test_tables_match_numpy_restatement():
    snr_lo, snr_hi, trans, em, cw = model_fields(MODEL)
    em_match, em_ins, tr, lcw = O.tables(MODEL, SNR)
    assert lcw == pytest.approx(math.log(cw), abs=0)
    for ctx in range(16):
        s = float(np.clip(SNR[ctx & 3], snr_lo[ctx & 3], snr_hi[ctx & 3]))
        x = np.exp(trans[ctx] @ np.array([1.0, s, s * s, s ** 3]))   # branch, stick, deletion
        den = 1.0 + x.sum()
        want = np.array([1 / den, x[2] / den, x[0] / den, x[1] / den], np.float32)
        assert np.allclose(tr[ctx], want, rtol=2e-7, atol=0)
        assert abs(tr[ctx].sum() - 1) < 1e-6
        for code in range(12):
            assert em_match[ctx][code] == np.float32(cw * em[0][ctx][code])
            cognate = (code & 3) == (ctx & 3)
            assert em_ins[ctx][code] == np.float32(cw * em[1 if cognate else 2][ctx][code])
    # each emission pmf sums to 1
    assert np.allclose(em.sum(axis=2), 1.0)
    # the folded factors are single fp32 products of the fp32 tables (DESIGN.md "Arrow model")
    mm, gg = O.folded(MODEL, SNR)
    for ctx in range(16):
        for code in range(12):
            assert mm[ctx][code] == np.float32(em_match[ctx][code]) * np.float32(tr[ctx][0])
            assert mm[20 + ctx][code] == em_match[20 + ctx][code]
            cognate = (code & 3) == (ctx & 3)
            assert gg[ctx][code] == np.float32(em_ins[ctx][code]) * np.float32(tr[ctx][2 if cognate else 3])
    assert np.all(mm[:, 12:] == 0) and np.all(gg[:, 12:] == 0)


def brute_force_ll(tpl, codes):
    """Sum over every path, no dynamic programming (SURVEY.md A.4 written as a recursion over moves)."""
    em_match, em_ins, tr, lcw = O.tables(MODEL, SNR)
    mm, gg = O.folded(MODEL, SNR)
    J, I = len(tpl), len(codes)

    def ctx(j):
        return 4 * tpl[j - 1] + tpl[j]

    def rec(i, j):
        # i read bases emitted, j template bases consumed, 1 <= j <= J-1, 1 <= i <= I-1
        total = 0.0
        if i == I - 1 and j == J - 1:
            total += em_match[20 + ctx(J - 1)][codes[I - 1]]          # pinned last match, no transition
        c = ctx(j)
        if i + 1 <= I - 1:                                            # insertion (branch / stick) in column j
            e = codes[i]
            total += gg[c][e] * rec(i + 1, j)
        if j + 1 <= J - 1:
            total += tr[c][1] * rec(i, j + 1)                         # deletion of t_j
            if i + 1 <= I - 1:
                total += mm[c][codes[i]] * rec(i + 1, j + 1)   # match
        return total

    p = em_match[16 + tpl[0]][codes[0]] * rec(1, 1)                   # pinned first match
    return (math.log(p) if p > 0 else -math.inf) - I * lcw


@pytest.mark.parametrize("seed", range(12))
def test_fill_alpha_equals_path_enumeration(seed):
    rng = np.random.default_rng(seed)
    J = int(rng.integers(2, 7)); I = int(rng.integers(2, 8))
    tpl = rng.integers(0, 4, J).astype(np.uint8)
    codes = rng.integers(0, 12, I).astype(np.uint8)
    want = brute_force_ll(tpl, codes)
    got = O.fill(MODEL, SNR, tpl, codes, W=32)
    if want == -math.inf:
        assert got["status"] == 3
        return
    assert got["status"] == 0
    assert got["ll_alpha"] == pytest.approx(want, abs=1e-12)
    assert got["ll_beta"] == pytest.approx(want, abs=1e-12)
    gf = O.fill(MODEL, SNR, tpl, codes, W=32, precision=1)
    assert gf["ll_alpha"] == pytest.approx(want, abs=1e-5)


def _pairs(n_zmw, insert, cfg_id=2, **kw):
    cfg = sim.get_config(cfg_id, insert_mean=insert, frac_low_snr=0.0, frac_few_passes=0.0, **kw)
    out = []
    for zi in range(n_zmw):
        z = sim.simulate_zmw(MODEL, cfg, zi)
        for k in range(z.n_reads):
            t = z.tpl[z.tstart[k]:z.tend[k]]
            t = t if z.strand[k] == 0 else (3 - t[::-1])
            out.append((z.snr, np.ascontiguousarray(t), z.read(k).copy()))
    return out


def test_alpha_beta_agree_and_band_is_lossless():
    for snr, t, r in _pairs(2, 3000, insert_sd=0)[:16]:
        a = O.fill(MODEL, snr, t, r, W=32)
        w = O.fill(MODEL, snr, t, r, W=256)
        assert a["status"] == 0 and w["status"] == 0
        assert a["ll_alpha"] == pytest.approx(a["ll_beta"], abs=1e-8)
        assert a["ll_alpha"] == pytest.approx(w["ll_alpha"], abs=1e-8)
        assert a["cells"] == 32 * (len(t) - 1)
        f = O.fill(MODEL, snr, t, r, W=32, precision=1)
        assert f["ll_alpha"] == pytest.approx(a["ll_alpha"], abs=1e-4)   # fp32 cells stay inside the north-star bound


def test_band_slides_in_steps_of_four_and_stays_lossless_at_10kb():
    """Quantised slide (DESIGN.md "Band rule"): starts are multiples of 4, move by 0 or 4 per column, and the
    32-row band still loses nothing against a 256-row band on long reads."""
    for snr, t, r in _pairs(1, 10000, insert_sd=0, snr_sd=1.5)[:4]:
        a = O.fill(MODEL, snr, t, r, W=32, dump=True)
        assert a["status"] == 0
        st = a["start"]
        assert np.all(st % 4 == 0)
        d = np.diff(st)
        assert set(np.unique(d)).issubset({0, 4})
        w = O.fill(MODEL, snr, t, r, W=256)
        assert abs(a["ll_alpha"] - w["ll_alpha"]) < 1e-9
        assert abs(a["ll_alpha"] - a["ll_beta"]) < 1e-7


def all_mutations(tpl, positions):
    J = len(tpl)
    muts = []
    for p in positions:
        if p < J:
            muts += [(O.MUT_SUB, p, b) for b in range(4) if b != tpl[p]]
            muts.append((O.MUT_DEL, p, 0))
        muts += [(O.MUT_INS, p, b) for b in range(4)]
    return muts


def test_incremental_mutation_ll_equals_refill():
    rng = np.random.default_rng(7)
    for snr, t, r in _pairs(1, 400, insert_sd=0)[1:5]:
        J = len(t)
        pos = sorted(set([0, 1, 2, 3, J - 3, J - 2, J - 1, J] + list(rng.integers(4, J - 4, 40))))
        muts = all_mutations(t, pos)
        inc, full = O.score(MODEL, snr, t, r, muts, W=32)
        assert np.all(np.isfinite(inc))
        assert np.max(np.abs(inc - full)) < 1e-7


def test_incremental_mutation_tiny_templates_vs_enumeration():
    rng = np.random.default_rng(3)
    for _ in range(6):
        J = int(rng.integers(3, 6)); I = int(rng.integers(3, 7))
        tpl = rng.integers(0, 4, J).astype(np.uint8)
        codes = rng.integers(0, 12, I).astype(np.uint8)
        muts = all_mutations(tpl, range(J + 1))
        inc, _ = O.score(MODEL, SNR, tpl, codes, muts, W=32)
        for (ty, p, b), got in zip(muts, inc):
            mt = list(tpl)
            if ty == O.MUT_SUB:
                mt[p] = b
            elif ty == O.MUT_INS:
                mt.insert(p, b)
            else:
                del mt[p]
            if len(mt) < 2:
                continue
            want = brute_force_ll(np.array(mt, np.uint8), codes)
            if want == -math.inf:
                assert got == -math.inf
            else:
                assert got == pytest.approx(want, abs=1e-10)


def _zmw_inputs(z, draft=None, mp=None):
    reads = [z.read(k) for k in range(z.n_reads)]
    if draft is None:
        return z.tpl, reads, z.strand.astype(np.int32), z.tstart.copy(), z.tend.copy()
    return draft, reads, z.strand.astype(np.int32), mp[z.tstart], mp[z.tend]


def test_strand_symmetry():
    cfg = sim.get_config(1, insert_mean=300)
    z = sim.simulate_zmw(MODEL, cfg, 5)
    tpl, reads, strand, ts, te = _zmw_inputs(z)
    d1, ll1 = O.score_all(MODEL, z.snr, tpl, reads, strand, ts, te)
    # reverse-complement the template and flip every read's strand/span: same molecule
    J = len(tpl)
    rc = (3 - tpl[::-1]).astype(np.uint8)
    snr = z.snr
    d2, ll2 = O.score_all(MODEL, snr, rc, reads, 1 - strand, J - te, J - ts)
    assert np.allclose(ll1, ll2, atol=1e-9)
    # SUB(p,b) <-> SUB(J-1-p, 3-b); DEL(p) <-> DEL(J-1-p); INS(p,b) <-> INS(J-p, 3-b)
    for p in range(2, J - 2):
        for b in range(4):
            assert d1[p, b] == pytest.approx(d2[J - 1 - p, 3 - b], abs=1e-8)
            assert d1[p, 5 + b] == pytest.approx(d2[J - p, 5 + 3 - b], abs=1e-8)
        assert d1[p, 4] == pytest.approx(d2[J - 1 - p, 4], abs=1e-8)


@pytest.mark.parametrize("index", [0, 1, 2])
def test_polish_recovers_truth(index):
    # 20 passes: the maximum-likelihood template is the truth (at 10 passes homopolymer
    # lengths are genuinely ambiguous, cf. docs/faq/accuracy-vs-passes.md:7-13 -- Q30 at 10 passes)
    cfg = sim.get_config(1, insert_mean=1200, passes_min=20, passes_max=20)
    z = sim.simulate_zmw(MODEL, cfg, index)
    draft, mp = sim.corrupt(z.tpl, 0.03, seed=index)
    assert not np.array_equal(draft, z.tpl)
    d, reads, strand, ts, te = _zmw_inputs(z, draft, mp)
    res = O.polish(MODEL, z.snr, d, reads, strand, ts, te)
    assert res["converged"]
    assert res["n_active"] == z.n_reads
    assert np.array_equal(res["consensus"], z.tpl)
    assert res["n_applied"] >= 20
    assert res["rq"] > 0.999
    assert np.median(res["qv"]) >= 30


@pytest.mark.parametrize("index", [0, 1, 2])
def test_polish_never_worse_than_truth_at_10_passes(index):
    cfg = sim.get_config(1, insert_mean=1200)
    z = sim.simulate_zmw(MODEL, cfg, index)
    draft, mp = sim.corrupt(z.tpl, 0.03, seed=index)
    d, reads, strand, ts, te = _zmw_inputs(z, draft, mp)
    res = O.polish(MODEL, z.snr, d, reads, strand, ts, te)
    assert res["converged"]
    _, ll_truth = O.score_all(MODEL, z.snr, z.tpl, reads, strand, z.tstart, z.tend)
    assert np.nansum(res["read_ll"]) >= np.nansum(ll_truth) - 1e-6
    # wherever the consensus departs from the truth the QV says so
    if not np.array_equal(res["consensus"], z.tpl):
        assert res["qv"].min() <= 10


def logspace_forward_ll(snr, tpl, codes):
    """Second, independent pin of the oracle's recursion: the UNBANDED forward algorithm in LOG space (numpy
    logaddexp, no scaling, no band, no folded-table indexing tricks), written from the move semantics of the
    brute-force enumeration above: state (i, j) = i read bases emitted, j template bases consumed."""
    em_match, em_ins, tr, lcw = O.tables(MODEL, snr)
    mm, gg = O.folded(MODEL, snr)
    J, I = len(tpl), len(codes)
    with np.errstate(divide="ignore"):
        lmm, lgg, ldel, lem = np.log(mm), np.log(gg), np.log(tr[:, 1]), np.log(em_match)
    ctx = lambda j: 4 * int(tpl[j - 1]) + int(tpl[j])
    NEG = -np.inf
    prev = np.full(I, NEG)                       # f(., j-1), rows 0..I-1 (row 0 unused)
    cur = np.full(I, NEG)
    cur[1] = lem[16 + int(tpl[0])][codes[0]]     # pinned first match: state (1, 1)
    c1 = ctx(1)
    for i in range(2, I):                        # insertions inside column 1
        cur[i] = cur[i - 1] + lgg[c1][codes[i - 1]]
    for j in range(2, J):
        prev, cur = cur, np.full(I, NEG)
        cp, cj = ctx(j - 1), ctx(j)
        base = np.full(I, NEG)
        base[1:] = prev[1:] + ldel[cp]                                                   # deletion of t_{j-1}
        base[2:] = np.logaddexp(base[2:], prev[1:I - 1] + lmm[cp][codes[1:I - 1]])       # match emits codes[i-1]
        gi = lgg[cj][codes[:I - 1]]                                                      # insertion emits codes[i-1]
        cur[1] = base[1]
        for i in range(2, I):
            cur[i] = np.logaddexp(base[i], cur[i - 1] + gi[i - 1])
    return cur[I - 1] + lem[20 + ctx(J - 1)][codes[I - 1]] - I * lcw


def test_oracle_fill_equals_unbanded_logspace_numpy_forward():
    """300-500 bp reads sampled from the HMM: the oracle's banded, scaled, probability-domain fill (the code the GPU
    kernels are checked against) equals an unbanded log-space forward written independently in numpy."""
    n = 0
    for ins in (300, 420, 500):
        for snr, t, r in _pairs(1, ins, insert_sd=0)[1:5]:
            want = logspace_forward_ll(snr, t, r)
            wide = O.fill(MODEL, snr, t, r, W=256)
            band = O.fill(MODEL, snr, t, r, W=32)
            assert wide["status"] == 0 and band["status"] == 0
            assert wide["ll_alpha"] == pytest.approx(want, abs=1e-9)
            assert wide["ll_beta"] == pytest.approx(want, abs=1e-9)
            assert band["ll_alpha"] == pytest.approx(want, abs=1e-9)      # the 32-row leading-edge band loses nothing
            n += 1
    assert n == 12
