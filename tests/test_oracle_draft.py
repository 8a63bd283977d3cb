"""CPU checks of the Draft Stage spec in the oracle (DESIGN.md "Draft stage"): the block-anchored band of the aligner keeps
what the per-row band kept -- partial passes that begin or end inside the molecule, reads with an insertion burst, both
strands -- and rejects unrelated reads; drafts are close to the truth."""
import numpy as np
import pytest

import oracle_lib as O
from ccs_b200 import sim

MODEL = sim.synthetic_model()


def _noisy(rng, tpl, sub=0.02, ins=0.05, dele=0.04):
    out = []
    for b in tpl:
        u = rng.random()
        if u < dele:
            continue
        if u < dele + ins:
            out.append(rng.integers(0, 4))
        out.append((b + 1 + rng.integers(0, 3)) % 4 if rng.random() < sub else b)
    r = np.array(out, np.uint8)
    return (4 * rng.integers(0, 3, len(r)) + r).astype(np.uint8)        # emission codes: pulse width x base


def _rc(codes):
    return ((codes[::-1] & 12) | (3 - (codes[::-1] & 3))).astype(np.uint8)


@pytest.mark.parametrize("J", [900, 4200])
def test_partial_passes_insertion_bursts_and_strands_map(J):
    rng = np.random.default_rng(J)
    tpl = rng.integers(0, 4, J).astype(np.uint8)
    full = [_noisy(rng, tpl) for _ in range(6)]
    reads = [f if k % 2 == 0 else _rc(f) for k, f in enumerate(full)]           # alternating strands
    cx = [3] * len(reads)
    # a partial first pass: the second ~60 % of the molecule; a partial last pass: the first ~60 %
    tail = _noisy(rng, tpl[int(0.4 * J):]); head = _noisy(rng, tpl[:int(0.6 * J)])
    # a full pass with a 24-base insertion burst in the middle, and an unrelated read of the same length
    burst = _noisy(rng, tpl)
    mid = len(burst) // 2
    burst = np.concatenate([burst[:mid], (4 * rng.integers(0, 3, 24) + rng.integers(0, 4, 24)).astype(np.uint8), burst[mid:]])
    junk = (4 * rng.integers(0, 3, J) + rng.integers(0, 4, J)).astype(np.uint8)
    reads = [tail] + reads + [burst, junk, _rc(head)]
    cx = [2] + cx + [3, 3, 1]
    d = O.draft_zmw(np.array([9, 16, 8.5, 13], np.float32), reads, np.array(cx, np.uint8))
    assert d["status"] == 16
    Jd = len(d["draft"])
    assert abs(Jd - J) <= 0.03 * J                                               # the draft is close to the molecule
    maps = d["maps"]
    n = len(reads)
    mapped = maps[:, 0].astype(bool)
    assert mapped[1:7].all(), "full passes of both strands must map"
    assert list(maps[1:7, 1]) == [0, 1, 0, 1, 0, 1]                              # strands from the k-mer vote
    for k in range(1, 7):
        assert maps[k, 2] <= 0.02 * J and maps[k, 3] >= Jd - 0.02 * J           # end to end on the draft
    # the partial passes map to their part of the draft
    assert mapped[0] and abs(maps[0, 2] - 0.4 * Jd) <= 0.03 * J and maps[0, 3] >= Jd - 0.02 * J
    assert mapped[n - 1] and maps[n - 1, 1] == 1 and maps[n - 1, 2] <= 0.02 * J and abs(maps[n - 1, 3] - 0.6 * Jd) <= 0.03 * J
    # the band follows a 24-base insertion burst; the unrelated read is not placed
    assert mapped[n - 3] and maps[n - 3, 2] <= 0.02 * J and maps[n - 3, 3] >= Jd - 0.02 * J
    assert not mapped[n - 2]


def test_simulated_zmws_give_drafts_near_the_truth():
    """'the draft consensus has a higher accuracy [than ~90 % subreads], but is still below 99 %'
    (docs/how-does-ccs-work.md:45-47): length within 2 % of the molecule, every kept read mapped."""
    cfg = sim.get_config(2, insert_mean=3000)
    n_ok = 0
    for seed in range(40, 46):
        z = sim.simulate_zmw(MODEL, cfg, seed)
        d = O.draft_zmw(z.snr, [z.read(k) for k in range(z.n_reads)], z.cx)
        if d["status"] != 16:
            continue
        n_ok += 1
        assert abs(len(d["draft"]) - len(z.tpl)) <= 0.02 * len(z.tpl)
        assert int(d["maps"][:, 0].sum()) >= z.n_reads - 2
    assert n_ok >= 4
