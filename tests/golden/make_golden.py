"""Regenerates tests/golden/*.npz from the CPU oracle on seeded synthetic inputs.

The reference ships no golden vectors for this path (/root/reference is documentation
only), so these fixtures pin THIS REPO's oracle against accidental drift: any change to the spec
(band rule, tie breaks, model numbers) shows up as a diff here and must be deliberate.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
from ccs_b200 import sim  # noqa: E402


def main():
    model = O.synthetic_model()
    # (1) per-read LLs + first-round delta-LLs: 50-base template x 5 reads
    cfg = sim.get_config(1, insert_mean=64, passes_min=5, passes_max=5, partials=0)
    z = sim.simulate_zmw(model, cfg, 11)
    reads = [z.read(k) for k in range(z.n_reads)]
    draft, mp = sim.corrupt(z.tpl, 0.05, seed=5)
    delta, rll = O.score_all(model, z.snr, draft, reads, z.strand.astype(np.int32), mp[z.tstart], mp[z.tend])
    np.savez_compressed(os.path.join(HERE, "arrow_small.npz"), snr=z.snr, tpl=z.tpl, draft=draft,
                        codes=z.codes, read_off=z.read_off, strand=z.strand, tstart=mp[z.tstart], tend=mp[z.tend],
                        delta=delta, read_ll=rll)
    # (2) whole path: 1 kb x 8 passes
    cfg = sim.get_config(1, insert_mean=1000, passes_min=8, passes_max=8)
    z = sim.simulate_zmw(model, cfg, 4)
    reads = [z.read(k) for k in range(z.n_reads)]
    d = O.draft_zmw(z.snr, reads, z.cx)
    r = O.ccs_zmw(model, z.snr, reads, z.cx)
    np.savez_compressed(os.path.join(HERE, "ccs_1kb.npz"), snr=z.snr, truth=z.tpl, codes=z.codes, read_off=z.read_off,
                        cx=z.cx, draft=d["draft"], maps=d["maps"], draft_status=d["status"], seq=r["seq"], qv=r["qv"],
                        rq=r["rq"], status=r["status"], np_=r["np"], iterations=r["iterations"],
                        n_applied=r["n_applied"], read_ll=r["read_ll"])
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
