"""Quantifies the QV-reuse optimisation: the same batch through ccsgpu_ccs with the full ConsensusQualities pass
(CCS_B200_REUSE_SCORES=0) and with stored delta-LLs reused outside a halo around every edit
(CCS_B200_REUSE_SCORES=1, CCS_B200_QV_HALO=h).  Prints one JSON line per halo."""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    from ccs_b200 import sim, api
    m = sim.synthetic_model()
    cfg = sim.get_config(int(sys.argv[3]))
    a = sim.simulate_batch(m, cfg, 5000, int(sys.argv[4]), -1.0, 8)
    b = api.Batch.from_arrays(a["zmw_read_off"], a["read_off"], a["codes"], a["snr"], a["cx"], a["hole"])
    ctx = api.Context(m)
    r = ctx.ccs(b)
    st = ctx.stats()
    np.savez(sys.argv[2], seq=r["seq"][:r["seq_off"][-1]], qv=r["qv"][:r["seq_off"][-1]], off=r["seq_off"], status=r["status"],
             rq=r["rq"], items=st["score_items"], ms_score=st["ms_score"])
else:
    cfg_id = sys.argv[1] if len(sys.argv) > 1 else "2"
    n = sys.argv[2] if len(sys.argv) > 2 else "200"
    halos = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [20, 32, 48, 64]
    def run(reuse, halo):
        env = dict(os.environ, CCS_B200_REUSE_SCORES=reuse, CCS_B200_QV_HALO=str(halo), CCS_B200_LANES="1")
        f = "/tmp/qv_reuse_%s_%d.npz" % (reuse, halo)
        subprocess.check_call([sys.executable, __file__, "--one", f, cfg_id, n], env=env)
        return np.load(f)
    a = run("0", 0)
    for h in halos:
        b = run("1", h)
        assert np.array_equal(a["seq"], b["seq"]) and np.array_equal(a["status"], b["status"])
        d = np.abs(a["qv"].astype(int) - b["qv"].astype(int))
        print(json.dumps({"config": cfg_id, "halo": h, "positions": int(d.size), "zmws": int(len(a["status"])), "max_abs_dqv": int(d.max()),
                          "n_diff": int((d > 0).sum()), "n_diff_gt1": int((d > 1).sum()),
                          "max_abs_drq": float(np.max(np.abs(a["rq"] - b["rq"]))),
                          "score_items_full": int(a["items"]), "score_items_reuse": int(b["items"]),
                          "ms_score_full": float(a["ms_score"]), "ms_score_reuse": float(b["ms_score"])}), flush=True)
