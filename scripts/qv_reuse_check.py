"""Quantifies the QV-reuse optimisation: same batch through ccsgpu_ccs with CCS_B200_REUSE_SCORES=0/1."""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    from ccs_b200 import sim, api
    m = sim.synthetic_model()
    cfg = sim.get_config(2, insert_mean=int(sys.argv[2]), insert_sd=200)
    a = sim.simulate_batch(m, cfg, 5000, int(sys.argv[3]), -1.0, 8)
    b = api.Batch.from_arrays(a["zmw_read_off"], a["read_off"], a["codes"], a["snr"], a["cx"], a["hole"])
    ctx = api.Context(m)
    r = ctx.ccs(b)
    np.savez(sys.argv[1], seq=r["seq"][:r["seq_off"][-1]], qv=r["qv"][:r["seq_off"][-1]], off=r["seq_off"], status=r["status"])
else:
    outs = []
    for reuse in ("0", "1"):
        env = dict(os.environ, CCS_B200_REUSE_SCORES=reuse)
        f = "/tmp/qv_reuse_%s.npz" % reuse
        subprocess.check_call([sys.executable, __file__, f, "6000", "300"], env=env)
        outs.append(np.load(f))
    a, b = outs
    assert np.array_equal(a["seq"], b["seq"]) and np.array_equal(a["status"], b["status"])
    d = np.abs(a["qv"].astype(int) - b["qv"].astype(int))
    print(json.dumps({"positions": int(d.size), "zmws": int(len(a["status"])), "max_abs_dqv": int(d.max()),
                      "n_diff": int((d > 0).sum()), "n_diff_gt1": int((d > 1).sum())}))
