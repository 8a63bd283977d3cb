"""A/B of kernel variants on one batch: the whole hot path on a single lane (serial kernels), per-kernel CUDA-event
times from the stats, results compared with the default variant.  Usage: kernel_ab.py ENVVAR v0,v1,... [config] [zmws]"""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "--one":
    from ccs_b200 import sim, api
    m = sim.synthetic_model()
    cfg = sim.get_config(int(sys.argv[3]))
    a = sim.simulate_batch(m, cfg, 7000, int(sys.argv[4]), -1.0, 8)
    b = api.Batch.from_arrays(a["zmw_read_off"], a["read_off"], a["codes"], a["snr"], a["cx"], a["hole"])
    ctx = api.Context(m)
    ctx.set_lanes(1)
    ctx.ccs(b)
    ctx.stats(reset=True)
    r = ctx.ccs(b)
    st = ctx.stats()
    np.savez(sys.argv[2], seq=r["seq"][:r["seq_off"][-1]], qv=r["qv"][:r["seq_off"][-1]], status=r["status"],
             st=json.dumps({k: st[k] for k in st if k.startswith("ms_") or k.startswith("top_") or k.startswith("bytes_")}))
else:
    var, vals = sys.argv[1], sys.argv[2].split(",")
    cfg_id = sys.argv[3] if len(sys.argv) > 3 else "2"
    n = sys.argv[4] if len(sys.argv) > 4 else "250"
    base = None
    for v in vals:
        f = "/tmp/kab_%s.npz" % v
        subprocess.check_call([sys.executable, __file__, "--one", f, cfg_id, n], env=dict(os.environ, **{var: v}))
        d = np.load(f)
        st = json.loads(str(d["st"]))
        same = None
        if base is None:
            base = d
        else:
            same = bool(np.array_equal(base["seq"], d["seq"]) and np.array_equal(base["status"], d["status"]) and
                        np.max(np.abs(base["qv"].astype(int) - d["qv"].astype(int))) <= 1)
        print(json.dumps({var: v, "same_as_first": same, **{k: round(st[k], 2) for k in
              ("ms_score", "ms_fill_alpha", "ms_fill_beta", "ms_poa_align", "ms_poa_map", "ms_poa_graph", "ms_resident", "ms_draft",
               "top_score_ms", "top_fill_alpha_ms", "top_fill_beta_ms")}}), flush=True)
