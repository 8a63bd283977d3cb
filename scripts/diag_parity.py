"""Diagnose a GPU-vs-oracle consensus mismatch of the scale parity test (config 5, 128 ZMWs): windows on / off on both sides."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle_lib as O
from ccs_b200 import sim, api
from concurrent.futures import ThreadPoolExecutor
MODEL = sim.synthetic_model()
CORES = os.cpu_count() or 8
cfg_id, n = 5, 128
cfg = sim.get_config(cfg_id)
a = sim.simulate_batch(MODEL, cfg, 20_000, n, -1.0, CORES)
b = api.Batch.from_arrays(a["zmw_read_off"], a["read_off"], a["codes"], a["snr"], a["cx"], a["hole"])
ctx = api.Context(MODEL)
res = {}
for w in (1024, 0):
    pc = ctx.default_polish_cfg(); pc.window_size = w
    res[w] = ctx.ccs(b, None, pc)
zs = []
for z in range(n):
    r0, r1 = a["zmw_read_off"][z], a["zmw_read_off"][z + 1]
    zs.append(dict(snr=a["snr"][4 * z:4 * z + 4], cx=a["cx"][r0:r1], reads=[a["codes"][a["read_off"][r]:a["read_off"][r + 1]] for r in range(r0, r1)]))
O.olib()
def orc(w):
    with ThreadPoolExecutor(max_workers=CORES) as ex:
        return list(ex.map(lambda z: O.ccs_zmw(MODEL, z["snr"], z["reads"], z["cx"], window_size=w), zs))
ora = {w: orc(w) for w in (1024, 0)}
for w in (1024, 0):
    bad = []
    for z in range(n):
        o = ora[w][z]; r = res[w]
        s0, s1 = r["seq_off"][z], r["seq_off"][z + 1]
        same = r["status"][z] == o["status"] and (o["status"] not in (16, 14, 13) or np.array_equal(r["seq"][s0:s1], o["seq"]))
        if not same: bad.append(z)
    print("window", w, "mismatching ZMWs:", bad)
    for z in bad:
        o = ora[w][z]; r = res[w]
        s0, s1 = r["seq_off"][z], r["seq_off"][z + 1]
        g = r["seq"][s0:s1]; q = o["seq"]
        k = 0
        while k < min(len(g), len(q)) and g[k] == q[k]: k += 1
        print("  zmw", z, "status", r["status"][z], o["status"], "len", len(g), len(q), "first diff at", k, "gpu", g[max(0,k-8):k+8], "oracle", q[max(0,k-8):k+8],
              "iters", r["iterations"][z], o["iterations"], "applied", r["n_applied"][z], o["n_applied"], "n_reads", len(zs[z]["reads"]))
# whole-template vs windowed on each side
for z in range(n):
    for side, get in (("gpu", lambda w: res[w]["seq"][res[w]["seq_off"][z]:res[w]["seq_off"][z + 1]]), ("oracle", lambda w: ora[w][z]["seq"])):
        if not np.array_equal(get(1024), get(0)):
            x, y = get(1024), get(0)
            k = 0
            while k < min(len(x), len(y)) and x[k] == y[k]: k += 1
            print("  windowed != whole on", side, "zmw", z, "len", len(x), len(y), "first diff", k)
ctx.close()
