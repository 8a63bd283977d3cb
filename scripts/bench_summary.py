"""Pretty-prints the interesting parts of one bench.py JSON line (stdin or file)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads(txt.strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "rounds", "score_items_per_step")}, "e2e", round(d["e2e"]["value"], 1),
      "h2d/step", int(d["e2e"]["h2d_bytes_per_step"]), d["config"]["workload"][:8], "L", d["config"]["lanes_per_gpu"], "C", d["config"]["contexts_per_gpu"])
r = d["roofline"]; print("roofline:", r["kernel"], round(r["achieved"], 1), round(r["frac"], 3))
for e in d["roofline_kernels"]:
    print(" ", e["kernel"][:40].ljust(40), "share", round(e["share_of_kernel_time"], 3),
          {k: (round(e[k]["ms"], 1), round(e[k]["frac"], 3)) for k in ("timed_region", "single_lane_all_launches", "largest_launch")})
print("kernel_ms:", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["kernel_ms"].items() if k != "note"})
for k, v in d.get("other_configs", {}).items():
    print(" ", k, "value", round(v["value"], 1), "e2e", round(v["e2e"], 1), "fill_alpha top GB/s", round(v["fill_alpha_largest_launch_GBps"], 1))
print("cpu_baseline:", d.get("cpu_baseline")); print("clocks:", d.get("clocks"))
