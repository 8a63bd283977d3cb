"""Turns gpurun_out/r2_full_*.ncu-rep + the launch list into profiles/r2_summary.md / .json (run here, no GPU).
Usage: python scripts/summarize_profiles.py [round-tag, default r2]"""
import collections
import csv
import io
import json
import subprocess
import sys

KERNELS = ["arrow_fill_alpha", "arrow_fill_beta", "arrow_score", "poa_align", "poa_map", "poa_traceback", "poa_commit", "poa_consensus",
           "arrow_pack_rowcodes"]
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tensor.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, data):
        if h in WANT:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            d[h] = x * UNIT.get(u, 1.0) if u in UNIT else x
    return d


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    samples = 0
    for r in data:
        if len(r) < len(hdr):
            continue
        samples += int(r[ix["# Samples"]])
        for s in names:
            tot[s] += int(r[ix[s]])
    return {k[6:]: round(100.0 * v / max(samples, 1), 1) for k, v in tot.most_common(6)}


def main():
    res = {}
    import os
    kernels = [k for k in KERNELS if os.path.exists("gpurun_out/%s_full_%s.ncu-rep" % (TAG, k))]
    for k in kernels:
        rep = "gpurun_out/%s_full_%s.ncu-rep" % (TAG, k)
        d = raw(rep)
        try:
            d["stall_pct"] = stalls(rep)
        except Exception:
            d["stall_pct"] = {}
        d["dram_bytes"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        res[k] = d
    # launch list
    rows = list(csv.reader(open("gpurun_out/%s_launches_bench.csv" % TAG)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].split("::")[-1]
            agg.setdefault(name, []).append(float(d["Metric Value"].replace(",", "")) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    res["launch_list"] = {n: {"launches": len(v), "total_ms": round(sum(v), 2), "share": round(sum(v) / tot, 3),
                              "max_ms": round(max(v), 2)} for n, v in agg.items()}
    out = {"config": 2, "kernels": {k + "_kernel" if not k.startswith("poa_map") else "poa_align_kernel(map)": res[k] for k in kernels},
           "launch_list": res["launch_list"]}
    json.dump(out, open("profiles/%s_summary.json" % TAG, "w"), indent=1)
    with open("profiles/%s_summary.md" % TAG, "w") as f:
        f.write("# Round-2 ncu summary (B200, config 2: 1000 ZMWs, 10 kb x 10 passes, full captures with `bench.py --config 2 --lanes 1 --contexts 1`)\n\n")
        f.write("Source: `gpurun_out/%s_full_*.ncu-rep` (`ncu --set full --clock-control none --import-source on`, first launch of "
                "each kernel = the full-population launch; `poa_map` = the 5th `poa_align_kernel` launch, the subread -> draft mapping) and "
                "`profiles/%s_launches_bench.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none` over the default "
                "`python bench.py --steps 2 --warmup 1 --no-cpu-baseline` = config 3, 2 contexts x 4 lanes; per-launch times "
                "are cold-cache and serialised, compare shares). Commands: `scripts/gpu_profile_%s.sh` (r2b: fill / score captures "
                "by `gpu_profile_r2b.sh`; Draft Stage kernels and the launch list re-captured on the final kernels by "
                "`gpu_profile_r2c.sh`); regenerate with `python scripts/summarize_profiles.py %s`.\n\n" % (TAG, TAG, TAG, TAG))
        f.write("| kernel | duration ms | DRAM read GB | DRAM write GB | DRAM %peak | issue-active % | warp-instr | regs | warps active % | top stalls (% of samples) |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for k in kernels:
            d = res[k]
            f.write("| %s | %.2f | %.3f | %.3f | %.1f | %.1f | %.3g | %d | %.1f | %s |\n" % (
                k, d["gpu__time_duration.sum"] * 1e3, d.get("dram__bytes_read.sum", 0) / 1e9,
                d.get("dram__bytes_write.sum", 0) / 1e9, d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
                d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0), d.get("smsp__inst_executed.sum", 0),
                int(d.get("launch__registers_per_thread", 0)), d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0),
                ", ".join("%s %.0f" % kv for kv in d["stall_pct"].items())))
        f.write("\nTensor pipe instructions: %s (none by design: no dense contraction on this path).\n\n" %
                {k: res[k].get("sm__inst_executed_pipe_tensor.sum", 0) for k in kernels})
        f.write("## Launch list of the bench command (share of summed kernel time)\n\n| kernel | launches | total ms | share | max ms |\n|---|---|---|---|---|\n")
        for n, v in res["launch_list"].items():
            f.write("| %s | %d | %.2f | %.3f | %.2f |\n" % (n, v["launches"], v["total_ms"], v["share"], v["max_ms"]))
    print(open("profiles/%s_summary.md" % TAG).read())


if __name__ == "__main__":
    main()
