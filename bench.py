#!/usr/bin/env python
"""bench.py -- ZMWs/s of the per-ZMW CCS hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            product arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N ...            reference arm (CPU oracle on host cores)

A step = one pass of the hot path over one batch of synthetic ZMWs (config 2 of BASELINE.json:
1 000 ZMWs, 10 kb insert x 10 passes, Sequel-II-shape reads sampled from the Arrow HMM).
N > 1 (torchrun): ZMW index ranges are sharded across ranks, no collective on the data path;
value = all ranks' ZMWs / max-over-ranks time ("weak": per-GPU work fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--zmws", type=int, default=1000, help="ZMWs per step per GPU")
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config id (2 = metric config)")
    ap.add_argument("--draft-error", type=float, default=0.02)
    ap.add_argument("--cpu-sample", type=int, default=0, help="ZMWs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=4, help="concurrent engine lanes per context (0 = library default)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per rank (0 = 2 x cpu count / world)")
    ap.add_argument("--contexts", type=int, default=2,
                    help="GPU contexts per rank; steps are dealt round-robin to the contexts and run concurrently "
                         "(pipelined batches, as a reader thread feeding two stage instances would)")
    ap.add_argument("--stage", default="ccs", choices=["ccs", "polish"],
                    help="ccs = whole per-ZMW hot path (filter + SparsePoa draft + Arrow polish + QVs); "
                         "polish = Polish Stage only on corrupted-truth drafts")
    return ap.parse_args()


METRIC = {"ccs": "ZMWs/sec through the per-ZMW hot path (filter -> SparsePoa draft -> Arrow polish -> QVs)",
          "polish": "ZMWs/sec (polish stage only: Arrow refinement + QVs of every ZMW)"}
STAGE_DESC = {"ccs": "whole hot path from raw subreads: FilterReads + SparsePoa draft + mapping + Arrow polish + QVs",
              "polish": "Polish Stage only; drafts = truth corrupted at 2 % (Draft Stage not timed)"}
WORKLOADS = {1: "config1: 1 ZMW 10kb x 10 passes", 2: "config2: 1000 ZMWs, 10 kb insert x 10 passes (+2 partial passes)",
             3: "config3: 15 kb insert, 5-20 passes", 5: "config5: 25 kb insert x 4 passes"}


def make_batch(model, cfg, first, n, draft_error, threads):
    from ccs_b200 import sim, api
    a = sim.simulate_batch(model, cfg, first, n, draft_error, threads)
    b = api.Batch.from_arrays(a["zmw_read_off"], a["read_off"], a["codes"], a["snr"], a["cx"], a["hole"], a["draft_off"],
                              a["draft"], a["strand"], a["dstart"], a["dend"])
    return b, a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_sample(model, arrays, idxs, threads, stage):
    """CPU oracle (checker) run of the ZMWs `idxs` of a simulated batch on `threads` host threads."""
    if stage == "polish":
        return oracle_polish_sample(model, arrays, idxs, threads)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor

    def one(z):
        r0, r1 = arrays["zmw_read_off"][z], arrays["zmw_read_off"][z + 1]
        reads = [arrays["codes"][arrays["read_off"][r]:arrays["read_off"][r + 1]] for r in range(r0, r1)]
        o = O.ccs_zmw(model, arrays["snr"][4 * z:4 * z + 4], reads, arrays["cx"][r0:r1])
        o["consensus"] = o["seq"]
        return o

    O.olib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, idxs))
    return time.perf_counter() - t0, res


def oracle_polish_sample(model, arrays, idxs, threads):
    """CPU oracle (checker) polish of the ZMWs `idxs` of a simulated batch on `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor

    def one(z):
        r0, r1 = arrays["zmw_read_off"][z], arrays["zmw_read_off"][z + 1]
        reads = [arrays["codes"][arrays["read_off"][r]:arrays["read_off"][r + 1]] for r in range(r0, r1)]
        d = arrays["draft"][arrays["draft_off"][z]:arrays["draft_off"][z + 1]]
        return O.polish(model, arrays["snr"][4 * z:4 * z + 4], d, reads, arrays["strand"][r0:r1].astype(np.int32),
                        arrays["dstart"][r0:r1], arrays["dend"][r0:r1])

    O.olib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, idxs))
    return time.perf_counter() - t0, res


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path.  /root/reference holds no source
    (docs only), so this is the CPU oracle port on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from ccs_b200 import sim
    model = sim.synthetic_model()
    cfg = sim.get_config(args.config)
    cores = os.cpu_count() or 1
    n = args.cpu_sample or max(cores, 4)
    times = []
    for step in range(args.warmup + args.steps):
        arrays = sim.simulate_batch(model, cfg, 1_000_000 + step * n, n, args.draft_error, cores)
        dt, _ = oracle_sample(model, arrays, list(range(n)), cores, args.stage)
        if step >= args.warmup:
            times.append(dt)
    T = sum(times)
    val = n * len(times) / T
    line = {"impl": "reference", "metric": METRIC[args.stage],
            "value": val, "unit": "ZMW/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * T / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS.get(args.config, str(args.config)), "zmws_per_step": n,
                       "stage": STAGE_DESC[args.stage]},
            "cpu_baseline": {"value": val, "unit": "ZMW/s", "cores": cores, "kind": "port",
                             "sample": f"{n} ZMWs of the same workload per step, {cores} threads"},
            "e2e": {"value": val, "unit": "ZMW/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the real stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    # libraries (NCCL's version banner, ...) print to stdout: keep fd 1 for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- ccs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from ccs_b200 import sim, api
    model = sim.synthetic_model()
    cfg = sim.get_config(args.config)
    # host threads of this rank: twice its share of the cores -- the stage threads spend much of their time blocked on
    # stream synchronisation, and 2x measured +6 % e2e over 1x on a 16-core box (profiles/r1_lanes_matrix.txt)
    threads = args.host_threads if args.host_threads > 0 else max(1, 2 * (os.cpu_count() or 8) // max(world, 1))
    # host threads of each stage context of this rank
    os.environ["CCS_B200_THREADS"] = str(max(1, threads // max(1, args.contexts)))
    free_b, _tot = torch.cuda.mem_get_info()
    budget = int(free_b * 0.85 / max(1, args.contexts))          # device bytes each stage context may use
    ctxs = [api.Context(model, device=local, budget_bytes=budget) for _ in range(max(1, args.contexts))]
    ctx = ctxs[0]
    if args.lanes > 0:
        for c in ctxs:
            c.set_lanes(args.lanes)
    pcfg = ctx.default_polish_cfg()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # distinct synthetic ZMW index range per rank and per step (working set >> L2: ~30 GB of DP bands)
    def step_batch(step):
        first = (step * world + rank) * args.zmws
        return make_batch(model, cfg, first, args.zmws, args.draft_error, threads)

    dcfg = ctx.default_draft_cfg()

    def run_step(b, c=None):
        c = c or ctx
        return c.ccs(b, dcfg, pcfg) if args.stage == "ccs" else c.polish(b, pcfg)

    for w in range(args.warmup):
        b, _ = step_batch(w)
        for c in ctxs:
            run_step(b, c)
    batches = [step_batch(args.warmup + k) for k in range(args.steps)]
    for c in ctxs:
        c.stats(reset=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    results = [None] * len(batches)
    step_s = [0.0] * len(batches)

    next_step = [0]
    qlock = threading.Lock()

    def worker(ci):   # each context takes the next unprocessed step as soon as it is free
        while True:
            with qlock:
                k = next_step[0]
                next_step[0] += 1
            if k >= len(batches):
                return
            ts = time.perf_counter()
            results[k] = run_step(batches[k][0], ctxs[ci])
            step_s[k] = time.perf_counter() - ts

    if len(ctxs) == 1:
        worker(0)
    else:
        ths = [threading.Thread(target=worker, args=(ci,)) for ci in range(len(ctxs))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    barrier()
    sts = [c.stats() for c in ctxs]
    st = {k: sum(x[k] for x in sts) for k in sts[0]}
    lanes = args.lanes if args.lanes > 0 else int(os.environ.get("CCS_B200_LANES", "4"))
    t_e2e = wall if len(ctxs) > 1 else st["ms_e2e"] / 1e3   # overlapping contexts: wall clock of the K steps
    # `value`: same run with the batch upload taken out (inputs resident): the initial H2D of the packed
    # read codes / templates is the only input traffic; its CUDA-event span (per lane, lanes overlap) is
    # subtracted from the wall time of the calls
    t_res = t_e2e - st["ms_h2d"] / 1e3 / max(lanes * len(ctxs), 1)
    # per-kernel timing pass: one more step on a single lane (kernels strictly serial, no overlap), used
    # only for the roofline object and the kernel_ms breakdown
    ctx.set_lanes(1)
    ctx.stats(reset=True)
    run_step(batches[0][0])
    torch.cuda.synchronize()
    st1 = ctx.stats()
    ctx.set_lanes(lanes)
    tt = torch.tensor([t_e2e, t_res, wall], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(args.zmws * args.steps), float(sum(int((r["status"] == 16).sum()) for r in results))],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_e2e, t_res, wall = [float(x) for x in tt.tolist()]
    n_total, n_hifi = [float(x) for x in cnt.tolist()]

    if rank == 0:
        peak, peak_src = peaks()
        # the step's full-population launch (every read of every ZMW): later launches of the same step only
        # refill the few ZMWs still being refined and are latency-, not bandwidth-, limited
        fa = st1["top_fill_alpha_bytes"]
        fa_ms = st1["top_fill_alpha_ms"]
        achieved = fa / (fa_ms * 1e-3) / 1e9 if fa_ms > 0 else 0.0
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of the same launch, `ncu --set full` (profiles/r1_summary.json)
            if args.config == 2 and args.zmws == 1000:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_summary.json")))["arrow_fill_alpha"]["dram_bytes"]
        except Exception:
            traffic = None
        kern_ms = {k: st1[k] for k in ("ms_fill_alpha", "ms_fill_beta", "ms_score", "ms_pick", "ms_qv", "ms_h2d",
                                       "ms_poa_align", "ms_draft", "ms_resident", "ms_e2e")}
        kern_ms["note"] = "one step, single lane (serial kernels); the timed region runs %d overlapping lanes" % lanes
        launches = sum(st[k] for k in st if k.startswith("launches"))
        line = {
            "metric": METRIC[args.stage],
            "value": n_total / t_res, "unit": "ZMW/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS.get(args.config, str(args.config)), "zmws_per_step_per_gpu": args.zmws,
                       "stage": STAGE_DESC[args.stage],
                       "l2": "inputs larger than L2 (tens of GB of DP bands per step, new ZMWs every step)",
                       "parallelism": f"zmw-range-shard x{world}, no collective", "lanes_per_gpu": lanes, "contexts_per_gpu": len(ctxs),
                       "value_def": "e2e wall time of the calls minus the batch-upload span (inputs resident)"},
            "hifi_zmws_per_s": n_hifi / t_res, "hifi_fraction": n_hifi / n_total,
            "e2e": {"value": n_total / t_e2e, "unit": "ZMW/s", "h2d_bytes_per_step": st["h2d_bytes"] / args.steps,
                    "d2h_bytes_per_step": st["d2h_bytes"] / args.steps, "wall_s": wall,
                    "step_s": [round(x, 4) for x in step_s]},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "arrow_fill_alpha_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": fa, "ms_per_launch": fa_ms,
                         "launch": "largest arrow_fill_alpha launch of one step (all reads of the batch), single lane",
                         "all_launches_GBps": st1["bytes_fill_alpha"] / max(st1["ms_fill_alpha"], 1e-9) / 1e6},
            "kernel_ms": kern_ms, "rounds": st["rounds"] / args.steps, "score_items_per_step": st["score_items"] / args.steps,
            "roofline_poa_align": {"kernel": "poa_align_kernel", "bound": "latency", "achieved":
                                   st1["bytes_poa_align"] / max(st1["ms_poa_align"], 1e-9) / 1e6, "unit": "GB/s",
                                   "ms_per_step": st1["ms_poa_align"], "tasks": st1["poa_tasks"]},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n = min(args.zmws, args.cpu_sample or max(6 * cores, 24))   # ~10 s of CPU work on the oracle
            _, arrays = batches[0]
            dt, ores = oracle_sample(model, arrays, list(range(n)), cores, args.stage)
            # the sample doubles as a parity spot check at full size
            r = results[0]
            same = sum(int(np.array_equal(r["seq"][r["seq_off"][z]:r["seq_off"][z + 1]], ores[z]["consensus"]))
                       for z in range(n))
            line["cpu_baseline"] = {"value": n / dt, "unit": "ZMW/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n} ZMWs of step 0, {cores} threads, {dt:.1f} s",
                                    "consensus_identical": f"{same}/{n}"}
        emit(line)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
