#!/usr/bin/env python
"""bench.py -- ZMWs/s of the per-ZMW CCS hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            product arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N ...            reference arm (CPU oracle on host cores)

A step = one pass of the hot path (FilterReads -> SparsePoa draft -> mapping -> Arrow polish -> QVs) over one batch
of synthetic ZMWs from HOST buffers.  The default workload is the largest single-GPU configuration of BASELINE.json,
config 3 (15 kb insert, mixed 5-20 passes; one batch = 600 ZMWs ~ 1.3e8 read bases); configs 2 (1 000 ZMWs, 10 kb x 10
passes) and 5 (25 kb x 4 passes) are measured in the same run with fewer steps and reported under "other_configs".
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

DEFAULT_ZMWS = {1: 1, 2: 1000, 3: 600, 4: 600, 5: 800}   # ZMWs per step per GPU: ~1.2e8 - 1.3e8 read bases each


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--zmws", type=int, default=0, help="ZMWs per step per GPU (0 = the config's default)")
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config id (3 = largest single-GPU config)")
    ap.add_argument("--other-configs", default="2,5", help="configs also measured (fewer steps) at N=1; '' = none")
    ap.add_argument("--draft-error", type=float, default=0.02)
    ap.add_argument("--deep-launch-zmws", type=int, default=2000,
                    help="N=1 only: after the timed runs, one config-2 batch of this many ZMWs on a single lane with the whole "
                         "device budget -- arrow_fill_alpha in one deep launch (0 = skip)")
    ap.add_argument("--max-poa-reads", type=int, default=0, help="override the Draft Stage's max_poa_reads (0 = library default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="ZMWs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=4, help="concurrent engine lanes per context (0 = library default)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per rank (0 = 2 x cpu count / world)")
    ap.add_argument("--contexts", type=int, default=2,
                    help="GPU contexts per rank; steps are dealt to the contexts and run concurrently "
                         "(pipelined batches, as the ccs reader feeds two stage instances per GPU)")
    ap.add_argument("--stage", default="ccs", choices=["ccs", "polish"],
                    help="ccs = whole per-ZMW hot path (filter + SparsePoa draft + Arrow polish + QVs); "
                         "polish = Polish Stage only on corrupted-truth drafts")
    ap.add_argument("--minutes", type=float, default=0.0,
                    help="time-boxed mode (config 4: the full-SMRT-Cell shape cannot be materialised): keep taking new "
                         "batches of the ZMW range for this many minutes and report the steady-state rate")
    return ap.parse_args()


METRIC = {"ccs": "ZMWs/sec through the per-ZMW hot path (filter -> SparsePoa draft -> Arrow polish -> QVs)",
          "polish": "ZMWs/sec (polish stage only: Arrow refinement + QVs of every ZMW)"}
STAGE_DESC = {"ccs": "whole hot path from raw subreads: FilterReads + SparsePoa draft + mapping + Arrow polish + QVs",
              "polish": "Polish Stage only; drafts = truth corrupted at 2 % (Draft Stage not timed)"}
WORKLOADS = {1: "config1: 1 ZMW 10kb x 10 passes", 2: "config2: 1000 ZMWs, 10 kb insert x 10 passes (+2 partial passes)",
             3: "config3: 15 kb insert, mixed 5-20 passes (+2 partial passes), batches of the 100 000-ZMW range",
             4: "config4: full SMRT Cell shape (config-3 distribution), ZMW range sharded across the GPUs, time-boxed",
             5: "config5: 25 kb insert x 4 passes (+2 partial passes)"}


def config_dict(args, cfg_id, zmws, world, lanes, contexts):
    """the `config` object -- identical for the product and the reference arm (the driver compares them)"""
    return {"workload": WORKLOADS.get(cfg_id, str(cfg_id)), "zmws_per_step_per_gpu": zmws,
            "stage": STAGE_DESC[args.stage],
            "l2": "inputs larger than L2 (tens of GB of DP bands per step, new ZMWs every step)",
            "parallelism": f"zmw-range-shard x{world}, no collective", "lanes_per_gpu": lanes, "contexts_per_gpu": contexts,
            "windowing": "drafts >= 2048 bases polished as 1024-base windows with 64 bases of overlap (library default), "
                         "on both arms",
            "value_def": "e2e wall time of the calls minus the batch-upload span (inputs resident)"}


def shard_first_index(step, rank, world, zmws):
    """first ZMW index of `rank` in `step`: every step deals `zmws` consecutive ZMWs to each rank, ranks in order (ZMW-range
    sharding) -- the ranges of all (step, rank) pairs tile the index space without overlap"""
    return (step * world + rank) * zmws


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
    (docs only), so this is the CPU oracle port on all host cores, a bounded sample of the workload per step.
    Loads the simulator library and the oracle only -- never the product library."""
    if rank != 0:
        return
    from ccs_b200 import sim
    model = sim.synthetic_model()
    cfg = sim.get_config(args.config)
    cores = os.cpu_count() or 1
    zmws = args.zmws or DEFAULT_ZMWS.get(args.config, 1000)
    n = min(zmws, args.cpu_sample or max(cores, 4))
    times = []
    for step in range(args.warmup + args.steps):
        arrays = sim.simulate_batch(model, cfg, 1_000_000 + step * n, n, args.draft_error, cores)
        dt, _ = oracle_sample(model, arrays, list(range(n)), cores, args.stage)
        if step >= args.warmup:
            times.append(dt)
    T = sum(times)
    val = n * len(times) / T
    lanes = args.lanes if args.lanes > 0 else 4
    line = {"impl": "reference", "metric": METRIC[args.stage],
            "value": val, "unit": "ZMW/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * T / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.config, zmws, max(world, 1), lanes, max(1, args.contexts)),
            "cpu_baseline": {"value": val, "unit": "ZMW/s", "cores": cores, "kind": "port",
                             "sample": f"{n} ZMWs of the workload per step (bounded sample of the {zmws}-ZMW batch), "
                                       f"{cores} threads; ms_per_step is the time of the sample"},
            "e2e": {"value": val, "unit": "ZMW/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the real stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


KERNELS = [  # (name, bytes key, ms key, top bytes key, top ms key)
    ("arrow_fill_alpha_kernel", "bytes_fill_alpha", "ms_fill_alpha", "top_fill_alpha_bytes", "top_fill_alpha_ms"),
    ("arrow_fill_beta_kernel", "bytes_fill_beta", "ms_fill_beta", "top_fill_beta_bytes", "top_fill_beta_ms"),
    ("arrow_score_kernel", "bytes_score", "ms_score", "top_score_bytes", "top_score_ms"),
    ("poa_align_kernel (SparsePoa rounds)", "bytes_poa_align", "ms_poa_align", "top_poa_align_bytes", "top_poa_align_ms"),
    ("poa_align_kernel (subread -> draft mapping)", "bytes_poa_map", "ms_poa_map", "top_poa_map_bytes", "top_poa_map_ms"),
]


def rooflines(st, st1, peak, peak_src, traffic):
    """One entry per heavy kernel: algorithmic bytes / CUDA-event time over (a) all launches of the timed region in the
    timed lane/context configuration (concurrent lanes: spans include the other lanes' interference), (b) all launches
    of one lane-sized chunk on a single lane (the timed launch shapes, strictly serial), (c) its largest launch."""
    tot = sum(st[k[2]] for k in KERNELS) + st["ms_pick"] + st["ms_qv"] + st["ms_poa_graph"]
    out = []
    for name, bk, mk, tbk, tmk in KERNELS:
        def gbps(b, ms):
            return b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        e = {"kernel": name, "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
             "share_of_kernel_time": st[mk] / tot if tot > 0 else 0.0,
             "timed_region": {"bytes": st[bk], "ms": st[mk], "achieved": gbps(st[bk], st[mk]),
                              "frac": gbps(st[bk], st[mk]) / peak},
             "single_lane_all_launches": {"bytes": st1[bk], "ms": st1[mk], "achieved": gbps(st1[bk], st1[mk]),
                                          "frac": gbps(st1[bk], st1[mk]) / peak},
             "largest_launch": {"bytes": st1[tbk], "ms": st1[tmk], "achieved": gbps(st1[tbk], st1[tmk]),
                                "frac": gbps(st1[tbk], st1[tmk]) / peak},
             "traffic": (traffic or {}).get(name.split(" ")[0])}
        out.append(e)
    return out


def measure(args, cfg_id, zmws, steps, warmup, ctxs, model, rank, world, local, threads, lanes, minutes=0.0):
    """W warm-up + K timed steps of one config through the stage contexts; returns raw timings and stats."""
    import torch
    import torch.distributed as dist
    from ccs_b200 import sim
    cfg = sim.get_config(cfg_id)
    ctx = ctxs[0]
    pcfg = ctx.default_polish_cfg()
    dcfg = ctx.default_draft_cfg()
    if args.max_poa_reads > 0:
        dcfg.max_poa_reads = args.max_poa_reads

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # distinct synthetic ZMW index range per rank and per step (working set >> L2: tens of GB of DP bands)
    def step_batch(step):
        first = shard_first_index(step, rank, world, zmws)
        return make_batch(model, cfg, first, zmws, args.draft_error, threads)

    def run_step(b, c=None):
        c = c or ctx
        return c.ccs(b, dcfg, pcfg) if args.stage == "ccs" else c.polish(b, pcfg)

    for w in range(warmup):
        b, _ = step_batch(w)
        for c in ctxs:
            run_step(b, c)
    batches = [step_batch(warmup + k) for k in range(steps)]
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
    deadline = t0 + 60.0 * minutes if minutes > 0 else None
    extra_steps = [0]

    def worker(ci):   # each context takes the next unprocessed step as soon as it is free
        while True:
            with qlock:
                k = next_step[0]
                next_step[0] += 1
            if k >= len(batches):
                if deadline is None or time.perf_counter() >= deadline:
                    return
                # time-boxed mode: keep cycling through the prepared batches (new ZMW ranges cannot be simulated as
                # fast as the GPU consumes them; the hot path does not cache anything between calls)
                run_step(batches[k % len(batches)][0], ctxs[ci])
                with qlock:
                    extra_steps[0] += 1
                continue
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
    for k in sts[0]:
        if k.startswith("top_"):
            st[k] = max(x[k] for x in sts)
    n_steps = steps + extra_steps[0]
    t_e2e = wall if len(ctxs) > 1 else st["ms_e2e"] / 1e3   # overlapping contexts: wall clock of the K steps
    # `value`: same run with the batch upload taken out (inputs resident): the H2D of the read codes is the only
    # input traffic; its CUDA-event span (per lane, lanes overlap) is subtracted from the wall time of the calls
    t_res = t_e2e - st["ms_h2d"] / 1e3 / max(lanes * len(ctxs), 1)
    # per-kernel timing pass: one lane-sized chunk (the launch shapes of the timed region) on a single lane, i.e. the
    # same kernels strictly serial, without the other lanes' interference
    sub, _ = make_batch(model, cfg, (steps + warmup + 7) * world * zmws + rank * zmws, max(1, zmws // max(lanes, 1)),
                        args.draft_error, threads)
    ctx.set_lanes(1)
    run_step(sub)                       # buffers of the single lane are sized here, not inside the measured pass
    ctx.stats(reset=True)
    run_step(sub)
    torch.cuda.synchronize()
    st1 = ctx.stats()
    ctx.set_lanes(lanes)
    n_hifi = float(sum(int((r["status"] == 16).sum()) for r in results if r is not None)) * n_steps / max(steps, 1)
    tt = torch.tensor([t_e2e, t_res, wall], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(zmws * n_steps), n_hifi], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_e2e, t_res, wall = [float(x) for x in tt.tolist()]
    n_total, n_hifi = [float(x) for x in cnt.tolist()]
    return dict(t_e2e=t_e2e, t_res=t_res, wall=wall, n_total=n_total, n_hifi=n_hifi, st=st, st1=st1, clocks=clocks,
                step_s=step_s, results=results, batches=batches, n_steps=n_steps)


def launches_of(st):
    return int(st["launches_fill_alpha"] + st["launches_fill_beta"] + st["launches_score"] + st["launches_pick"] +
               st["launches_qv"] + st["launches_draft"] + st["launches_pack"])


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
    # host threads of this rank: twice its share of the cores -- the stage threads spend much of their time blocked on
    # stream synchronisation
    threads = args.host_threads if args.host_threads > 0 else max(1, 2 * (os.cpu_count() or 8) // max(world, 1))
    os.environ["CCS_B200_THREADS"] = str(max(1, threads // max(1, args.contexts)))
    free_b, _tot = torch.cuda.mem_get_info()
    budget = int(free_b * 0.85 / max(1, args.contexts))          # device bytes each stage context may use
    ctxs = [api.Context(model, device=local, budget_bytes=budget) for _ in range(max(1, args.contexts))]
    lanes = args.lanes if args.lanes > 0 else int(os.environ.get("CCS_B200_LANES", "4"))
    for c in ctxs:
        c.set_lanes(lanes)
    zmws = args.zmws or DEFAULT_ZMWS.get(args.config, 1000)
    M = measure(args, args.config, zmws, args.steps, args.warmup, ctxs, model, rank, world, local, threads, lanes,
                minutes=args.minutes)
    others = {}
    if world == 1 and args.minutes <= 0:
        for oc in [int(x) for x in args.other_configs.split(",") if x.strip()]:
            if oc == args.config:
                continue
            oz = DEFAULT_ZMWS.get(oc, 1000)
            R = measure(args, oc, oz, max(2, args.steps // 2), max(1, args.warmup // 2), ctxs, model, rank, world, local,
                        threads, lanes)
            others["config%d" % oc] = {
                "workload": WORKLOADS.get(oc, str(oc)), "zmws_per_step": oz, "steps": R["n_steps"],
                "value": R["n_total"] / R["t_res"], "e2e": R["n_total"] / R["t_e2e"], "unit": "ZMW/s",
                "hifi_fraction": R["n_hifi"] / R["n_total"],
                "fill_alpha_largest_launch_GBps": R["st1"]["top_fill_alpha_bytes"] / max(R["st1"]["top_fill_alpha_ms"], 1e-9) / 1e6,
                "kernel_ms_single_lane": {k: R["st1"][k] for k in ("ms_fill_alpha", "ms_fill_beta", "ms_score", "ms_poa_align",
                                                                  "ms_poa_map", "ms_poa_graph")}}

    if rank == 0:
        st, st1 = M["st"], M["st1"]
        peak, peak_src = peaks()
        traffic, traffic_any = None, {}
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch of the full-population launches (`ncu --set full`)
            pj = os.path.join(ROOT, "profiles", "r2b_summary.json")
            tj = json.load(open(pj if os.path.exists(pj) else os.path.join(ROOT, "profiles", "r2_summary.json")))
            traffic_any = {k: v.get("dram_bytes") for k, v in tj.get("kernels", {}).items()}
            if tj.get("config") == args.config and zmws == 1000:
                traffic = traffic_any
        except Exception:
            traffic = None
        rl = rooflines(st, st1, peak, peak_src, traffic)
        for e in rl:   # the ncu capture is of a config-2, 1000-ZMW, single-lane launch: quoted with its own context
            e["traffic_ncu_config2_1000zmw_launch"] = traffic_any.get(e["kernel"].split(" ")[0])
        dom = max(rl, key=lambda e: e["timed_region"]["ms"])
        n_l = {"arrow_fill_alpha_kernel": "launches_fill_alpha", "arrow_fill_beta_kernel": "launches_fill_beta",
               "arrow_score_kernel": "launches_score"}.get(dom["kernel"])
        kern_ms = {k: st1[k] for k in ("ms_fill_alpha", "ms_fill_beta", "ms_score", "ms_pick", "ms_qv", "ms_h2d",
                                       "ms_poa_align", "ms_poa_map", "ms_poa_graph", "ms_draft", "ms_resident", "ms_e2e")}
        kern_ms["note"] = ("one lane-sized chunk (%d ZMWs) on a single lane (serial kernels); the timed region runs %d such "
                           "chunks per step on %d overlapping lanes x %d contexts" % (max(1, zmws // lanes), lanes, lanes, len(ctxs)))
        line = {
            "metric": METRIC[args.stage],
            "value": M["n_total"] / M["t_res"], "unit": "ZMW/s", "n_gpus": world, "steps": M["n_steps"], "warmup": args.warmup,
            "ms_per_step": 1e3 * M["t_res"] / M["n_steps"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, args.config, zmws, world, lanes, len(ctxs)),
            "hifi_zmws_per_s": M["n_hifi"] / M["t_res"], "hifi_fraction": M["n_hifi"] / M["n_total"],
            "e2e": {"value": M["n_total"] / M["t_e2e"], "unit": "ZMW/s", "h2d_bytes_per_step": st["h2d_bytes"] / M["n_steps"],
                    "d2h_bytes_per_step": st["d2h_bytes"] / M["n_steps"], "wall_s": M["wall"],
                    "step_s": [round(x, 4) for x in M["step_s"]]},
            "gpu_launches": launches_of(st),
            # the dominant kernel of the timed region: algorithmic bytes per launch / average launch duration, both over
            # ALL its launches inside the timed region (CUDA events on the launching streams)
            "roofline": {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["timed_region"]["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": dom["timed_region"]["frac"], "traffic": dom["traffic"],
                         "peak_source": peak_src, "launches": int(st[n_l]) if n_l else None,
                         "bytes_all_launches": dom["timed_region"]["bytes"], "ms_all_launches": dom["timed_region"]["ms"],
                         "largest_launch_frac_single_lane": dom["largest_launch"]["frac"],
                         "note": "timed lane/context configuration: spans of concurrent lanes overlap, so each launch is "
                                 "slowed by the others; see roofline_kernels for the serial figures"},
            # all kernels of the timed region against its wall clock: algorithmic bytes of the fills, the scoring and the
            # aligner launches / wall time of the K steps (the per-kernel spans above overlap; this one does not)
            "roofline_whole_step": (lambda b: {"bytes": b, "wall_s": M["wall"], "achieved": b / M["wall"] / 1e9, "peak": peak,
                                               "unit": "GB/s", "frac": b / M["wall"] / 1e9 / peak})(
                float(st["bytes_fill_alpha"] + st["bytes_fill_beta"] + st["bytes_score"] + st["bytes_poa_align"] + st["bytes_poa_map"])),
            "roofline_kernels": rl,
            "kernel_ms": kern_ms, "rounds": st["rounds"] / M["n_steps"], "score_items_per_step": st["score_items"] / M["n_steps"],
            "other_configs": others,
            "clocks": M["clocks"],
        }
        if args.minutes > 0:
            line["time_boxed"] = {"minutes": args.minutes, "steps_run": M["n_steps"],
                                  "note": "steady-state rate over the time box; prepared batches are cycled"}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n = min(zmws, args.cpu_sample or max(4 * cores, 24))   # ~10-30 s of CPU work on the oracle
            _, arrays = M["batches"][0]
            dt, ores = oracle_sample(model, arrays, list(range(n)), cores, args.stage)
            # the sample doubles as a parity spot check at full size
            r = M["results"][0]
            same = sum(int(np.array_equal(r["seq"][r["seq_off"][z]:r["seq_off"][z + 1]], ores[z]["consensus"]))
                       for z in range(n))
            qv_ok = sum(int(len(ores[z]["qv"]) == r["seq_off"][z + 1] - r["seq_off"][z] and
                            (len(ores[z]["qv"]) == 0 or
                             np.max(np.abs(r["qv"][r["seq_off"][z]:r["seq_off"][z + 1]].astype(int) - ores[z]["qv"].astype(int))) <= 1))
                        for z in range(n))
            line["cpu_baseline"] = {"value": n / dt, "unit": "ZMW/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n} ZMWs of step 0, {cores} threads, {dt:.1f} s",
                                    "consensus_identical": f"{same}/{n}", "qv_within_1": f"{qv_ok}/{n}"}
        if world == 1 and args.minutes <= 0 and args.deep_launch_zmws > 0:
            # arrow_fill_alpha (and beta) in ONE deep launch: the stage contexts of the timed runs are closed, a fresh
            # context gets the whole device budget and a single lane, and a config-2 batch goes through it once
            for c in ctxs:
                c.close()
            ctxs = []
            try:
                big = api.Context(model, device=local, budget_bytes=int(torch.cuda.mem_get_info()[0] * 0.85))
                big.set_lanes(1)
                bb, _ = make_batch(model, sim.get_config(2), 7_000_000, args.deep_launch_zmws, args.draft_error, threads)
                big.stats(reset=True)
                big.ccs(bb, big.default_draft_cfg(), big.default_polish_cfg())
                torch.cuda.synchronize()
                sb = big.stats()
                big.close()

                def _top(bk, mk):
                    gb = sb[bk] / max(sb[mk], 1e-9) / 1e6
                    return {"bytes": sb[bk], "ms": sb[mk], "achieved": gb, "unit": "GB/s", "peak": peak, "frac": gb / peak}
                line["deep_launch"] = {
                    "workload": "config 2, %d ZMWs in one chunk on one lane (largest launch of each fill kernel, CUDA events)" % args.deep_launch_zmws,
                    "arrow_fill_alpha_kernel": _top("top_fill_alpha_bytes", "top_fill_alpha_ms"),
                    "arrow_fill_beta_kernel": _top("top_fill_beta_bytes", "top_fill_beta_ms")}
            except Exception as e:      # the probe must never cost the bench line
                line["deep_launch"] = {"error": str(e)[:200]}
        emit(line)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
