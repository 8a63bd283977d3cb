"""CPU-only checks (-m "not gpu"): the C ABI loads and exports every symbol include/ccsgpu.h
declares, compute entry points fail loudly without a device (no CPU fallback), the simulator is
deterministic and well-formed, the oracle reproduces the committed golden fixtures, and the doc-derived
known-answer tables hold."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from ccs_b200 import sim, api, lib, LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def declared_functions(header="ccsgpu.h"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b(ccs(?:gpu)?_[a-z0-9_]+)\s*\(", hdr)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    L = lib()
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ccsgpu.h but not exported by {LIB_PATH}"
    # the synthetic-data generator is a library of its own (include/ccssim.h): the product exports none of it
    from ccs_b200 import simlib, SIM_LIB_PATH
    S = simlib()
    sim_names = declared_functions("ccssim.h")
    assert len(sim_names) >= 9
    for n in sim_names:
        assert hasattr(S, n), f"{n} declared in include/ccssim.h but not exported by {SIM_LIB_PATH}"
        if n.startswith("ccs_sim_"):
            assert not hasattr(L, n), f"{n} (test infrastructure) leaked into the product library"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.CcsGpuError) as e:
        api.Context(sim.synthetic_model())
    assert "-4" in str(e.value) or "no usable CUDA device" in str(e.value)


def test_cfg_defaults():
    L = lib()
    p = api.CPolishCfg(); L.ccs_polish_cfg_default(C.byref(p))
    assert (p.max_iterations, p.separation, p.neighborhood) == (40, 10, 20)
    assert p.min_rq == pytest.approx(0.99)          # rq >= 0.99 <=> HiFi, docs/faq/reads-bam.md:38
    assert (p.window_size, p.window_overlap) == (1024, 64)   # windowing, docs/how-does-ccs-work.md:57-61
    d = api.CDraftCfg(); L.ccs_draft_cfg_default(C.byref(d))
    assert (d.min_passes, d.top_passes) == (3, 60)  # --top-passes 60, docs/faq/accuracy-vs-passes.md:48-52
    assert d.min_snr == pytest.approx(2.5)


def test_status_enum_order_matches_docs():
    # /root/reference/docs/faq/reports-aux-files.md:143-159
    want = ["POOR_SNR", "NO_SUBREADS", "TOO_FEW_PASSES", "LOW_PASS_SHORTCUT", "HETERODUPLEXES", "COVERAGE_DROPS",
            "INSUFFICIENT_SPANS", "TOO_FEW_PASSES_AFTER_DRAFT_ALIGNMENT", "DRAFT_FAILURE", "TOO_LONG", "TOO_SHORT",
            "TOO_MANY_UNUSABLE", "EMPTY_WINDOW_DURING_POLISHING", "NON_CONVERGENT", "POOR_QUALITY", "EXCEPTION_THROWN",
            "SUCCESS"]
    assert api.ZMW_STATUS == want
    hdr = open(os.path.join(ROOT, "include", "ccsgpu.h")).read()
    got = re.findall(r"CCS_ZMW_([A-Z_]+)", hdr.split("typedef enum ccs_zmw_status")[1].split("}")[0])
    assert got == want


def test_simulator_is_deterministic_and_shaped():
    m = sim.synthetic_model()
    cfg = sim.get_config(2, insert_mean=1200)
    a = sim.simulate_zmw(m, cfg, 7); b = sim.simulate_zmw(m, cfg, 7); c = sim.simulate_zmw(m, cfg, 8)
    assert np.array_equal(a.codes, b.codes) and np.array_equal(a.tpl, b.tpl)
    assert not np.array_equal(a.tpl, c.tpl)
    assert a.codes.max() < 12 and a.tpl.max() < 4
    assert a.n_reads == 12 and list(a.cx) == [2] + [3] * 10 + [1]      # partial, 10 full passes, partial
    assert list(a.strand[:4]) in ([0, 1, 0, 1], [1, 0, 1, 0])          # strands alternate
    # ~10-13 % longer/shorter reads than the insert (indel dominated error model)
    lens = np.diff(a.read_off)[1:-1]
    assert np.all(np.abs(lens / len(a.tpl) - 1.0) < 0.08)
    # batch generator agrees with the single-ZMW call
    arr = sim.simulate_batch(m, cfg, 7, 2, 0.02, 2)
    assert np.array_equal(arr["codes"][:len(a.codes)], a.codes)
    assert arr["zmw_read_off"].tolist() == [0, 12, 24]


def test_oracle_reproduces_golden_arrow_small():
    g = np.load(os.path.join(GOLD, "arrow_small.npz"))
    model = O.synthetic_model()
    reads = [g["codes"][g["read_off"][k]:g["read_off"][k + 1]] for k in range(len(g["read_off"]) - 1)]
    delta, rll = O.score_all(model, g["snr"], g["draft"], reads, g["strand"].astype(np.int32), g["tstart"], g["tend"])
    assert np.allclose(rll, g["read_ll"], rtol=0, atol=1e-9)
    fin = np.isfinite(g["delta"])
    assert np.array_equal(fin, np.isfinite(delta))
    assert np.allclose(delta[fin], g["delta"][fin], rtol=0, atol=1e-9)


def test_oracle_reproduces_golden_ccs_1kb():
    g = np.load(os.path.join(GOLD, "ccs_1kb.npz"))
    model = O.synthetic_model()
    reads = [g["codes"][g["read_off"][k]:g["read_off"][k + 1]] for k in range(len(g["read_off"]) - 1)]
    d = O.draft_zmw(g["snr"], reads, g["cx"])
    assert d["status"] == int(g["draft_status"]) and np.array_equal(d["draft"], g["draft"])
    assert np.array_equal(d["maps"], g["maps"])
    r = O.ccs_zmw(model, g["snr"], reads, g["cx"])
    assert r["status"] == int(g["status"]) and np.array_equal(r["seq"], g["seq"]) and np.array_equal(r["qv"], g["qv"])
    assert r["np"] == int(g["np_"]) and r["iterations"] == int(g["iterations"]) and r["n_applied"] == int(g["n_applied"])
    assert np.allclose(r["read_ll"], g["read_ll"], equal_nan=True, atol=1e-9)
    # the draft is "still below 99 %" while the polished read is HiFi (docs/how-does-ccs-work.md:46-47, :106)
    assert r["rq"] >= 0.99 and r["status"] == 16


def test_qv_and_rq_formulas():
    # rq = mean per-base accuracy; QV range 0..93 (docs/how-does-ccs-work.md:103-106, docs/faq/qv-binning.md:25-31)
    g = np.load(os.path.join(GOLD, "ccs_1kb.npz"))
    qv = g["qv"].astype(float)
    assert qv.min() >= 0 and qv.max() <= 93
    assert float(g["rq"]) == pytest.approx(1.0 - np.mean(10 ** (-qv / 10)), abs=1e-12)


def test_filter_reads_rule():
    # <50 % or >200 % of the median subread length are removed (docs/how-does-ccs-work.md:23-25)
    m = sim.synthetic_model()
    cfg = sim.get_config(1, insert_mean=600)
    z = sim.simulate_zmw(m, cfg, 1)
    reads = [z.read(k) for k in range(z.n_reads)]
    reads[3] = np.concatenate([reads[3]] * 3)       # missed adapters: > 200 % of the median
    reads[0] = reads[0][:200]                       # short partial: < 50 %
    d = O.draft_zmw(z.snr, reads, z.cx)
    assert d["status"] == 16
    assert d["maps"][3][0] == 0 and d["maps"][0][0] == 0
    assert d["maps"][1][0] == 1
    # fewer than --min-passes full-length subreads -> TOO_FEW_PASSES (status 2)
    d2 = O.draft_zmw(z.snr, reads[:3], z.cx[:3])
    assert d2["status"] == 2
    # SNR below --min-snr -> POOR_SNR (status 0)
    assert O.draft_zmw(np.array([2.0, 9, 9, 9], np.float32), reads, z.cx)["status"] == 0


def test_model_json_round_trip(tmp_path):
    """chemistry-bundle style model injection: save -> load reproduces the parameter blob bit for bit"""
    L = lib()
    m = sim.synthetic_model()
    p = str(tmp_path / "model.json").encode()
    assert L.ccs_model_save_json(m.ctypes.data_as(C.c_void_p), p) == 0
    import json
    j = json.load(open(p.decode()))
    assert j["ModelForm"] == "PwSnr" and len(j["TransitionParameters"]) == 16 and len(j["EmissionParameters"]) == 3
    back = np.zeros_like(m)
    assert L.ccs_model_load_json(p, back.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(back, m)
    bad = tmp_path / "bad.json"
    bad.write_text('{"ModelForm": "Marginal"}')
    assert L.ccs_model_load_json(str(bad).encode(), back.ctypes.data_as(C.c_void_p)) == -7     # CCS_ERR_CHEMISTRY


def test_device_poa_graph_ops_match_oracle(tmp_path):
    """Graph routines of the Draft Stage kernels (seed chain, CommitAdd on the id/order/rank layout, FindConsensus,
    hashed k-mer vote: ccs_b200/csrc/cuda/poa_graph_ops.cuh -- the code the CUDA kernels run) instantiated over a host
    execution context and checked against the oracle's independent graph, driven by the oracle's CPU alignments.
    Runs single-threaded, as a team of 8 threads (barrier = __syncthreads), and as a team under ThreadSanitizer
    (races between the phases of a routine would be races between the CTA's threads)."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    srcs = [os.path.join(ROOT, "tests", "host", "poa_device_graph_parity.cpp"), os.path.join(ROOT, "oracle", "poa_oracle.cpp")]
    exe = str(tmp_path / "poa_device_graph_parity")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-pthread", "-o", exe] + srcs)
    for team in ("1", "8"):
        out = subprocess.run([exe, "30", team], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr + out.stdout
        assert out.stdout.startswith("ok:")
    tsan = str(tmp_path / "poa_device_graph_parity_tsan")
    if subprocess.call([cxx, "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-pthread", "-o", tsan] + srcs,
                       stderr=subprocess.DEVNULL) == 0:
        out = subprocess.run([tsan, "4", "6"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "data race" not in out.stderr, out.stderr[-3000:] + out.stdout


def test_draft_host_helpers_match_oracle(tmp_path):
    """FilterReads, the hashed k-mer orientation vote and orient() of ccs_b200/csrc/host/draft_host.h against the
    oracle's restatements (sorted k-mer lists), on ragged read sets, both strands and unrelated reads."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "draft_host_parity")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "draft_host_parity.cpp"),
                           os.path.join(ROOT, "oracle", "pipeline_oracle.cpp"),
                           os.path.join(ROOT, "oracle", "poa_oracle.cpp"),
                           os.path.join(ROOT, "oracle", "arrow_oracle.cpp")])
    out = subprocess.run([exe, "150"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("ok:")


def test_window_host_matches_oracle(tmp_path):
    """Windowing of the Polish Stage (ccs_b200/csrc/host/window_host.h: window layout, read slices from the mapping grid,
    core borders, empty windows) against the oracle's make_windows / window_reads."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "window_host_parity")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "window_host_parity.cpp"),
                           os.path.join(ROOT, "oracle", "pipeline_oracle.cpp"),
                           os.path.join(ROOT, "oracle", "poa_oracle.cpp"),
                           os.path.join(ROOT, "oracle", "arrow_oracle.cpp")])
    out = subprocess.run([exe, "60"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("ok:")


def test_oracle_windowed_polish_equals_whole_template_polish():
    """Windowing is a pure decomposition (docs/how-does-ccs-work.md:57-61,108-110): on a 4 kb molecule the concatenated
    cores of four independently polished windows are the consensus -- and, within 1, the QVs -- of polishing the whole
    template at once."""
    model = O.synthetic_model()
    cfg = sim.get_config(2, insert_mean=4200)
    for seed in (0, 1):
        z = sim.simulate_zmw(model, cfg, seed)
        reads = [z.read(k) for k in range(z.n_reads)]
        whole = O.ccs_zmw(model, z.snr, reads, z.cx, window_size=0)
        win = O.ccs_zmw(model, z.snr, reads, z.cx, window_size=1024, window_overlap=64)
        assert whole["status"] == win["status"] == 16
        assert np.array_equal(whole["seq"], win["seq"])
        assert np.max(np.abs(whole["qv"].astype(int) - win["qv"].astype(int))) <= 1
        assert np.array_equal(whole["read_status"], win["read_status"])


def test_polish_host_helpers_match_oracle(tmp_path):
    """Candidate order + BestMutations (separation), Template::ApplyMutations and the de-duplicated candidate count of
    ccs_b200/csrc/host/polish_host.h against the oracle's best_mutations / apply_mutations / mutation_is_canonical."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "polish_host_parity")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "polish_host_parity.cpp"),
                           os.path.join(ROOT, "oracle", "arrow_oracle.cpp")])
    out = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("ok:")


def test_bam_io_round_trip(tmp_path):
    """PacBio BAM reader / writer (ccs_b200/csrc/host/bam_io.*): block-parallel deflate -> block-parallel inflate
    round trip of a synthetic subreads.bam, thread-count independent bytes on disk, clean stop on a truncated file."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "bam_io_roundtrip")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "bam_io_roundtrip.cpp"),
                           os.path.join(ROOT, "ccs_b200", "csrc", "host", "bam_io.cpp"), "-lz"])
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("ok:")
