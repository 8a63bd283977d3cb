"""The `ccs` command-line surface (subreads.bam -> hifi_reads.bam): BAM container, PacBio tags,
fatal / non-fatal error behaviour.  CPU part here; the end-to-end run is in the gpu-marked test."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bam_util import read_bam
from ccs_b200 import sim, simlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CCS = os.path.join(ROOT, "ccs_b200", "bin", "ccs")
MODEL = sim.synthetic_model()


def write_subreads(path, cfg, first, n, chem=True):
    rc = simlib().ccs_sim_write_subreads_bam(path.encode(), b"m64000_000000_000000", MODEL.ctypes.data_as(C.c_void_p),
                                          C.byref(cfg), C.c_int64(first), C.c_int32(n), C.c_int32(int(chem)))
    assert rc == 0


def test_subreads_bam_container_and_tags(tmp_path):
    cfg = sim.get_config(1, insert_mean=300)
    p = str(tmp_path / "m.subreads.bam")
    write_subreads(p, cfg, 0, 3)
    text, recs = read_bam(p)
    assert "READTYPE=SUBREAD" in text and "BINDINGKIT=" in text and "PU:m64000_000000_000000" in text
    z = sim.simulate_zmw(MODEL, cfg, 0)
    first = [r for r in recs if r["tags"]["zm"] == 1]
    assert len(first) == z.n_reads
    for k, r in enumerate(first):
        codes = z.read(k)
        assert r["flag"] == 4 and r["ref"] == -1
        assert r["name"] == "m64000_000000_000000/1/%d_%d" % (r["tags"]["qs"], r["tags"]["qe"])
        assert r["seq"] == "".join("ACGT"[c & 3] for c in codes)
        assert r["tags"]["pw"] == [int(c >> 2) + 1 for c in codes]          # pw tag, docs/faq/bam-output.md:20
        assert r["tags"]["cx"] == int(z.cx[k])
        assert np.allclose(r["tags"]["sn"], z.snr)                           # sn tag, docs/faq/bam-output.md:28
        assert r["tags"]["qe"] - r["tags"]["qs"] == len(codes)


def test_cli_aborts_without_chemistry(tmp_path):
    # "abort if chemistry information is missing in the BAM header" (docs/changelog.md:66)
    cfg = sim.get_config(1, insert_mean=200)
    p = str(tmp_path / "nochem.subreads.bam")
    write_subreads(p, cfg, 0, 1, chem=False)
    r = subprocess.run([CCS, p, str(tmp_path / "o.bam")], capture_output=True, text=True)
    assert r.returncode == 1 and "chemistry" in r.stderr


def test_cli_usage_and_bad_input(tmp_path):
    r = subprocess.run([CCS], capture_output=True, text=True)
    assert r.returncode == 2 and "Usage: ccs" in r.stderr
    bad = tmp_path / "x.bam"
    bad.write_bytes(b"not a bam")
    r = subprocess.run([CCS, str(bad), str(tmp_path / "o.bam")], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([CCS, "--chunk", "3/2", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 2


def test_cli_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = sim.get_config(1, insert_mean=200)
    p = str(tmp_path / "m.subreads.bam")
    write_subreads(p, cfg, 0, 1)
    r = subprocess.run([CCS, p, str(tmp_path / "o.bam")], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_cli_end_to_end_matches_api(tmp_path):
    from ccs_b200 import api
    cfg = sim.get_config(2, insert_mean=900, insert_sd=50, frac_low_snr=0.2, frac_few_passes=0.1)
    n = 24
    p = str(tmp_path / "m.subreads.bam")
    out = str(tmp_path / "m.hifi_reads.bam")
    write_subreads(p, cfg, 0, n)
    r = subprocess.run([CCS, p, out, "--batch-size", "10", "--log-level", "INFO"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text, recs = read_bam(out)
    assert "READTYPE=CCS" in text and "@PG\tID:ccs" in text
    zs = [sim.simulate_zmw(MODEL, cfg, i) for i in range(n)]
    ctx = api.Context(MODEL)
    res = ctx.ccs(api.Batch(zs))
    ctx.close()
    want = {}
    for zi in range(n):
        if res["status"][zi] == api.ZMW_SUCCESS:
            s0, s1 = res["seq_off"][zi], res["seq_off"][zi + 1]
            want[zi + 1] = ("".join("ACGT"[b] for b in res["seq"][s0:s1]), list(res["qv"][s0:s1]), res["n_passes"][zi],
                            float(res["rq"][zi]))
    assert 0 < len(want) < n                       # some ZMWs fail filters, some pass
    assert [r["tags"]["zm"] for r in recs] == sorted(want)          # ZMW order preserved
    for r in recs:
        seq, qv, npass, rq = want[r["tags"]["zm"]]
        assert r["name"] == "m64000_000000_000000/%d/ccs" % r["tags"]["zm"]      # docs/faq/mode-by-strand.md:11-14
        assert r["seq"] == seq and r["qual"] == qv
        assert r["tags"]["np"] == npass and abs(r["tags"]["rq"] - rq) < 1e-6 and r["tags"]["rq"] >= 0.99
        assert len(r["tags"]["sn"]) == 4 and "ec" in r["tags"] and r["tags"]["RG"]
    rep = open(str(tmp_path / "m.hifi_reads.ccs_report.txt")).read()
    assert "ZMWs input                    : %d" % n in rep
    assert "ZMWs pass filters             : %d" % len(want) in rep
    assert "Below SNR threshold" in rep and "HiFi Reads                    : %d" % len(want) in rep
    # zmw_metrics.json.gz: one entry per input ZMW, status names in the reference's vocabulary
    import gzip
    import json
    met = json.loads(gzip.open(str(tmp_path / "m.hifi_reads.zmw_metrics.json.gz")).read())["zmws"]
    assert len(met) == n and [m["zmw"] for m in met] == ["m64000_000000_000000/%d" % (i + 1) for i in range(n)]
    assert sum(m["status"] == "SUCCESS" for m in met) == len(want)
    assert {m["status"] for m in met} <= set(api.ZMW_STATUS)
    for m in met:
        assert (m["predicted_accuracy"] >= 0.99) if m["status"] == "SUCCESS" else True
        assert m["insert_size"] > 0 and m["polymerase_length"] >= m["insert_size"]
    # chunked runs concatenate to the full run (docs/faq/parallelize.md:15-28)
    parts = []
    for i in (1, 2):
        o = str(tmp_path / ("c%d.bam" % i))
        rr = subprocess.run([CCS, p, o, "--chunk", "%d/2" % i], capture_output=True, text=True)
        assert rr.returncode == 0, rr.stderr
        parts += [(x["name"], x["seq"], x["qual"]) for x in read_bam(o)[1]]
    assert parts == [(x["name"], x["seq"], x["qual"]) for x in recs]


@pytest.mark.gpu
def test_cli_pipeline_by_strand_reports_and_damaged_input(tmp_path):
    """The pipeline (reader -> queue -> stage workers -> ordered writer) gives the same bytes whatever the batch size,
    the number of stage instances and -j; --by-strand emits one read per strand named .../ccs/fwd|rev
    (docs/faq/mode-by-strand.md:10-24); --report-json / --hifi-summary-json are written; a cut input file or an
    unwritable output is an error exit, not a short run."""
    import json
    cfg = sim.get_config(2, insert_mean=800, insert_sd=40, frac_low_snr=0.1, frac_few_passes=0.1)
    n = 30
    p = str(tmp_path / "m.subreads.bam")
    write_subreads(p, cfg, 100, n)
    outs = []
    for k, extra in enumerate((["--batch-size", "7", "--pipeline", "1", "-j", "2"],
                               ["--batch-size", "16", "--pipeline", "3", "--gpus", "1"],
                               ["--batch-size", "256"])):
        o = str(tmp_path / ("p%d.bam" % k))
        r = subprocess.run([CCS, p, o, "--report-json", str(tmp_path / ("p%d.json" % k)),
                            "--hifi-summary-json", str(tmp_path / ("h%d.json" % k))] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append([(x["name"], x["seq"], x["qual"], x["tags"]["np"]) for x in read_bam(o)[1]])
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 0
    rj = json.load(open(str(tmp_path / "p0.json")))
    assert rj["input"] == n and rj["pass_filters"] == len(outs[0]) and rj["hifi"]["reads"] == len(outs[0])
    assert sum(rj["exclusive_failed_counts"].values()) == rj["fail_filters"] == n - len(outs[0])
    hs = json.load(open(str(tmp_path / "h0.json")))
    assert hs["hifi_reads"] == len(outs[0]) and hs["hifi_yield_bp"] == sum(len(x[1]) for x in outs[0])
    rep = open(str(tmp_path / "p0.ccs_report.txt")).read()
    for label in ("<Q20 Reads", ">=Q30 Reads", "Base quality >=Q30 (bp)", "ZMWs with tandem repeats", "ZMWs missing adapters"):
        assert label in rep                       # docs/faq/reports-aux-files.md:16-72
    # --by-strand
    ob = str(tmp_path / "bs.bam")
    cfg16 = sim.get_config(2, insert_mean=800, insert_sd=40, frac_low_snr=0.1, frac_few_passes=0.1, passes_min=16, passes_max=16)
    p16 = str(tmp_path / "m16.subreads.bam")            # 16 passes = 8 per strand: enough for >= Q20 per strand
    write_subreads(p16, cfg16, 300, 12)
    r = subprocess.run([CCS, p16, ob, "--by-strand"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    recs = read_bam(ob)[1]
    names = [x["name"] for x in recs]
    assert names and all(nm.endswith("/ccs/fwd") or nm.endswith("/ccs/rev") for nm in names)
    by_zmw = {}
    for x in recs:
        by_zmw.setdefault(x["tags"]["zm"], {})[x["name"].rsplit("/", 1)[1]] = x
    both = [z for z, d in by_zmw.items() if len(d) == 2]
    assert both                                    # 8 passes per strand >= --min-passes 3
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for z in both:
        f, rv = by_zmw[z]["fwd"]["seq"], by_zmw[z]["rev"]["seq"]
        assert abs(len(f) - len(rv)) <= 0.06 * len(f)      # draft ends may be clipped where local alignments start late
        rc = "".join(comp[c] for c in reversed(rv))
        k = 12                                     # the two strands describe the same molecule
        kf = {f[i:i + k] for i in range(len(f) - k)}
        assert sum(rc[i:i + k] in kf for i in range(len(rc) - k)) > 0.8 * (len(rc) - k)
        assert by_zmw[z]["fwd"]["tags"]["np"] <= 8 and by_zmw[z]["rev"]["tags"]["np"] <= 8
    assert "Single-Strand Reads input" in open(str(tmp_path / "bs.ccs_report.txt")).read()
    # damaged input: exit code 1 and a message, the ZMWs in front of the damage are still written
    data = open(p, "rb").read()
    cut = str(tmp_path / "cut.subreads.bam")
    open(cut, "wb").write(data[:len(data) * 3 // 5])
    r = subprocess.run([CCS, cut, str(tmp_path / "cut.bam")], capture_output=True, text=True)
    assert r.returncode == 1 and ("BGZF" in r.stderr or "truncated" in r.stderr), r.stderr
    r = subprocess.run([CCS, p, "/nonexistent-dir/out.bam"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot write" in r.stderr
