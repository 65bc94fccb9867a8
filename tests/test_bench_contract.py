"""CPU: bench.py's reference arm (the CPU port of the reference's path, `--impl reference`) prints ONE JSON line with the
contract fields, for both workloads, on a tiny shape; the synthetic batch helpers produce the documented formats."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload", ["encoder", "hotpath", "config3"])
def test_reference_arm_json_contract(workload):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--batch", "4",
           "--n-atoms", "10", "--smiles-len", "6", "--cpu-sample", "2", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                                       # stdout carries the JSON line only
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_molecules_per_sec" and d["unit"] == "molecules/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["per_gpu_batch"] == 4 and d["config"]["seq_len"] == 12 and "b4x10atoms" in d["config"]["workload"]
    assert ("infonce" in d["config"]["workload"]) == (workload != "encoder")
    assert d["scaling"] == ("strong" if workload == "config3" else "weak")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "2 of the 4 molecules" in cb["sample"]
    assert d["config"]["cpu_sample_molecules_per_step"] == 2 and d["config"]["global_batch"] == 4
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_batch_helpers():
    sys.path.insert(0, ROOT)
    import bench
    try:
        bench.set_shape(6, 12, 9)
        tokens, dist, et, g, coord = bench.make_batch(1234)
        assert tokens.shape == (6, 14) and dist.shape == (6, 14, 14) and et.shape == (6, 14, 14) and g.shape == (6, 14, 512)
        assert coord.shape == (6, 14, 3)
        smiles, y, w, stats = bench.make_head_batch(1234)
        assert smiles.shape == (6, 9, 512) and y.shape == (6, 1) and w.shape == (6,)
        assert abs(float(w.mean()) - 1.0) < 1e-6                         # sample weights normalised to mean 1
        assert all(v.shape == (bench.FDS_BUCKETS, 512) for v in stats.values()) and (stats["running_var_last_epoch"] > 0).all()
        # every rank draws the SAME FDS statistics, but its own molecules
        _, y2, _, stats2 = bench.make_head_batch(1235)
        assert not torch.equal(y, y2) and all(torch.equal(stats[k], stats2[k]) for k in stats)
    finally:
        bench.set_shape(128, 64, 64)
