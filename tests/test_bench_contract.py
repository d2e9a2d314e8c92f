"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs the unmodified reference
(oracle/_ref/spk_ref replicas; the C oracle port where that binary is absent) on the host cores and
prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "site_updates_per_sec" and line["unit"] == "Gsites/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == "potts_neighonly_sc26n_1024"  # BASELINE configs[2], the headline
    # the arm states what it really ran: a bounded sample, the reference's own random stream
    rs = line["reference_sample"]
    assert rs["per_replica_cells"] == [96, 96, 96] and rs["scaled_from_per_gpu_cells"] == [1024, 1024, 1024]
    assert rs["replicas"] == line["cpu_baseline"]["cores"] and "rng" in rs
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Gsites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: on a box without CUDA the product arm of bench.py refuses to run."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-cpu", "--no-extra",
                          "--no-plugin"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "needs a CUDA device" in (out.stderr + out.stdout)


def test_workload_tables_are_consistent():
    """every bench workload has a traffic figure from a committed ncu capture and a CPU sample size;
    the multi-GPU check's cases split into even slabs at 2, 4 and 8 ranks."""
    sys.path.insert(0, ROOT)
    import bench
    from spparks_b200.selfcheck import cases
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert bench.DEFAULT_WORKLOAD == "potts_neighonly_sc26n_1024"
    for name, w in bench.WORKLOADS.items():
        assert traffic.get(name, 0) > 0, name
        assert bench.nsites(w["lattice"], w["cpu_n"]) * 148 < 4e9  # the reference needs ~148 B/site on the host
        cfg = bench.make_config(name, w, 1)
        assert cfg["B_alg_bytes_per_site_update"] == 4 * (bench.NCOLORS[w["lattice"]] + 1)
    for world in (2, 4, 8):
        for model, lat, n, q, T in cases(world):
            d = n[2] if lat.startswith("sc") else n[1]
            assert d % world == 0 and (d // world) % 2 == 0 and d // world >= (2 if lat.startswith("sc") else 4)
