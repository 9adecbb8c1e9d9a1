"""CPU test of the bench.py contract: the reference arm (`--impl reference`, the oracle port on the host cores — the one
place besides `cpu_baseline` where bench.py may execute oracle/) prints ONE JSON line with the keys the driver reads, on
the same metric / unit / workload string as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "leapfrog_grad_evals_per_sec" and d["unit"] == "grad-evals/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 1e3 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("simulation sweep: 1024 datasets in total")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["ess"]["ess_per_s_min_coord"] > 0 and cb["c_port_value"] > cb["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_workload_strings_and_flop_model():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.parse_args.__globals__["argparse"].Namespace(datasets=1024, weak=False, chains=4, K=20, n_per_group=10, method="hmc",
                                                          num_results=20000, num_burnin=5000, leapfrog=10)
    assert bench.workload_name(a) == ("simulation sweep: 1024 datasets in total, sharded over the GPUs x 4 chains, K=20 N=20 D=1, "
                                      "sample_hmc(20000,5000) L=10")
    a.weak = True
    assert "1024 datasets per GPU" in bench.workload_name(a)
    # 64 NK + 48 N + 4 NKD per gradient (28,160 at config 5), 28 (2NK + 2N) per value (23,520)
    assert bench.flops_per_grad(20, 20, 1) == 28160 and bench.flops_per_logp(20, 20) == 23520
