"""bench.py --impl reference: the CPU arm of the bench contract (the oracle prover at the FULL configuration, never the GPU
library). Runs here without a GPU: the line's keys, the proofs actually run, and the torchrun form where rank 0 alone works."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _line(out):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def _env():
    env = dict(os.environ)
    env["WF_REF_BUDGET_S"] = "5"      # one proof of the configuration, then stop
    env["CUDA_VISIBLE_DEVICES"] = ""  # the arm must not need a device
    return env


def test_reference_arm_line_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=_env(), cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    assert REQUIRED <= set(d), sorted(REQUIRED - set(d))
    assert d["impl"] == "reference" and d["metric"] == "prover_ms" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["gpu_launches"] == 0 and d["n_gpus"] == 1
    assert 1 <= d["steps"] <= 2 and d["steps_requested"] == 2           # the proofs actually run inside the wall budget
    assert d["value"] == d["ms_per_step"] == d["cpu_baseline"]["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "FULL configuration" in cb["sample"] and "2^20 rows x 8 columns" in cb["sample"]
    assert "cfg2" in d["config"]["workload"]


def test_reference_arm_under_torchrun_only_rank0_works():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--config", "cfg2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=_env(), cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)                                                   # exactly one line: rank 1 printed nothing
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong" and d["steps"] == 1


def test_sharded_line_helper_reproduces_the_recorded_8gpu_keys():
    # bench.sharded_records builds the roofline / throughput keys of an N > 1 line from rank 0's stage times: fed with the
    # stage times of the committed 8-GPU line it must give that line's own numbers (and the ALU-pipe roofline beside them)
    sys.path.insert(0, ROOT)
    import bench
    d = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r2_bench_line_8gpu.json")) if l.startswith("{")][0])
    r = bench.sharded_records(d["stage_ms"], 22, 32, 8, d["roofline"]["peak"], d["roofline"]["peak_source"], d["clocks"]["sm_mhz"])
    assert r["ntt_gelem_per_s"] == d["ntt_gelem_per_s"] and r["lde_commit_fri_ms"] == d["lde_commit_fri_ms"]
    assert r["merkle_leaves_per_s"] == d["merkle_leaves_per_s"]
    for k in ("achieved", "frac", "algorithmic_bytes", "kernel_ms", "aggregate_GBps"):
        assert r["roofline"][k] == d["roofline"][k], k
    alu = r["roofline"]["alu_pipe"]
    assert alu["bound"] == "alu_pipe" and 0.3 < alu["frac"] < 1.0
