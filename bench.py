#!/usr/bin/env python3
"""bench.py — one "step" = one pass of the STARK proving hot path over one synthetic trace:
   trace commit (interpolate 8 columns of 2^20 rows, LDE at blowup 8, Blake3_256 row hashes, Merkle
   tree) + FRI commit phase (folding 4, remainder max degree 31) over 2^23 evaluations.
   That is BASELINE.json configs[1] (2^20 x 8 Goldilocks, blowup 8, Blake3_256), the configuration
   the headline metric is quoted on that fits one GPU.

   python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-n L] [--cols C]

Prints ONE JSON line (rank 0). `value` = ms per step with the trace already resident in HBM;
`e2e` = ms per step through the C ABI with HOST buffers (pinned trace columns copied H2D inside the
timed region, roots read back). N > 1: every rank proves its own independent trace (weak scaling,
no data-path collective); value = max over ranks of ms per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 0xFFFFFFFF00000001
METRIC = "prover_ms_lde_commit_fri"
FOLDING, REM_MAX_DEG, LOG_BLOWUP = 4, 31, 3


def rand_trace(cols, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=(cols, n), dtype=np.uint64)
    a[a >= np.uint64(P)] -= np.uint64(P)   # values in [0, p)
    return a


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max((int(s[1]) for s in self.samples if s[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


def workload_name(log_n, cols):
    return (f"cfg2: 2^{log_n}x{cols} Goldilocks trace, blowup 8, Blake3_256: interpolate+LDE+row-hash+Merkle (trace commit) "
            f"+ FRI commit phase (folding {FOLDING}, remainder max degree {REM_MAX_DEG}) over 2^{log_n + LOG_BLOWUP} evaluations "
            "of a degree<n codeword (LDE column 0)")


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference's algorithms, OpenMP with the reference's
# `concurrent` decomposition) on a bounded sample of the same workload.
# --------------------------------------------------------------------------------------------------
def cpu_step(o, trace, log_b):
    polys = o.interpolate_columns(trace)
    lde = o.lde_rows(polys, 1 << log_b)
    leaves = o.hash_rows(o.BLAKE3, lde)
    nodes = o.merkle_nodes(o.BLAKE3, leaves)
    roots, rem, _ = o.fri_build_layers(o.BLAKE3, np.ascontiguousarray(lde[:, 0]), FOLDING, REM_MAX_DEG, 1 << log_b)
    return nodes[1].tobytes(), roots


def cpu_sample(log_n, cols, sample_log_n, steps, warmup):
    from oracle import oracle as o
    o.lib()
    cores = os.cpu_count() or 1
    o.set_threads(cores)
    tr = rand_trace(cols, 1 << sample_log_n, 7)
    for _ in range(warmup):
        cpu_step(o, tr, LOG_BLOWUP)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(o, tr, LOG_BLOWUP)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    scale = 1 << (log_n - sample_log_n)
    return dt * 1e3 * scale, cores, (f"oracle (C++ restatement of winterfell v0.13.1, OpenMP {cores} threads) on 2^{sample_log_n}x{cols} rows "
                                     f"(1/{scale} of the workload), time scaled x{scale}")


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample_log_n = min(args.log_n, 16)
    ms, cores, sample = cpu_sample(args.log_n, args.cols, sample_log_n, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(ms, 3), "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": {"workload": workload_name(args.log_n, args.cols)},
        "cpu_baseline": {"value": round(ms, 3), "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ms, 3), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--cols", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import winterfell_b200 as wf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    stream = torch.cuda.Stream()
    ctx = wf.Context(local_rank, stream.cuda_stream)
    log_n, cols = args.log_n, args.cols
    n = 1 << log_n
    N = n << LOG_BLOWUP
    trace = rand_trace(cols, n, 1234 + rank)
    host = torch.from_numpy(trace.view(np.int64)).pin_memory()          # pinned host trace (ColMatrix columns)
    host_np = host.numpy().view(np.uint64)
    dev = host.cuda(non_blocking=False)                                  # resident copy for the kernel-only arm
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > L2 (126 MB)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step_resident(stage_events=None):
        """Hot path on device-resident input. Returns (root, fri roots)."""
        def mark(name):
            if stage_events is not None:
                e = ev(); e.record(stream); stage_events.append((name, e))
        mark("start")
        m = ctx.mat_from_device_columns(dev.data_ptr(), cols, n); mark("layout")
        polys = m.interpolate(); mark("interpolate")
        lde = polys.lde(LOG_BLOWUP); mark("lde")
        tree = ctx.commit_rows(wf.HASH_BLAKE3_256, lde); mark("commit")
        # FRI codeword: column 0 of the LDE (degree < n) as its own 1-column matrix
        fm = lde.select_columns(0, 1); mark("fri_input")
        f, roots = ctx.fri_build_layers_default(wf.HASH_BLAKE3_256, fm, 1, FOLDING, REM_MAX_DEG, 1 << LOG_BLOWUP); mark("fri")
        root = tree.root()
        for h in (m, polys, lde, tree, fm, f):
            h.free()
        return root, roots

    def step_e2e():
        """Same path through the host-buffer entry points (what the Rust shim calls)."""
        m = ctx.mat_from_host_columns(host_np)           # H2D of the 8 columns inside the timed region
        polys = m.interpolate()
        lde = polys.lde(LOG_BLOWUP)
        tree = ctx.commit_rows(wf.HASH_BLAKE3_256, lde)
        fm = lde.select_columns(0, 1)
        f, roots = ctx.fri_build_layers_default(wf.HASH_BLAKE3_256, fm, 1, FOLDING, REM_MAX_DEG, 1 << LOG_BLOWUP)
        root = tree.root()                               # D2H of the commitment
        for h in (m, polys, lde, tree, fm, f):
            h.free()
        return root, roots

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            r_res = step_resident()
        r_e2e = step_e2e()
        assert r_res[0] == r_e2e[0] and (r_res[1] == r_e2e[1]).all(), "resident and e2e arms disagree"

        # ---- kernel-resident arm: K steps, each bracketed by events, L2 flushed between steps ----
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        l0 = ctx.launches
        total_ms = 0.0
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            a, b = ev(), ev()
            a.record(stream)
            step_resident()
            b.record(stream)
            b.synchronize()
            total_ms += a.elapsed_time(b)
        barrier()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3
        launches = int(ctx.launches - l0)
        ms_step = total_ms / args.steps

        # ---- e2e arm ----
        barrier()
        e2e_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            a, b = ev(), ev()
            a.record(stream)
            step_e2e()
            b.record(stream)
            b.synchronize()
            e2e_ms += a.elapsed_time(b)
        barrier()
        e2e_step = e2e_ms / args.steps
        sampler.stop_flag = True
        sampler.join(timeout=2)

        # ---- stage breakdown (one extra step with events between the stages) ----
        flush.zero_()
        stages = []
        step_resident(stages)
        torch.cuda.synchronize()
        breakdown = {stages[i][0]: round(stages[i - 1][1].elapsed_time(stages[i][1]), 4) for i in range(1, len(stages))}

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_step, e2e_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel: ntt_pass_kernel (K1 + K2 = interpolate + LDE). Algorithmic bytes per base
        # column = 8n(2 + b): read trace, write polys, write LDE.
        ntt_ms = breakdown["interpolate"] + breakdown["lde"]
        alg_bytes = 8.0 * n * (2 + (1 << LOG_BLOWUP)) * cols
        achieved = alg_bytes / (ntt_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(ms_step, 4), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(log_n, cols), "l2": "256 MiB memset between timed steps (L2 flush)",
                       "parallelism": f"{world} independent traces, one per GPU" if world > 1 else "single GPU"},
            "e2e": {"value": round(e2e_step, 4), "unit": "ms", "h2d_bytes_per_step": int(trace.nbytes),
                    "d2h_bytes_per_step": int(32 * (1 + len(r_res[1])))},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "ntt_pass_kernel (interpolate + LDE launches)", "achieved": round(achieved, 1), "peak": hbm,
                         "unit": "GB/s", "frac": round(achieved / hbm, 4), "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes": int(alg_bytes), "kernel_ms": round(ntt_ms, 4)},
            "stage_ms": breakdown,
            "ntt_gelem_per_s": round(N * cols / (ntt_ms * 1e-3) / 1e9, 3),
            "merkle_leaves_per_s": round(N / (breakdown["commit"] * 1e-3), 1),
            "wall_ms_per_step_incl_flush": round(wall_ms / args.steps, 3),
        }
        if not args.no_cpu_baseline:
            ms, cores, sample = cpu_sample(log_n, cols, min(log_n, 16), 2, 1)
            line["cpu_baseline"] = {"value": round(ms, 2), "unit": "ms", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
