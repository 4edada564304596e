#!/usr/bin/env python3
"""bench.py — one "step" = one complete STARK proof of BASELINE.json configs[1]: the 8-column
   "FibSmall x 4" AIR (4 copies of the reference's fib_small AIR side by side) on a 2^20-row Goldilocks
   trace, blowup 8, Blake3_256, base field, 32 queries, FRI folding 4 / remainder max degree 31,
   grinding 16 — trace interpolation + LDE + row hashing + Merkle commitment, constraint evaluation,
   composition polynomial LDE + commitment, OOD frames, DEEP composition, FRI commit phase, PoW
   grinding, query openings and proof serialization (Prover::prove, prover/src/lib.rs:250-492).
   The emitted proof is byte-identical to the CPU oracle's (tests/test_gpu_prover.py).

   python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-n L] [--pairs K]

Prints ONE JSON line (rank 0). `value` = ms per proof with the trace already resident in HBM;
`e2e` = ms per proof through the C ABI with HOST buffers (pinned trace columns copied H2D inside the
timed region, proof bytes returned to the host). N > 1: every rank proves its own independent trace
(weak scaling, no data-path collective); value = max over ranks of ms per proof.
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
# dram__bytes_read.sum + dram__bytes_write.sum of the four ntt_pass_kernel launches of the cfg2 trace interpolate + LDE,
# from the committed ncu --set full capture (profiles/r1_ntt_pass_v2_summary.txt, launches 0-3)
NCU_TRAFFIC_BYTES = 2495322112
METRIC = "prover_ms"
FOLDING, REM_MAX_DEG, LOG_BLOWUP, NUM_QUERIES, GRINDING = 4, 31, 3, 32, 16


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nv, self.h = self._nvml()  # initialised here, before the timed region starts

    def _nvml(self):
        """NVML in-process (the library nvidia-smi itself reads): no process spawn inside the timed region —
        a cold `nvidia-smi` start takes a driver lock for several ms and once stretched a 10-step mean by 14 %."""
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None, None

    def run(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                if nv is not None:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    bit = lambda name: "Active" if r & getattr(nv, name, 0) else "Not Active"
                    self.samples.append([str(sm), str(mx), bit("nvmlClocksThrottleReasonHwSlowdown"),
                                         bit("nvmlClocksThrottleReasonHwThermalSlowdown"), bit("nvmlClocksThrottleReasonSwThermalSlowdown"),
                                         bit("nvmlClocksThrottleReasonSwPowerCap")])
                    time.sleep(0.2)
                    continue
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


def workload_name(log_n, pairs):
    return (f"cfg2: full STARK proof of FibSmall x {pairs} ({2 * pairs} columns) on a 2^{log_n}-row Goldilocks trace, blowup 8, "
            f"Blake3_256, base field, {NUM_QUERIES} queries, FRI folding {FOLDING}, remainder max degree {REM_MAX_DEG}, grinding {GRINDING}: "
            "trace LDE+commit, constraint evaluation, composition LDE+commit, OOD, DEEP, FRI, grinding, openings")


def proof_opts():
    return np.array([NUM_QUERIES, 1 << LOG_BLOWUP, GRINDING, 1, FOLDING, REM_MAX_DEG, 0, 0, 0], dtype=np.uint32)


def host_cores():
    """Usable host cores: CPU affinity, capped by the cgroup CPU quota (a container with a quota of
    8 cores on a 128-thread host must not run 128 OpenMP threads)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference's prover, OpenMP with the reference's
# `concurrent` decomposition where it has one) on a bounded sample of the same workload.
# --------------------------------------------------------------------------------------------------
def cpu_sample(log_n, pairs, sample_log_n, steps, warmup):
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    from oracle import oracle as o
    o.lib()
    cores = host_cores()
    o.set_threads(cores)
    tr, res = o.build_fib_trace(pairs, 1 << sample_log_n)
    opts = proof_opts()
    for _ in range(warmup):
        o.prove_fib(tr, res, opts)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.prove_fib(tr, res, opts)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    scale = 1 << (log_n - sample_log_n)
    return dt * 1e3 * scale, cores, (f"oracle prover (C++ restatement of winterfell v0.13.1 generate_proof, OpenMP {cores} threads) on "
                                     f"2^{sample_log_n} rows (1/{scale} of the workload), time scaled x{scale} (linear; favours the CPU by log n)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample_log_n = min(args.log_n, 15)
    ms, cores, sample = cpu_sample(args.log_n, args.pairs, sample_log_n, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(ms, 3), "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": {"workload": workload_name(args.log_n, args.pairs)},
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--pairs", type=int, default=4)
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
    log_n, pairs = args.log_n, args.pairs
    cols = 2 * pairs
    n = 1 << log_n
    N = n << LOG_BLOWUP
    opts = proof_opts()
    # a valid trace: the AIR recurrence from the pair starts (j+1, j+1)
    trace = np.zeros((cols, n), dtype=np.uint64)
    results = np.zeros(pairs, dtype=np.uint64)
    for j in range(pairs):
        va, vb = j + 1, j + 1
        ca, cb = [0] * n, [0] * n
        for i in range(n):
            ca[i], cb[i] = va, vb
            va = va + vb
            if va >= P:
                va -= P
            vb = vb + va
            if vb >= P:
                vb -= P
        trace[2 * j] = np.array(ca, dtype=np.uint64)
        trace[2 * j + 1] = np.array(cb, dtype=np.uint64)
        results[j] = cb[n - 1]
    host = torch.from_numpy(trace.view(np.int64)).pin_memory()          # pinned host trace (ColMatrix columns)
    host_np = host.numpy().view(np.uint64)
    dev = host.cuda(non_blocking=False)                                  # resident copy for the kernel-only arm
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > L2 (126 MB)
    out_buf = np.zeros(1 << 22, dtype=np.uint8)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step_resident():
        return ctx.prove_fib_dev(dev.data_ptr(), pairs, log_n, results, opts, out_buf)

    e2e_buf = np.zeros(1 << 23, dtype=np.uint8)  # proof bytes land here (allocated once, like a caller's buffer)

    def step_e2e():
        return ctx.prove_fib(host_np, results, opts, out_buf=e2e_buf)     # H2D of the trace and D2H of the proof inside

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            p_res = step_resident()
        p_e2e = step_e2e()
        assert p_res == p_e2e, "resident and e2e arms produced different proofs"

        import gc
        gc.collect()
        gc.disable()  # no collector pauses inside the timed regions (a single 30 ms host stall moves a 30-step mean by 1 ms)
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
            if os.environ.get("WF_BENCH_TRACE"):
                print(f"resident step {a.elapsed_time(b):.3f} ms", file=sys.stderr)
        barrier()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3
        launches = int(ctx.launches - l0)
        ms_step = total_ms / args.steps

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
            if os.environ.get("WF_BENCH_TRACE"):
                print(f"e2e step {a.elapsed_time(b):.3f} ms", file=sys.stderr)
        barrier()
        e2e_step = e2e_ms / args.steps
        sampler.stop_flag = True
        sampler.join(timeout=2)
        gc.enable()

        # stage breakdown: one extra proof with the library's stage events on
        flush.zero_()
        ctx.set_profiling(True)
        step_resident()
        breakdown = {k: round(v, 4) for k, v in ctx.stage_times()}
        ctx.set_profiling(False)

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
        ntt_ms = breakdown["trace_interpolate"] + breakdown["trace_lde"]
        alg_bytes = 8.0 * n * (2 + (1 << LOG_BLOWUP)) * cols
        achieved = alg_bytes / (ntt_ms * 1e-3) / 1e9
        lcf = sum(breakdown[k] for k in ("trace_interpolate", "trace_lde", "trace_commit", "composition_lde", "composition_commit", "fri_layers"))
        line = {
            "metric": METRIC, "value": round(ms_step, 4), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(log_n, pairs), "l2": "256 MiB memset between timed steps (L2 flush)",
                       "parallelism": f"{world} independent proofs, one per GPU" if world > 1 else "single GPU"},
            "e2e": {"value": round(e2e_step, 4), "unit": "ms", "h2d_bytes_per_step": int(trace.nbytes), "d2h_bytes_per_step": len(p_e2e)},
            "gpu_launches": launches // max(args.steps, 1),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "ntt_pass_kernel (trace interpolate + LDE launches)", "achieved": round(achieved, 1), "peak": hbm,
                         "unit": "GB/s", "frac": round(achieved / hbm, 4),
                         "traffic": NCU_TRAFFIC_BYTES if (log_n, cols) == (20, 8) else None, "peak_source": peak_src,
                         "algorithmic_bytes": int(alg_bytes), "kernel_ms": round(ntt_ms, 4),
                         "launches": 4,
                         "note": "achieved / traffic / algorithmic_bytes are sums over the 4 launches of the group (2 iNTT passes, "
                                 "2 LDE passes with all 8 cosets in grid.z). traffic = dram__bytes_read.sum + dram__bytes_write.sum "
                                 "of the same 4 launches in profiles/r1_ntt_pass_v2_summary.txt (ncu --set full): 3.7x the algorithmic "
                                 "bytes because the two-pass schedule writes and re-reads the intermediate. The kernel is integer-ALU-"
                                 "bound (ALU pipe 55-59 %, FMA 28-32 %, issue slots 66-68 %, DRAM 21 % of peak); see DESIGN.md 4"},
            "stage_ms": breakdown,
            "lde_commit_fri_ms": round(lcf, 4),
            "proof_bytes": len(p_e2e),
            "ntt_gelem_per_s": round(N * cols / (ntt_ms * 1e-3) / 1e9, 3),
            "merkle_leaves_per_s": round(N / (breakdown["trace_commit"] * 1e-3), 1),
            "wall_ms_per_step_incl_flush": round(wall_ms / args.steps, 3),
        }
        if not args.no_cpu_baseline:
            # in a fresh process: torch has already initialised libgomp in this one with the default
            # (spinning) wait policy, which thrashes under a cgroup CPU quota
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                                      "--log-n", str(log_n), "--pairs", str(pairs)], capture_output=True, text=True, timeout=600,
                                     env={**os.environ, "OMP_WAIT_POLICY": "PASSIVE", "RANK": "0", "WORLD_SIZE": "1"}).stdout.strip().splitlines()[-1]
                line["cpu_baseline"] = json.loads(out)["cpu_baseline"]
            except Exception as e:  # the reported baseline must never take the GPU line down
                line["cpu_baseline"] = {"value": None, "unit": "ms", "cores": host_cores(), "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
