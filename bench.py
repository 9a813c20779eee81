#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200: KKT factor+solve FP64 TFLOP/s (cfg3, n=8192
normal-equations reduction) plus batched IPM solves/s (cfg4) in the same JSON line.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    the reference's own CPU path

A "step" of the headline metric is one pass of the hot path over the resident condensed
matrix K = Hx + M^T W M of the cfg3 QP: out-of-place LDL^T (root-free Cholesky) + two
triangular solves (predictor and corrector right-hand sides), i.e. N^3/3 + 2*2N^2 flops --
the linear algebra of one IPM iteration.  `value` is that with K resident in HBM; `e2e` is
the same metric through the solver C ABI from pinned HOST buffers (ipmz_create uploads the
QP, ipmz_solve runs the whole IPM loop incl. assembly, ipmz_get_iterate reads the answer).
N > 1: the single large KKT does not shard ("replicas only") so every rank runs a
replica; the batched workload shards 4096 QPs by problem index with no collective.
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
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "kkt_factor_solve_fp64_tflops"
UNIT = "TFLOP/s"
CFG3 = dict(n=8192, m=4096, seed=3)
CFG4 = dict(n=256, m=128, count=4096, seed0=1000)


def flops_factor_solve(N, nrhs=2):
    return N ** 3 / 3.0 + nrhs * 2.0 * N * N


def make_cfg3(n, m, seed):
    """cfg3 inputs: Q = 3I + sym N(0,1/n), A ~ N(0,1/n), l/u = A x0 -+ 0.25, box [-1,1]."""
    rng = np.random.default_rng(seed)
    S = rng.standard_normal((n, n)) / np.sqrt(n)
    Q = 3.0 * np.eye(n) + 0.5 * (S + S.T)
    del S
    c = rng.standard_normal(n)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    x0 = rng.uniform(-0.5, 0.5, n)
    mid = A @ x0
    return dict(Q=Q, c=c, A=A, l_A=mid - 0.25, u_A=mid + 0.25, l_x=-np.ones(n), u_x=np.ones(n))


def static_config(n, m, N, world):
    """The `config` both arms print (identical keys and values: the driver compares them); run-dependent numbers
    (residuals, wall time, the reference arm's sample) live in `details` / `cpu_baseline.sample`."""
    return {"workload": "cfg3: dense QP n=%d m=%d, normal-equations reduction; step = out-of-place LDL^T "
                        "(root-free Cholesky) of the condensed KKT + 2 triangular solves" % (n, m),
            "N": N, "flops_per_step": flops_factor_solve(N), "parallelism": "replicas x%d" % world,
            "l2": "input matrix %.0f MB > 126 MB L2, re-read from HBM every step" % (N * N * 8 / 1e6)}


def condensed_block(d, ns):
    """Leading ns x ns block of K = Hx + A^T W A at the reference's initial point (all slacks
    and duals 1: Hx = Q + 2I, W = 2I) -- the CPU sample matrix of the reference arm."""
    A = d["A"][:, :ns]
    return np.ascontiguousarray(d["Q"][:ns, :ns] + 2.0 * np.eye(ns) + 2.0 * (A.T @ A))


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    POLLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
try:
    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
except Exception:
    mx = -1
print("max", mx, flush=True)
while True:
    try:
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        print("%.6f %d %d" % (time.time(), sm, rs), flush=True)
    except Exception:
        pass
    time.sleep(0.001)
"""

    def __init__(self, device):
        # NVML polled every millisecond by a separate process (no GIL contention with the bench thread);
        # start() / stop() only mark the time window whose samples are reported
        self.device, self.rows, self.proc, self.mx = device, [], None, None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device
        if vis:
            parts = vis.split(",")
            if device < len(parts) and parts[device].strip().isdigit():
                idx = int(parts[device])
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", self.POLLER, str(idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        self.t0 = self.t1 = None

    def _read(self):
        for line in self.proc.stdout:
            t = line.split()
            if t and t[0] == "max":
                self.mx = int(t[1])
            elif len(t) == 3:
                self.rows.append((float(t[0]), int(t[1]), int(t[2])))

    def wait_ready(self, timeout=90.0):
        """Block until the helper process has delivered its first sample (a fresh box pages the interpreter and
        pynvml in slowly), so that the timed region is never entered with a poller that is not polling yet."""
        t_end = time.time() + timeout
        while self.proc and self.proc.poll() is None and not self.rows and time.time() < t_end:
            time.sleep(0.01)
        return bool(self.rows)

    def start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["NVML poller unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for r in self.rows if self.t0 <= r[0] <= self.t1]
        how = "NVML polled every ms by a helper process; samples inside the timed region"
        if not rows and self.rows:  # a region shorter than one NVML query: the samples that bracket it
            before = [r for r in self.rows if r[0] < self.t0][-1:]
            after = [r for r in self.rows if r[0] > self.t1][:1]
            rows = before + after
            how = "NVML polled by a helper process; the timed region was shorter than one query: nearest samples before/after it"
        sm = [r[1] for r in rows]
        bits = 0
        for r in rows:
            bits |= r[2]
        names = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        reasons = [nm for nm, bit in names if bits & bit]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx if self.mx and self.mx > 0 else None,
                "samples": len(sm), "reasons": sorted(reasons), "how": how}


# ------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref = unmodified sources),
    one single-threaded replica per host core, on a bounded sample of the cfg3 workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import oracle_lib as ol
    kind = "reference" if ol.have_ref() else "port"
    ns = args.sample_n
    d = make_cfg3(CFG3["n"] if not args.quick else 2 * ns, CFG3["m"] if not args.quick else ns, CFG3["seed"])
    K = condensed_block(d, ns)
    b = np.random.default_rng(11).standard_normal(ns)
    cores = args.cores or os.cpu_count() or 1
    steps, warm = args.steps, args.warmup

    def worker(q):
        L = ol.ref() if kind == "reference" else ol.port()
        fac = L.ref_ldlt if kind == "reference" else L.orc_ldlt
        sol = L.ref_solve_ldlt if kind == "reference" else L.orc_solve_ldlt
        Lm, D = np.zeros((ns, ns)), np.zeros(ns)
        times = []
        for it in range(warm + steps):
            t0 = time.perf_counter()
            fac(ns, ol._ptr(K), ol._ptr(Lm), ol._ptr(D))
            for _ in range(2):
                x = b.copy()
                sol(ns, ol._ptr(Lm), ol._ptr(D), ol._ptr(x))
            times.append(time.perf_counter() - t0)
        q.put((sum(times[warm:]), float(np.max(np.abs(K @ x - b)))))

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(q,)) for _ in range(cores)]
    t0 = time.perf_counter()
    for p in procs:
        p.start()
    outs = [q.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.perf_counter() - t0
    tmax = max(o[0] for o in outs)
    value = cores * steps * flops_factor_solve(ns) / tmax * 1e-12
    sample = ("each step = LinearSolvers::ldlt_decomposition + 2x overwriting_solve_ldlt on the leading %dx%d block "
              "of the cfg3 condensed KKT (the full N = 8192 step is ~90 s per core: bounded sample, same TFLOP/s "
              "metric, flops counted for the sample size), one single-threaded replica per host core" % (ns, ns))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": tmax / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": static_config(CFG3["n"], CFG3["m"], CFG3["n"], world) if not args.quick else
        static_config(2 * ns, ns, 2 * ns, world),
        "details": {"residual": max(o[1] for o in outs), "wall_s": wall, "sample_n": ns},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import ipm_zoo_b200 as z

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available() and z.device_count() > 0, "bench needs a GPU: no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n, m = (CFG3["n"], CFG3["m"]) if not args.quick else (1024, 512)
    N = n
    d = make_cfg3(n, m, CFG3["seed"])
    peak = z.fp64_peak_tflops(local)

    # ---- headline: device-resident factor + 2 solves on the condensed cfg3 matrix ----
    prob = z.Problem(d["Q"], d["c"], d["A"], d["l_A"], d["u_A"], None, None, d["l_x"], d["u_x"])
    opt = z.Options(reduction=z.NORMAL, device=local)
    s = z.Solver(prob, opt)
    Kc = s.assemble()
    s.close()
    sampler = ClockSampler(local)  # helper process: polling long before the timed region starts
    fac = z.Factor(N, device=local)
    fac.set_matrix(Kc)
    rhs = np.random.default_rng(11).standard_normal(N)
    fac.set_rhs(rhs)
    for _ in range(args.warmup):
        fac.run(1, 2)
    x = fac.solution()
    resid = float(np.max(np.abs(Kc @ x - rhs)) / np.max(np.abs(rhs)))
    launches0 = z.launch_count()
    sampler.wait_ready()
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ms_dev = fac.run(args.steps, 2)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = z.launch_count() - launches0
    ms_max = allmax(ms_dev)
    flops_step = flops_factor_solve(N)
    value = world * args.steps * flops_step / (ms_max * 1e-3) * 1e-12

    prof = fac.profile()
    info = fac.info()
    syrk_tf = prof["syrk_flops"] / (prof["syrk_ms"] * 1e-3) * 1e-12 if prof["syrk_ms"] > 0 else 0.0
    step_ms = ms_dev / args.steps
    factor_only_ms = fac.run(args.steps, 0) / args.steps
    peak_src = {"used": "live DMMA issue-rate probe (ipmz_fp64_peak_probe), this run", "probe_tflops": peak,
                "cublas_dgemm_8192_tflops": 36.0, "cusolver_dpotrf_8192_tflops": 23.8,
                "library_numbers_from": "profiles/r01_fp64_ceilings.log (tools/fp64_probe.cu on this pool)",
                "note": "MEASURED_PEAKS.json has no FP64 entry; against cuBLAS DGEMM the fractions are x%.3f" %
                        (peak / 36.0 if peak else 0.0)}
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")))
        k = tj["k_ldlt_dataflow"]
        if k["n"] == N and info["dataflow"]:
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
    except Exception:
        traffic = None
    if info["dataflow"]:
        # the whole factorization is ONE persistent kernel: algorithmic N^3/3 flops over its duration
        roofline = {
            "bound": "tensor", "kernel": "k_ldlt_dataflow (persistent dataflow LDL^T: DMMA tile updates + panel tasks)",
            "achieved": syrk_tf, "peak": peak, "unit": "TFLOP/s", "frac": syrk_tf / peak if peak else None,
            "traffic": traffic, "traffic_unit": "DRAM bytes per launch (profiles/r01_traffic.json)",
            "peak_source": peak_src,
            "flops_per_launch": prof["syrk_flops"], "launches_per_step": 1,
            "ms_per_launch": prof["syrk_ms"], "share_of_step": prof["syrk_ms"] / step_ms if step_ms else None,
            "tasks_per_launch": info["ntasks"], "simulated_schedule_ms": info["simulated_us"] * 1e-3,
            "factor_ms": factor_only_ms, "solves_ms": step_ms - factor_only_ms,
            "whole_step_frac_of_peak": (flops_step / (step_ms * 1e-3) * 1e-12) / peak if peak else None,
        }
    else:
        roofline = {
            "bound": "tensor", "kernel": "k_syrk_ldl (FP64 DMMA trailing update)",
            "achieved": syrk_tf, "peak": peak, "unit": "TFLOP/s", "frac": syrk_tf / peak if peak else None,
            "traffic": None, "peak_source": peak_src,
            "flops_per_launch": prof["syrk_flops"] / max(1, prof["syrk_launches"]),
            "launches_per_step": prof["syrk_launches"],
            "ms_per_launch": prof["syrk_ms"] / max(1, prof["syrk_launches"]),
            "share_of_step": prof["syrk_ms"] / step_ms if step_ms else None,
            "diag_ms": prof["diag_ms"], "panel_ms": prof["panel_ms"], "syrk_ms": prof["syrk_ms"],
            "whole_step_frac_of_peak": (flops_step / (step_ms * 1e-3) * 1e-12) / peak if peak else None,
        }
    fac.close()
    del Kc

    # ---- e2e: whole IPM solve through the C ABI from pinned host buffers ----
    pin = {}
    for k in ("Q", "c", "A", "l_A", "u_A", "l_x", "u_x"):
        pin[k] = z.pinned_empty(d[k].shape)
        pin[k][...] = d[k]
    pprob = z.Problem(pin["Q"], pin["c"], pin["A"], pin["l_A"], pin["u_A"], None, None, pin["l_x"], pin["u_x"])
    h2d = sum(int(v.nbytes) for v in pin.values())
    e2e_runs = []
    for rep in range(1 + max(1, min(args.steps, 3))):  # one untimed + up to three timed repetitions, median reported
        barrier()
        t0 = time.perf_counter()
        sv = z.Solver(pprob, opt)
        r = sv.solve()
        it = sv.iterate()
        t1 = time.perf_counter() - t0
        if rep == 0:  # untimed repetition: HBM-side evidence of the streaming kernels at cfg3 size
            probe = sv.probe_kernels(20)
        sv.close()
        if rep > 0:
            e2e_runs.append((t1, r))
    # median: one repetition in three occasionally pays 20-40 ms of cudaMalloc / page-pinning noise inside ipmz_create
    t_e2e = allmax(float(np.median([t for t, _ in e2e_runs])))
    r = e2e_runs[-1][1]
    e2e_val = world * r.iterations * flops_factor_solve(N) / t_e2e * 1e-12
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(it.nbytes),
           "seconds": t_e2e, "iterations": r.iterations, "converged": r.converged, "f": r.f,
           "device_loop_ms": r.solve_ms,
           "call": "ipmz_create (H2D of Q, A, bounds) + ipmz_solve (full Mehrotra loop, normal reduction) + "
                   "ipmz_get_iterate; flops counted = iterations x (N^3/3 + 4N^2), assembly flops not counted"}

    # achieved HBM GB/s of the streaming kernels (algorithmic bytes / CUDA-event time) against the measured copy peak
    hbm_peak, hbm_src = 6547.2, "fallback: MEASURED_PEAKS.json absent"
    try:
        hbm_peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                     "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    solve_ms = (step_ms - factor_only_ms) / 2.0
    hbm_kernels = [{"kernel": "k_trsv_fused (forward + backward sweep of one solve, N=%d)" % N, "ms": solve_ms,
                    "bytes": float(N) * N * 8.0, "GBps": float(N) * N * 8.0 / (solve_ms * 1e-3) * 1e-9,
                    "frac": float(N) * N * 8.0 / (solve_ms * 1e-3) * 1e-9 / hbm_peak,
                    "note": "dependent chain of N/128 block steps: latency-, not bandwidth-bound"}]
    for name, ms, by in probe:
        if ms > 0:
            hbm_kernels.append({"kernel": name, "ms": ms, "bytes": by, "GBps": by / (ms * 1e-3) * 1e-9,
                                "frac": by / (ms * 1e-3) * 1e-9 / hbm_peak})
    try:  # Bunch-Kaufman (indefinite KKT of EqualityHandling::None): bandwidth-bound unblocked pivoting
        nb, mb = (2048, 1024) if not args.quick else (256, 128)
        rngb = np.random.default_rng(3)
        Mb = rngb.standard_normal((nb, nb))
        Cb = rngb.standard_normal((mb, nb)) / np.sqrt(nb)
        Kb = np.block([[Mb @ Mb.T / nb + np.eye(nb), Cb.T], [Cb, np.zeros((mb, mb))]])
        ms_bk = z.bk_factor_time(Kb, 2)
        by_bk = float(nb + mb) ** 3 / 3.0 * 16.0
        hbm_kernels.append({"kernel": "k_bk_factor (Bunch-Kaufman, saddle-point KKT N=%d, bit-exact vs the reference)" % (nb + mb),
                            "ms": ms_bk, "bytes": by_bk, "GBps": by_bk / (ms_bk * 1e-3) * 1e-9,
                            "frac": by_bk / (ms_bk * 1e-3) * 1e-9 / hbm_peak,
                            "note": "bytes = the reference's N^3/3 multiply-subtract pairs x 16 B; the symmetric in-place "
                                    "layout moves 2x that; the 75 MB matrix is L2-resident"})
        del Kb, Mb, Cb
    except Exception as e:  # evidence only
        hbm_kernels.append({"kernel": "k_bk_factor", "error": repr(e)})
    hbm = {"peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src, "n": n, "m": m,
           "how": "algorithmic bytes per launch / CUDA-event time over 20 back-to-back launches on the library stream "
                  "(ipmz_probe_kernels); the three vector passes move < 1 MB and measure launch latency",
           "kernels": hbm_kernels}

    # ---- the other BASELINE.json configs, end to end through the C ABI (rank 0, single GPU) ----
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = bench_configs(z, args, local)

    # ---- batched IPM solves/s (cfg4), sharded by problem index, no collective ----
    batched = None
    if not args.no_batched:
        batched = bench_batched(z, args, world, rank, local, barrier, allmax, allsum)

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference arm in a subprocess ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                  "--warmup", "0", "--sample-n", str(args.sample_n)] + (["--quick"] if args.quick else []),
                                 capture_output=True, text=True, timeout=900)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the baseline is reported, never substituted
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": static_config(n, m, N, world),
            "details": {"solution_residual": resid, "wall_s": wall},
            "scaling_note": "value at n_gpus > 1 is n_gpus independent replicas of the single-GPU factorization (a "
                            "single large KKT does not shard: DESIGN section 5); the sharded workload is `batched` "
                            "(cfg4, strong scaling)",
            "configs": configs,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "hbm": hbm, "batched": batched,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_configs(z, args, local):
    """cfg1 (n=200, m_eq=100, box; normal-equations reduction), cfg2 (n=2048, m=1024; augmented quasi-definite LDL^T)
    and cfg5 (n=4096 portfolio, eps=1e-6; augmented): whole solve from pinned host buffers through the C ABI
    (ipmz_create: H2D; ipmz_solve; ipmz_get_iterate: D2H), best of 4 after one warm-up."""
    import workloads as P  # plain numpy generators at the repo root: the GPU arm imports nothing from tests/ or oracle/
    cases = [("cfg1", "n=200 m_eq=100 eq(SlackedSlacks)+box, normal equations", lambda: P.eq_box(200, 100, 1), z.NORMAL),
             ("cfg2", "n=2048 m=1024 ineq+box, augmented quasi-definite LDL^T", lambda: P.ineq_box(2048, 1024, 2, kind="shift"),
              z.AUGMENTED),
             ("cfg5", "n=4096 portfolio eps=1e-6 near convergence, augmented LDL^T", lambda: P.portfolio(4096, 32, 1e-6, 5),
              z.AUGMENTED)]
    if args.quick:
        cases = cases[:1]
    out = {}
    for name, what, make, red in cases:
        q = make()
        pin = {}
        for k in ("Q", "c", "A", "l_A", "u_A", "C", "d", "l_x", "u_x"):
            v = getattr(q, k)
            if v is not None:
                pin[k] = z.pinned_empty(v.shape)
                pin[k][...] = v
        pr = z.Problem(pin["Q"], pin["c"], pin.get("A"), pin.get("l_A"), pin.get("u_A"), pin.get("C"), pin.get("d"),
                       pin["l_x"], pin["u_x"], q.ineq_bounds, q.var_bounds, q.equalities)
        best, r = None, None
        for rep in range(5):  # best of 4 after one warm-up (host-side allocation times vary from box to box)
            t0 = time.perf_counter()
            sv = z.Solver(pr, z.Options(reduction=red, device=local))
            r = sv.solve()
            it = sv.iterate()
            sv.close()
            dt = time.perf_counter() - t0
            if rep > 0 and (best is None or dt < best):
                best = dt
        Nred = q.n if red == z.NORMAL else q.N
        out[name] = {"workload": what, "reduction": "normal" if red == z.NORMAL else "augmented", "N": Nred,
                     "e2e_ms": best * 1e3, "device_loop_ms": r.solve_ms, "iterations": r.iterations,
                     "converged": bool(r.converged), "ms_per_iteration": r.solve_ms / max(1, r.iterations), "f": r.f,
                     "h2d_bytes": sum(int(v.nbytes) for v in pin.values()), "d2h_bytes": int(it.nbytes),
                     "factor_solve_tflops_device": r.iterations * flops_factor_solve(Nred) / (r.solve_ms * 1e-3) * 1e-12}
    return out


def _cpu_solve_worker(args_):
    """One host process of the batched CPU baseline: the reference's own Optimizer::solve on cfg4 problems."""
    seeds, kind = args_
    import oracle_lib as ol
    import problems as P
    t0 = time.perf_counter()
    its = []
    for sd in seeds:
        q = P.ineq_box(CFG4["n"], CFG4["m"], sd, kind="shift")
        tr = ol.ref_solve(q, quiet=True, steps=False) if kind == "reference" else ol.port_solve(q, steps=False)
        its.append((tr.iterations, float(tr.f[min(tr.iterations, tr.n_logged - 1)]) if kind != "reference" else None))
    return time.perf_counter() - t0, its


def batched_cpu_baseline(per_core=2):
    """one reference process per host core over cfg4 problems (Optimizer ctor + solve(), stdout
    disabled), core count stated.  Also the same-run parity gate's oracle values (port: it exposes f)."""
    import multiprocessing as mp
    import oracle_lib as ol
    kind = "reference" if ol.have_ref() else "port"
    cores = os.cpu_count() or 1
    seeds = [[CFG4["seed0"] + c * per_core + j for j in range(per_core)] for c in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        outs = pool.map(_cpu_solve_worker, [(sd, kind) for sd in seeds])
        wall = time.perf_counter() - t0
        gate = pool.map(_cpu_solve_worker, [([CFG4["seed0"] + i], "port") for i in range(16)])
    nsolved = cores * per_core
    return ({"value": nsolved / wall, "unit": "solves/s", "cores": cores, "kind": kind,
             "sample": "%d cfg4 QPs (%d per core), one single-threaded %s process per host core, Optimizer ctor + solve() "
                       "with stdout disabled; wall %.2f s" % (nsolved, per_core, kind, wall),
             "per_core_solves_per_s": nsolved / wall / cores},
            [g[1][0] for g in gate])


def bench_batched(z, args, world, rank, local, barrier, allmax, allsum):
    import workloads as P  # the GPU arm's inputs; the CPU baseline / parity gate below wrap the same arrays for the oracle
    import torch
    total = CFG4["count"] if not args.quick else 64
    n, m = CFG4["n"], CFG4["m"]
    lo, hi = z.shard_range(total, world, rank)
    cnt = hi - lo
    if cnt <= 0:  # more ranks than problems: this rank only takes part in the barriers / reductions
        for _ in range(9):
            barrier()
        allmax(0.0); allmax(0.0); allmax(0.0); allsum(0.0); allsum(0.0); allmax(0.0)
        return None
    keys = ("Q", "c", "A", "l_A", "u_A", "l_x", "u_x")
    shapes = dict(Q=(cnt, n, n), c=(cnt, n), A=(cnt, m, n), l_A=(cnt, m), u_A=(cnt, m), l_x=(cnt, n), u_x=(cnt, n))
    pin = {k: z.pinned_empty(shapes[k]) for k in keys}
    for i in range(cnt):
        q = P.ineq_box(n, m, CFG4["seed0"] + lo + i, kind="shift")
        for k in keys:
            pin[k][i] = getattr(q, k)
    red = z.NORMAL if args.batch_reduction == "normal" else z.AUGMENTED
    # one handle per rank: the whole share runs in ONE persistent kernel (one CTA per problem at a time, problems
    # pulled from a ticket counter).  End to end the kernel is launched first and consumes problems as the copy
    # stream delivers them in `chunks` groups (ipmz_batch_solve_streamed): upload and solve overlap.
    sp = z.Problem(*(pin[k] for k in ("Q", "c", "A", "l_A", "u_A")), None, None, pin["l_x"], pin["u_x"])
    bs = z.BatchSolver(sp, cnt, z.Options(reduction=red, device=local))
    chunks = max(1, min(args.batch_chunks, cnt // 16))
    xout = z.pinned_empty((cnt, n))
    h2d = sum(int(v.nbytes) for v in pin.values())
    dev_ms, e2e_s, up_s, iters, conv = [], [], [], None, 0
    launches0 = z.launch_count()
    for rep in range(3):
        # end to end: H2D of the problem data + solve + D2H of x, all ranks at once
        barrier()
        t0 = time.perf_counter()
        bs.solve_streamed(chunks=chunks, per_problem=False)
        bs.x(xout)
        e2e_s.append(time.perf_counter() - t0)
        # the upload alone, all ranks at once: the H2D floor of the end-to-end number
        barrier()
        t0 = time.perf_counter()
        bs.upload()
        torch.cuda.synchronize()
        up_s.append(time.perf_counter() - t0)
        # device-resident: data already in HBM when the timed region starts
        barrier()
        dev_ms.append(bs.solve(per_problem=False)[1])
        if rep == 2:
            res = bs.results()
            iters = [r.iterations for r in res]
            conv = sum(1 for r in res if r.converged)
            xs = xout.copy()
            fs = [r.f for r in res[:16]]
    launches = z.launch_count() - launches0
    bs.close()
    t_dev = allmax(min(dev_ms[1:])) * 1e-3
    t_e2e = allmax(min(e2e_s[1:]))
    t_up = allmax(min(up_s[1:]))
    nconv = allsum(conv)
    it_sum = allsum(float(np.sum(iters)))
    it_max = int(allmax(float(np.max(iters))))
    floor = max(t_dev, t_up)
    out = {"metric": "ipm_solves_per_sec", "workload": "cfg4: %d independent QPs n=%d m=%d (ineq + box), sharded by "
           "problem index over %d GPU(s), no collective; one persistent kernel per GPU (k_ipm_batch: one CTA per "
           "problem, whole Mehrotra loop on the device, no host round trip)" % (total, n, m, world),
           "reduction": args.batch_reduction,
           "value": total / t_dev, "unit": "solves/s", "scaling": "strong",
           "e2e": {"value": total / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": int(xout.nbytes), "seconds": t_e2e, "upload_chunks": chunks,
                   "call": "ipmz_batch_solve_streamed (kernel launched first, problems consumed as the copy stream "
                           "delivers them) + ipmz_batch_get_x",
                   "upload_only_seconds": t_up, "h2d_gbs_per_rank": h2d / t_up * 1e-9,
                   "h2d_gbs_aggregate": world * h2d / t_up * 1e-9,
                   "floor_seconds": floor, "floor": "max(device time, concurrent upload time of all ranks)",
                   "e2e_over_floor": t_e2e / floor},
           "device_seconds": t_dev, "converged": int(nconv), "iterations_mean": it_sum / total,
           "iterations_max": it_max,
           "gpu_launches": launches,
           "gpu_launches_note": "3 repetitions x (streamed solve: 1 kernel; upload: transpose + initial point; "
                                "device-resident solve: 1 kernel)",
           "algorithmic": {"flops_per_problem_iteration": n * n * m + n ** 3 / 3.0,
                           "note": "condensed assembly n^2 m + LDL^T n^3/3 (FP64 DMMA); input bytes per problem %d"
                                   % ((n * n + m * n + 3 * n + 2 * m) * 8)}}
    out["tflops_device"] = it_sum * out["algorithmic"]["flops_per_problem_iteration"] / t_dev * 1e-12
    # roofline of the one kernel of the batched path: the contraction flops (assembly + LDL^T) against the live DMMA probe;
    # DRAM traffic per problem from the committed ncu capture of the same kernel (profiles/r02_traffic.json)
    try:
        peak = float(z.fp64_peak_tflops(local)) * world
    except Exception:
        peak = None
    traffic = None
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")))["k_ipm_batch"]
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["problems"]
    except Exception:
        traffic = None
    out["roofline"] = {"bound": "tensor", "kernel": "k_ipm_batch (whole Mehrotra solve of one QP per CTA)",
                       "achieved": out["tflops_device"], "peak": peak, "unit": "TFLOP/s",
                       "frac": out["tflops_device"] / peak if peak else None,
                       "traffic": traffic, "traffic_unit": "DRAM bytes per problem (profiles/r02_traffic.json); algorithmic "
                       "input bytes per problem %d: Q, M and the factor of the ~300 problems in flight (1.1 MB each) exceed "
                       "the 126 MB L2 and are re-streamed every phase" % ((n * n + m * n + 3 * n + 2 * m) * 8),
                       "launches_per_step": 1}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu, gate_its = batched_cpu_baseline()
            out["cpu_baseline"] = cpu
            import oracle_lib as ol
            import problems as PO
            dfs = []
            for i in range(16):
                q = PO.ineq_box(n, m, CFG4["seed0"] + i, kind="shift")
                tr = ol.port_solve(q, steps=False)
                dfs.append(abs(fs[i] - tr.f[tr.iterations]) / max(1.0, abs(tr.f[tr.iterations])))
                gate_its[i] = tr.iterations
            out["parity_gate"] = {"problems": 16, "oracle": "port (bit-for-bit the reference on every golden case)",
                                  "iterations_equal": [int(v) for v in iters[:16]] == [int(v) for v in gate_its],
                                  "max_rel_df": float(max(dfs)), "tolerance": 1e-8,
                                  "pass": bool([int(v) for v in iters[:16]] == [int(v) for v in gate_its] and
                                               max(dfs) <= 1e-8)}
        except Exception as e:  # reported, never substituted
            out["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": 0, "kind": "reference",
                                   "sample": "failed: %r" % (e,)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sample-n", type=int, default=3072, help="CPU sample size of the reference arm")
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--batch-reduction", default="normal", choices=["normal", "augmented"])
    ap.add_argument("--batch-chunks", type=int, default=64, help="upload groups of the streamed end-to-end batch solve")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1 / cfg2 / cfg5 end-to-end solves")
    ap.add_argument("--quick", action="store_true", help="small sizes (debug only; not a valid bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
