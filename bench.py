#!/usr/bin/env python
"""bench.py — BASELINE.json metric: SC(vx) iterations/sec, batched RocketQuat K=50, at 1/2/4/8 B200 vs the CPU path.

One "step" = one pass of the hot path over one batch: SCAlgorithm::solve() for `--batch` perturbed RocketQuat instances
per GPU (every outer iteration = K1 multiple shooting + K2 SOCP solve + convergence logic).  Unit of work = one
instance-iteration (one pass of SCAlgorithm::iterate for one instance).

  python bench.py --gpus N --steps K --warmup W            our CUDA engine (torchrun launches one rank per GPU for N > 1)
  python bench.py --impl reference ...                     the CPU restatement of the reference path (oracle/) on host cores

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM, CUDA events);
`e2e` = the same metric through the public API with host buffers (H2D of the boundary states and D2H of the
trajectories inside the timed region).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "sc_instance_iterations_per_sec"
UNIT = "instance-iterations/s"
K_BENCH = 50
NX, NU = 14, 4


def algorithmic_bytes(K, which):
    """per instance-iteration, unfused discretize -> solve: 8*[2(K-1)(nx^2+2 nx nu+2 nx) + 3K(nx+nu)]"""
    dd = 8 * (K - 1) * (NX * NX + 2 * NX * NU + 2 * NX)
    tr = 8 * K * (NX + NU)
    if which == "k_discretize":
        return dd + tr            # read X,U ; write A,B,C,s,z
    if which == "k_solve":
        return dd + 2 * tr        # read A..z, read linearisation point ; write new X,U
    return 2 * dd + 3 * tr


def workload_name(K, batch):
    return (f"RocketQuat 6-DoF landing, free-final-time SC, K={K}, batch={batch} perturbed initial states per GPU "
            f"(reference Monte-Carlo recipe, seed 0x5C99), max_iterations=15")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(n_inst, nthreads, first=0):
    """times the oracle (CPU restatement of the reference path) on n_inst perturbed instances, one instance per thread"""
    import ctypes as C
    import orc_py as O
    O.build()
    p, rpy = O.falcon9()
    cfg = O.sc_config(K=K_BENCH)
    arr = (O.RQParams * n_inst)()
    for i in range(n_inst):
        arr[i] = O.rq_perturb(p, rpy, 0x5C99, first + i)
    iters = np.zeros(n_inst, np.int32)
    L = O.lib()
    L.orc_sc_solve_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    t0 = time.perf_counter()
    total = L.orc_sc_solve_batch(O.ROCKETQUAT, n_inst, C.byref(arr), C.sizeof(O.RQParams), C.byref(cfg), iters.ctypes.data_as(C.c_void_p), None, None, None, None, nthreads)
    dt = time.perf_counter() - t0
    return total, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import orc_py as O
    cores = os.cpu_count() or 1
    n_inst = max(1, min(cores, 128))
    for _ in range(args.warmup):
        cpu_reference_run(min(n_inst, cores), cores)
    tot_it, tot_t = 0, 0.0
    for s in range(args.steps):
        it, dt = cpu_reference_run(n_inst, cores, first=s * n_inst)
        tot_it += it; tot_t += dt
    value = tot_it / tot_t
    sample = f"{n_inst} perturbed RocketQuat K={K_BENCH} instances per step (full SC solve, max 15 iterations), one instance per thread"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(K_BENCH, args.batch), "K": K_BENCH, "batch_per_gpu": args.batch,
                       "reference_sample_per_step": n_inst,
                       "note": "CPU restatement of the reference path (oracle/, ECOS-equivalent IPM), not ECOS itself; each step solves a bounded "
                               "sample of the same perturbed instances, one instance per host thread"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU (weak scaling)")
    ap.add_argument("--K", type=int, default=K_BENCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--algorithm", default="SC", choices=["SC", "SCvx"], help="SC (default, the measured path) or the SCvx variant (no CPU baseline arm)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import scpp_b200 as S
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    if S.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")

    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=args.K, algorithm=args.algorithm)
    if args.algorithm == "SCvx":
        args.no_cpu_baseline = True
    # interior warm start of the sub-problems (engine knob, same optimum; parity-tested in tests/test_gpu_parity.py);
    # SCPP_WARM=0 gives ECOS-style cold starts
    cfg.ipm.warm = float(os.environ.get("SCPP_WARM", "0.995"))
    cfg.solver = int(os.environ.get("SCPP_SOLVER", "1"))      # 1: CTA-per-instance solver (round 2); 0: warp-per-instance rounds (round 1)
    cfg.ipm_slice = int(os.environ.get("SCPP_SLICE", "1"))    # interior-point iterations per K2 launch (0: lock-step outer iterations)
    rpy = np.deg2rad([-20.0, 20.0, 0.0])      # rpy_init of configs/RocketQuat/model.info
    n_local = args.batch
    xi = S.perturbed_initial_states(x_init, rpy, n_local, first=rank * n_local)
    eng = S.SCAlgorithm(model, params, cfg, n_local, device=local)
    if world > 1:
        obj = [S.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        eng.comm_init(world, rank, obj[0])

    def sync_all():
        if dist is not None:
            dist.barrier()

    # ---- warm-up
    eng.set_boundary_states(xi, x_final)
    for _ in range(args.warmup):
        eng.solve()
    res = eng.get_solution()

    # ---- device-resident timing: inputs already in HBM, CUDA events inside the engine, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    dev_ms, disc_ms, socp_ms, launches, inst_iters, outer, rounds, inst_rounds = [], 0.0, 0.0, 0, 0, 0, 0, 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.solve()
        t = eng.last_timing()
        dev_ms.append(t["ms_total"]); disc_ms += t["ms_discretize"]; socp_ms += t["ms_socp"]
        launches += t["kernel_launches"]; inst_iters += t["instance_iterations"]; outer += t["outer_iterations"]
        r = eng.last_rounds(); rounds += r["rounds"]; inst_rounds += r["instance_rounds"]
    sync_all()
    wall_dev = time.perf_counter() - t_wall0
    # ---- end-to-end through the public API with host buffers
    h2d = 2 * n_local * NX * 8
    d2h = n_local * (args.K * (NX + NU) * 8 + 8 + 4 + 4)
    sync_all()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        eng.set_boundary_states(xi, x_final)
        eng.solve()
        res = eng.get_solution(res)
        e2e_iters += int(res["iterations"].sum())
    sync_all()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    dev_total_ms = float(np.sum(dev_ms))
    stats = np.array([dev_total_ms, wall_e2e, wall_dev], dtype=np.float64)
    sums = np.array([inst_iters, e2e_iters, launches], dtype=np.float64)
    if dist is not None:
        import torch
        tmax = torch.tensor(stats, device="cuda"); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = torch.tensor(sums, device="cuda"); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        stats, sums = tmax.cpu().numpy(), tsum.cpu().numpy()
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        value = sums[0] / (stats[0] * 1e-3)
        e2e_value = sums[1] / stats[1]
        # roofline of the dominant kernel (k_solve; one launch = one ROUND = one interior-point iteration of every unfinished
        # instance): algorithmic bytes of the instance-iterations finished in the timed region, spread over its launches, divided
        # by the mean launch duration (CUDA events on the engine stream, rank 0's launches)
        socp_launches = max(1, rounds)
        bytes_per_launch = algorithmic_bytes(args.K, "k_solve") * (inst_iters / socp_launches)
        ach = bytes_per_launch / (socp_ms / socp_launches * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_solve", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": None, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "algorithmic_bytes_per_instance_iteration": algorithmic_bytes(args.K, "k_solve"),
                "launches_per_step": socp_launches / args.steps, "ms_per_launch": socp_ms / socp_launches,
                "share_of_step": socp_ms / max(1e-9, float(np.sum(dev_ms)))}
        whole = algorithmic_bytes(args.K, "all") * inst_iters / (float(np.sum(dev_ms)) * 1e-3) / 1e9
        roof_it = {"bound": "hbm", "scope": "k_discretize + k_solve (whole iteration, SURVEY §8d: 285024 B at K=50)", "achieved": whole,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": whole / peaks["hbm_gbs"]}
        tpath = os.path.join(ROOT, "profiles", "k_solve_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("K") == args.K:      # ncu dram bytes of one launch per instance it advanced x mean instances per launch here
                roof["traffic"] = tj["dram_bytes_per_instance_round"] * (inst_rounds / socp_launches)
                roof["traffic_gbs"] = roof["traffic"] / (socp_ms / socp_launches * 1e-3) / 1e9
                roof["traffic_frac_of_peak"] = roof["traffic_gbs"] / peaks["hbm_gbs"]
                roof["traffic_source"] = "profiles/k_solve_traffic.json (ncu dram__bytes_read+write of one full-batch launch)"
        prof = os.path.join(ROOT, "profiles", "fp64_peak.json")
        fp64 = None
        if os.path.exists(prof):
            with open(prof) as f:
                fp64 = json.load(f)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": stats[0] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.K, n_local) if args.algorithm == "SC" else workload_name(args.K, n_local).replace("free-final-time SC", "fixed-final-time SCvx").replace("max_iterations=15", f"max_iterations={cfg.max_iterations}"),
                           "batch_per_gpu": n_local, "global_batch": n_local * world, "K": args.K, "parallelism": f"instances sharded x{world}",
                           "l2": f"working set {eng.device_bytes() / 1e6:.0f} MB per GPU >> 126 MB L2 (no flush needed)",
                           "integrator": (f"RK4 x {cfg.nsub}" if cfg.nsub > 0 else f"RK4 x {-cfg.nsub} and x {-2 * cfg.nsub}, Richardson-extrapolated") + " (reference RKF78 x 5)", "ipm_tol": cfg.ipm.feastol, "ipm_warm": cfg.ipm.warm, "ipm_slice": cfg.ipm_slice},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": 1e3 * stats[1] / args.steps},
                "gpu_launches": int(sums[2]),
                "kernel_ms": {"k_discretize": disc_ms / args.steps, "k_solve": socp_ms / args.steps, "step_total": float(np.mean(dev_ms))},
                "instance_iterations_per_step": inst_iters / args.steps, "outer_iterations_per_step": outer / args.steps,
                "converged_fraction": float((res["flags"] == 1).mean()), "failed_fraction": float((res["flags"] == 2).mean()),
                "roofline": roof, "roofline_iteration": roof_it, "fp64_peak": fp64, "clocks": clocks}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = max(1, min(cores, 128))
            it, dt = cpu_reference_run(n_s, cores)
            line["cpu_baseline"] = {"value": it / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_s} of the same perturbed instances (full SC solve), one instance per thread, {dt:.1f} s wall"}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
