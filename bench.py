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


def workload_name(K, batch, config="RocketQuat"):
    veh = "Starship parameters (model.info:1-105), " if config == "RocketQuatStarship" else ""
    return (f"RocketQuat 6-DoF landing, {veh}free-final-time SC, K={K}, batch={batch} perturbed initial states per GPU "
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


def cpu_reference_run(n_inst, nthreads, first=0, K=K_BENCH, config="RocketQuat"):
    """times the oracle (CPU restatement of the reference path) on n_inst perturbed instances, one instance per thread"""
    import ctypes as C
    import orc_py as O
    O.build()
    p, rpy = O.starship() if config == "RocketQuatStarship" else O.falcon9()
    cfg = O.sc_config(K=K)
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
        cpu_reference_run(min(n_inst, cores), cores, K=args.K, config=args.config)
    tot_it, tot_t = 0, 0.0
    for s in range(args.steps):
        it, dt = cpu_reference_run(n_inst, cores, first=s * n_inst, K=args.K, config=args.config)
        tot_it += it; tot_t += dt
    value = tot_it / tot_t
    sample = f"{n_inst} perturbed RocketQuat K={args.K} instances per step (full SC solve, max 15 iterations), one instance per thread"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.K, args.batch, args.config), "K": args.K, "batch_per_gpu": args.batch,
                       "reference_sample_per_step": n_inst,
                       "note": "CPU restatement of the reference path (oracle/, ECOS-equivalent IPM), not ECOS itself; each step solves a bounded "
                               "sample of the same perturbed instances, one instance per host thread"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


K1_JACOBIAN = 2


def make_engine(S, args, rank, local, world, warm, solver):
    name = args.config
    model, params, x_init, x_final, cfg = S.load_model(name, K=args.K, algorithm=args.algorithm)
    cfg.ipm.warm = warm
    cfg.ipm.stalled_step = int(os.environ.get("SCPP_STALLED_STEP", "0"))      # experiment knob (scpp_b200.h): default off
    cfg.solver = solver
    cfg.jacobian = int(os.environ.get("SCPP_JACOBIAN", str(K1_JACOBIAN)))      # K1 path (scpp_b200.h): 2 = hand-derived Jacobian shared by the columns of an interval
    cfg.ipm_slice = int(os.environ.get("SCPP_SLICE", "1"))    # solver 0: interior-point iterations per K2 launch (0: lock-step outer iterations)
    rpy = np.deg2rad([70.0, 0.0, 0.0]) if name == "RocketQuatStarship" else np.deg2rad([-20.0, 20.0, 0.0])      # rpy_init of configs/<name>/model.info
    xi = S.perturbed_initial_states(x_init, rpy, args.batch, first=rank * args.batch)
    eng = S.SCAlgorithm(model, params, cfg, args.batch, device=local)
    return eng, cfg, xi, x_final


def device_steps(eng, steps):
    """`steps` solves with the inputs resident in HBM; CUDA-event times and counters summed from the engine"""
    out = dict(ms=[], disc=0.0, socp=0.0, launches=0, inst_iters=0, outer=0, rounds=0, inst_rounds=0)
    for _ in range(steps):
        eng.solve()
        t = eng.last_timing(); r = eng.last_rounds()
        out["ms"].append(t["ms_total"]); out["disc"] += t["ms_discretize"]; out["socp"] += t["ms_socp"]
        out["launches"] += t["kernel_launches"]; out["inst_iters"] += t["instance_iterations"]; out["outer"] += t["outer_iterations"]
        out["rounds"] += r["rounds"]; out["inst_rounds"] += r["instance_rounds"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU (weak scaling)")
    ap.add_argument("--K", type=int, default=K_BENCH)
    ap.add_argument("--config", default="RocketQuat", choices=["RocketQuat", "RocketQuatStarship"], help="parameter set under configs/ (Starship: BASELINE configs[4], use --K 100 --batch 4096)")
    ap.add_argument("--solver", type=int, default=int(os.environ.get("SCPP_SOLVER", "0")), choices=[0, 1, 2],
                    help="K2 mapping: 0 warp per instance in rounds, 1 CTA per instance with the factor in shared memory, 2 = 0 with the tail of a solve on 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cold-start and other-solver measurements (sweeps)")
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
    if args.algorithm == "SCvx":
        args.no_cpu_baseline = True

    # main measurement: interior warm start of the sub-problems (engine knob, same optimum; parity-tested in tests/test_gpu_parity.py).
    # value_cold (below) is the like-for-like figure against the CPU arm: ECOS-style cold start of every sub-problem, the library default.
    warm = float(os.environ.get("SCPP_WARM", "0.995"))
    eng, cfg, xi, x_final = make_engine(S, args, rank, local, world, warm, args.solver)
    n_local = args.batch
    if world > 1:
        from scpp_b200.sharding import broadcast_unique_id
        eng.comm_init(world, rank, broadcast_unique_id(dist, S.comm_unique_id, rank))

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
    t_wall0 = time.perf_counter()
    D = device_steps(eng, args.steps)
    sync_all()
    wall_dev = time.perf_counter() - t_wall0
    info = eng.get_info()
    dev_ms, disc_ms, socp_ms, launches, inst_iters, outer, rounds, inst_rounds = D["ms"], D["disc"], D["socp"], D["launches"], D["inst_iters"], D["outer"], D["rounds"], D["inst_rounds"]
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
    dev_bytes = eng.device_bytes()
    eng.close()

    # ---- the same workload with cold-started sub-problems, and with the other K2 mapping (single-GPU figures, rank 0's shard on every rank)
    extras = {}
    if not args.no_extras:
        for key, w, sv in (("cold", 0.0, args.solver), ("other_solver", warm, 1 if args.solver != 1 else 0)):
            try:
                e2, c2, _, _ = make_engine(S, args, rank, local, world, w, sv)
            except S.ScppError as ex:
                extras[key] = {"unavailable": str(ex)}
                continue
            e2.set_boundary_states(xi, x_final)
            e2.solve()
            sync_all()
            d2 = device_steps(e2, max(1, min(2, args.steps)))
            sync_all()
            extras[key] = {"ms": float(np.sum(d2["ms"])), "inst_iters": d2["inst_iters"], "steps": len(d2["ms"])}
            e2.close()

    dev_total_ms = float(np.sum(dev_ms))
    stats = np.array([dev_total_ms, wall_e2e, wall_dev] + [extras.get(k, {}).get("ms", 0.0) for k in ("cold", "other_solver")], dtype=np.float64)
    sums = np.array([inst_iters, e2e_iters, launches] + [extras.get(k, {}).get("inst_iters", 0) for k in ("cold", "other_solver")], dtype=np.float64)
    if dist is not None:
        from scpp_b200.sharding import reduce_timing
        stats, sums = reduce_timing(dist, stats, sums, device="cuda")      # max over ranks of the times, sum over ranks of the counts
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        value = sums[0] / (stats[0] * 1e-3)
        e2e_value = sums[1] / stats[1]
        kname = "k_solve" if args.solver == 0 else "k_solve_cta"
        # roofline of the dominant kernel: algorithmic bytes (SURVEY §8d share of K2: tiles + linearisation point in, X,U out = 146 112 B per
        # instance-iteration at K = 50) of the instance-iterations finished in the timed region, spread over its launches (solver 0: one launch
        # per ROUND = one interior-point iteration of every unfinished instance; solver 1: one launch per outer iteration), divided by the mean
        # launch duration (CUDA events on the engine stream, rank 0's launches)
        socp_launches = max(1, rounds)
        bytes_per_launch = algorithmic_bytes(args.K, "k_solve") * (inst_iters / socp_launches)
        ach = bytes_per_launch / (socp_ms / socp_launches * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": None, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "algorithmic_bytes_per_instance_iteration": algorithmic_bytes(args.K, "k_solve"),
                "launches_per_step": socp_launches / args.steps, "ms_per_launch": socp_ms / socp_launches,
                "share_of_step": socp_ms / max(1e-9, float(np.sum(dev_ms)))}
        # the whole iteration against SURVEY §8(d)'s figure (discretize -> solve, Jacobian tensors round-trip HBM once: 285 024 B at K = 50)
        whole = algorithmic_bytes(args.K, "all") * inst_iters / (float(np.sum(dev_ms)) * 1e-3) / 1e9
        roof_it = {"bound": "hbm", "scope": "k_discretize + K2 (whole iteration, SURVEY §8d)", "algorithmic_bytes_per_instance_iteration": algorithmic_bytes(args.K, "all"),
                   "achieved": whole, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": whole / peaks["hbm_gbs"]}
        ipm_iters = float(info[:, :, 5].sum())      # interior-point iterations of the last step (rank 0's shard)
        prof = {}
        for nm in ("r02_k2_traffic.json", "fp64_peak.json", "r02_fp64_flops.json"):
            pth = os.path.join(ROOT, "profiles", nm)
            if os.path.exists(pth):
                with open(pth) as f:
                    prof[nm] = json.load(f)
        tr = prof.get("r02_k2_traffic.json")
        if tr and tr.get("K") == args.K and kname in tr:      # ncu dram bytes per instance and interior-point iteration x the iterations one launch advances
            per_launch_ipm = ipm_iters / max(1.0, socp_launches / args.steps)
            roof["traffic"] = tr[kname]["dram_bytes_per_instance_ipm_iteration"] * per_launch_ipm
            roof["traffic_gbs"] = roof["traffic"] / (socp_ms / socp_launches * 1e-3) / 1e9
            roof["traffic_frac_of_peak"] = roof["traffic_gbs"] / peaks["hbm_gbs"]
            roof["traffic_source"] = tr[kname].get("source", "profiles/r02_k2_traffic.json (ncu dram__bytes_read+write of one launch)")
        fp64 = prof.get("fp64_peak.json")
        # the bound that actually binds: FP64 issue.  Counted flops (2 per DFMA, 1 per DADD/DMUL, 512 per DMMA.8x8x4 warp
        # instruction, from the ncu SASS page of one launch, profiles/r02_fp64_flops.json) x the work of the last step / its device time
        roof_fp64 = None
        fl = prof.get("r02_fp64_flops.json")
        if fl and fp64 and fl.get("K") == args.K and kname in fl:
            flops_step = fl[kname]["flop_per_instance_ipm_iteration"] * ipm_iters + fl["k_discretize_shared" if cfg.jacobian == 2 and "k_discretize_shared" in fl else "k_discretize"]["flop_per_instance_discretization"] * (inst_iters / args.steps)
            tf = flops_step / (float(np.mean(dev_ms)) * 1e-3) / 1e12
            roof_fp64 = {"bound": "fp64", "achieved": tf, "peak": fp64["dfma_tflops"], "unit": "TFLOP/s", "frac": tf / fp64["dfma_tflops"],
                         "flop_per_instance_iteration": flops_step / max(1.0, inst_iters / args.steps),
                         "peak_source": "profiles/fp64_peak.json (tools/fp64_peak.cu, measured DFMA rate on this pool's B200)",
                         "count_source": "profiles/r02_fp64_flops.json"}
        # what the counted iterations are: the SC loop of the shipped RocketQuat weights stalls (trajectory stops moving while the virtual control
        # stays above nu_tol: DESIGN.md); an instance-iteration whose trust-region radii sum below delta_tol re-solves an unchanged sub-problem
        its = res["iterations"]
        moved = sum(int((info[i, :n, 1] > cfg.delta_tol).sum()) for i, n in enumerate(its)) if args.algorithm == "SC" else int(its.sum())
        st = np.concatenate([info[i, :n, 6] for i, n in enumerate(its)]).astype(int)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": stats[0] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.K, n_local, args.config) if args.algorithm == "SC" else workload_name(args.K, n_local, args.config).replace("free-final-time SC", "fixed-final-time SCvx").replace("max_iterations=15", f"max_iterations={cfg.max_iterations}"),
                           "batch_per_gpu": n_local, "global_batch": n_local * world, "K": args.K, "parallelism": f"instances sharded x{world}",
                           "l2": f"working set {dev_bytes / 1e6:.0f} MB per GPU >> 126 MB L2 (no flush needed)",
                           "k1_jacobian": {0: "hand-derived, one thread per column", 1: "dual numbers over the flow map, one thread per column (library default)",
                                           2: "hand-derived, linearisation shared by the columns of an interval (k_discretize_shared; parity-tested against 0 and 1)"}[cfg.jacobian],
                           "integrator": (f"RK4 x {cfg.nsub}" if cfg.nsub > 0 else f"RK4 x {-cfg.nsub} and x {-2 * cfg.nsub}, Richardson-extrapolated") + " (reference RKF78 x 5)",
                           "ipm_tol": cfg.ipm.feastol, "ipm_warm": warm, "ipm_slice": cfg.ipm_slice, "k2_solver": args.solver,
                           "note": "value: sub-problems warm-started from the previous interior point (engine knob, same optimum); value_cold: every sub-problem "
                                   "cold-started like ECOS does (library default) = the like-for-like figure against cpu_baseline / --impl reference"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": 1e3 * stats[1] / args.steps},
                "gpu_launches": int(sums[2]),
                "kernel_ms": {"k_discretize": disc_ms / args.steps, kname: socp_ms / args.steps, "step_total": float(np.mean(dev_ms))},
                "instance_iterations_per_step": inst_iters / args.steps, "outer_iterations_per_step": outer / args.steps,
                "interior_point_iterations_per_instance_iteration": ipm_iters / max(1.0, inst_iters / args.steps),
                "iterations_split": {"trajectory_moved": moved, "stalled": int(its.sum()) - moved, "criterion": "sum of the trust-region radii of the iteration > delta_tol (SC.info)"},
                "subproblem_exit_status": {"optimal": float((st == 0).mean()), "reduced_accuracy": float((st == 3).mean()), "failed": float(((st == 1) | (st == 2)).mean())},
                "converged_fraction": float((res["flags"] == 1).mean()), "failed_fraction": float((res["flags"] == 2).mean()),
                "roofline": roof, "roofline_iteration": roof_it, "roofline_fp64": roof_fp64, "fp64_peak": fp64, "clocks": clocks}
        if "cold" in extras and "ms" in extras["cold"]:
            line["value_cold"] = sums[3] / (stats[3] * 1e-3)
            line["warm_start_gain"] = value / line["value_cold"]
        if "other_solver" in extras:
            line["value_other_solver"] = {"k2_solver": 1 if args.solver != 1 else 0, "value": (sums[4] / (stats[4] * 1e-3)) if "ms" in extras["other_solver"] else None,
                                          "note": extras["other_solver"].get("unavailable", "same workload and warm start, the other K2 mapping")}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = max(1, min(cores, 128))
            it, dt = cpu_reference_run(n_s, cores, K=args.K, config=args.config)
            line["cpu_baseline"] = {"value": it / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_s} of the same perturbed instances (full SC solve, cold-started sub-problems), one instance per thread, {dt:.1f} s wall"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
