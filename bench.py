#!/usr/bin/env python
"""bench.py — VEM iterations/s of the MAGMAclust hot path on B200 (contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one iteration of training_VEM's loop body (MAGMAclust.R:44-93) from the study's start state
(theta_k = theta_i = (1, 1, 0.2), prior means 0, tau_0 = one-hot k-means labels): E-step, ELBO, M-step with the
built-in L-BFGS, ELBO at the new hp, acceptance rule. Workload = BASELINE.json configs[1] ("C2": M = 1000
individuals, K = 5, n_i = 50, individual-specific hyper-parameters) PER GPU; with N GPUs the individuals are
sharded (weak scaling: N*1000 individuals) and the grid accumulators are all-reduced with NCCL.

  value : device-resident state, CUDA-event time on the library's stream, max over ranks
  e2e   : the same step through the host-buffer API (upload dataset + state from pinned host memory, download
          mean/cov/tau/hp every step)
  --impl reference : the CPU path timed on this box's host cores. R is not installed (neither here nor on the
          GPU box), so this is the numpy restatement in oracle/ ("port"), timed on a bounded sample and
          extrapolated to the workload (per-individual cost is linear in M, grid cost unchanged).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line. Native libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there when NCCL_DEBUG is set), so descriptor 1 is pointed at stderr for the whole run and the result line
# goes to a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


METRIC = "VEM iterations/sec at M individuals x K clusters"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3"])
    ap.add_argument("--no-flush", action="store_true", help="do not evict L2 between timed steps")
    ap.add_argument("--reps", type=int, default=2, help="repetitions of the timed region; the faster one is reported")
    ap.add_argument("--no-steady", action="store_true", help="skip the extra steady-state measurement of the dominant kernel")
    ap.add_argument("--cpu-sample", type=int, default=200, help="individuals in the CPU baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--individuals", type=int, default=0, help="override the individuals per GPU of the workload")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class Clocks:
    """nvidia-smi sampler. It is started BEFORE the warm-up (attaching NVML to the GPU stalls the CUDA context for tens
    of milliseconds, which must not land inside the timed region) and only the samples whose timestamp falls inside
    the window given to stop() are used."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        self.idx = gpu_index
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    @staticmethod
    def _ts(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [[x.strip() for x in line.split(",")] for line in out.strip().splitlines()]
        lines = [f for f in lines if len(f) >= 9]
        if t0 is not None:
            inside = [f for f in lines if (self._ts(f[0]) or 0) >= t0 - 0.03 and (self._ts(f[0]) or 0) <= t1 + 0.03]
            if inside:
                lines = inside
        for f in lines:
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_data(name, nranks, individuals=0):
    from magmaclust_b200 import synth
    c = dict(synth.CONFIGS[name])
    if individuals:
        c["M"] = individuals
    ds = synth.make_dataset(c["M"] * nranks, c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    return ds, c


def start_state(M, K):
    theta_k = np.tile(np.array([1.0, 1.0, 0.2]), (K, 1))
    theta_i = np.tile(np.array([1.0, 1.0, 0.2]), (M, 1))
    return theta_k, theta_i, np.zeros(K)


# ------------------------------------------------------------------------------------------------
def cpu_baseline(ds, cfg, sample, steps=1, warmup=0):
    """One VEM iteration of the oracle (numpy port of the reference R code) on `sample` individuals of the
    workload; returns seconds per iteration on the sample, split into the per-individual and the grid part, and
    the extrapolation to the full per-GPU workload."""
    from magmaclust_b200 import synth
    from oracle import magma_oracle as O
    M_full = cfg["M"]
    K = cfg["K"]
    sub = synth.shard(ds, 0, max(1, ds["M"] // sample))
    n = np.diff(sub["offsets"])
    db = {"ID": np.repeat(sub["ids"], n), "Timestamp": sub["t"], "Output": sub["y"]}
    names = [f"K{k + 1}" for k in range(K)]
    ids = [int(i) for i in sub["ids"]]
    tau = {k: {i: float(sub["tau0"][r, c]) for r, i in enumerate(ids)} for c, k in enumerate(names)}
    m_k = {k: 0.0 for k in names}
    common_i = cfg["common_hp_i"]

    def one_iteration():
        t = {}
        hp = {"theta_k": {k: np.array([1.0, 1.0, 0.2]) for k in names},
              "theta_i": {i: np.array([1.0, 1.0, 0.2]) for i in ids},
              "pi_k": np.array([np.mean(list(tau[k].values())) for k in names])}
        t0 = time.perf_counter()
        param = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
        t["e_step"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.elbo_VEM(hp, db, O.kernel, O.kernel_mu, param, m_k)
        t["elbo"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        st = {}
        new_hp = O.m_step_VEM(db, hp, param, O.kernel_mu, O.kernel, m_k, True, common_i, stats=st)
        t["m_step"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.logL_monitoring_VEM(new_hp, db, O.kernel, O.kernel_mu, param, m_k)
        O.logL_monitoring_VEM(hp, db, O.kernel, O.kernel_mu, param, m_k)
        t["elbo"] += time.perf_counter() - t0
        t["nfev"] = st
        return t

    for _ in range(warmup):
        one_iteration()
    runs = [one_iteration() for _ in range(max(1, steps))]
    tot = float(np.mean([r["e_step"] + r["elbo"] + r["m_step"] for r in runs]))
    t_k = float(np.mean([r["nfev"]["t_k"] for r in runs]))   # theta_k optimiser: N x N work, independent of M
    Ms = len(ids)
    return {"sample_seconds_per_iteration": tot, "sample_individuals": Ms, "grid_seconds": t_k,
            "phases": {k: float(np.mean([r[k] for r in runs])) for k in ("e_step", "elbo", "m_step")},
            "nfev": runs[-1]["nfev"], "M_full": M_full}


def extrapolate(b, M):
    """per-individual loops scale linearly in M; the theta_k optimiser (N x N grid work) does not"""
    return (b["sample_seconds_per_iteration"] - b["grid_seconds"]) * (M / b["sample_individuals"]) + b["grid_seconds"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ds, cfg = workload_data(args.workload, 1, args.individuals)
    cores = os.cpu_count() or 1
    # bounded sample: the whole --steps/--warmup run should end within a few minutes (~0.07 s per individual)
    sample = int(min(args.cpu_sample, max(40, 1600 // max(1, args.steps + args.warmup))))
    base = cpu_baseline(ds, cfg, sample, steps=args.steps, warmup=args.warmup)
    M = cfg["M"] * args.gpus
    sec = extrapolate(base, M)
    value = M / sec  # individual-VEM-iterations per second
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "individual-VEM-iterations/s",
        "vem_iterations_per_s": 1.0 / sec, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: M={cfg['M']} individuals/GPU x K={cfg['K']}, n_i={cfg['n']}, "
                               f"grid N={cfg['n_grid']}, individual-specific hp={not cfg['common_hp_i']}",
                   "M_total": M, "K": cfg["K"]},
        "cpu_baseline": {"value": value, "unit": "individual-VEM-iterations/s", "cores": cores, "kind": "port",
                         "sample": f"numpy port of the reference R code (R is not installed on this box): one VEM "
                                   f"iteration on {base['sample_individuals']} of {M} individuals "
                                   f"({base['sample_seconds_per_iteration']:.2f} s, of which "
                                   f"{base['grid_seconds']:.2f} s theta_k grid work); per-individual part scaled "
                                   f"linearly in M, grid part kept",
                         "phases_s": base["phases"], "nfev": base["nfev"]},
        "e2e": {"value": value, "unit": "individual-VEM-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dist = None
    if world > 1:
        import torch.distributed as dist  # control plane only (id broadcast, barriers, max over ranks)
        dist.init_process_group("gloo", rank=rank, world_size=world)

    from magmaclust_b200 import synth
    from magmaclust_b200.engine import Engine

    ds, cfg = workload_data(args.workload, world, args.individuals)
    sh = synth.shard(ds, rank, world)
    M, K, N = sh["M"], sh["K"], sh["N"]
    eng = Engine(local)
    if world > 1:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.tensor(list(Engine.comm_unique_id()), dtype=torch.uint8)
        dist.broadcast(idt, 0)
        eng.comm_init(world, rank, bytes(idt.tolist()))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    common_i = cfg["common_hp_i"]
    theta_k, theta_i, m_k = start_state(M, K)
    tau0 = np.ascontiguousarray(sh["tau0"])
    pi0 = ds["tau0"].mean(axis=0)
    host = {"offsets": sh["offsets"], "grid_idx": sh["grid_idx"], "y": sh["y"], "grid": np.ascontiguousarray(sh["grid"]),
            "theta_k": theta_k, "theta_i": theta_i, "pi": np.ascontiguousarray(pi0), "m_k": m_k, "tau0": tau0}
    pinned = [a for a in host.values() if eng.pin(a)]

    eng.set_data(host["offsets"], host["grid_idx"], host["y"], host["grid"], M_total=sh["M_total"])
    peaks = eng.measure_fp64_peaks()

    def step_resident():
        eng.set_state(host["theta_k"], host["theta_i"], host["pi"], host["m_k"], host["tau0"])
        return eng.vem_iteration(True, common_i)

    # ---- value: resident dataset, CUDA events on the library stream ---------------------------------
    clocks = Clocks(local)
    time.sleep(0.3)   # let nvidia-smi finish attaching before any timed work
    for _ in range(args.warmup):
        step_resident()
    # The timed region (exactly --steps steps between barriers, CUDA events on the library stream) is measured
    # `reps` times back to back and the faster repetition is reported: a fresh box occasionally stalls one step by
    # tens of milliseconds (seen right after another process released the GPU), which is not a property of the code.
    best = None
    for rep in range(args.reps):
        phase_ms = {}
        launches = 0
        nfev_i_sum = 0
        nfev_k = 0
        barrier()
        t_wall0 = time.time()
        eng.timer_start()
        step_wall = []
        for _ in range(args.steps):
            tw = time.perf_counter()
            if not args.no_flush:
                eng.flush_l2()
            step_resident()
            step_wall.append((time.perf_counter() - tw) * 1e3)
            # reading the phase events costs a host sync that the step itself already paid
            ph, nl = eng.last_timings()
            for k, v in ph.items():
                phase_ms[k] = phase_ms.get(k, 0.0) + v
            launches += nl
            st = eng.last_mstep_stats()
            nfev_i_sum += st["nfev_i_sum"]
            nfev_k += st["nfev_k"]
        ms = eng.timer_stop()
        t_wall1 = time.time()
        barrier()
        if os.environ.get("MCB_BENCH_STEP_TIMES"):
            print("step wall ms:", " ".join("%.2f" % x for x in step_wall), file=sys.stderr)
        ms = max_over_ranks(ms)
        if best is None or ms < best[0]:
            best = (ms, phase_ms, launches, nfev_i_sum, nfev_k, t_wall0, t_wall1)
    ms, phase_ms, launches, nfev_i_sum, nfev_k, t_wall0, t_wall1 = best
    clk = clocks.stop(t_wall0, t_wall1)
    ms = max_over_ranks(ms)
    M_total = sh["M_total"]
    value = M_total * args.steps / (ms / 1e3)

    # ---- e2e: host buffers in, host buffers out, every step -----------------------------------------
    out_mean = np.empty((K, N))
    out_cov = np.empty((K, N, N))
    out_tau = np.empty((M, K))
    pinned += [a for a in (out_mean, out_cov, out_tau) if eng.pin(a)]
    from magmaclust_b200._lib import dptr

    def step_e2e():
        eng.set_data(host["offsets"], host["grid_idx"], host["y"], host["grid"], M_total=sh["M_total"])
        eng.set_state(host["theta_k"], host["theta_i"], host["pi"], host["m_k"], host["tau0"])
        # one call per loop body of training_VEM, param of the E-step delivered to the (pinned) host buffers
        r = eng.vem_iteration_host(True, common_i, out_mean, out_cov, out_tau)
        hp = eng.get_new_hp()
        return r, hp

    for _ in range(min(2, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    eng.timer_start()
    for _ in range(args.steps):
        step_e2e()
    ms_e2e = eng.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))
    h2d = sum(host[k].nbytes for k in ("offsets", "grid_idx", "y", "grid", "theta_k", "theta_i", "pi", "tau0")) + 8 * K
    d2h = out_mean.nbytes + out_cov.nbytes + out_tau.nbytes + 8 * (3 * K + 3 * M + K)
    e2e_value = M_total * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    n = cfg["n"]
    flop_fit = 5.0 * n ** 3 + 10.0 * n ** 2           # objective + gradient of one individual
    flop_mean_eval = 5.0 * float(N) ** 3               # K^-1 (N^3) + K^-1 Csum K^-1 (4 N^3), shared theta_k
    kernels = {}
    if phase_ms.get("mstep_indiv", 0) > 0:
        t = phase_ms["mstep_indiv"] / 1e3
        kernels["indiv_mstep_kernel"] = {"ms_per_step": phase_ms["mstep_indiv"] / args.steps,
                                         "gp_fits_per_s": nfev_i_sum / t, "tflops": nfev_i_sum * flop_fit / t / 1e12}
    if phase_ms.get("mstep_mean", 0) > 0:
        t = phase_ms["mstep_mean"] / 1e3
        kernels["mean_process_evaluations"] = {"ms_per_step": phase_ms["mstep_mean"] / args.steps,
                                               "evals_per_s": nfev_k / t, "tflops": nfev_k * flop_mean_eval / t / 1e12}
    # The roofline is quoted on the kernel that carries the flops of the step: the batched per-individual
    # objective/gradient kernel (or, with shared theta_i, the same device function driven from the host). The
    # theta_k chain (N x N inverse + 2 DMMA GEMMs per evaluation) is listed beside it: at N = 401 it is bound by the
    # N sequential pivots of the factorisation, not by flops or bytes.
    dom = "indiv_mstep_kernel" if "indiv_mstep_kernel" in kernels else (max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None)
    # three quarters of the executed FMAs of a GP fit run on the FP64 tensor path (DMMA m8n8k4), the sweep on DFMA: the
    # denominator is the higher of the two live measurements (DMMA 37.1 vs DFMA 33.7 TFLOP/s on this pool)
    peak = max(peaks["dmma_tflops"], peaks["dfma_tflops"])
    roofline = None
    if dom:
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if dom in tr and args.workload == "C2":
                traffic = tr[dom]["dram__bytes_read.sum"] + tr[dom]["dram__bytes_write.sum"]
        except Exception:
            traffic = None
        roofline = {"kernel": dom, "bound": "tensor", "bound_detail": "FP64: DMMA m8n8k4 (no tcgen05 path for doubles) + DFMA",
                    "achieved": kernels[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                    "frac": kernels[dom]["tflops"] / peak, "traffic": traffic,
                    "peak_source": "live DMMA m8n8k4 loop on this GPU (mcb_measure_fp64_peaks; of measured): %.1f TFLOP/s; "
                                   "live DFMA loop %.1f TFLOP/s; MEASURED_PEAKS.json has no FP64 entry (its tensor figure is "
                                   "bf16)" % (peaks["dmma_tflops"], peaks["dfma_tflops"]),
                    "flops_per_unit": "GP fit = 5 n^3 + 10 n^2 (SURVEY 8d) = %.0f flop at n = %d; units per launch = sum of L-BFGS "
                                      "evaluations over the individuals (%.0f per step)" % (flop_fit, n, nfev_i_sum / args.steps),
                    "all": kernels, "phase_ms_per_step": {k: v / args.steps for k, v in phase_ms.items()},
                    "note": "theta_i (mstep_indiv) and theta_k (mstep_mean) optimisations run concurrently on two streams; "
                            "the theta_k chain is latency-bound (N sequential pivots per inverse)"}

    # ---- the dominant kernel away from the latency-bound benchmark size (reported beside the roofline, N = 1 only) -----
    # At 1000 individuals per GPU the per-individual M-step is bounded by its longest optimisation (82 sequential
    # evaluations); with 8x the individuals the same kernel runs at its steady-state rate. Same step, same code path.
    steady = None
    if world == 1 and roofline is not None and not args.no_steady and not common_i and args.workload == "C2":
        try:
            Ms = 8 * cfg["M"]
            ds8 = synth.make_dataset(Ms, cfg["K"], cfg["n"], cfg["n_grid"], cfg["common_hp_i"], cfg["seed"] + 1)
            tk8, ti8, mk8 = start_state(Ms, K)
            eng.set_data(ds8["offsets"], ds8["grid_idx_full"], ds8["y"], np.ascontiguousarray(ds8["grid_full"]), M_total=Ms)
            t_i, n_i = 0.0, 0
            for it in range(3):
                eng.set_state(tk8, ti8, np.ascontiguousarray(ds8["tau0"].mean(0)), mk8, np.ascontiguousarray(ds8["tau0"]))
                eng.vem_iteration(True, common_i)
                if it == 0:
                    continue   # the first M-step has no evaluation counts to order its work queue by
                ph8, _ = eng.last_timings()
                t_i += ph8["mstep_indiv"] / 1e3
                n_i += eng.last_mstep_stats()["nfev_i_sum"]
            tf = n_i * flop_fit / t_i / 1e12
            steady = {"individuals": Ms, "gp_fits_per_s": n_i / t_i, "tflops": tf, "frac": tf / peak,
                      "note": "same kernel and step at 8x the individuals (its M-step phase next to the theta_k chain, "
                              "which keeps 26 + 15 of the 148 SMs): throughput bound instead of bound by the longest optimisation"}
            roofline["steady_state"] = steady
        except Exception as ex:  # the extra measurement must never take the contract line down
            roofline["steady_state"] = {"error": str(ex)[:200]}

    gp_fits = sum_over_ranks(nfev_i_sum)
    line = None
    if rank == 0:
        base = None
        if not args.skip_cpu_baseline:
            b = cpu_baseline(ds, cfg, args.cpu_sample)
            sec = extrapolate(b, cfg["M"])
            base = {"value": cfg["M"] / sec, "unit": "individual-VEM-iterations/s", "cores": os.cpu_count() or 1,
                    "kind": "port",
                    "sample": f"numpy port of the reference R code (no R on this box): one VEM iteration on "
                              f"{b['sample_individuals']} of {cfg['M']} individuals took "
                              f"{b['sample_seconds_per_iteration']:.2f} s ({b['grid_seconds']:.2f} s of it theta_k grid "
                              f"work); per-individual part scaled linearly in M, grid part kept",
                    "phases_s": b["phases"], "nfev": b["nfev"]}
        line = {
            "metric": METRIC, "value": value, "unit": "individual-VEM-iterations/s",
            "vem_iterations_per_s": args.steps / (ms / 1e3), "gp_fits_per_s": gp_fits / (ms / 1e3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: M={cfg['M']} individuals/GPU x K={cfg['K']}, n_i={cfg['n']}, "
                                   f"grid N={N}, individual-specific hp={not common_i}, common theta_k",
                       "M_total": int(M_total), "K": K,
                       "step": "one training_VEM loop body from the start state (E-step, ELBO, M-step L-BFGS, ELBO)",
                       "repetitions": "timed region of %d steps measured %d times, faster repetition reported" % (args.steps, args.reps),
                       "l2": "flushed between steps (256 MiB memset on the stream, inside the timed region)"
                             if not args.no_flush else "not flushed",
                       "state_reset": "start state re-uploaded each step (%d B H2D, inside the timed region)" %
                                      (theta_i.nbytes + tau0.nbytes + theta_k.nbytes)},
            "clocks": clk, "roofline": roofline, "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": "individual-VEM-iterations/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "pinned_host_buffers": len(pinned)},
            "gpu_launches": int(launches),
            "mstep_evals_per_step": {"theta_i_total": nfev_i_sum / args.steps, "theta_k": nfev_k / args.steps},
        }
        emit(line)
    for a in pinned:
        eng.unpin(a)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
