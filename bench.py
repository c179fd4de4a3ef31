#!/usr/bin/env python
"""bench.py — VEM iterations/s of the MAGMAclust hot path on B200 (contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one iteration of training_VEM's loop body (MAGMAclust.R:44-93) from the study's start state
(theta_k = theta_i = (1, 1, 0.2), prior means 0, tau_0 = one-hot k-means labels): E-step, ELBO, M-step with the
built-in L-BFGS, ELBO at the new hp, acceptance rule. Headline workload = BASELINE.json configs[1] ("C2": M = 1000
individuals, K = 5, n_i = 50, individual-specific hyper-parameters) PER GPU; with N GPUs the individuals are
sharded (weak scaling: N*1000 individuals) and the grid accumulators are all-reduced with NCCL.

  value : device-resident state, CUDA-event time on the library's stream, max over ranks
  e2e   : the same step through the host-buffer API (upload dataset + state from pinned host memory, download
          mean/cov/tau/hp every step)

Next to the headline, the same line carries (every --gpus N, unless --no-extras):
  sharded_vs_single : N > 1 only. After the timed region every rank runs ONE more sharded step, rank 0 re-runs the same
                      step UNSHARDED (all M_total individuals on its own GPU) and the relative errors of mean / cov / tau /
                      ELBO go into the line; above 1e-8 / 1e-6 (north_star's tolerances) the run FAILS (exit 1)
  c3_strong         : BASELINE.json configs[2]: M_total = 100,000 individuals FIXED (K = 5, n_i = 30, N = 201, shared hp),
                      sharded over the N GPUs: ms per VEM iteration (strong scaling), phases, and its own sharded_vs_single
  c4_owner          : configs[3]: dense grid N = 8192, K = 8 (M = 4096 x 32 obs): ms per E-step; with N > 1 cluster k is
                      owned by rank k mod N (one cluster per GPU at N = 8), A_k reduced to its owner, C_k broadcast back
  c5_predict        : configs[4]: 10,000 new individuals x 20 obs onto a 2,000-point grid, K = 5, all targets: new individuals
                      predicted per second through host buffers (strong scaling: B / N per rank, no exchange step)
  c1_training       : configs[0] (the reference's CPU-runnable case, M = 50, K = 3): ms per whole training_VEM
  training_run      : one real multi-iteration training_VEM trajectory (until the reference's stopping rule, <= 15
                      iterations) of the headline workload, timed once

  --impl reference : the CPU path timed on this box's host cores. R is not installed (neither here nor on the
          GPU box), so this is the numpy restatement in oracle/ ("port"). Each step is one VEM iteration on a BOUNDED
          SAMPLE of the workload's individuals; value is the throughput measured on that sample (not an extrapolation).
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
UNIT = "individual-VEM-iterations/s"
TOL_POST, TOL_LL = 1e-8, 1e-6     # north_star: hyper-posterior means / covariances; log-likelihoods / responsibilities


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
    ap.add_argument("--no-extras", action="store_true", help="skip c3_strong / c4_owner / c5_predict / c1_training / training_run")
    ap.add_argument("--cpu-sample", type=int, default=200, help="individuals in the CPU baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--individuals", type=int, default=0, help="override the individuals per GPU of the workload")
    ap.add_argument("--data", default="", help="directory written by data/gen.py for --workload: read the dataset from its files "
                    "(the bytes baseline/run_reference.R reads) instead of generating it in memory; N = 1 only")
    return ap.parse_args()


def config_block(name, cfg, n_gpus):
    """the `config` object, IDENTICAL in both arms (the driver compares them)"""
    return {"workload": f"{name}: M={cfg['M']} individuals/GPU x K={cfg['K']}, n_i={cfg['n']}, grid of {cfg['n_grid']} points, "
                        f"individual-specific hp={not cfg['common_hp_i']}, common theta_k; one training_VEM loop body "
                        f"(E-step, ELBO, M-step L-BFGS, ELBO) from the study's start state",
            "M_total": int(cfg["M"] * n_gpus), "K": int(cfg["K"])}


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


DATA_DIR = ""


def workload_data(name, nranks, individuals=0, strong=False):
    """the synthetic dataset of a BASELINE config. Weak scaling: M individuals PER rank; strong: M in total."""
    from magmaclust_b200 import synth
    if DATA_DIR and nranks == 1 and not individuals and not strong:
        from data import gen
        ds, c = gen.load(DATA_DIR)
        if c.get("name") != name:
            raise SystemExit(f"--data {DATA_DIR} holds {c.get('name')}, not {name}")
        return ds, c
    c = dict(synth.CONFIGS[name])
    if individuals:
        c["M"] = individuals
    M_total = c["M"] if strong else c["M"] * nranks
    ds = synth.make_dataset(M_total, c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"],
                            equispaced_cover=(name == "C4"))
    return ds, c


def start_state(M, K):
    theta_k = np.tile(np.array([1.0, 1.0, 0.2]), (K, 1))
    theta_i = np.tile(np.array([1.0, 1.0, 0.2]), (M, 1))
    return theta_k, theta_i, np.zeros(K)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


# ------------------------------------------------------------------------------------------------
def cpu_baseline(ds, cfg, sample, steps=1, warmup=0):
    """VEM iterations of the oracle (numpy port of the reference R code) on `sample` individuals of the workload;
    seconds per iteration on the sample (mean over `steps`), split into phases."""
    from magmaclust_b200 import synth
    from oracle import magma_oracle as O
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
    t_all0 = time.perf_counter()
    runs = [one_iteration() for _ in range(max(1, steps))]
    wall = (time.perf_counter() - t_all0) / max(1, steps)
    tot = float(np.mean([r["e_step"] + r["elbo"] + r["m_step"] for r in runs]))
    t_k = float(np.mean([r["nfev"]["t_k"] for r in runs]))   # theta_k optimiser: N x N work, independent of M
    return {"sample_seconds_per_iteration": tot, "wall_seconds_per_iteration": wall, "sample_individuals": len(ids),
            "grid_seconds": t_k, "phases": {k: float(np.mean([r[k] for r in runs])) for k in ("e_step", "elbo", "m_step")},
            "nfev": runs[-1]["nfev"]}


def baseline_record(b, cfg, cores):
    Ms, sec = b["sample_individuals"], b["sample_seconds_per_iteration"]
    full = (sec - b["grid_seconds"]) * (cfg["M"] / Ms) + b["grid_seconds"]
    return {"value": Ms / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"numpy port of the reference R code (R is not installed on this box): one VEM iteration on the first "
                      f"{Ms} of the workload's {cfg['M']} individuals takes {sec:.2f} s on {cores} host cores ({b['grid_seconds']:.2f} s "
                      f"of it is theta_k grid work that does not grow with M); value = {Ms} individuals / {sec:.2f} s, measured, "
                      f"not extrapolated",
            "phases_s": b["phases"], "nfev": b["nfev"],
            "extrapolated_full_workload_s_per_iteration": full}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ds, cfg = workload_data(args.workload, 1, args.individuals)
    cores = os.cpu_count() or 1
    # bounded sample: the whole --steps/--warmup run should end within a few minutes (~0.07 s per individual and iteration)
    sample = int(min(args.cpu_sample, max(40, 1600 // max(1, args.steps + args.warmup))))
    base = cpu_baseline(ds, cfg, sample, steps=args.steps, warmup=args.warmup)
    rec = baseline_record(base, cfg, cores)
    sec = base["wall_seconds_per_iteration"]
    value = base["sample_individuals"] / sec
    rec["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "vem_iterations_per_s": 1.0 / sec, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_block(args.workload, cfg, args.gpus),
        "cpu_baseline": rec,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
class Dist:
    """control plane only (id broadcast, barriers, max / gather over ranks): torch.distributed over gloo"""

    def __init__(self, rank, world):
        self.rank, self.world, self.d = rank, world, None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("gloo", rank=rank, world_size=world)
            self.d = dist

    def barrier(self):
        if self.d is not None:
            self.d.barrier()

    def _red(self, x, op):
        if self.d is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.d.all_reduce(t, op=op)
        return float(t[0])

    def max(self, x):
        return self._red(x, self.d.ReduceOp.MAX) if self.d is not None else x

    def sum(self, x):
        return self._red(x, self.d.ReduceOp.SUM) if self.d is not None else x

    def gather_rows(self, a):
        """concatenation over ranks (rank order) of a 2-D float64 array, on every rank"""
        if self.d is None:
            return a
        parts = [None] * self.world
        self.d.all_gather_object(parts, np.ascontiguousarray(a))
        return np.concatenate(parts)

    def comm_init(self, eng, Engine):
        if self.d is None:
            return
        import torch
        idt = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            idt = torch.tensor(list(Engine.comm_unique_id()), dtype=torch.uint8)
        self.d.broadcast(idt, 0)
        eng.comm_init(self.world, self.rank, bytes(idt.tolist()))

    def close(self):
        if self.d is not None:
            self.d.barrier()
            self.d.destroy_process_group()


def sharded_vs_single(D, eng, local, ds, sh, common_i, what="vem"):
    """One more sharded step on every rank; rank 0 repeats it unsharded on a second handle (all M_total individuals on its GPU).
    Relative errors of the hyper-posterior mean / covariance, responsibilities, ELBO; `ok` against north_star's tolerances."""
    from magmaclust_b200.engine import Engine
    K = sh["K"]
    tk, ti, mk = start_state(sh["M"], K)
    pi0 = np.ascontiguousarray(ds["tau0"].mean(axis=0))
    eng.set_state(tk, ti, pi0, mk, np.ascontiguousarray(sh["tau0"]))
    if what == "vem":
        elbo = eng.vem_iteration(True, common_i)[0]
    else:
        eng.e_step()
        elbo = eng.elbo()[1]
    mean = eng.get_mean()
    cov = eng.get_cov(K - 1)          # one cluster's covariance travels (the last one: owned by the last rank in owner mode)
    tau_all = D.gather_rows(eng.get_tau_new())
    rec = None
    if D.rank == 0:
        ref = Engine(local)
        try:
            ref.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], np.ascontiguousarray(ds["grid_full"]))
            tk1, ti1, mk1 = start_state(ds["M"], K)
            ref.set_state(tk1, ti1, pi0, mk1, np.ascontiguousarray(ds["tau0"]))
            if what == "vem":
                relbo = ref.vem_iteration(True, common_i)[0]
            else:
                ref.e_step()
                relbo = ref.elbo()[1]
            rt = ref.get_tau_new()
            rec = {"mean": rel(mean, ref.get_mean()), "cov": rel(cov, ref.get_cov(K - 1)),
                   "tau": float(np.max(np.abs(tau_all - rt))), "elbo": abs(elbo - relbo) / abs(relbo),
                   "hard_assignments_identical": bool(np.array_equal(tau_all.argmax(1), rt.argmax(1))),
                   "individuals": int(ds["M"]), "ranks": D.world}
            rec["ok"] = bool(rec["mean"] < TOL_POST and rec["cov"] < TOL_POST and rec["tau"] < TOL_LL and
                             rec["elbo"] < TOL_LL and rec["hard_assignments_identical"])
        finally:
            ref.close()
    D.barrier()
    return rec


def timed_steps(D, eng, step, steps, warmup, reps, flush):
    """W warm-up steps, then `reps` repetitions of [barrier, CUDA-event timer, K steps, timer, barrier]; max over ranks"""
    for _ in range(warmup):
        step()
    out = []
    for _ in range(reps):
        phase_ms, launches, nfev_i_sum, nfev_k, walls = {}, 0, 0, 0, []
        D.barrier()
        t_wall0 = time.time()
        eng.timer_start()
        for _ in range(steps):
            tw = time.perf_counter()
            if flush:
                eng.flush_l2()
            step()
            walls.append((time.perf_counter() - tw) * 1e3)
            ph, nl = eng.last_timings()      # reading the phase events costs a host sync that the step itself already paid
            for k, v in ph.items():
                phase_ms[k] = phase_ms.get(k, 0.0) + v
            launches += nl
            st = eng.last_mstep_stats()
            nfev_i_sum += st["nfev_i_sum"]
            nfev_k += st["nfev_k"]
        ms = eng.timer_stop()
        t_wall1 = time.time()
        D.barrier()
        out.append(dict(ms=D.max(ms), phase_ms=phase_ms, launches=launches, nfev_i_sum=nfev_i_sum, nfev_k=nfev_k,
                        t0=t_wall0, t1=t_wall1, walls=walls))
    return out


def run_c3_strong(D, local, args):
    """BASELINE configs[2]: M_total = 100,000 fixed, sharded over the ranks (strong scaling), shared hyper-parameters"""
    from magmaclust_b200 import synth
    from magmaclust_b200.engine import Engine
    ds, cfg = workload_data("C3", D.world, strong=True)
    sh = synth.shard(ds, D.rank, D.world)
    eng = Engine(local)
    try:
        D.comm_init(eng, Engine)
        K = cfg["K"]
        tk, ti, mk = start_state(sh["M"], K)
        pi0 = np.ascontiguousarray(ds["tau0"].mean(axis=0))
        tau0 = np.ascontiguousarray(sh["tau0"])
        eng.set_data(sh["offsets"], sh["grid_idx"], sh["y"], np.ascontiguousarray(sh["grid"]), M_total=sh["M_total"])

        def step():
            eng.set_state(tk, ti, pi0, mk, tau0)
            return eng.vem_iteration(True, True)

        steps = max(10, min(args.steps, 12))
        r = timed_steps(D, eng, step, steps, 3, 1, not args.no_flush)[0]
        n = cfg["n"]
        evals = r["nfev_i_sum"]          # shared theta_i: evaluations x local individuals
        t_i = r["phase_ms"].get("mstep_indiv", 0.0) / 1e3
        out = {"workload": f"C3: M_total={ds['M']} individuals (fixed) x K={K}, n_i={n}, grid of {cfg['n_grid']} points, shared hp",
               "M_total": int(ds["M"]), "individuals_per_gpu": int(sh["M"]), "steps": steps, "warmup": 3,
               "ms_per_step": r["ms"] / steps, "value": ds["M"] * steps / (r["ms"] / 1e3), "unit": UNIT, "scaling": "strong",
               "phase_ms": {k: v / steps for k, v in r["phase_ms"].items()},
               "theta_i_evaluations_per_step": evals / max(1, sh["M"]) / steps, "theta_k_evaluations_per_step": r["nfev_k"] / steps,
               "gp_fits_per_s_this_rank": (evals / t_i) if t_i > 0 else None,
               "gp_fit_tflops_this_rank": (evals * (5.0 * n ** 3 + 10.0 * n ** 2) / t_i / 1e12) if t_i > 0 else None}
        if D.world > 1:
            out["sharded_vs_single"] = sharded_vs_single(D, eng, local, ds, sh, True, "vem")
        return out
    finally:
        eng.close()


def run_c4_owner(D, local, args):
    """BASELINE configs[3]: N = 8192 grid, K = 8. E-step only (the N x N hyper-posterior solves are the path's one dense
    contraction): with N > 1 ranks cluster k is owned by rank k mod N"""
    from magmaclust_b200 import synth
    from magmaclust_b200.engine import Engine
    ds, cfg = workload_data("C4", D.world, strong=True)
    sh = synth.shard(ds, D.rank, D.world)
    eng = Engine(local)
    try:
        D.comm_init(eng, Engine)
        K, N = cfg["K"], cfg["n_grid"]
        tk, ti, mk = start_state(sh["M"], K)
        pi0 = np.ascontiguousarray(ds["tau0"].mean(axis=0))
        tau0 = np.ascontiguousarray(sh["tau0"])
        eng.set_data(sh["offsets"], sh["grid_idx"], sh["y"], np.ascontiguousarray(sh["grid"]), M_total=sh["M_total"])
        ph_sum, steps = {}, 3

        def step():
            eng.set_state(tk, ti, pi0, mk, tau0)
            eng.e_step()

        step()   # warm-up (allocations, NCCL channels)
        D.barrier()
        eng.timer_start()
        for _ in range(steps):
            step()
            ph, _ = eng.last_timings()
            for k, v in ph.items():
                ph_sum[k] = ph_sum.get(k, 0.0) + v
        ms = D.max(eng.timer_stop())
        D.barrier()
        owned = len(range(D.rank, K, D.world))
        t_dense = ph_sum.get("dense_hyperpost", 0.0) / steps / 1e3
        out = {"workload": f"C4: dense union grid N={N}, K={K}, M={ds['M']} individuals x {cfg['n']} obs; E-step",
               "N": N, "K": K, "steps": steps, "warmup": 1, "ms_per_estep": ms / steps,
               "clusters_owned_by_this_rank": owned, "ownership": "cluster k on rank k mod n_gpus" if D.world > 1 else "single GPU",
               "phase_ms": {k: v / steps for k, v in ph_sum.items()},
               # one distinct K_k^-1 (shared theta_k, every rank) + the owned A_k^-1, N^3 flop each
               "dense_tflops_this_rank": ((1 + owned) * float(N) ** 3 / t_dense / 1e12) if t_dense > 0 else None}
        if D.world > 1:
            out["sharded_vs_single"] = sharded_vs_single(D, eng, local, ds, sh, True, "estep")
        return out
    finally:
        eng.close()


def run_c5_predict(D, local, args):
    """BASELINE configs[4]: pred_magmaclust-style prediction, 10,000 new individuals (20 observations each) onto a 2,000-point
    grid with the cluster-mixture posterior (K = 5): posterior_mu_k once, then update_tau_star_k_EM + pred_gp_clust for the
    whole batch in one call. New individuals shard with NO exchange step: every rank holds the prediction
    posterior and takes B / n_gpus of them (strong scaling). Host buffers in, page-locked host buffers out (reused across
    calls); 16 B K T bytes come back per new individual."""
    from magmaclust_b200 import synth
    from magmaclust_b200.engine import Engine
    B_total, T, K, nstar = 10000, 2000, 5, 20
    ds = synth.make_dataset(M=1000, K=K, n=30, n_grid=T, common_hp_i=True, seed=synth.CONFIGS["C5"]["seed"])
    eng = Engine(local)
    try:
        M = ds["M"]
        eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
        tk, ti, mk = start_state(M, K)
        eng.set_state(tk, ti, np.ascontiguousarray(ds["tau0"].mean(0)), mk, np.ascontiguousarray(ds["tau0"]))
        eng.timer_start()
        eng.posterior_mu_k(ds["grid_full"], tau=ds["tau0"], want_mean=False, want_cov=False)
        ms_post = eng.timer_stop()
        rng = np.random.default_rng(7)
        obs = np.sort(np.argsort(rng.random((B_total, T)), axis=1)[:, :nstar], axis=1).astype(np.int32)
        mu = ds["mu_true"][rng.integers(0, K, B_total)]
        y = np.take_along_axis(mu, obs, axis=1) + 0.3 * rng.standard_normal((B_total, nstar))
        lo, hi = B_total * D.rank // D.world, B_total * (D.rank + 1) // D.world
        B = hi - lo
        offsets = (np.arange(B + 1) * nstar).astype(np.int32)
        o_idx, o_y = np.ascontiguousarray(obs[lo:hi].ravel()), np.ascontiguousarray(y[lo:hi].ravel())
        pred_idx = np.arange(T, dtype=np.int32)
        theta_new = np.array([1.0, 1.0, 0.2])
        out = (eng.pinned_empty((B, K)), eng.pinned_empty((B, K, T)), eng.pinned_empty((B, K, T)))
        eng.predict(offsets, o_idx, o_y, theta_new, pred_idx, out=out)   # warm-up (device allocations)
        steps, ker = 3, 0.0
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.predict(offsets, o_idx, o_y, theta_new, pred_idx, out=out)
            ker += eng.last_timings()[0]["predict"]
        wall = D.max((time.perf_counter() - t0) * 1e3)
        D.barrier()
        tau, mean, var = out
        flop = B * K * (nstar ** 3 + 2.0 * nstar ** 2 * T + 4.0 * nstar * T)     # SURVEY 8d per (new individual, cluster)
        return {"workload": f"C5: {B_total} new individuals (fixed) x {nstar} obs onto a {T}-point grid, K={K}, all {T} targets",
                "new_individuals_this_rank": B, "steps": steps, "warmup": 1, "ms_per_call_e2e": wall / steps,
                "value": B_total * steps / (wall / 1e3), "unit": "new individuals predicted/s (host buffers in, page-locked host buffers out)",
                "scaling": "strong", "posterior_mu_k_ms": ms_post, "predict_kernels_ms_this_rank": ker / steps,
                "predict_kernel_tflops_this_rank": flop / (ker / steps / 1e3) / 1e12 if ker > 0 else None,
                "h2d_bytes_per_call": int(offsets.nbytes + o_idx.nbytes + o_y.nbytes + pred_idx.nbytes),
                "d2h_bytes_per_call": int(tau.nbytes + mean.nbytes + var.nbytes),
                "checks": {"tau_rowsum_err": float(np.abs(tau.sum(1) - 1).max()), "var_min": float(var.min())}}
    finally:
        eng.close()


def run_c1(D, local, args):
    """BASELINE configs[0] (the reference's own CPU-runnable case: M = 50, K = 3, ~30 obs, common hp): one whole training_VEM
    on this rank (replicas only: 50 individuals do not shard usefully). The reference's stored R timing for a training of this
    shape is 97.8 s."""
    from magmaclust_b200 import synth
    from magmaclust_b200.engine import Engine
    c = synth.CONFIGS["C1"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    eng = Engine(local)
    try:
        tk, ti, mk = start_state(ds["M"], c["K"])
        eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], np.ascontiguousarray(ds["grid_full"]))
        best = None
        for rep in range(3):
            eng.set_state(tk, ti, np.ascontiguousarray(ds["tau0"].mean(0)), mk, np.ascontiguousarray(ds["tau0"]))
            eng.timer_start()
            trace, conv = [], False
            for it in range(15):
                elbo, eps, conv = eng.vem_iteration(True, True)
                trace.append(elbo)
                if conv:
                    break
            ms = eng.timer_stop()
            if rep > 0 and (best is None or ms < best[0]):
                best = (ms, len(trace), bool(conv), trace[-1])
        return {"workload": "C1: M=50 individuals x K=3, n_i=30, grid of 201 points, common hp; whole training_VEM from the study's "
                            "start state", "ms_per_training": best[0], "iterations": best[1], "converged": best[2],
                "ms_per_iteration": best[0] / best[1], "elbo_last": best[3], "replicas": D.world,
                "reference_stored_r_seconds_for_this_shape": 97.8}
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    D = Dist(rank, world)

    from magmaclust_b200 import synth
    from magmaclust_b200.engine import Engine

    ds, cfg = workload_data(args.workload, world, args.individuals)
    sh = synth.shard(ds, rank, world)
    M, K, N = sh["M"], sh["K"], sh["N"]
    eng = Engine(local)
    D.comm_init(eng, Engine)

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
    # The timed region (exactly --steps steps between barriers, CUDA events on the library stream) is measured
    # `reps` times back to back and the faster repetition is reported (the others are listed): a fresh box occasionally
    # stalls one step by tens of milliseconds (seen right after another process released the GPU), which is not a property
    # of the code.
    runs = timed_steps(D, eng, step_resident, args.steps, args.warmup, args.reps, not args.no_flush)
    best = min(runs, key=lambda r: r["ms"])
    ms, phase_ms, launches, nfev_i_sum, nfev_k = best["ms"], best["phase_ms"], best["launches"], best["nfev_i_sum"], best["nfev_k"]
    if os.environ.get("MCB_BENCH_STEP_TIMES"):
        print("step wall ms:", " ".join("%.2f" % x for x in best["walls"]), file=sys.stderr)
    clk = clocks.stop(best["t0"], best["t1"])
    M_total = sh["M_total"]
    value = M_total * args.steps / (ms / 1e3)

    # ---- e2e: host buffers in, host buffers out, every step -----------------------------------------
    out_mean = np.empty((K, N))
    out_cov = np.empty((K, N, N))
    out_tau = np.empty((M, K))
    pinned += [a for a in (out_mean, out_cov, out_tau) if eng.pin(a)]

    def step_e2e():
        eng.set_data(host["offsets"], host["grid_idx"], host["y"], host["grid"], M_total=sh["M_total"])
        eng.set_state(host["theta_k"], host["theta_i"], host["pi"], host["m_k"], host["tau0"])
        # one call per loop body of training_VEM, param of the E-step delivered to the (pinned) host buffers
        r = eng.vem_iteration_host(True, common_i, out_mean, out_cov, out_tau)
        hp = eng.get_new_hp()
        return r, hp

    for _ in range(min(2, args.warmup)):
        step_e2e()
    D.barrier()
    t0 = time.perf_counter()
    eng.timer_start()
    for _ in range(args.steps):
        step_e2e()
    ms_e2e = eng.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = D.max(max(ms_e2e, wall_e2e))
    h2d = sum(host[k].nbytes for k in ("offsets", "grid_idx", "y", "grid", "theta_k", "theta_i", "pi", "tau0")) + 8 * K
    d2h = out_mean.nbytes + out_cov.nbytes + out_tau.nbytes + 8 * (3 * K + 3 * M + K)
    e2e_value = M_total * args.steps / (ms_e2e / 1e3)

    # ---- one real training_VEM trajectory (MC.R:44-93: until eps < 1e-2, at most 15 iterations), timed once ------------
    training_run = None
    if not args.no_extras:
        eng.set_state(host["theta_k"], host["theta_i"], host["pi"], host["m_k"], host["tau0"])
        D.barrier()
        eng.timer_start()
        trace, conv = [], False
        for _ in range(15):
            elbo, eps, conv = eng.vem_iteration(True, common_i)
            trace.append(elbo)
            if conv:
                break
        t_train = D.max(eng.timer_stop())
        training_run = {"iterations": len(trace), "converged": bool(conv), "ms_total": t_train,
                        "ms_per_iteration": t_train / len(trace), "elbo_first": trace[0], "elbo_last": trace[-1]}
        eng.set_state(host["theta_k"], host["theta_i"], host["pi"], host["m_k"], host["tau0"])

    # ---- sharded == unsharded, in the driver's own record (N > 1) ----------------------------------------------------------
    parity = None
    if world > 1:
        parity = sharded_vs_single(D, eng, local, ds, sh, common_i, "vem")

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
                                               "evals_per_s": nfev_k / t, "us_per_evaluation": 1e6 * t / max(1, nfev_k),
                                               "tflops": nfev_k * flop_mean_eval / t / 1e12}
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
        for name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", name)))
                key = dom if dom in tr else ("indiv_mstep2_kernel" if "indiv_mstep2_kernel" in tr else None)
                if key and args.workload == "C2":
                    traffic = tr[key]["dram__bytes_read.sum"] + tr[key]["dram__bytes_write.sum"]
                    break
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
        if not common_i:
            # the persistent theta_i kernel leaves whole SMs to the slab inverse of the theta_k chain (api.cu, m_step_impl:
            # three / two / one column part per 32-row slab while that stays <= 40 SMs); `peak` above is the whole GPU's
            nblk = (int(sh["N"]) + 31) // 32
            reserve = 3 * nblk if 3 * nblk <= 40 else (2 * nblk if 2 * nblk <= 40 else nblk)
            import torch
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            roofline["sm_share"] = {"sms": sms, "left_to_theta_k_chain": reserve, "kernel_sms": sms - reserve,
                                    "frac_on_its_sms": roofline["frac"] * sms / (sms - reserve)}

    # ---- the dominant kernel away from the latency-bound benchmark size (reported beside the roofline, N = 1 only) -----
    # At 1000 individuals per GPU the per-individual M-step is bounded by its longest optimisation (82 sequential
    # evaluations); with 8x the individuals the same kernel runs at its steady-state rate. Same step, same code path.
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
            roofline["steady_state"] = {"individuals": Ms, "gp_fits_per_s": n_i / t_i, "tflops": tf, "frac": tf / peak,
                                        "note": "same kernel and step at 8x the individuals (its M-step phase next to the theta_k "
                                                "chain): throughput bound instead of bound by the longest optimisation"}
        except Exception as ex:  # the extra measurement must never take the contract line down
            roofline["steady_state"] = {"error": str(ex)[:200]}

    gp_fits = D.sum(nfev_i_sum)
    for a in pinned:
        eng.unpin(a)
    eng.close()

    # ---- the other BASELINE configurations, in the same record -----------------------------------------------------------
    c1 = c3 = c4 = c5 = None
    if not args.no_extras and args.workload == "C2" and not args.individuals:
        try:
            c3 = run_c3_strong(D, local, args)
        except Exception as ex:
            c3 = {"error": str(ex)[:300]}
        try:
            c4 = run_c4_owner(D, local, args)
        except Exception as ex:
            c4 = {"error": str(ex)[:300]}
        try:
            c5 = run_c5_predict(D, local, args)
        except Exception as ex:
            c5 = {"error": str(ex)[:300]}
        try:
            c1 = run_c1(D, local, args)
        except Exception as ex:
            c1 = {"error": str(ex)[:300]}

    failed = False
    if rank == 0:
        base = None
        if not args.skip_cpu_baseline:
            base = baseline_record(cpu_baseline(ds, cfg, args.cpu_sample), cfg, os.cpu_count() or 1)
        reps_ms = [r["ms"] / args.steps for r in runs]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT,
            "vem_iterations_per_s": args.steps / (ms / 1e3), "gp_fits_per_s": gp_fits / (ms / 1e3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "ms_per_step_repetitions": reps_ms, "ms_per_step_median_of_repetitions": float(np.median(reps_ms)),
            "step_wall_ms_median": float(np.median(best["walls"])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.workload, cfg, world),
            "timing": {"repetitions": "timed region of %d steps measured %d times, fastest repetition reported as value, all "
                                      "listed in ms_per_step_repetitions" % (args.steps, args.reps),
                       "l2": "flushed between steps (256 MiB memset on the stream, inside the timed region)"
                             if not args.no_flush else "not flushed",
                       "state_reset": "start state re-uploaded each step (%d B H2D, inside the timed region)" %
                                      (theta_i.nbytes + tau0.nbytes + theta_k.nbytes)},
            "clocks": clk, "roofline": roofline, "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "pinned_host_buffers": len(pinned)},
            "gpu_launches": int(launches),
            "mstep_evals_per_step": {"theta_i_total": nfev_i_sum / args.steps, "theta_k": nfev_k / args.steps},
            "training_run": training_run, "sharded_vs_single": parity, "c3_strong": c3, "c4_owner": c4,
            "c5_predict": c5, "c1_training": c1,
        }
        for rec in (parity, (c3 or {}).get("sharded_vs_single"), (c4 or {}).get("sharded_vs_single")):
            if rec is not None and not rec.get("ok", False):
                failed = True
        if failed:
            line["parity_failed"] = True
        emit(line)
    D.close()
    return 1 if failed else 0


if __name__ == "__main__":
    a = parse()
    if a.data:
        globals()["DATA_DIR"] = a.data
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
