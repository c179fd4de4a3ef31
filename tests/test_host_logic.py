"""CPU: host-side logic — canonicalisation, the L-BFGS of csrc/lbfgs.h (compiled with g++), synthetic data,
and the sharded accumulation over world_size-2 gloo."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import small_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_canon_matches_reference_ordering():
    from magmaclust_b200.magma import Canon
    db, hp, tau, m_k = small_problem(seed=3)
    # shuffle the individuals' blocks: unique(db$ID) order must be first appearance, rows kept per individual
    per = O.split_db(db)
    order = [5, 1, 9, 3] + [i for i in per if i not in (5, 1, 9, 3)]
    db2 = {k: np.concatenate([per[i][k] for i in order]) for k in ("ID", "Timestamp", "Output")}
    c = Canon(db2)
    assert c.ids == order
    assert np.array_equal(c.grid, np.unique(db["Timestamp"]))
    for r, i in enumerate(order):
        seg = slice(c.offsets[r], c.offsets[r + 1])
        assert np.array_equal(c.grid[c.grid_idx[seg]], per[i]["Timestamp"])
        assert np.array_equal(c.y[seg], per[i]["Output"])


def test_canon_rejects_unsorted_individual():
    from magmaclust_b200.magma import Canon
    db = {"ID": np.array([1, 1, 1]), "Timestamp": np.array([0.0, 2.0, 1.0]), "Output": np.zeros(3)}
    with pytest.raises(ValueError):
        Canon(db)


def test_only_reference_kernels_are_accepted():
    from magmaclust_b200 import magma
    with pytest.raises(ValueError):
        magma._check_kernels(lambda m, t: m)
    magma._check_kernels(magma.kernel, magma.kernel_mu)
    assert np.allclose(magma.kernel(np.array([0.0, 1.0]), (1.0, 0.5)), O.kernel(np.array([0.0, 1.0]), (1.0, 0.5)))


LBFGS_DRIVER = r"""
#include "lbfgs.h"
#include <cstdio>
using namespace mcb;
// f(x) = sum_i 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2  (chained Rosenbrock, D = 3)
static void fg(const double* x, double& f, double* g) {
  f = 0; g[0] = g[1] = g[2] = 0;
  for (int i = 0; i < 2; ++i) {
    double a = x[i + 1] - x[i] * x[i], b = 1 - x[i];
    f += 100 * a * a + b * b; g[i] += -400 * x[i] * a - 2 * b; g[i + 1] += 200 * a;
  }
}
int main() {
  Lbfgs o; double x0[3] = {-1.2, 1.0, 0.5}; lb_init(o, 3, x0);
  double f, g[3]; int st = LB_EVAL;
  while (st == LB_EVAL) { fg(o.x, f, g); st = lb_advance(o, f, g); }
  printf("%d %d %d %.17g %.17g %.17g %.17g\n", st, o.iter, o.nfev, o.x[0], o.x[1], o.x[2], o.f);
  return 0;
}
"""


def test_lbfgs_header_follows_lbfgsb_trajectory():
    """csrc/lbfgs.h (host + device) reproduces scipy's L-BFGS-B (m = 5, factr = 1e7, pgtol = 0) on an
    unconstrained problem: same iteration and evaluation counts, same minimiser"""
    from scipy.optimize import minimize
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        open(src, "w").write(LBFGS_DRIVER)
        exe = os.path.join(td, "t")
        subprocess.run(["g++", "-O2", "-I", os.path.join(ROOT, "magmaclust_b200", "csrc"), "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    st, it, nfev = int(out[0]), int(out[1]), int(out[2])
    x = np.array([float(v) for v in out[3:6]])

    def f(x):
        a = x[1:] - x[:-1] ** 2
        b = 1 - x[:-1]
        return np.sum(100 * a * a + b * b)

    def g(x):
        gr = np.zeros(3)
        for i in range(2):
            a = x[i + 1] - x[i] ** 2
            b = 1 - x[i]
            gr[i] += -400 * x[i] * a - 2 * b
            gr[i + 1] += 200 * a
        return gr

    r = minimize(f, [-1.2, 1.0, 0.5], jac=g, method="L-BFGS-B",
                 options={"maxcor": 5, "ftol": 1e7 * np.finfo(float).eps, "gtol": 0, "maxiter": 100})
    assert st == 1
    assert (it, nfev) == (r.nit, r.nfev)
    assert np.allclose(x, r.x, rtol=0, atol=1e-9)


LBFGS_TWIN_DRIVER = r"""
#include "lbfgs.h"
#include <cstdio>
#include <cstring>
#include <random>
using namespace mcb;
static int D;
static double A[3][3], bvec[3], c4, wall;
static void fg(const double* x, double& f, double* g) {
  f = 0; for (int i = 0; i < D; ++i) g[i] = 0;
  for (int i = 0; i < D; ++i) {
    double ax = 0; for (int j = 0; j < D; ++j) ax += A[i][j] * x[j];
    f += 0.5 * x[i] * ax - bvec[i] * x[i]; g[i] += ax - bvec[i];
  }
  for (int i = 0; i + 1 < D; ++i) { double a = x[i + 1] - x[i] * x[i]; f += c4 * a * a; g[i] += -4 * c4 * x[i] * a; g[i + 1] += 2 * c4 * a; }
  // outside the wall the objective is not finite (a covariance that stopped being positive definite): bisection branch
  for (int i = 0; i < D; ++i) if (fabs(x[i]) > wall) f = INFINITY;
}
int main() {
  std::mt19937_64 rng(5); std::normal_distribution<double> nd;
  int mism = 0, runs = 0, status_seen[8] = {0}; long long evals = 0;
  for (int rep = 0; rep < 4000; ++rep) {
    D = 2 + rep % 2;
    double R[3][3]; for (auto& r : R) for (auto& v : r) v = nd(rng);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = (i == j) ? 0.3 : 0.0; for (int k = 0; k < 3; ++k) A[i][j] += R[i][k] * R[j][k]; }
    for (auto& v : bvec) v = nd(rng);
    c4 = (rep % 3 == 0) ? 0.0 : 5.0 * fabs(nd(rng));
    wall = (rep % 5 == 0) ? 2.5 : 1e300;
    double x0[3] = {nd(rng), nd(rng), nd(rng)};
    Lbfgs a, b; memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    lb_init(a, D, x0, 100); lb_init(b, D, x0, 100);
    double f, g[3]; int sa = LB_EVAL, sb = LB_EVAL, n = 0;
    while (sa == LB_EVAL && n < 5000) {
      fg(a.x, f, g); sa = lb_advance_any(a, f, g);
      fg(b.x, f, g); sb = lb_advance(b, f, g);
      ++n;
      if (sa != sb || memcmp(a.x, b.x, sizeof(double) * D) || a.col != b.col || a.iter != b.iter || a.nfev != b.nfev) { ++mism; break; }
    }
    if (memcmp(&a.f, &b.f, 8)) ++mism;
    ++status_seen[sa & 7]; ++runs; evals += n;
  }
  printf("%d %d %lld", runs, mism, evals);
  for (int s = 0; s < 6; ++s) printf(" %d", status_seen[s]);
  printf("\n");
  return 0;
}
"""


def test_compile_time_d_lbfgs_walks_the_same_trajectory_as_the_generic_one():
    """lbfgs.h: lb_advance (D and the memory depth compile-time constants: the form the kernels run) against lb_advance_any
    (run-time D) on 4000 random problems, D = 2 and 3, a fifth of them with a non-finite region: identical iterates, counts
    and end states, bit for bit"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        open(src, "w").write(LBFGS_TWIN_DRIVER)
        exe = os.path.join(td, "t")
        subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "magmaclust_b200", "csrc"), "-o", exe, src], check=True)
        out = [int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    runs, mism, evals = out[:3]
    assert runs == 4000 and mism == 0 and evals > 20000
    assert out[3 + 1] > 3000          # most runs converge (LB_CONVERGED = 1)


def test_synth_configs_have_the_named_shapes():
    from magmaclust_b200 import synth
    ds = synth.config("C1")
    assert (ds["M"], ds["K"], ds["n"]) == (50, 3, 30) and ds["tau0"].shape == (50, 3)
    assert np.all(np.diff(ds["offsets"]) == 30)
    assert np.all(ds["tau0"].sum(1) == 1)
    seg = ds["grid_idx"][:30]
    assert np.all(np.diff(seg) > 0)
    ds2 = synth.config("C1")
    assert np.array_equal(ds["y"], ds2["y"])  # seeded
    sh = synth.shard(ds, 1, 2)
    assert sh["M"] == 25 and sh["offsets"][0] == 0 and sh["N"] == 201


def test_shard_plan_balances_cubic_cost():
    from magmaclust_b200 import shard
    n = np.array([10] * 50 + [40] * 10)
    b = shard.plan(n, 4)
    assert b[0] == 0 and b[-1] == 60 and np.all(np.diff(b) >= 1)
    cost = np.array([(n[b[r]:b[r + 1]] ** 3).sum() for r in range(4)], dtype=float)
    assert cost.max() / cost.mean() < 1.6
    assert np.array_equal(shard.plan(n, 1), [0, 60])


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
from oracle import magma_oracle as O
from helpers import small_problem
from magmaclust_b200 import shard
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
db, hp, tau, m_k = small_problem(M=11, seed=6)
per = O.split_db(db)
ids = list(per)
names = list(m_k)
grid = np.sort(O.unique_stable(db['Timestamp']))
b = shard.plan([len(per[i]['Timestamp']) for i in ids], world)
mine = ids[b[rank]:b[rank + 1]]
N, K = len(grid), len(names)
acc = np.zeros((K, N, N)); w = np.zeros((K, N))
for c, k in enumerate(names):
    inv_k = O.kern_to_inv(grid, O.kernel_mu, hp['theta_k'][k][:2], hp['theta_k'][k][2])
    acc[c] = shard.prior_scale(rank) * inv_k.m
    w[c] = shard.prior_scale(rank) * (inv_k.m @ np.repeat(m_k[k], N))
    for i in mine:
        d = per[i]
        inv_i = O.kern_to_inv(d['Timestamp'], O.kernel, hp['theta_i'][i][:2], hp['theta_i'][i][2])
        r = inv_k.pos(d['Timestamp'])
        acc[c][np.ix_(r, r)] += tau[k][i] * inv_i.m
        w[c][r] += tau[k][i] * (inv_i.m @ d['Output'])
ta, tw = torch.from_numpy(acc), torch.from_numpy(w)
dist.all_reduce(ta); dist.all_reduce(tw)
ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
ok = True
for c, k in enumerate(names):
    cov = np.linalg.inv(ta[c].numpy())
    mean = cov @ tw[c].numpy()
    ok &= np.max(np.abs(cov - ref['cov'][k].m)) / np.max(np.abs(ref['cov'][k].m)) < 1e-9
    ok &= np.max(np.abs(mean - ref['mean'][k]['Output'])) / np.max(np.abs(ref['mean'][k]['Output'])) < 1e-9
# pi_k = (sum over ranks of local tau sums) / M_total
s = torch.tensor([sum(tau[k][i] for i in mine) for k in names])
dist.all_reduce(s)
ok &= np.allclose(s.numpy() / len(ids), [np.mean(list(tau[k].values())) for k in names])
print('RANK', rank, 'OK' if ok else 'FAIL', flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_sharded_accumulation_world_size_2_gloo():
    """the N > 1 path on CPU: individuals split by shard.plan, prior precision contributed by rank 0 only, one
    all-reduce(sum) of (A_k, w_k) reproduces the single-process hyper-posterior"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "w.py")
        open(src, "w").write(GLOO_WORKER)
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
        procs = [subprocess.Popen([sys.executable, src, ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=300)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o
            assert "OK" in o
