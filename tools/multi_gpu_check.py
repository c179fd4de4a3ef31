"""Multi-GPU parity check (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/multi_gpu_check.py

Every rank runs its shard through the NCCL-coupled library; rank 0 additionally runs the SAME problem unsharded on
a second, communicator-free handle and compares: hyper-posterior means / covariances (rel 1e-8), responsibilities
(1e-6), hard assignments (identical), the ELBO and one full VEM iteration's hyper-parameters. Two shapes: a small
grid (replicated solves after an all-reduce) and a grid above 768 points (cluster ownership: reduce to the owner,
invert there, broadcast back).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclust_b200 import synth  # noqa: E402
from magmaclust_b200.engine import Engine  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = Engine(local)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.tensor(list(Engine.comm_unique_id()), dtype=torch.uint8)
    dist.broadcast(idt, 0)
    eng.comm_init(world, rank, bytes(idt.tolist()))
    ok = True
    for name, (M, K, n, n_grid, common_i) in {"small grid (all-reduce, replicated solves)": (240, 4, 24, 201, False),
                                               "large grid (cluster ownership)": (512, 4, 24, 1024, True)}.items():
        ds = synth.make_dataset(M, K, n, n_grid, common_i, seed=77)
        sh = synth.shard(ds, rank, world)
        rng = np.random.default_rng(5)
        theta_i = np.stack([rng.uniform(0.5, 1.5, M), rng.uniform(0.5, 1.5, M), rng.uniform(0.2, 0.5, M)], 1)
        if common_i:
            theta_i[:] = theta_i[0]
        theta_k = np.tile([1.2, 0.8, 0.3], (K, 1))
        w = rng.dirichlet(np.ones(K), M)
        pi = w.mean(0)
        per = (M + world - 1) // world
        lo, hi = rank * per, min(M, (rank + 1) * per)
        eng.set_data(sh["offsets"], sh["grid_idx"], sh["y"], sh["grid"], M_total=M)
        eng.set_state(theta_k, theta_i[lo:hi], pi, np.zeros(K), w[lo:hi])
        eng.e_step()
        mean, tau = eng.get_mean(), eng.get_tau_new()
        cov0 = eng.get_cov(K - 1)
        elbo = eng.elbo()
        it = eng.vem_iteration(True, common_i)
        tk_new, ti_new, pi_new = eng.get_new_hp()
        taus = [None] * world
        dist.all_gather_object(taus, tau)
        tis = [None] * world
        dist.all_gather_object(tis, ti_new)
        if rank == 0:
            ref = Engine(local)
            ref.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
            ref.set_state(theta_k, theta_i, pi, np.zeros(K), w)
            ref.e_step()
            rm, rt, rc = ref.get_mean(), ref.get_tau_new(), ref.get_cov(K - 1)
            relbo = ref.elbo()
            rit = ref.vem_iteration(True, common_i)
            rtk, rti, rpi = ref.get_new_hp()
            tau_all, ti_all = np.concatenate(taus), np.concatenate(tis)
            checks = {
                "mean": rel(mean, rm) < 1e-8, "cov": rel(cov0, rc) < 1e-8, "tau": np.max(np.abs(tau_all - rt)) < 1e-6,
                "hard": bool(np.array_equal(tau_all.argmax(1), rt.argmax(1))),
                "elbo": abs(elbo[1] - relbo[1]) < 1e-6 * abs(relbo[1]),
                "vem_elbo": abs(it[0] - rit[0]) < 1e-6 * abs(rit[0]),
                "theta_k": np.allclose(tk_new, rtk, atol=1e-4), "theta_i": np.allclose(ti_all, rti, atol=5e-3),  # optimiser end points: loose by construction (factr = 1e7)
                "pi": np.allclose(pi_new, rpi, atol=1e-12),
            }
            print(name, "world", world, {k: bool(v) for k, v in checks.items()},
                  "rel mean %.2e cov %.2e dtau %.2e" % (rel(mean, rm), rel(cov0, rc), np.max(np.abs(tau_all - rt))), flush=True)
            ok = ok and all(checks.values())
            ref.close()
        dist.barrier()
    eng.close()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-GPU PARITY", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
