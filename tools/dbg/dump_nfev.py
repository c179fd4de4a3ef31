"""debug: per-individual L-BFGS evaluation counts of one C2 M-step (input of the scheduling simulation, tools/dbg/sched_sim.py)"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from magmaclust_b200 import synth
from magmaclust_b200.engine import Engine
c = synth.CONFIGS["C2"]
ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
e = Engine(0)
e.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
K, M = c["K"], c["M"]
e.set_state(np.tile([1.0, 1.0, 0.2], (K, 1)), np.tile([1.0, 1.0, 0.2], (M, 1)), ds["tau0"].mean(0), np.zeros(K), ds["tau0"])
e.e_step()
e.m_step(True, False)
nfev, nit, st = e.last_mstep_indiv()
print(json.dumps({"nfev": nfev.tolist(), "status": np.bincount(st).tolist()}))
