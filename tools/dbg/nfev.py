import sys, numpy as np
sys.path.insert(0, '/root/repo')
from magmaclust_b200 import synth
from magmaclust_b200.engine import Engine
c = dict(synth.CONFIGS['C2'])
ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
sh = synth.shard(ds, 0, 1)
M, K = sh["M"], sh["K"]
eng = Engine(0)
eng.set_data(sh["offsets"], sh["grid_idx"], sh["y"], np.ascontiguousarray(sh["grid"]), M_total=M)
tk = np.tile(np.array([1.0, 1.0, 0.2]), (K, 1)); ti = np.tile(np.array([1.0, 1.0, 0.2]), (M, 1))
for it in range(2):
    eng.set_state(tk, ti, np.ascontiguousarray(ds["tau0"].mean(0)), np.zeros(K), np.ascontiguousarray(sh["tau0"]))
    eng.vem_iteration(True, False)
    print(eng.last_mstep_stats(), eng.last_timings()[0])
