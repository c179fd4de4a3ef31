"""Print ms/step and the per-phase split of one bench.py run (helper for environment-variable sweeps)."""
import json
import subprocess
import sys

out = subprocess.run([sys.executable, "bench.py", "--skip-cpu-baseline", "--no-extras"] + sys.argv[1:], capture_output=True, text=True)
if out.returncode != 0:
    print("bench failed:", out.stderr[-2000:])
    sys.exit(1)
d = json.loads(out.stdout.strip().splitlines()[-1])
p = d["roofline"]["phase_ms_per_step"]
print("ms/step %.3f  e2e %.3f  theta_i %.3f  theta_k %.3f  estep %.3f  elbo %.3f  evals %s" % (
    d["ms_per_step"], d["e2e"]["ms_per_step"], p["mstep_indiv"], p["mstep_mean"],
    p["indiv_factor_scatter"] + p["dense_hyperpost"] + p["indiv_tau"], p["elbo"], d["mstep_evals_per_step"]))
