"""Residual of K (K^-1 V) - V for the SE kernel matrix on n equispaced points (kern_to_inv through the C ABI)."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
from magmaclust_b200.engine import Engine
eng = Engine(0)
for n in [int(a) for a in sys.argv[1:]] or [401, 801, 1500, 2048, 4096, 8192]:
    x = np.linspace(0.0, 10.0, n)
    th, sg = np.array([1.0, 1.0]), 0.2
    Kinv, ld = eng.kern_to_inv(x, th, sg)
    d = x[:, None] - x[None, :]
    K = np.exp(th[0] - 0.5 * np.exp(-th[1]) * d * d)
    K[np.diag_indices(n)] = np.exp(th[0]) + sg ** 2
    V = np.random.default_rng(0).standard_normal((n, 4))
    R = K @ (Kinv @ V) - V
    sign, ldr = np.linalg.slogdet(K) if n <= 4096 else (1, float('nan'))
    print(n, 'resid', np.abs(R).max(), 'sym', np.abs(Kinv - Kinv.T).max(), 'logdet', ld, ldr, flush=True)
eng.close()
