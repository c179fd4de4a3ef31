import sys, os
import numpy as np
sys.path.insert(0, '/root/repo')
from magmaclust_b200.engine import Engine
n = int(sys.argv[1]); mode = sys.argv[2]
eng = Engine(0)
x = np.linspace(0.0, 10.0, n)
Kinv, ld = eng.kern_to_inv(x, np.array([1.0, 1.0]), 0.2)
if mode == "save":
    np.save('/tmp/inv_ref.npy', Kinv)
else:
    ref = np.load('/tmp/inv_ref.npy')
    E = np.abs(Kinv - ref)
    B = 256
    nb = (n + B - 1) // B
    blk = np.array([[E[i*B:(i+1)*B, j*B:(j+1)*B].max() for j in range(nb)] for i in range(nb)])
    np.set_printoptions(linewidth=250, precision=1)
    print('max ref', np.abs(ref).max(), 'logdet', ld)
    print(np.log10(blk + 1e-300).round(0).astype(int))
eng.close()
