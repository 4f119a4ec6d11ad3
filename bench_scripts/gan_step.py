"""a few eager GAN iterations (examples/t4_40b.4th:60-67, N=1024) for an ncu launch list: `ncu --metrics gpu__time_duration.sum ... python bench_scripts/gan_step.py`;
prints the launch count per iteration so that the last iteration can be cut out of the list"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4, host as th
th.init(0)
L = t4.load()
N = 1024
L.t4k_rand_seed(4321)
D, G = th.gan_discriminator(N, 0.3), th.gan_generator(N)
rng = np.random.default_rng(200)
real = th.Tensor.from_numpy((rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32))
REAL, FAKE = th.Tensor.tensor(N, 1, 1, 1, np.ones((N, 1), np.float32)), th.Tensor.tensor(N, 1, 1, 1, np.zeros((N, 1), np.float32))
z1, z2 = th.Tensor.tensor(N, 128, 1, 1), th.Tensor.tensor(N, 128, 1, 1)
its = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(its):
    n0 = L.t4k_launch_count()
    z1.randn(); z2.randn()
    th.gan_iteration(D, G, real, z1, z2, REAL, FAKE, losses=False)
    th.sync()
    print("iteration %d: %d launches" % (i, L.t4k_launch_count() - n0), flush=True)
