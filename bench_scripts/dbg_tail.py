import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tensorforth_b200 import lib as t4
from oracle import oracle as orc
L = t4.load()
rng = np.random.default_rng(1)
rnd = lambda *s: (rng.random(s, dtype=np.float32) * 2 - 1).astype(np.float32)
p = lambda t: C.c_void_p(t.data_ptr())
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
for (N, E1, EH, E0) in ((32, 1960, 100, 10), (64, 1960, 100, 10), (512, 1960, 100, 10)):
    layer = t4.L_RELU
    X, W1, B1 = rnd(N, E1), rnd(EH, E1) * 0.05, rnd(EH)
    W2, B2 = rnd(E0, EH) * 0.3, rnd(E0)
    T = orc.onehot(np.arange(N) % E0, E0)
    Xd, W1d, B1d, W2d, B2d, Td = dev(X), dev(W1), dev(B1), dev(W2), dev(B2), dev(T)
    z = lambda *s: torch.zeros(*s, device="cuda")
    y1a, a1a, f1a, y2a, pa, pda = z(N, EH), z(N, EH), z(N, EH), z(N, E0), z(N, E0), z(N, E0)
    rc = L.t4k_linear_act_head_fwd(layer, p(Xd), p(W1d), p(B1d), p(y1a), p(a1a), p(f1a), 0.1, p(W2d), p(B2d), p(y2a), p(pa), p(pda), N, EH, E1, E0, None)
    nf = L.t4k_head_train_scratch_floats(layer, N, EH, E1, E0)
    y1b, a1b, f1b, y2b, pb, pdb = z(N, EH), z(N, EH), z(N, EH), z(N, E0), z(N, E0), z(N, E0)
    scratch = z(max(int(nf), 4)); ncta = C.c_int(0)
    rc2 = L.t4k_linear_act_head_train(layer, p(Xd), p(W1d), p(B1d), p(y1b), p(a1b), p(f1b), 0.1, p(W2d), p(B2d), p(y2b), p(pb), p(pdb), p(Td), p(scratch), C.byref(ncta), N, EH, E1, E0, None)
    torch.cuda.synchronize()
    print(N, "rc", rc, rc2, "nf", nf, "ncta", ncta.value, "Pdup maxdiff", float((pda - pdb).abs().max()), "F1 diff", float((f1a - f1b).abs().max()),
          "P row sums", pdb.sum(1)[:4].tolist(), "d check", float((pb - (pda - Td)).abs().max()))
