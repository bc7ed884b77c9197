import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from zpc_b200 import api, synth
from zpc_b200.solver import MpmSolver
P = synth.elastic_cube(10, 32, jitter_F=0.03, jitter_C=0.3, seed=5)
P["v"][:] = P["v"] * 8.0
n0 = P["m"].shape[0]
P["m"] = (P["m"] * (1.0 + 0.1 * np.arange(n0) / n0)).astype(np.float32)
dx = P["dx"]
res = {}
for var in (4, 6):
    api.set_tuning(var, 1)
    sol = MpmSolver(P, dx, P["volume"], synth.DT * 10, synth.GRAVITY, mode=1, layout="binned", rebin_every=5)
    out = []
    for step in range(7):
        sol.substep()
        torch.cuda.synchronize()
        Q = sol.particles_host()
        o = np.argsort(Q["m"], kind="stable")
        out.append({k: Q[k][o] for k in "xvCF"})
        # grid after this step's p2g is gone; compare particles only
    res[var] = out
for step in range(7):
    a, b = res[4][step], res[6][step]
    for k in "xvCF":
        d = np.abs(a[k] - b[k]).reshape(n0, -1).max(1)
        i = int(d.argmax())
        print("step", step, k, "max abs diff %.3e at particle %d (val %s), count>1e-5*scale: %d" % (d.max(), i, a[k][i].ravel()[:3], int((d > 1e-5 * np.abs(a[k]).max()).sum())))
