"""Wall-clock breakdown of one logLike(grad=True, exact_grad=True) with a device sync after every phase
(diagnostic; run on the GPU box).  The phases are run serially here; the real logLike overlaps the prior's 2-D
work with the 3-D factorisation and the k=1 solve with the Takahashi pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bench
from spdepy_b200.engine import Engine, to_dev

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
inp = bench.make_inputs(name)
mod = bench.build_ours(inp); m = mod.mod
m.initFit(inp["data"], idx=inp["idx"])
par = inp["theta"]
for _ in range(3):
    m.logLike(par, grad=True, exact_grad=True)
eng = m.engine
def T(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    print("%-34s %8.2f ms" % (label, best)); return out
r, nobs = m.r, m._obs["nobs"]; obs, cnt = m._obs["nodes"], m._obs["cnt"]; tau = float(np.exp(par[-1]))
data = to_dev(m.data.reshape(nobs, r))
st = T("assemble (host fields + K2/K3)", lambda: m._assemble(par))
Q = st["Q"]
T("factor Qc (3-D)", lambda: eng.factorize(1, Q, cnt, tau))
prior = T("prior: 2-D factors + selinvs", lambda: m._prior_collapsed(st, want_grad=True))
T("logdet Qc", lambda: eng.logdet(1))
mu = T("solve mu_c (k=1)", lambda: eng.solve(1, eng.scatter_obs(data, obs, tau)))
T("quad + resid", lambda: (Engine.dot(mu, eng.q_apply(Q, mu)), Engine.residual_ss(data, mu, obs)))
W = T("selinv Qc (Takahashi)", lambda: eng.selinv(1))
W = W * (-0.5 * r)
W = T("sddmm mu mu^T", lambda: eng.sddmm(mu, mu, -0.5, W), reps=1)
prior["c"] = 0.5 * r
T("grad_from_weights (adjoint, chain rule)", lambda: m._grad_from_weights(st, W, prior))
T("full logLike(grad, exact)", lambda: m.logLike(par, grad=True, exact_grad=True))
T("full logLike(grad=False)", lambda: m.logLike(par, grad=False))
