"""Wall-clock breakdown of one logLike(grad=True, exact_grad=True) with a device sync after every phase
(diagnostic; run on the GPU box)."""
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
for _ in range(2):
    m.logLike(par, grad=True, exact_grad=True)
eng = m.engine
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print("%-28s %8.2f ms" % (label, (time.perf_counter() - t0) * 1e3)); return out
r, nobs = m.r, m._obs["nobs"]; obs, cnt = m._obs["nodes"], m._obs["cnt"]; tau = float(np.exp(par[-1]))
data = to_dev(m.data.reshape(nobs, r))
st = T("assemble", lambda: m._assemble(par))
Q = st["Q"]
T("factor Q", lambda: eng.factorize(0, Q)); T("logdet", lambda: eng.logdet(0))
T("factor Qc", lambda: eng.factorize(1, Q, cnt, tau))
mu = T("solve mu_c", lambda: eng.solve(1, eng.scatter_obs(data, obs, tau)))
T("quad+resid", lambda: (Engine.dot(mu, eng.q_apply(Q, mu)), Engine.residual_ss(data, mu, obs)))
Z = T("selinv Q", lambda: eng.selinv(0)); Zc = T("selinv Qc", lambda: eng.selinv(1))
W = T("W = (Z-Zc)*r/2", lambda: (Z - Zc) * (0.5 * r))
W = T("sddmm mu", lambda: eng.sddmm(mu, mu, -0.5, W))
T("grad_from_weights", lambda: m._grad_from_weights(st, W))
T("full logLike", lambda: m.logLike(par, grad=True, exact_grad=True))
