"""GPU-vs-interpreter comparison of the streamed evaluator, segment by segment (debug aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import plan_emulator as pe
from scipy import sparse
from spdepy_b200 import _lib
from spdepy_b200.engine import Engine, to_dev

M, N, T, bc, thr, k = [int(a) for a in sys.argv[1:7]]
warm = int(sys.argv[7]) if len(sys.argv) > 7 else 0
emul = int(sys.argv[8]) if len(sys.argv) > 8 else 1
eng = Engine.get(M, N, T, bc)
pat, n = eng.pattern, eng.n
rng = np.random.default_rng(M + k)
W = pat.to_csc(rng.normal(size=pat.nslots * n)); A = (W + W.T) * 0.5
A = sparse.csc_matrix(A + sparse.diags(np.abs(A).sum(axis=1).A1 + 1.0)); flat = pat.from_sparse(A)
rng = np.random.default_rng(k); cnt = np.zeros(n); cnt[rng.choice(n, n // 4, replace=False)] = 1.0; tau = 3.0
if warm:      # dirty the allocator first
    junk = torch.randn(warm * 1024 * 1024 // 8, dtype=torch.float64, device="cuda"); del junk; torch.cuda.empty_cache()
ooc = _lib.OocHandle(eng.plan, thr, True, True)
B = rng.normal(size=(n, k))
X = to_dev(B.copy()); Z = torch.empty(eng.nslots * n, dtype=torch.float64, device="cuda")
Qd, cd = to_dev(flat), to_dev(cnt)
ld = ooc.run(Qd.data_ptr(), cd.data_ptr(), tau, X.data_ptr(), k, 15, Z.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("logdet gpu %.12f" % ld, flush=True)
if emul:
    em = pe.OocEmulator(eng.plan, ooc)
    ld_e, X_e, Z_e = em.evaluate(flat, cnt, tau, B, 15, True)
    print("logdet emu %.12f" % ld_e)
    for s in em.order:
        g = ooc.info_d(16 + int(s)); e = em.last_ld[int(s)]
        flag = "" if abs(g - e) <= 1e-10 * max(1, abs(e)) else "   <-- differs"
        print("seg %3d top %d root %4d par %3d cols %5d-%5d u %7d  gpu %.10f emu %.10f%s" % (s, em.segs[s]["top"], em.segs[s]["root"], em.segs[s]["parent"], em.segs[s]["col0"], em.segs[s]["col1"], em.segs[s]["u_size"], g, e, flag))
    print("solve relerr", np.abs(X.cpu().numpy() - X_e).max() / np.abs(X_e).max(), "Z relerr", np.abs(Z.cpu().numpy() - Z_e).max() / np.abs(Z_e).max())
