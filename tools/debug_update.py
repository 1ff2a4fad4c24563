import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from helpers import load_golden, make_grids
import spdepy_b200 as sp
from spdepy_b200.engine import to_dev
d = load_golden("ad_iso_bc3_ext")
g, g0 = make_grids(d)
m0 = sp.model(grid=g0, spde=d["mod0_spde"], ha=d["ha"], anisotropic=d["ani"], bc=d["bc"], parameters=d["mod0_par"])
mod = sp.model(grid=g, spde=d["spde"], ha=d["ha"], anisotropic=d["ani"], bc=d["bc"], mod0=m0)
mod.mod.setQ(d["par"]); mod.setModel()
eng = mod.mod.engine; n = eng.n
def diag(tag):
    torch.cuda.synchronize()
    dd = mod._Qdev[21*n:22*n].cpu().numpy()
    print(tag, "nonzero diag:", (dd != 0).sum(), "of", n, dd[:3])
diag("after setModel")
X = mod.sample(n=4, seed=3, simple=True); diag("after sample")
nodes = np.asarray(g.obs_nodes(d["idx"]), dtype=np.int64)
cnt = to_dev(np.bincount(nodes, minlength=n).astype(np.float64))
eng.add_diag(mod._Qdev, cnt, mod.tau); diag("after add_diag")
eng.factorize(0, mod._Qdev); diag("after factorize")
b = eng.scatter_obs(to_dev(np.ones((nodes.size,1))), to_dev(nodes, torch.int64), 1.0); diag("after scatter")
eng.solve(0, b); diag("after solve")
print("Q export diag nonzero:", (eng.to_scipy(mod._Qdev).diagonal() != 0).sum())
