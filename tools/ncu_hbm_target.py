"""ncu target for the HBM-bound kernels: one warm logLike, then ONE logLike(grad=True, exact_grad=True) bracketed by
cudaProfilerStart/Stop (plain launches, SPDE_GRAPHS=0), so that
    ncu --profile-from-start off --set full -k regex:'k_fill_spacetime|k_scatter_q|...' python tools/ncu_hbm_target.py c3
captures each assembly / reduction / helper kernel of exactly one evaluation."""
import os, sys
os.environ["SPDE_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
inp = bench.make_inputs(name)
mod = bench.build_ours(inp); m = mod.mod
m.initFit(inp["data"], idx=inp["idx"])
m.logLike(inp["theta"], grad=True, exact_grad=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
like, jac = m.logLike(inp["theta"], grad=True, exact_grad=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", like)
