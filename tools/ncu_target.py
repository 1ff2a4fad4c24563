"""ncu target: warm everything, then bracket ONE numeric factorisation (or one Takahashi pass) with
cudaProfilerStart/Stop so that `ncu --profile-from-start off -k regex:k_gemm_grouped -s N -c M` counts GEMM
launches from the start of that schedule (plain launches: SPDE_GRAPHS=0)."""
import os, sys
os.environ["SPDE_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
what = sys.argv[2] if len(sys.argv) > 2 else "factor"
inp = bench.make_inputs(name)
mod = bench.build_ours(inp); m = mod.mod
st = m._assemble(inp["theta"])
eng = m.engine
eng.factorize(0, st["Q"])
if what == "selinv":
    eng.selinv(0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if what == "selinv":
    eng.selinv(0)
else:
    eng.factorize(0, st["Q"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", eng.logdet(0))
