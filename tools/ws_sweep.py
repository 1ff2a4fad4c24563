"""cp.async ring (cfg 2 = 64x64, production of round 1) against the warp-specialised bulk-async kernel (cfg 3) through
spde_gemm_single, on the shapes of the supernodal schedules (N/N layout, C -= A*B^T)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spdepy_b200._lib import check, lib
n = 12288
A = torch.randn(n * 8192, dtype=torch.float64, device="cuda"); B = torch.randn(n * 8192, dtype=torch.float64, device="cuda")
C = torch.zeros(n * n, dtype=torch.float64, device="cuda")
for (M, N, K) in ((8192, 8192, 8192), (8192, 8192, 512), (10000, 10000, 2401), (12274, 64, 1024), (4097, 4097, 577), (8192, 2048, 128),
                  (2960, 64, 448), (1500, 1500, 512), (6000, 64, 64)):
    out = []
    for cfg, name in ((2, "ring 64x64"), (3, "ws 128x64")):
        t = ctypes.c_float(); best = 1e30
        for _ in range(3):
            check(lib.spde_gemm_single(cfg, 0, 0, 1 << 11, M, N, K, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n, 3, ctypes.byref(t), None))
            best = min(best, t.value)
        out.append("%s: %.2f TFLOP/s (%.3f ms)" % (name, 2.0 * M * N * K / best / 1e9, best))
    print((M, N, K), " | ".join(out), flush=True)
