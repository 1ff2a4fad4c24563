"""Pipeline depth of the 64x64 cp.async ring kernel on the SKINNY products of the panel chain (about one CTA per SM, so
the k-loop of a CTA is exposed to memory latency): spde_gemm_single, N/N layout, C -= A*B^T."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spdepy_b200._lib import check, lib
n = 16384
A = torch.randn(n * 1024, dtype=torch.float64, device="cuda"); B = torch.randn(n * 1024, dtype=torch.float64, device="cuda")
C = torch.zeros(n * 64, dtype=torch.float64, device="cuda")
names = {2: "k32 s2 (production)", 10: "k32 s3", 12: "k32 s4", 15: "k32 s6", 13: "k64 s2", 14: "k64 s3", 11: "k16 s4", 3: "ws 128x64"}
for (M, N, K) in ((15000, 64, 448), (15000, 64, 256), (15000, 64, 64), (8000, 64, 448), (8000, 64, 64), (2000, 64, 320), (2000, 64, 64),
                  (500, 64, 192), (4000, 448, 64), (15000, 512, 512)):
    out = []
    for cfg in (2, 10, 12, 15, 13, 14, 11, 3):
        t = ctypes.c_float(); best = 1e30
        for _ in range(3):
            check(lib.spde_gemm_single(cfg, 0, 0, 1 << 11, M, N, K, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n, 20, ctypes.byref(t), None))
            best = min(best, t.value)
        out.append("%s %.1f us" % (names[cfg], best * 1e3))
    print((M, N, K), " | ".join(out), flush=True)
