"""Time the grouped DMMA GEMM for several tile shapes on square and skinny problems (tuning aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spdepy_b200._lib import check, lib
names = {0: "128x128 2x4w", 1: "128x64 4x2w", 2: "64x64 2x2w", 3: "128x64 2x2w", 4: "128x128 4x4w", 5: "64x128 2x4w", 6: "128x128 4x2w", 7: "256x64 8x2w", 8: "128x64 2x4w", 9: "64x64 bk32 s2", 10: "64x64 bk32 s3", 11: "64x64 bk16 s4", 12: "64x128 bk32 s2", 13: "64x64 bk8 s4", 14: "64x64 1x4w", 15: "64x64 4x1w"}
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda"); C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
for (M, N, K) in ((8192, 8192, 8192), (8192, 8192, 512), (8192, 64, 8192), (4096, 4096, 2048), (8192, 2048, 128)):
    out = []
    for cfg in (2, 5, 9, 10, 11, 12, 13, 14, 15):
        t = ctypes.c_float(); best = 1e30
        for _ in range(2):
            check(lib.spde_gemm_single(cfg, 0, 0, 1 << 11, M, N, K, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n, 3, ctypes.byref(t), None))
            best = min(best, t.value)
        out.append("%s: %.1f" % (names[cfg], 2.0 * M * N * K / best / 1e9))
    print((M, N, K), " | ".join(out))
