"""FP64 roofline probe (run on the GPU box): cuBLAS DGEMM through torch.matmul vs the library's
grouped DMMA GEMM, plus a device copy.  Prints one JSON line; bench.py reads profiles/fp64_peak.json."""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spdepy_b200._lib import check, lib


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = min(timed(lambda: torch.matmul(A, B), 3) for _ in range(3))
    out["cublas_dgemm_tflops"] = 2 * n ** 3 / ms / 1e9
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    for cfg, name in ((0, "dmma_128x128"), (1, "dmma_128x64"), (2, "dmma_64x64")):
        t = ctypes.c_float()
        best = 1e30
        for _ in range(3):
            check(lib.spde_gemm_single(cfg, 0, 0, 1 << 11, n, n, n, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n, 3,
                                       ctypes.byref(t), None))
            best = min(best, t.value)
        out[name + "_tflops"] = 2 * n ** 3 / best / 1e9
    for (M, N, K) in ((8192, 8192, 64), (8192, 64, 8192), (2048, 2048, 2048)):
        t = ctypes.c_float()
        check(lib.spde_gemm_single(0 if N > 64 else 1, 0, 0, 1 << 11, M, N, K, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n, 5,
                                   ctypes.byref(t), None))
        out["dmma_%dx%dx%d_tflops" % (M, N, K)] = 2 * M * N * K / t.value / 1e9
    x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ms = min(timed(lambda: y.copy_(x), 5) for _ in range(3))
    out["copy_gbs"] = 2 * x.numel() * 8 / ms / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
