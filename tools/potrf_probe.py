"""Latency probe of k_potrf (run on the GPU box): microseconds per launch and SM-clock phase breakdown."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from spdepy_b200._lib import lib, check
torch.zeros(1, device="cuda")
for b, ld, nt in ((64, 64, 1), (64, 15000, 1), (64, 15000, 4), (32, 64, 1), (8, 64, 1), (64, 512, 296), (64, 512, 2960), (16, 128, 4000)):
    us = ctypes.c_float()
    clk = np.zeros(5, np.int64)
    check(lib.spde_potrf_bench(b, ld, nt, 200, ctypes.byref(us), clk.ctypes.data, None))
    d = np.diff(clk)
    print("b=%2d ld=%5d tasks=%4d: %7.2f us/launch; clocks load %d sweep %d inverse %d store %d (total %d)" % (b, ld, nt, us.value, d[0], d[1], d[2], d[3], clk[4] - clk[0]))
