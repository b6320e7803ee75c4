"""Bring-up aid (run on the GPU box): per-role cycle breakdown of the tcgen05 GEMM kernel for the hot shapes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffuscene_b200 import capi   # noqa: E402

lib = capi.load()
names = ["prod_wait_empty", "prod_total", "mma_wait_tmem", "mma_wait_full", "mma_total", "epi_wait_tfull", "epi_total", "tiles"]


def run(M, N, K, n_obj=0, res=False, tag=""):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    gamma = torch.ones(N).cuda()
    beta = torch.zeros(N).cuda()
    r = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda() if res else None
    d = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    tr = np.zeros((256, 8), dtype=np.uint64)
    us = C.c_float()
    rc = lib.ds_test_gemm_trace(a.data_ptr(), w.data_ptr(), bias.data_ptr(), None if r is None else r.data_ptr(),
                                d.data_ptr(), M, N, K, n_obj, gamma.data_ptr(), beta.data_ptr(), 20,
                                tr.ctypes.data, C.byref(us))
    if rc:
        print("FAILED", lib.ds_last_error().decode())
        return
    t = tr[:148].astype(np.float64)
    tf = 2.0 * M * N * K / us.value / 1e6
    print("%-28s M=%d N=%d K=%d n_obj=%d res=%d: %.1f us  %.0f TFLOP/s" % (tag, M, N, K, n_obj, res, us.value, tf))
    print("   " + "  ".join("%s=%.0f" % (n, t[:, i].mean()) for i, n in enumerate(names)))


cl = os.environ.get("DS_TC_CLUSTER", "1")
print("cluster size", cl)
M = 49152
run(M, 512, 512, tag="plain 512x512")
run(M, 512, 1024, tag="plain 512x1024")
run(M, 512, 512, res=True, tag="plain+res")
run(M, 3072, 512, tag="dec.l0")
run(M, 1024, 512, tag="enc.l1")
run(M, 512, 512, n_obj=12, tag="GN")
run(M, 512, 512, n_obj=12, res=True, tag="GN+res")
run(M, 512, 1024, n_obj=12, res=True, tag="GN+res K=1024")
run(M, 384, 512, tag="qkv (BN=128)")
run(M, 512, 128, tag="to_out K=128")
