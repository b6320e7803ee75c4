"""Bring-up aid (GPU box): alternate the device-resident and the host-buffer sampling calls and print each call's
wall time, to separate API overhead from clock drift."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from diffuscene_b200.engine import DenoiserEngine  # noqa: E402
from diffuscene_b200.schedule import get_betas, make_tables  # noqa: E402
from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs  # noqa: E402

kw, N, D, T, B, prec, F, label = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "bed_d97"]
T = int(os.environ.get("T", "300"))
spec = NetSpec.from_net_kwargs(kw)
eng = DenoiserEngine(spec, N, T, precision=prec)
eng.load_state_dict(seeded_state_dict(unet1d_param_specs(spec), seed=0))
eng.set_schedule(make_tables(get_betas("linear", 1e-4, 0.02, T), "v", "fixedsmall"))
pos = torch.randn(N, 128).pin_memory()
xh = torch.randn(B, N, D).pin_memory()
eng.set_context(pos.cuda(), shared=True)
xd = xh.cuda()
for i in range(2):
    eng.sample(B, x_init=xd, seed=i)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    eng.sample(B, x_init=xd, seed=10 + rep)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    eng.set_context(pos.to("cuda", non_blocking=True), shared=True)
    t2 = time.perf_counter()
    x = xh.to("cuda", non_blocking=True)
    t3 = time.perf_counter()
    out = eng.sample(B, x_init=x, seed=10 + rep, host_output=True)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("rep %d: resident %.1f ms | set_context %.1f ms, H2D %.1f ms, host-output sample %.1f ms | graph builds %d" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), eng.graph_build_count()))
