"""Probe (GPU box): does running two half-batches concurrently (two engines, two streams, two host threads)
hide kernel tails / launch overheads better than one full batch?"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffuscene_b200.engine import DenoiserEngine          # noqa: E402
from diffuscene_b200.schedule import get_betas, make_tables  # noqa: E402
from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs  # noqa: E402

BED = dict(dim=512, dim_mults=[1, 1, 1, 1], channels=62, objectness_dim=0, class_dim=22, angle_dim=2,
           objfeat_dim=32, context_dim=0, instanclass_dim=128, seperate_all=True)
spec = NetSpec.from_net_kwargs(BED)
sd = seeded_state_dict(unet1d_param_specs(spec), seed=0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
tables = make_tables(get_betas("linear", 1e-4, 0.02, 1000), "v", "fixedsmall")
ctx = torch.randn(12, 128).cuda()


def make():
    e = DenoiserEngine(spec, 12, 1000, precision="bf16")
    e.load_state_dict(sd)
    e.set_schedule(tables)
    e.set_context(ctx, shared=True)
    return e


def timed(fn):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


e0 = make()
one = timed(lambda: e0.sample(4096, num_steps=T, seed=1))
print("1 engine  x 4096 scenes, %d steps: %.3f s  -> %.1f scenes/s (1000-step equiv)" % (T, one, 4096 / one * T / 1000))
for nsplit in (2, 3, 4):
    engs = [e0] + [make() for _ in range(nsplit - 1)]
    per = 4096 // nsplit

    def run_all():
        th = [threading.Thread(target=lambda e=e, i=i: e.sample(per, num_steps=T, seed=1, scene_offset=i * per))
              for i, e in enumerate(engs)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    two = timed(run_all)
    print("%d engines x %d scenes concurrently: %.3f s -> %.1f scenes/s" % (nsplit, per, two, per * nsplit / two * T / 1000))
    for e in engs[1:]:
        e.close()
