"""Helpers shared by the GPU parity tests (engines are cached per (case, precision, backend))."""
import os

import numpy as np
import torch

from diffuscene_b200.engine import DenoiserEngine
from diffuscene_b200.schedule import get_betas, make_tables
from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs
from tests.cases import CASES, make_inputs

_ENGINES = {}


def case_tables(case):
    dk = case["diffusion_kwargs"]
    betas = get_betas(dk["schedule_type"], dk["beta_start"], dk["beta_end"], dk["time_num"])
    return make_tables(betas, dk["model_mean_type"], dk["model_var_type"])


def get_engine(name, precision="fp32", backend="auto", fuse=None):
    key = (name, precision, backend, fuse)
    if key in _ENGINES:
        return _ENGINES[key]
    case = CASES[name]
    spec = NetSpec.from_net_kwargs(case["net_kwargs"])
    eng = DenoiserEngine(spec, case["N"], case["diffusion_kwargs"]["time_num"], precision=precision,
                         gemm_backend=backend, fuse_level=fuse)
    eng.load_state_dict(seeded_state_dict(unet1d_param_specs(spec), seed=case["seed"]))
    eng.set_schedule(case_tables(case))
    inp = make_inputs(case, spec)
    shared = case.get("shared_context", True)
    eng.set_context(inp["context"][0] if shared else inp["context"], shared=shared)
    if spec.text_condition:
        eng.set_context_cross(inp["context_cross"])
    _ENGINES[key] = (eng, case, spec, inp)
    # keep at most a few engines alive (each holds ~0.5-1 GB of weights + tables)
    while len(_ENGINES) > 4:
        k0 = next(iter(_ENGINES))
        _ENGINES.pop(k0)[0].close()
    return _ENGINES[key]


def gold(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def cuda(t):
    return None if t is None else t.cuda()
