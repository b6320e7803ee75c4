"""Generate golden vectors by running the UNMODIFIED reference (CPU, fp32) on seeded inputs.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The reference ships no golden vectors, so these are
the pin for oracle/ and, through it, for the CUDA path.  Weights are not stored: they are a
deterministic function of parameter names (diffuscene_b200.weights.seeded_state_dict), loaded
into the reference with load_state_dict(strict=True), which also checks the key inventory.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

for _name, _attrs in [("tkinter", {}), ("tkinter.messagebox", {"NO": 0}), ("tkinter.tix", {"Tree": object}),
                      ("clip", {})]:
    _m = types.ModuleType(_name)
    for _k, _v in _attrs.items():
        setattr(_m, _k, _v)
    sys.modules[_name] = _m
sys.path.insert(0, "/root/reference")

from scene_synthesis.networks.denoise_net import Unet1D                     # noqa: E402
from scene_synthesis.networks.diffusion_ddpm import DiffusionPoint          # noqa: E402

from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs   # noqa: E402
from tests.cases import CASES, make_inputs, noise_stream, STATS                        # noqa: E402


def build_reference(case):
    net_kwargs = dict(case["net_kwargs"])
    cfg = dict(case["net_cfg"])
    stats_path = os.path.join("/tmp", "ds_b200_stats.json")
    with open(stats_path, "w") as f:
        json.dump(STATS, f)
    dk = dict(case["diffusion_kwargs"])
    if dk.get("loss_iou", False):
        dk["train_stats_file"] = stats_path
    unet = Unet1D(**net_kwargs)
    dp = DiffusionPoint(denoise_net=unet, config=cfg, **dk)
    spec = NetSpec.from_net_kwargs(net_kwargs)
    # seeded under the checkpoint names ("diffusion.model.*"); DiffusionPoint itself sees "model.*"
    sd = seeded_state_dict(unet1d_param_specs(spec), seed=case["seed"])
    dp.load_state_dict({k[len("diffusion."):]: v for k, v in sd.items()}, strict=True)
    dp.eval()
    return dp, spec


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    for name, case in CASES.items():
        dp, spec = build_reference(case)
        gd = dp.diffusion
        inp = make_inputs(case, spec)
        out = {}
        x, t, ctx, cross = inp["x"], inp["t"], inp["context"], inp["context_cross"]
        out["fwd"] = dp._denoise(x, t, ctx, cross).numpy()

        # one reverse step at the sampled t, with and without clipping
        nz = noise_stream(case["seed"] + 77)
        fn = lambda size, dtype=None, device=None: nz(tuple(size))
        out["step_clip"] = gd.p_sample(dp._denoise, x, t, ctx, cross, fn, clip_denoised=True).numpy()
        nz = noise_stream(case["seed"] + 77)
        out["step_noclip"] = gd.p_sample(dp._denoise, x, t, ctx, cross, fn, clip_denoised=False).numpy()
        t0 = torch.zeros_like(t)
        nz = noise_stream(case["seed"] + 78)
        out["step_t0"] = gd.p_sample(dp._denoise, x, t0, ctx, cross, fn, clip_denoised=True).numpy()

        # short full loop (T = case time_num, kept small for loop cases)
        if case.get("loop", False):
            shape = tuple(x.shape)
            nz = noise_stream(case["seed"] + 100)
            if case["net_cfg"].get("room_arrange_condition", False):
                boxes = inp["boxes"]
                full_shape = tuple(boxes.shape)
                out["loop_arrange"] = gd.p_sample_loop_arrange(
                    dp._denoise, full_shape, "cpu", ctx, cross, noise_fn=fn, clip_denoised=True,
                    input_boxes=boxes).numpy()
            else:
                out["loop"] = gd.p_sample_loop(dp._denoise, shape, "cpu", ctx, cross, noise_fn=fn,
                                               clip_denoised=True).numpy()
                nz = noise_stream(case["seed"] + 100)
                traj = gd.p_sample_loop_trajectory(dp._denoise, shape, "cpu", 4, ctx, cross, noise_fn=fn,
                                                   clip_denoised=True)
                out["traj"] = np.stack([a.numpy() for a in traj])
                nz = noise_stream(case["seed"] + 101)
                partial = inp["partial"]
                out["loop_complete"] = gd.p_sample_loop_complete(
                    dp._denoise, shape, "cpu", ctx, cross, noise_fn=fn, clip_denoised=True,
                    partial_boxes=partial).numpy()

        # training loss
        if case.get("loss", False):
            losses, ld = gd.p_losses(dp._denoise, inp["x0"], inp["t_loss"], inp["noise_loss"], ctx, cross)
            out["losses"] = losses.numpy()
            for k, v in ld.items():
                out["ld." + k] = np.asarray(v.numpy())
        path = os.path.join(HERE, name + ".npz")
        np.savez(path, **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
