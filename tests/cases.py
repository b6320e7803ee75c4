"""Seeded parity cases shared by the golden generator, the oracle tests and the GPU tests.

Each case names a Unet1D instantiation (the reference's `net_kwargs`), the `config` dict the
reference's GaussianDiffusion reads its attribute dims from, and its `diffusion_kwargs`.
Inputs are functions of the case seed only, so nothing but outputs needs to be stored.
"""
from __future__ import annotations

import torch

# synthetic dataset bounds for the IoU term (reference reads them from train_stats_file,
# diffusion_ddpm.py:137-151)
STATS = {
    "bounds_translations": [-2.7, 0.04, -2.75, 2.8, 3.6, 2.9],
    "bounds_sizes": [0.04, 0.02, 0.01, 2.9, 1.8, 2.6],
    "bounds_angles": [-3.14159, 3.14159],
}

_BED = dict(dim=512, dim_mults=[1, 1, 1, 1], channels=62, objectness_dim=0, class_dim=22, angle_dim=2,
            objfeat_dim=32, context_dim=0, instanclass_dim=128, seperate_all=True)
_BED_CFG = dict(objectness_dim=0, class_dim=22, angle_dim=2, objfeat_dim=32)


def _dk(T=1000, mean="v", var="fixedsmall", iou=False, sep=True):
    return dict(schedule_type="linear", beta_start=0.0001, beta_end=0.02, time_num=T, loss_type="mse",
                model_mean_type=mean, model_var_type=var, loss_separate=sep, loss_iou=iou,
                train_stats_file=None)


CASES = {
    # shipped bedroom config (config/uncond/diffusion_bedrooms_instancond_lat32_v.yaml)
    "bed62": dict(seed=11, B=3, N=12, net_kwargs=_BED, net_cfg=_BED_CFG,
                  diffusion_kwargs=_dk(iou=True), loss=True),
    # same network, short schedule, full sampling loops (plain / trajectory / completion)
    "bed62_loop": dict(seed=12, B=2, N=12, net_kwargs=_BED, net_cfg=_BED_CFG,
                       diffusion_kwargs=_dk(T=12), loop=True),
    # BASELINE.json's synthetic d=97 shape, eps-prediction
    "bed97": dict(seed=13, B=2, N=12,
                  net_kwargs=dict(_BED, channels=97, angle_dim=4, class_dim=23, objfeat_dim=64),
                  net_cfg=dict(objectness_dim=0, class_dim=23, angle_dim=4, objfeat_dim=64),
                  diffusion_kwargs=_dk(mean="eps", sep=False), loss=True),
    # living / dining room shape, x0-prediction, fixedlarge variance
    "liv65": dict(seed=14, B=2, N=21,
                  net_kwargs=dict(_BED, channels=65, class_dim=25),
                  net_cfg=dict(objectness_dim=0, class_dim=25, angle_dim=2, objfeat_dim=32),
                  diffusion_kwargs=_dk(mean="x0", var="fixedlarge", iou=True), loss=True),
    # reference default attribute dims: objectness channel on, no objfeat, scalar angle, 256-d context
    "obj29": dict(seed=17, B=2, N=12,
                  net_kwargs=dict(dim=512, dim_mults=[1, 1, 1, 1], channels=29, objectness_dim=1, class_dim=21,
                                  angle_dim=1, objfeat_dim=0, context_dim=256, instanclass_dim=0,
                                  seperate_all=True),
                  net_cfg=dict(objectness_dim=1, class_dim=21, angle_dim=1, objfeat_dim=0),
                  diffusion_kwargs=_dk(iou=True), loss=True),
    # text-conditioned bedroom: cross-attention on a synthetic [B, L, 512] prefix
    "text62": dict(seed=15, B=2, N=12, L=7,
                   net_kwargs=dict(_BED, text_condition=True, text_dim=512),
                   net_cfg=_BED_CFG, diffusion_kwargs=_dk(), loss=False),
    # the same text-conditioned network through full sampling loops (plain / trajectory / completion), short schedule
    "text62_loop": dict(seed=18, B=2, N=12, L=7,
                        net_kwargs=dict(_BED, text_condition=True, text_dim=512),
                        net_cfg=_BED_CFG, diffusion_kwargs=_dk(T=8), loop=True),
    # re-arrangement network: 5 diffused channels, joint (non-separate) head, 512-d per-object condition
    "arr5": dict(seed=16, B=2, N=12,
                 net_kwargs=dict(dim=512, dim_mults=[1, 1, 1, 1], channels=5, objectness_dim=0, class_dim=22,
                                 angle_dim=2, objfeat_dim=32, context_dim=0, instanclass_dim=512),
                 net_cfg=dict(_BED_CFG, room_arrange_condition=True),
                 diffusion_kwargs=_dk(T=10), loop=True, loss=True, shared_context=False),
}


def noise_stream(seed: int):
    """Counter-based deterministic N(0,1) source: call k uses torch.Generator(seed + k)."""
    state = {"k": 0}

    def draw(shape):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 7919 + state["k"])
        state["k"] += 1
        return torch.randn(tuple(shape), generator=g, dtype=torch.float32)

    return draw


def _clean_scene(g, B, N, spec):
    """A structurally valid clean x0: U(-1,1) boxes / feats, unit (cos, sin) angles, +-1 one-hot class."""
    tr = torch.rand(B, N, spec.translation_dim, generator=g) * 2 - 1
    sz = torch.rand(B, N, spec.size_dim, generator=g) * 2 - 1
    an = torch.randn(B, N, spec.angle_dim, generator=g)
    an = an / an.norm(dim=-1, keepdim=True)
    cls = torch.randint(0, spec.class_dim, (B, N), generator=g)
    oh = torch.nn.functional.one_hot(cls, spec.class_dim).float() * 2 - 1
    parts = [tr, sz, an, oh]
    if spec.objectness_dim > 0:
        parts.append(torch.rand(B, N, spec.objectness_dim, generator=g) * 2 - 1)
    if spec.objfeat_dim > 0:
        parts.append(torch.rand(B, N, spec.objfeat_dim, generator=g) * 2 - 1)
    return torch.cat(parts, dim=-1)


def make_inputs(case, spec):
    g = torch.Generator(device="cpu")
    g.manual_seed(case["seed"])
    B, N = case["B"], case["N"]
    T = case["diffusion_kwargs"]["time_num"]
    d = spec.point_dim
    out = {}
    out["x"] = torch.randn(B, N, d, generator=g)
    out["t"] = torch.randint(0, T, (B,), generator=g, dtype=torch.int64)
    if case.get("shared_context", True):
        out["context"] = torch.randn(N, spec.cond_dim, generator=g)[None].expand(B, N, spec.cond_dim).contiguous()
    else:
        out["context"] = torch.randn(B, N, spec.cond_dim, generator=g)
    out["context_cross"] = torch.randn(B, case["L"], spec.text_dim, generator=g) if spec.text_condition else None
    arrange = case["net_cfg"].get("room_arrange_condition", False)
    if arrange:
        from diffuscene_b200.weights import NetSpec
        full = NetSpec(class_dim=case["net_cfg"]["class_dim"], angle_dim=case["net_cfg"]["angle_dim"],
                       objfeat_dim=case["net_cfg"]["objfeat_dim"], objectness_dim=case["net_cfg"]["objectness_dim"])
        boxes = _clean_scene(g, B, N, full)
        out["boxes"] = boxes
        td, sd, ad = full.translation_dim, full.size_dim, full.angle_dim
        out["x0"] = torch.cat([boxes[..., :td], boxes[..., td + sd:td + sd + ad]], dim=-1).contiguous()
    else:
        out["x0"] = _clean_scene(g, B, N, spec)
        out["partial"] = out["x0"][:, :3, :].contiguous()
    out["t_loss"] = torch.randint(0, T, (B,), generator=g, dtype=torch.int64)
    out["noise_loss"] = torch.randn(out["x0"].shape, generator=g)
    return out
