"""Seeded class-level parity cases (the reference's `DiffusionSceneLayout_DDPM` public methods), shared by
tests/golden/make_golden_class.py (runs the unmodified reference) and the CPU / GPU tests of
diffuscene_b200.networks.  Everything is a function of the case seed; only outputs are stored."""
from __future__ import annotations

import copy
import json
import os
import zlib

import torch
import yaml

from diffuscene_b200.weights import seeded_tensor
from tests.cases import STATS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CLASS_CASES = {
    # shipped unconditional bedroom config; short schedule so that full loops stay cheap
    "cls_bed": dict(seed=31, B=2, N=12, point_dim=62, n_classes=23, kind="uncond", T=6,
                    yaml="uncond/diffusion_bedrooms_instancond_lat32_v.yaml"),
    # shipped re-arrangement config (fc_arrange_condition, 5 diffused channels)
    "cls_arr": dict(seed=32, B=2, N=12, point_dim=5, n_classes=23, kind="arrange", T=5,
                    yaml="rearrange/diffusion_bedrooms_instancond_lat32_v_rearrange.yaml"),
    # completion with the partial-scene condition MLP and the non-learnable instance embedding
    # (fc_partial_condition + fc_instance_condition; no shipped YAML sets these, the class supports them)
    "cls_part": dict(seed=33, B=2, N=12, point_dim=62, n_classes=23, kind="partial", T=5, partial_num_points=3,
                     yaml="uncond/diffusion_bedrooms_instancond_lat32_v.yaml",
                     override=dict(learnable_embedding=False, room_partial_condition=True, partial_num_points=3,
                                   partial_emb_dim=64),
                     net_override=dict(instanclass_dim=192)),
    # shipped text config: fc_text_f on (stand-in) BERT hidden states, cross-attention in the loop
    "cls_text": dict(seed=34, B=2, N=12, point_dim=62, n_classes=23, kind="uncond", T=5, text=True,
                     yaml="text/diffusion_bedrooms_instancond_lat32_v_bert.yaml"),
}


def class_config(case, stats_path):
    with open(os.path.join(ROOT, "config", case["yaml"])) as f:
        cfg = yaml.safe_load(f.read().replace("\r", ""))
    cfg = copy.deepcopy(cfg)
    with open(stats_path, "w") as f:
        json.dump(STATS, f)
    net = cfg["network"]
    net.update(case.get("override", {}))
    net["net_kwargs"].update(case.get("net_override", {}))
    net["diffusion_kwargs"]["time_num"] = case["T"]
    net["diffusion_kwargs"]["train_stats_file"] = stats_path
    return cfg


def class_state_dict(like_sd, seed):
    """Deterministic value for every key of a (reference or repo) state dict, a function of (key, shape, seed)
    only; norm scales / shifts are moved away from (1, 0) so that the affine paths are exercised."""
    out = {}
    den = {}
    for k, v in like_sd.items():
        if k.startswith(("bertmodel.", "clip_model.")):
            continue
        if k.startswith("diffusion.model."):
            den[k] = tuple(v.shape)
            continue
        kind = "n" if k == "positional_embedding" else ("w" if v.dim() >= 2 else "b:64")
        out[k] = seeded_tensor(k, tuple(v.shape), kind, seed)
    for k, shape in den.items():
        # kinds by name, as in weights.unet1d_param_specs
        if k.endswith("norm.weight") or k.endswith(".g"):
            kind = "g"
        elif k.endswith("norm.bias"):
            kind = "beta"
        elif k.endswith(".weight"):
            kind = "w"
        else:
            kind = "b:%d" % 512
        out[k] = seeded_tensor(k, shape, kind, seed)
    return out


def class_batch(case):
    """`sample_params` as the data layer hands them over."""
    g = torch.Generator().manual_seed(case["seed"] * 13 + 5)
    B, N = case["B"], case["N"]
    cls = torch.randint(0, 22, (B, N), generator=g)
    sp = dict(translations=torch.rand(B, N, 3, generator=g) * 2 - 1, sizes=torch.rand(B, N, 3, generator=g) * 2 - 1,
              angles=torch.nn.functional.normalize(torch.randn(B, N, 2, generator=g), dim=-1),
              class_labels=torch.nn.functional.one_hot(cls, 22).float() * 2 - 1,
              objfeats_32=torch.rand(B, N, 32, generator=g) * 2 - 1, room_layout=torch.zeros(B, 1, 64, 64))
    if case.get("text"):
        sp["description"] = ["a bedroom with a double bed and two nightstands", "wardrobe next to the desk"][:B]
    return sp


def fake_bert_hidden(texts, L=9):
    """Stand-in for the frozen encoder's last_hidden_state: [B, L, 768], a function of the strings only."""
    rows = []
    for s in texts:
        g = torch.Generator().manual_seed(zlib.crc32(s.encode()))
        rows.append(torch.empty(L, 768).normal_(generator=g))
    return torch.stack(rows)


def crafted_samples(case):
    """Network samples for the post-processing tests: the emptiness channel of scene 0 carries exact zeros,
    both signs and a negative zero."""
    g = torch.Generator().manual_seed(case["seed"] * 17 + 3)
    B, N, d = case["B"], case["N"], 62
    s = torch.randn(B, N, d, generator=g)
    e = 8 + 22 - 1
    s[0, :, e] = torch.tensor([0.0, -0.0, 0.5, -0.5, 1e-30, -1e-30, 2.0, -2.0, 0.0, 0.25, -0.25, -1.0])[:N]
    return s
