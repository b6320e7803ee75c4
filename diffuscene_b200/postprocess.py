"""Batched attribute scaling around the sampler, on whatever device the tensors live on.

The reference encodes / decodes scene attributes on the CPU with numpy, one scene at a time
(scene_synthesis/datasets/threed_front_dataset.py: `Scale_CosinAngle_ObjfeatsNorm.scale` :483-489, `.descale`
:491-495, `__getitem__` :497-513, `post_process` :515-535).  Here the same arithmetic runs on the whole [B, N, .]
batch that `network.sample()` returns, so finished scenes never leave the GPU between the sampler and retrieval:

  encode   translations / sizes: clip to [lo, hi], map to [-1, 1];  angles: theta -> (cos, sin);
           objfeats: bounds entries 1 and 2 are (lo, hi) (entry 0 is the std the reference ignores here)
  decode   the inverse: angles = atan2(sin, cos); class labels, layouts and text pass through.

`bounds[k]` is a sequence of array-likes exactly as the reference's dataset objects hold them.
"""
from __future__ import annotations

from typing import Dict, Mapping, Sequence

import torch

_PASS_THROUGH = ("room_layout", "class_labels", "relations", "description", "desc_emb")
_OBJFEAT_KEYS = ("objfeats", "objfeats_32")


def _lo_hi(bounds: Mapping[str, Sequence], key: str, like: torch.Tensor):
    b = bounds[key]
    lo, hi = (b[1], b[2]) if key in _OBJFEAT_KEYS else (b[0], b[1])
    as_t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=like.device)
    return as_t(lo), as_t(hi)


def scale(x: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """clip to [lo, hi] then affine map onto [-1, 1] (reference :483-489)."""
    x = torch.minimum(torch.maximum(x.float(), lo), hi)
    return 2.0 * ((x - lo) / (hi - lo)) - 1.0


def descale(x: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """inverse of `scale` without the clip (reference :491-495)."""
    return (x + 1.0) / 2.0 * (hi - lo) + lo


def encode_batch(params: Mapping[str, torch.Tensor], bounds: Mapping[str, Sequence]) -> Dict[str, torch.Tensor]:
    """World-space attributes -> network space for a whole batch (reference `__getitem__` :497-513)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in params.items():
        if k == "angles":
            out[k] = torch.cat([torch.cos(v), torch.sin(v)], dim=-1)
        elif k in _OBJFEAT_KEYS or k in bounds:
            out[k] = scale(v, *_lo_hi(bounds, k, v))
        else:
            out[k] = v
    return out


def post_process_batch(params: Mapping[str, torch.Tensor], bounds: Mapping[str, Sequence]) -> Dict[str, torch.Tensor]:
    """Network-space samples -> world-space attributes for a whole batch (reference `post_process` :515-535)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in params.items():
        if k in _PASS_THROUGH:
            out[k] = v
        elif k == "angles":
            out[k] = torch.atan2(v[..., 1:2], v[..., 0:1])
        else:
            out[k] = descale(v, *_lo_hi(bounds, k, v))
    return out
