"""Batched scene collate for training: ragged scenes -> the padded `sample_params` batch.

The reference builds every training sample on the CPU, one scene at a time, inside DataLoader workers
(scene_synthesis/datasets/threed_front_dataset.py: `Permutation.__getitem__` :576-584 shuffles the objects of a scene,
`Diffusion.__getitem__` :888-925 drops the start-label column, pads every attribute to `max_length` rows -- class rows
with the end label, everything else with zeros -- and maps the class one-hots to {-1, +1}; `collate_fn` :927-935 is
torch's default stacking).  At 8 x 1024 scenes per iteration that path is DataLoader-bound, so here the same
transformation runs on the whole batch with a handful of device ops: one padded scatter per attribute, one batched
argsort for the permutation augmentation.

Input: a list of per-scene dicts (numpy arrays or tensors) with `class_labels [L, C + 1]` (one-hot including the start
and end label columns, as the reference's encoders emit them) and any of `translations / sizes / angles / objfeats /
objfeats_32 [L, .]`.  Output: dict of `[B, max_length, .]` float32 tensors on `device` plus `length [B]`.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional

import torch

_ROW_KEYS = ("translations", "sizes", "angles", "objfeats", "objfeats_32", "objectness")


def collate_scenes(scenes: List[Mapping[str, object]], max_length: int, device="cpu", permute: bool = False,
                   generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    B = len(scenes)
    dev = torch.device(device)
    lengths = torch.tensor([int(torch.as_tensor(s["class_labels"]).shape[0]) for s in scenes], dtype=torch.int64)
    if int(lengths.max()) > max_length:
        raise ValueError("a scene has %d objects, max_length is %d" % (int(lengths.max()), max_length))
    # flat (scene, slot) index of every real object; with `permute` the slots of a scene are a random permutation
    scene_id = torch.repeat_interleave(torch.arange(B), lengths)
    start = torch.cumsum(lengths, 0) - lengths
    slot = torch.arange(int(lengths.sum())) - start[scene_id]
    if permute:
        # rank of a random key inside its scene = a uniform random permutation of that scene's objects
        # (reference :576-584, np.random.permutation per scene); the key sort is one batched argsort
        key = torch.rand(int(lengths.sum()), generator=generator)
        order = torch.argsort(scene_id.double() * 2.0 + key.double())
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel())
        slot = rank - start[scene_id]
    scene_id, slot = scene_id.to(dev), slot.to(dev)

    def flat(key):
        return torch.cat([torch.as_tensor(s[key], dtype=torch.float32).reshape(int(l), -1) for s, l in zip(scenes, lengths)]).to(dev)

    out: Dict[str, torch.Tensor] = {"length": lengths.to(dev)}
    cl = flat("class_labels")
    cl = torch.cat([cl[:, :-2], cl[:, -1:]], dim=1)                    # drop the start label, keep the end label last
    C = cl.shape[1]
    padded = torch.zeros(B, max_length, C, device=dev)
    padded[:, :, -1] = 1.0                                               # padding rows carry the end label
    padded[scene_id, slot] = cl
    out["class_labels"] = padded * 2.0 - 1.0                             # one-hot -> {-1, +1}
    for k in _ROW_KEYS:
        if all(k in s for s in scenes):
            v = flat(k)
            p = torch.zeros(B, max_length, v.shape[1], device=dev)       # padding rows are zero
            p[scene_id, slot] = v
            out[k] = p
    return out
