"""Parameter inventory of the denoiser and deterministic seeded initialisation.

The inventory reproduces the reference's state-dict key names and shapes
(reference: scene_synthesis/networks/denoise_net.py:336-482 builds the modules,
SURVEY.md appendix A.1 lists the resulting keys) so checkpoints interchange.
Nothing here touches the GPU.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass
class NetSpec:
    """Shape description of one Unet1D instance (reference denoise_net.py:336-362)."""

    dim: int = 512
    channels: int = 62
    seperate_all: bool = True
    objectness_dim: int = 0
    class_dim: int = 22
    translation_dim: int = 3
    size_dim: int = 3
    angle_dim: int = 2
    objfeat_dim: int = 32
    cond_dim: int = 128          # context_dim + instanclass_dim
    text_condition: bool = False
    text_dim: int = 512
    n_stages: int = 4
    heads: int = 4
    dim_head: int = 32
    groups: int = 8

    @property
    def bbox_dim(self) -> int:
        return self.translation_dim + self.size_dim + self.angle_dim

    @property
    def point_dim(self) -> int:
        if self.seperate_all:
            return self.bbox_dim + self.class_dim + self.objectness_dim + self.objfeat_dim
        return self.channels

    @property
    def time_dim(self) -> int:
        return self.dim * 4

    @property
    def attn_hidden(self) -> int:
        return self.heads * self.dim_head

    @staticmethod
    def from_net_kwargs(kw: dict) -> "NetSpec":
        """Map the reference's `net_kwargs` (config/*/*.yaml) onto a NetSpec.

        Keys the reference accepts and ignores (self_condition, merge_bbox,
        modulate_time_context_instanclass, learned_variance) are accepted and ignored here too.
        """
        kw = dict(kw)
        mults = list(kw.get("dim_mults", (1, 2, 4, 8)))
        if any(m != 1 for m in mults):
            raise NotImplementedError(
                "diffuscene_b200 supports dim_mults of all ones (every shipped config); got %r" % (mults,))
        if kw.get("learned_sinusoidal_cond", False) or kw.get("random_fourier_features", False):
            raise NotImplementedError("learned / random sinusoidal time embeddings are not supported")
        if kw.get("resnet_block_groups", 8) != 8:
            raise NotImplementedError("resnet_block_groups must be 8")
        if kw.get("learned_variance", False) and not kw.get("seperate_all", False):
            raise NotImplementedError("learned_variance is not supported")
        return NetSpec(
            dim=kw.get("dim", 256),
            channels=kw.get("channels", 3),
            seperate_all=kw.get("seperate_all", False),
            objectness_dim=kw.get("objectness_dim", 1),
            class_dim=kw.get("class_dim", 21),
            translation_dim=kw.get("translation_dim", 3),
            size_dim=kw.get("size_dim", 3),
            angle_dim=kw.get("angle_dim", 1),
            objfeat_dim=kw.get("objfeat_dim", 0),
            cond_dim=kw.get("context_dim", 256) + kw.get("instanclass_dim", 0),
            text_condition=kw.get("text_condition", False),
            text_dim=kw.get("text_dim", 256),
            n_stages=len(mults),
        )


ParamSpec = Tuple[str, Tuple[int, ...], str]   # (name, shape, kind)


def _conv(name: str, cout: int, cin: int, bias: bool = True) -> List[ParamSpec]:
    out = [(name + ".weight", (cout, cin, 1), "w")]
    if bias:
        out.append((name + ".bias", (cout,), "b:%d" % cin))
    return out


def _linear(name: str, cout: int, cin: int) -> List[ParamSpec]:
    return [(name + ".weight", (cout, cin), "w"), (name + ".bias", (cout,), "b:%d" % cin)]


def _resblock(name: str, cin: int, cout: int, emb: int) -> List[ParamSpec]:
    p: List[ParamSpec] = []
    p += _linear(name + ".mlp.1", 2 * cout, emb)
    for blk, ci in (("block1", cin), ("block2", cout)):
        p += _conv("%s.%s.proj" % (name, blk), cout, ci)
        p.append(("%s.%s.norm.weight" % (name, blk), (cout,), "g"))
        p.append(("%s.%s.norm.bias" % (name, blk), (cout,), "beta"))
    if cin != cout:
        p += _conv(name + ".res_conv", cout, cin)
    return p


def _linattn(name: str, dim: int, hidden: int) -> List[ParamSpec]:
    p: List[ParamSpec] = [(name + ".fn.norm.g", (1, dim, 1), "g")]
    p += _conv(name + ".fn.fn.to_qkv", 3 * hidden, dim, bias=False)
    p += _conv(name + ".fn.fn.to_out.0", dim, hidden)
    p.append((name + ".fn.fn.to_out.1.g", (1, dim, 1), "g"))
    return p


def _crossattn(name: str, dim: int, text_dim: int, hidden: int) -> List[ParamSpec]:
    p: List[ParamSpec] = [(name + ".fn.norm.g", (1, dim, 1), "g")]
    p += _conv(name + ".fn.fn.to_q", hidden, dim, bias=False)
    p += _conv(name + ".fn.fn.to_kv", 2 * hidden, text_dim, bias=False)
    p += _conv(name + ".fn.fn.to_out.0", dim, hidden)
    p.append((name + ".fn.fn.to_out.1.g", (1, dim, 1), "g"))
    return p


def _mlp3(name: str, d0: int, d1: int, d2: int, d3: int) -> List[ParamSpec]:
    return _conv(name + ".0", d1, d0) + _conv(name + ".2", d2, d1) + _conv(name + ".4", d3, d2)


def unet1d_param_specs(s: NetSpec, prefix: str = "diffusion.model.") -> List[ParamSpec]:
    """Ordered (name, shape, kind) list of every Unet1D parameter for spec `s`."""
    C, H = s.dim, s.attn_hidden
    p: List[ParamSpec] = []
    if s.seperate_all:
        if s.objectness_dim > 0:
            p += _mlp3("objectness_embedf", s.objectness_dim, C, 2 * C, C)
        if s.objfeat_dim > 0:
            p += _mlp3("objfeat_embedf", s.objfeat_dim, C, 2 * C, C)
        p += _mlp3("class_embedf", s.class_dim, C, 2 * C, C)
        p += _mlp3("bbox_embedf", s.bbox_dim, C, 2 * C, C)
        p += _conv("init_conv", C, C)
    else:
        p += _conv("init_conv", C, s.channels)
    p += _linear("time_mlp.1", s.time_dim, C)
    p += _linear("time_mlp.3", s.time_dim, s.time_dim)
    for i in range(s.n_stages):
        d = "downs.%d" % i
        p += _resblock(d + ".0", C, C, s.cond_dim)
        p += _resblock(d + ".1", C, C, s.time_dim)
        if s.text_condition:
            p += _crossattn(d + ".2", C, s.text_dim, H)
        p += _resblock(d + ".3", C, C, s.time_dim)
        p += _linattn(d + ".4", C, H)
        if i == s.n_stages - 1:
            p += _conv(d + ".5", C, C)
    p += _resblock("mid_block0", C, C, s.cond_dim)
    p += _resblock("mid_block1", C, C, s.time_dim)
    if s.text_condition:
        p += _crossattn("mid_attn_cross", C, s.text_dim, H)
    p.append(("mid_attn.fn.norm.g", (1, C, 1), "g"))
    p += _conv("mid_attn.fn.fn.to_qkv", 3 * H, C, bias=False)
    p += _conv("mid_attn.fn.fn.to_out", C, H)
    p += _resblock("mid_block2", C, C, s.time_dim)
    for i in range(s.n_stages):
        u = "ups.%d" % i
        p += _resblock(u + ".0", C, C, s.cond_dim)
        p += _resblock(u + ".1", 2 * C, C, s.time_dim)
        if s.text_condition:
            p += _crossattn(u + ".2", C, s.text_dim, H)
        p += _resblock(u + ".3", 2 * C, C, s.time_dim)
        p += _linattn(u + ".4", C, H)
        if i == s.n_stages - 1:
            p += _conv(u + ".5", C, C)
    p += _resblock("final_res_block", 2 * C, C, s.time_dim)
    if s.seperate_all:
        if s.objectness_dim > 0:
            p += _mlp3("objectness_hidden2output", C, 2 * C, C, s.objectness_dim)
        if s.objfeat_dim > 0:
            p += _mlp3("objfeat_hidden2output", C, 2 * C, C, s.objfeat_dim)
        p += _mlp3("class_hidden2output", C, 2 * C, C, s.class_dim)
        p += _mlp3("bbox_hidden2output", C, 2 * C, C, s.bbox_dim)
    else:
        p += _conv("final_conv", s.channels, C)
    return [(prefix + n, shp, k) for (n, shp, k) in p]


def seeded_tensor(name: str, shape: Tuple[int, ...], kind: str, seed: int, perturb_norms: bool = True) -> torch.Tensor:
    """Deterministic value for one parameter, a function of (name, shape, kind, seed) only.

    Distributions follow torch's defaults for the layer types (U(+-1/sqrt(fan_in)) for conv /
    linear weights and biases).  With perturb_norms (the parity tests), norm scales / shifts are moved away
    from (1, 0) so that the affine paths are exercised; without it they start at torch's (1, 0) like the
    reference's freshly constructed modules.
    """
    if not perturb_norms and kind == "g":
        return torch.ones(shape, dtype=torch.float32)
    if not perturb_norms and kind == "beta":
        return torch.zeros(shape, dtype=torch.float32)
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    if kind == "w":
        fan_in = int(math.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    if kind.startswith("b:"):
        bound = 1.0 / math.sqrt(int(kind[2:]))
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    if kind == "g":
        return 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "beta":
        return 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "n":
        return torch.randn(shape, generator=g, dtype=torch.float32)
    raise ValueError(kind)


def seeded_state_dict(specs: List[ParamSpec], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {n: seeded_tensor(n, shp, k, seed) for (n, shp, k) in specs}


def count_params(specs: List[ParamSpec]) -> int:
    return sum(int(math.prod(shp)) for (_, shp, _) in specs)
