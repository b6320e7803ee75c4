"""Class-level golden vectors: the UNMODIFIED reference `DiffusionSceneLayout_DDPM`
(scene_synthesis/networks/diffusion_scene_layout_ddpm.py:14-482) run on CPU / fp32 through its public methods.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden_class.py
Writes tests/golden/cls_*.npz.  Pins what the DiffusionPoint-level goldens (make_golden.py) do not reach:
the attribute pack order of get_loss, the conditioning builders (positional embedding / fc_instance_condition,
fc_partial_condition, fc_arrange_condition, fc_text_f), sample()'s dispatch, and
delete_empty_from_network_samples / delete_empty_boxes (including a score of exactly 0.0).

Randomness: the reference draws t / noise / x_T / step noise from torch's global generator.  While a reference
method runs, torch.randn / torch.randint are replaced by recording wrappers around a seeded stream; the recorded
draws are stored next to the outputs so that the tests can inject exactly the same numbers.
Weights: a deterministic function of the parameter names (diffuscene_b200.weights.seeded_tensor), loaded with
load_state_dict(strict=True).  The frozen BERT encoder of the text config (no checkpoint offline) is replaced by a
deterministic stand-in that returns seeded `last_hidden_state` [B, L, 768]: what is pinned is everything after it.
"""
import copy
import json
import os
import sys
import types

import numpy as np
import torch
import yaml

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

from tests.cases_class import CLASS_CASES, class_config, class_state_dict, class_batch, fake_bert_hidden, \
    crafted_samples   # noqa: E402
from tests.cases import noise_stream   # noqa: E402


class Recorder:
    """Replaces torch.randn / torch.randint by seeded, recorded draws for the duration of a `with` block."""

    def __init__(self, seed):
        self.seed = seed
        self.randn_draws, self.randint_draws = [], []

    def __enter__(self):
        self._randn, self._randint = torch.randn, torch.randint
        rec = self

        def randn(*size, **kw):
            if "size" in kw:
                size = kw["size"]
            if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
                size = tuple(size[0])
            g = torch.Generator().manual_seed(rec.seed * 7919 + len(rec.randn_draws))     # = tests.cases.noise_stream
            out = rec._randn(tuple(size), generator=g, dtype=torch.float32)
            rec.randn_draws.append(out.clone())
            return out

        def randint(low, high, size, **kw):
            g = torch.Generator().manual_seed(977 + len(rec.randint_draws))
            out = rec._randint(low, high, size, generator=g)
            rec.randint_draws.append(out.clone())
            return out

        torch.randn, torch.randint = randn, randint
        # the samplers bind `noise_fn=torch.randn` as a DEFAULT ARGUMENT at import time (diffusion_ddpm.py:356,374,
        # 448,479,775-799): swap those defaults too (runtime objects only; the reference files are untouched)
        import scene_synthesis.networks.diffusion_ddpm as M
        self._patched = []
        for cls in (M.GaussianDiffusion, M.DiffusionPoint):
            for fn in vars(cls).values():
                d = getattr(fn, "__defaults__", None)
                if d and any(x is self._randn for x in d):
                    self._patched.append((fn, d))
                    fn.__defaults__ = tuple(randn if x is self._randn else x for x in d)
        return self

    def __exit__(self, *a):
        torch.randn, torch.randint = self._randn, self._randint
        for fn, d in self._patched:
            fn.__defaults__ = d


def build_reference(case):
    import transformers
    from scene_synthesis.networks.diffusion_scene_layout_ddpm import DiffusionSceneLayout_DDPM

    class _Tok:
        def __call__(self, text, return_tensors="pt", padding=True):
            class _Enc(dict):
                def to(self, device):
                    return self
            return _Enc(text=list(text))

    class _Bert(torch.nn.Module):
        def forward(self, text=None, **kw):
            out = types.SimpleNamespace()
            out.last_hidden_state = fake_bert_hidden(text)
            return out

    orig = (transformers.BertTokenizer.from_pretrained, transformers.BertModel.from_pretrained)
    transformers.BertTokenizer.from_pretrained = staticmethod(lambda *a, **k: _Tok())
    transformers.BertModel.from_pretrained = staticmethod(lambda *a, **k: _Bert())
    try:
        cfg = class_config(case, "/tmp/ds_b200_cls_stats.json")
        net = DiffusionSceneLayout_DDPM(case["n_classes"], None, cfg["network"])
    finally:
        transformers.BertTokenizer.from_pretrained, transformers.BertModel.from_pretrained = orig
    sd = class_state_dict(net.state_dict(), case["seed"])
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net, cfg


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    import contextlib
    import io
    for name, case in CLASS_CASES.items():
        net, cfg = build_reference(case)
        out = {}
        B, N, d = case["B"], case["N"], case["point_dim"]
        sp = class_batch(case)
        room = torch.zeros(B, 1, 64, 64)
        quiet = contextlib.redirect_stdout(io.StringIO())
        # ---- get_loss -------------------------------------------------------------------------------------
        with Recorder(case["seed"] + 1) as r, quiet:
            loss, ld = net.get_loss(sp)
        out["loss"] = np.asarray(loss.numpy())
        for k, v in ld.items():
            out["ld." + k] = np.asarray(v.numpy())
        out["loss_t"] = r.randint_draws[0].numpy()
        out["loss_noise"] = r.randn_draws[0].numpy()
        # ---- sampling through the public methods -------------------------------------------------------
        kw = {}
        if case.get("text"):
            kw["text"] = sp["description"]
        if case["kind"] == "arrange":
            boxes = torch.cat([sp["translations"], sp["sizes"], sp["angles"], sp["class_labels"], sp["objfeats_32"]], -1)
            with Recorder(case["seed"] + 2) as r, quiet:
                # `point_dim` here is the FULL attribute width: p_sample_loop_arrange asserts the interleaved result
                # against the shape it was given (diffusion_ddpm.py:505)
                s = net.sample(room, N, boxes.shape[-1], batch_size=B, input_boxes=boxes, clip_denoised=True)
            out["input_boxes"] = boxes.numpy()
        elif case["kind"] == "partial":
            full = torch.cat([sp["translations"], sp["sizes"], sp["angles"], sp["class_labels"], sp["objfeats_32"]], -1)
            partial = full[:, :case["partial_num_points"]].contiguous()
            with Recorder(case["seed"] + 2) as r, quiet:
                s = net.sample(room, N, d, batch_size=B, partial_boxes=partial, clip_denoised=True)
            out["partial_boxes"] = partial.numpy()
        else:
            with Recorder(case["seed"] + 2) as r, quiet:
                s = net.sample(room, N, d, batch_size=B, clip_denoised=True, **kw)
        out["sample"] = s.numpy()
        # draw 0 is the shape-only torch.randn at the top of sample() (reference :232); the rest are the loop's
        out["sample_draws"] = np.stack([a.numpy() for a in r.randn_draws[1:] if tuple(a.shape) == tuple(r.randn_draws[1].shape)])
        if case["kind"] == "partial":      # re-noising draws of the given objects (diffusion_ddpm.py:461), one per step
            out["sample_draws_partial"] = np.stack([a.numpy() for a in r.randn_draws[1:] if a.shape[1] == case["partial_num_points"]])
        if case["kind"] == "uncond":
            with Recorder(case["seed"] + 3) as r, quiet:
                traj = net.sample(room, N, d, batch_size=B, clip_denoised=True, ret_traj=True, freq=2, **kw)
            out["traj"] = np.stack([a.numpy() for a in traj])
            with Recorder(case["seed"] + 2) as r, quiet:
                # the reference's post-processing concatenates onto batch-1 buffers (:364-371): batch_size must be 1
                kw1 = {k: v[:1] for k, v in kw.items()}
                boxes = net.generate_layout(room[:1], N, d, batch_size=1, clip_denoised=True, **kw1)
            out["layout_draws"] = np.stack([a.numpy() for a in r.randn_draws[1:]])     # batch-1 draws of this call
            for k, v in boxes.items():
                out["layout." + k] = v.numpy()
        # ---- post-processing on a crafted batch (zeros, signs, ties) --------------------------------------
        if case["kind"] != "arrange":
            cs = crafted_samples(case)[:1]
            out["crafted"] = cs.numpy()
            for keep in (False, True):
                boxes = net.delete_empty_from_network_samples(cs, device="cpu", keep_empty=keep)
                for k, v in boxes.items():
                    out["del%d.%s" % (int(keep), k)] = v.numpy()
            td, sd_, bb, cd = 3, 3, net.bbox_dim, net.class_dim
            sdict = {"translations": cs[:, :, :td], "sizes": cs[:, :, td:td + sd_], "angles": cs[:, :, td + sd_:bb],
                     "class_labels": cs[:, :, bb:bb + cd]}
            if net.objfeat_dim > 0:
                sdict["objfeats"] = cs[:, :, bb + cd:]
            boxes = net.delete_empty_boxes(sdict, device="cpu")
            for k, v in boxes.items():
                out["delb." + k] = v.numpy()
        np.savez(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
