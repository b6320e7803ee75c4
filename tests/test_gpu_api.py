"""GPU: the drop-in Python surface (scene_synthesis.networks.build_network & co) on the CUDA engine."""
import copy
import json
import os

import numpy as np
import pytest
import torch
import yaml

from tests.cases import STATS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config(tmp_path, name="uncond/diffusion_bedrooms_instancond_lat32_v.yaml", T=20):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", name)))
    stats = tmp_path / "stats.json"
    stats.write_text(json.dumps(STATS))
    cfg["network"]["diffusion_kwargs"]["train_stats_file"] = str(stats)
    cfg["network"]["diffusion_kwargs"]["time_num"] = T
    return cfg


def _batch(B, device, seed=0):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randint(0, 22, (B, 12), generator=g)
    sp = dict(translations=torch.rand(B, 12, 3, generator=g) * 2 - 1, sizes=torch.rand(B, 12, 3, generator=g) * 2 - 1,
              angles=torch.nn.functional.normalize(torch.randn(B, 12, 2, generator=g), dim=-1),
              class_labels=torch.nn.functional.one_hot(cls, 22).float() * 2 - 1,
              objfeats_32=torch.rand(B, 12, 32, generator=g) * 2 - 1, room_layout=torch.zeros(B, 1, 64, 64))
    return {k: v.to(device) for k, v in sp.items()}


def test_build_network_sample_and_checkpoint_roundtrip(tmp_path):
    from scene_synthesis.networks import build_network
    cfg = _config(tmp_path)
    torch.manual_seed(0)
    net, train_on_batch, validate_on_batch = build_network(30, 23, cfg, None, device="cuda", precision="fp32")
    sd = net.state_dict()
    assert "positional_embedding" in sd and "diffusion.model.downs.0.0.mlp.1.weight" in sd and len(sd) == 389
    # checkpoint round trip through the reference file format
    path = tmp_path / "model_00000"
    torch.save(sd, path)
    net2, _, _ = build_network(30, 23, cfg, str(path), device="cuda", precision="fp32")
    room = torch.zeros(3, 1, 64, 64, device="cuda")
    a = net.sample(room, 12, 62, batch_size=3, clip_denoised=True, seed=5)
    b = net2.sample(room, 12, 62, batch_size=3, clip_denoised=True, seed=5)
    assert a.shape == (3, 12, 62) and torch.isfinite(a).all() and torch.equal(a, b)
    boxes = net.generate_layout(room[:1], 12, 62, batch_size=1, clip_denoised=True, device="cpu")
    n = boxes["class_labels"].shape[1]
    assert boxes["class_labels"].shape == (1, n, 21) and boxes["translations"].shape == (1, n, 3)
    assert boxes["angles"].shape == (1, n, 2) and boxes["objfeats"].shape == (1, n, 32)
    per_scene = net.delete_empty_batched(a)
    assert len(per_scene) == 3 and per_scene[0]["class_index"].dtype == torch.int64
    with pytest.raises(NotImplementedError):
        build_network(30, 23, {"network": {"type": "autoregressive_transformer"}}, None, "cuda")


def test_native_loss_matches_autograd_loss_and_training_step(tmp_path):
    from scene_synthesis.networks import build_network, optimizer_factory
    cfg = _config(tmp_path, T=1000)
    torch.manual_seed(1)
    net, train_on_batch, validate_on_batch = build_network(30, 23, cfg, None, device="cuda", precision="fp32")
    sp = _batch(4, "cuda")
    torch.manual_seed(7)
    native = validate_on_batch(net, sp, cfg)                  # CUDA-library p_losses (no grad)
    torch.manual_seed(7)
    loss, ld = net.get_loss(sp)                                # autograd functional path, same t / noise draws
    assert abs(native - float(loss)) < 2e-4 * max(1.0, abs(native))
    opt = optimizer_factory(cfg["training"], net.parameters())
    w0 = net.state_dict()["diffusion.model.init_conv.weight"].clone()
    l1 = train_on_batch(net, opt, sp, cfg)
    l2 = train_on_batch(net, opt, sp, cfg)
    assert np.isfinite(l1) and np.isfinite(l2)
    assert not torch.equal(w0, net.state_dict()["diffusion.model.init_conv.weight"])
    # the engine picks up the updated weights
    room = torch.zeros(2, 1, 64, 64, device="cuda")
    out = net.sample(room, 12, 62, batch_size=2, clip_denoised=True, ddim=True, ddim_steps=5, seed=1)
    assert torch.isfinite(out).all()


def test_text_and_arrange_configs_run(tmp_path):
    from scene_synthesis.networks import build_network
    cfg = _config(tmp_path, "text/diffusion_bedrooms_instancond_lat32_v_bert.yaml", T=8)
    cfg["network"]["diffusion_kwargs"]["loss_iou"] = False
    net, _, _ = build_network(30, 23, cfg, None, device="cuda")
    room = torch.zeros(2, 1, 64, 64, device="cuda")
    prefix = torch.randn(2, 9, 512, device="cuda")            # projected text prefix (BERT is outside the loop)
    out = net.sample(room, 12, 62, batch_size=2, text=prefix, clip_denoised=True, seed=3)
    assert out.shape == (2, 12, 62) and torch.isfinite(out).all()
    cfg = _config(tmp_path, "rearrange/diffusion_bedrooms_instancond_lat32_v_rearrange.yaml", T=8)
    cfg["network"]["diffusion_kwargs"]["loss_iou"] = False
    net, _, _ = build_network(30, 23, cfg, None, device="cuda")
    boxes = torch.rand(2, 12, 62, device="cuda") * 2 - 1
    out = net.sample(room, 12, 5, batch_size=2, input_boxes=boxes, clip_denoised=True, seed=3)
    assert out.shape == (2, 12, 62) and torch.equal(out[..., 3:6], boxes[..., 3:6]) and torch.equal(out[..., 8:], boxes[..., 8:])
