"""CPU: host logic of the drop-in API (no GPU): state-dict inventory, LR schedules, optimizer factory,
error behaviour -- the parts of scene_synthesis/networks/__init__.py:15-169 that do not touch the device."""
import math
import os

import pytest
import torch
import yaml

from scene_synthesis.networks import (adjust_learning_rate, build_network, optimizer_factory, schedule_factory)
from diffuscene_b200.schedule import get_betas, make_tables
from diffuscene_b200.weights import NetSpec, unet1d_param_specs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(name="uncond/diffusion_bedrooms_instancond_lat32_v.yaml"):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", name)))
    cfg["network"]["diffusion_kwargs"]["loss_iou"] = False
    return cfg


def test_state_dict_keys_are_the_reference_keys():
    net, _, _ = build_network(30, 23, _cfg(), None, "cpu")
    sd = net.state_dict()
    spec = NetSpec.from_net_kwargs(_cfg()["network"]["net_kwargs"])
    want = {"positional_embedding": (12, 128)}
    want.update({n: shp for (n, shp, _) in unet1d_param_specs(spec)})
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    # a fresh module starts with identity norms like the reference's nn.GroupNorm / LayerNorm
    assert torch.all(sd["diffusion.model.downs.0.0.block1.norm.weight"] == 1)
    assert torch.all(sd["diffusion.model.downs.0.4.fn.norm.g"] == 1)
    net.load_state_dict(sd, strict=True)


def test_cpu_device_has_no_compute_path():
    net, _, validate = build_network(30, 23, _cfg(), None, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        net.sample(torch.zeros(1, 1, 64, 64), 12, 62, batch_size=1)


def test_unknown_network_type_and_schedules_raise():
    with pytest.raises(NotImplementedError):
        build_network(30, 23, {"network": {"type": "something_else"}}, None, "cpu")
    with pytest.raises(NotImplementedError):
        schedule_factory({"schedule": "exponential"})
    with pytest.raises(NotImplementedError):
        optimizer_factory({"optimizer": "LBFGS"}, [torch.nn.Parameter(torch.zeros(1))])
    with pytest.raises(NotImplementedError):
        get_betas("cosine", 1e-4, 0.02, 10)         # broken in the reference too


def test_lr_schedules_follow_the_reference_formulas():
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = optimizer_factory({"optimizer": "Adam", "lr": 2e-4}, p)
    assert isinstance(opt, torch.optim.Adam) and opt.defaults["weight_decay"] == 0.0
    s = schedule_factory({"schedule": "step", "lr": 2e-4, "lr_step": 10000, "lr_decay": 0.5})
    for epoch, want in ((0, 2e-4), (9999, 2e-4), (10000, 1e-4), (25000, 5e-5)):
        adjust_learning_rate(s, opt, epoch)
        assert opt.param_groups[0]["lr"] == pytest.approx(want)
    s = schedule_factory({"schedule": "lambda", "start_epoch": 10, "end_epoch": 20, "start_lr": 1e-3, "end_lr": 1e-4})
    assert s.get_learning_rate(5) == pytest.approx(1e-3)
    assert s.get_learning_rate(15) == pytest.approx(1e-3 * (0.5 + 0.5 * 0.1))
    assert s.get_learning_rate(30) == pytest.approx(1e-4)
    s = schedule_factory({"schedule": "warmupcosine", "warmup_epochs": 10, "total_epochs": 110, "lr": 2e-4, "min_lr": 1e-6})
    assert s.get_learning_rate(10) == pytest.approx(2e-4)
    assert s.get_learning_rate(60) == pytest.approx(1e-6 + (2e-4 - 1e-6) * 0.5 * (1 + math.cos(math.pi * 0.5)))


def test_schedule_tables_match_oracle_tables():
    from oracle import diffusion_ref as D
    for mean, var in (("v", "fixedsmall"), ("eps", "fixedlarge"), ("x0", "fixedsmall")):
        t = make_tables(get_betas("linear", 1e-4, 0.02, 1000), mean, var)
        o = D.make_schedule(D.make_betas("linear", 1e-4, 0.02, 1000), mean, var)
        assert torch.equal(t["sqrt_ac"], o.sqrt_ac) and torch.equal(t["coef1"], o.coef1)
        assert torch.equal(t["coef2"], o.coef2) and torch.equal(t["loss_weight"], o.loss_weight)
        sig = torch.exp(0.5 * D.step_logvar(o))
        sig[0] = 0
        assert torch.equal(t["sigma"], sig)
    w = get_betas("warm0.1", 1e-4, 0.02, 100)
    assert w[9] == pytest.approx(0.02) and w[0] == pytest.approx(1e-4) and (w[10:] == 0.02).all()
