"""GPU: the native training step (ds_train_step: forward + hand-written backward + fused Adam) against torch autograd
over the CPU ORACLE (oracle/unet1d_ref.py + oracle/diffusion_ref.py, themselves pinned to the reference's goldens):
loss value, the 9 dict entries and the gradient of EVERY parameter (SURVEY 8d gate iv), fp32 parity mode; bf16 mode
against the same gradients with a stated tolerance; optimizer step against torch.optim.Adam + clip_grad_norm_."""
import numpy as np
import pytest
import torch

from diffuscene_b200.engine import DenoiserEngine
from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs
from oracle import diffusion_ref as D
from oracle.unet1d_ref import unet1d_forward
from tests.cases import CASES, STATS, make_inputs
from tests.gpu_common import case_tables

pytestmark = pytest.mark.gpu


def _oracle_grads(case, spec, sd, inp, ctx):
    """autograd through the oracle: loss = p_losses(...).mean(); returns (losses, dict, grads by name, d(context))."""
    dk = case["diffusion_kwargs"]
    sched = D.make_schedule(D.make_betas(dk["schedule_type"], dk["beta_start"], dk["beta_end"], dk["time_num"]),
                            dk["model_mean_type"], dk["model_var_type"])
    cfg = case["net_cfg"]
    ls = D.LossSpec(class_dim=cfg["class_dim"], angle_dim=cfg["angle_dim"], objfeat_dim=cfg["objfeat_dim"],
                    objectness_dim=cfg["objectness_dim"], loss_separate=dk["loss_separate"], loss_iou=dk["loss_iou"],
                    bounds_translations=STATS["bounds_translations"], bounds_sizes=STATS["bounds_sizes"],
                    room_arrange_condition=cfg.get("room_arrange_condition", False))
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    c = ctx.clone().requires_grad_(True)
    B, N = case["B"], case["N"]
    cfull = c[None].expand(B, N, -1) if c.dim() == 2 else c
    with torch.enable_grad():
        den = lambda xx, tt: unet1d_forward(sdg, spec, xx, tt, cfull, None)
        losses, ld = D.p_losses(sched, ls, den, inp["x0"], inp["t_loss"], inp["noise_loss"])
        losses.mean().backward()
    return losses.detach(), ld, {k: v.grad for k, v in sdg.items()}, c.grad


def _native(case, spec, sd, inp, ctx, shared, prec):
    eng = DenoiserEngine(spec, case["N"], case["diffusion_kwargs"]["time_num"], precision=prec)
    eng.set_schedule(case_tables(case))
    layout = eng.flat_layout()
    total = sum(n for _, n in layout.values())
    flat = torch.empty(total, device="cuda")
    for k, (off, n) in layout.items():
        flat[off:off + n] = sd[k].reshape(-1).cuda()
    grads = torch.full((total,), float("nan"), device="cuda")
    dk = case["diffusion_kwargs"]
    bounds = STATS["bounds_translations"] + STATS["bounds_sizes"]
    losses, ld, dctx = eng.train_step(flat, inp["x0"].cuda(), inp["t_loss"].cuda(), inp["noise_loss"].cuda(), ctx.cuda(),
                                      shared, dk["loss_separate"], dk["loss_iou"], bounds, flat_grads=grads)
    torch.cuda.synchronize()
    out = {k: grads[off:off + n].reshape(sd[k].shape).cpu() for k, (off, n) in layout.items()}
    eng.close()
    return losses.cpu(), {k: float(v) for k, v in ld.items()}, out, dctx.cpu()


@pytest.mark.parametrize("name", ["bed62", "liv65", "obj29", "bed97", "arr5"])
def test_gradients_match_oracle_autograd_fp32(name):
    case = CASES[name]
    spec = NetSpec.from_net_kwargs(case["net_kwargs"])
    sd = seeded_state_dict(unet1d_param_specs(spec), seed=case["seed"])
    inp = make_inputs(case, spec)
    shared = case.get("shared_context", True)
    ctx = inp["context"][0] if shared else inp["context"]
    ref_losses, ref_ld, ref_g, ref_dc = _oracle_grads(case, spec, sd, inp, ctx)
    losses, ld, g, dc = _native(case, spec, sd, inp, ctx, shared, "fp32")
    np.testing.assert_allclose(losses.numpy(), ref_losses.numpy(), rtol=1e-3, atol=1e-4)
    for k, v in ref_ld.items():
        np.testing.assert_allclose(ld[k], float(v), rtol=1e-3, atol=1e-4, err_msg=k)
    assert set(g) == set(ref_g)
    worst = []
    for k in sorted(g):
        a, b = g[k], ref_g[k]
        assert torch.isfinite(a).all(), k
        scale = max(b.abs().max().item(), 1e-6)
        err = (a - b).abs().max().item() / scale
        worst.append((err, k))
    worst.sort(reverse=True)
    print("worst relative gradient errors:", worst[:5])
    assert worst[0][0] < 2e-3, worst[:8]
    np.testing.assert_allclose(dc.numpy(), ref_dc.numpy(), rtol=2e-3, atol=2e-3 * ref_dc.abs().max().item())


def test_gradients_bf16_mode_close_to_fp32_oracle():
    """bf16 storage of activations / activation gradients, fp32 accumulation and fp32 parameter gradients: the
    cosine between every parameter's gradient and the oracle's stays above 0.99 (measured: see the printed table)."""
    case = CASES["bed62"]
    spec = NetSpec.from_net_kwargs(case["net_kwargs"])
    sd = seeded_state_dict(unet1d_param_specs(spec), seed=case["seed"])
    inp = make_inputs(case, spec)
    ctx = inp["context"][0]
    _, _, ref_g, _ = _oracle_grads(case, spec, sd, inp, ctx)
    losses, ld, g, dc = _native(case, spec, sd, inp, ctx, True, "bf16")
    cos = []
    for k in sorted(g):
        a, b = g[k].flatten().double(), ref_g[k].flatten().double()
        if b.norm() < 1e-10:
            continue
        cos.append((float(a @ b / (a.norm() * b.norm() + 1e-30)), k))
    cos.sort()
    print("lowest gradient cosines (bf16 mode):", cos[:5])
    assert cos[0][0] > 0.97 and np.mean([c for c, _ in cos]) > 0.995


def test_train_on_batch_native_matches_torch_adam(tmp_path):
    """Two optimizer iterations through the drop-in API: native backward + fused Adam + device-side clip against the
    autograd functional + torch.optim.Adam + clip_grad_norm_ on a twin model (fp32 mode, same t / noise draws)."""
    import copy
    import json
    import os
    import yaml
    from scene_synthesis.networks import build_network, optimizer_factory
    from diffuscene_b200.optim import NativeAdam
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "config", "uncond/diffusion_bedrooms_instancond_lat32_v.yaml")))
    stats = tmp_path / "stats.json"
    stats.write_text(json.dumps(STATS))
    cfg["network"]["diffusion_kwargs"]["train_stats_file"] = str(stats)
    cfg["training"]["max_grad_norm"] = 0.05          # small enough that the clip is active
    torch.manual_seed(3)
    net_a, train_on_batch, _ = build_network(30, 23, cfg, None, device="cuda", precision="fp32")
    net_b, _, _ = build_network(30, 23, cfg, None, device="cuda", precision="fp32")
    net_b.load_state_dict(copy.deepcopy(net_a.state_dict()))
    net_b.native_backward = False
    opt_a = optimizer_factory(cfg["training"], net_a.parameters())
    assert isinstance(opt_a, NativeAdam)
    opt_b = torch.optim.Adam(net_b.parameters(), lr=cfg["training"]["lr"], weight_decay=0.0)
    g = torch.Generator().manual_seed(0)
    B = 4
    cls = torch.randint(0, 22, (B, 12), generator=g)
    sp = dict(translations=torch.rand(B, 12, 3, generator=g) * 2 - 1, sizes=torch.rand(B, 12, 3, generator=g) * 2 - 1,
              angles=torch.nn.functional.normalize(torch.randn(B, 12, 2, generator=g), dim=-1),
              class_labels=torch.nn.functional.one_hot(cls, 22).float() * 2 - 1,
              objfeats_32=torch.rand(B, 12, 32, generator=g) * 2 - 1, room_layout=torch.zeros(B, 1, 64, 64))
    sp = {k: v.cuda() for k, v in sp.items()}
    for it in range(2):
        torch.manual_seed(100 + it)
        la = train_on_batch(net_a, opt_a, sp, cfg)
        torch.manual_seed(100 + it)
        lb = train_on_batch(net_b, opt_b, sp, cfg)
        assert abs(la - lb) < 1e-3 * max(1.0, abs(lb)), (it, la, lb)
    sa, sb = net_a.state_dict(), net_b.state_dict()
    worst = max(((sa[k] - sb[k]).abs().max().item(), k) for k in sa)
    print("largest parameter difference after 2 iterations:", worst)
    # Adam normalises the step to ~lr per element: a sign flip of a tiny gradient moves a weight by 2 lr at most
    assert worst[0] < 3 * cfg["training"]["lr"], worst
    mean_diff = np.mean([(sa[k] - sb[k]).abs().mean().item() for k in sa])
    assert mean_diff < 0.05 * cfg["training"]["lr"], mean_diff
    # the sampling engine picks the trained weights up
    room = torch.zeros(2, 1, 64, 64, device="cuda")
    assert torch.isfinite(net_a.sample(room, 12, 62, batch_size=2, clip_denoised=True, ddim=True, ddim_steps=3, seed=1)).all()
