"""GPU parity, fp32 mode: the CUDA path (through the C ABI) against the golden vectors of the unmodified
reference and against the oracle.  Tolerance is the north star's: rtol 1e-3 / atol 1e-4, class-argmax exact."""
import numpy as np
import pytest
import torch

from tests.cases import CASES, STATS, noise_stream
from tests.gpu_common import cuda, get_engine, gold

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("name", list(CASES))
def test_forward_parity(name, golden_dir):
    eng, case, spec, inp = get_engine(name, "fp32")
    g = gold(golden_dir, name)
    out = eng.forward(cuda(inp["x"]), cuda(inp["t"])).cpu().numpy()
    np.testing.assert_allclose(out, g["fwd"], **TOL)
    if spec.seperate_all:
        b0 = spec.bbox_dim
        assert (out[..., b0:b0 + spec.class_dim - 1].argmax(-1) == g["fwd"][..., b0:b0 + spec.class_dim - 1].argmax(-1)).all()
    # host-buffer entry point gives the same numbers
    out_h = eng.forward_host(inp["x"], inp["t"]).numpy()
    np.testing.assert_array_equal(out_h, out)


@pytest.mark.parametrize("name", list(CASES))
def test_p_sample_step_parity(name, golden_dir):
    eng, case, spec, inp = get_engine(name, "fp32")
    g = gold(golden_dir, name)
    for key, clip, t, seed in (("step_clip", True, inp["t"], 77), ("step_noclip", False, inp["t"], 77),
                               ("step_t0", True, torch.zeros_like(inp["t"]), 78)):
        nz = noise_stream(case["seed"] + seed)(inp["x"].shape)
        out = eng.p_sample_step(cuda(inp["x"]), cuda(t), cuda(nz), clip).cpu().numpy()
        np.testing.assert_allclose(out, g[key], **TOL)


def _stack_noise(stream, n, shape):
    return torch.stack([stream(shape) for _ in range(n)])


@pytest.mark.parametrize("name", ["bed62_loop", "text62_loop"])
def test_sampling_loops_parity(name, golden_dir):
    eng, case, spec, inp = get_engine(name, "fp32")
    g = gold(golden_dir, name)
    T = case["diffusion_kwargs"]["time_num"]
    shape = tuple(inp["x"].shape)
    loop_tol = dict(rtol=2e-3, atol=2e-4)
    for use_graph in (True, False):
        nz = noise_stream(case["seed"] + 100)
        x_T = nz(shape)
        noise = _stack_noise(nz, T, shape)
        out = eng.sample(shape[0], clip_denoised=True, x_init=x_T, noise=noise, use_graph=use_graph)
        np.testing.assert_allclose(out.cpu().numpy(), g["loop"], **loop_tol)
    out, traj = eng.sample(shape[0], clip_denoised=True, x_init=x_T, noise=noise, traj_freq=4)
    got = np.concatenate([x_T[None].numpy(), traj.cpu().numpy()])
    np.testing.assert_allclose(got, g["traj"], **loop_tol)
    # completion: per step the reference draws the partial noise first, then the step noise
    nz = noise_stream(case["seed"] + 101)
    x_T = nz(shape)
    pn, sn = [], []
    for _ in range(T):
        pn.append(nz(inp["partial"].shape))
        sn.append(nz(shape))
    out = eng.sample(shape[0], clip_denoised=True, x_init=x_T, noise=torch.stack(sn), partial=inp["partial"],
                     partial_noise=torch.stack(pn))
    np.testing.assert_allclose(out.cpu().numpy(), g["loop_complete"], **loop_tol)


def test_arrange_loop_parity(golden_dir):
    eng, case, spec, inp = get_engine("arr5", "fp32")
    g = gold(golden_dir, "arr5")
    T = case["diffusion_kwargs"]["time_num"]
    small = (case["B"], case["N"], 5)
    nz = noise_stream(case["seed"] + 100)
    x_T = nz(small)
    noise = _stack_noise(nz, T, small)
    x = eng.sample(case["B"], clip_denoised=True, x_init=x_T, noise=noise).cpu()
    boxes = inp["boxes"]
    full = torch.cat([x[..., :3], boxes[..., 3:6], x[..., 3:], boxes[..., 8:]], dim=-1)   # diffusion_ddpm.py:496-503
    np.testing.assert_allclose(full.numpy(), g["loop_arrange"], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c.get("loss")])
def test_p_losses_parity(name, golden_dir):
    eng, case, spec, inp = get_engine(name, "fp32")
    g = gold(golden_dir, name)
    dk = case["diffusion_kwargs"]
    bounds = STATS["bounds_translations"] + STATS["bounds_sizes"]
    losses, ld = eng.p_losses(cuda(inp["x0"]), cuda(inp["t_loss"]), cuda(inp["noise_loss"]), dk["loss_separate"],
                              dk["loss_iou"], bounds)
    np.testing.assert_allclose(losses.cpu().numpy(), g["losses"], **TOL)
    for k, v in ld.items():
        if "ld." + k in g.files:
            np.testing.assert_allclose(float(v), float(g["ld." + k]), **TOL)
    xq = eng.q_sample(cuda(inp["x0"]), cuda(inp["t_loss"]), cuda(inp["noise_loss"])).cpu()
    from oracle import diffusion_ref as D
    from tests.gpu_common import case_tables
    tb = case_tables(case)
    ref = tb["sqrt_ac"][inp["t_loss"]].reshape(-1, 1, 1) * inp["x0"] + \
        tb["sqrt_1mac"][inp["t_loss"]].reshape(-1, 1, 1) * inp["noise_loss"]
    np.testing.assert_allclose(xq.numpy(), ref.numpy(), rtol=1e-6, atol=1e-7)


def test_intermediate_taps_match_plan_interpreter():
    """Localises any kernel bug: every op output of the CUDA path against the CPU plan interpreter."""
    from diffuscene_b200 import capi
    from diffuscene_b200.weights import seeded_state_dict, unet1d_param_specs
    from oracle.plan_interp import run_plan
    eng, case, spec, inp = get_engine("bed62", "fp32")
    eng.enable_taps(True)
    try:
        from tests.gpu_common import case_tables
        eng.set_context(inp["context"][0], shared=True)
        eng.forward(cuda(inp["x"]), cuda(inp["t"]))
        plan = capi.plan_export(eng.cfg, no_reuse=True)
        sd = {k[len("diffusion.model."):]: v for k, v in
              seeded_state_dict(unet1d_param_specs(spec), seed=case["seed"]).items()}
        taps = {}
        run_plan(plan, sd, inp["x"], inp["t"], inp["context"], inp["context_cross"], taps=taps)
        M = case["B"] * case["N"]
        bad = []
        for op in plan["ops"]:
            if op["kind"] == 0:
                continue
            got = eng.read_tap(op["name"], M)
            ref = taps[op["name"]]
            c0, c1 = op["out_col"], op["out_col"] + (op["N"] if op["kind"] == 1 else ref.shape[1])
            got, ref = got[:, c0:c1], ref[:, c0:c1]      # grouped GEMMs fill column blocks of a shared buffer
            err = (got - ref).abs().max().item()
            if err > 2e-3 * max(1.0, ref.abs().max().item()):
                bad.append((op["name"], err))
        assert not bad, bad[:10]
    finally:
        eng.enable_taps(False)
        eng.set_context(inp["context"][0], shared=True)


def test_philox_sampling_properties():
    """In-kernel RNG: deterministic per seed, independent of how scenes are sharded, N(0,1) moments."""
    eng, case, spec, inp = get_engine("bed62_loop", "fp32")
    a = eng.sample(4, seed=7)
    b = eng.sample(4, seed=7)
    assert torch.equal(a, b)
    c = eng.sample(2, seed=7, scene_offset=2)          # scenes 2,3 of the batch above, run as their own shard
    np.testing.assert_allclose(c.cpu().numpy(), a[2:].cpu().numpy(), rtol=1e-5, atol=1e-6)
    d = eng.sample(4, seed=8)
    assert not torch.equal(a, d)
    assert torch.isfinite(a).all() and a.abs().max() < 50
    # raw generator moments through a zero-step "loop" is not exposed; check x_T statistics via a 1-step DDIM
    z = eng.sample(64, seed=3, num_steps=1, ddim=True)   # = clamp(x0 estimate): finite and clamped
    assert z.abs().max() <= 1.0 + 1e-6


def test_ddim_matches_oracle():
    """DDIM is a working restatement of the formula at diffusion_ddpm.py:401-444 (dead code in the reference):
    parity target is the oracle.  eta = 0 (deterministic) and eta = 0.5 with injected noise."""
    from oracle import diffusion_ref as D
    from oracle.unet1d_ref import unet1d_forward
    from diffuscene_b200.weights import seeded_state_dict, unet1d_param_specs
    eng, case, spec, inp = get_engine("bed62_loop", "fp32")
    dk = case["diffusion_kwargs"]
    sched = D.make_schedule(D.make_betas("linear", dk["beta_start"], dk["beta_end"], dk["time_num"]), "v", "fixedsmall")
    sd = seeded_state_dict(unet1d_param_specs(spec), seed=case["seed"])
    den = lambda x, t: unet1d_forward(sd, spec, x, t, inp["context"], None)
    shape = tuple(inp["x"].shape)
    S = 5
    times = [tp[0] for tp in D.ddim_times(dk["time_num"], S)] + [-1]
    for eta in (0.0, 0.5):
        nz = noise_stream(case["seed"] + 300)
        ref = D.ddim_sample_loop(sched, den, shape, nz, steps=S, eta=eta)
        nz = noise_stream(case["seed"] + 300)
        x_T = nz(shape)
        noise = torch.stack([nz(shape) if tp[1] >= 0 else torch.zeros(shape) for tp in D.ddim_times(dk["time_num"], S)])
        out = eng.sample(shape[0], ddim=True, num_steps=S, ddim_eta=eta, x_init=x_T, noise=noise, ddim_times=times)
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-3, atol=2e-4)
