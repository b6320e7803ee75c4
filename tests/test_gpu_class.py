"""GPU: the drop-in model class against goldens taken from the UNMODIFIED reference class
`DiffusionSceneLayout_DDPM` through its public methods (tests/golden/make_golden_class.py): get_loss (native value
and the training path), sample / generate_layout / trajectory, completion with fc_partial_condition, arrangement
with fc_arrange_condition, text conditioning through fc_text_f.  fp32 parity mode; the reference's random draws
(t, noise, x_T, per-step noise) are stored in the goldens and injected."""
import numpy as np
import pytest
import torch

from tests.cases_class import CLASS_CASES, class_batch, class_config, class_state_dict, fake_bert_hidden

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4          # north-star fp32 gate


def _net(name, tmp_path):
    from scene_synthesis.networks import build_network
    case = CLASS_CASES[name]
    cfg = class_config(case, str(tmp_path / "stats.json"))
    net, _, validate = build_network(30, case["n_classes"], cfg, None, device="cuda", precision="fp32")
    sd = class_state_dict(net.state_dict(), case["seed"])
    net.load_state_dict(sd, strict=True)
    return net, case, cfg


def _sp(case):
    sp = class_batch(case)
    out = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in sp.items()}
    if case.get("text"):
        out["context_cross"] = None
        out.pop("context_cross")
        # BERT itself is outside the path (frozen, once per scene): feed its hidden states [B, L, 768]
        out["description"] = fake_bert_hidden(sp["description"]).cuda()
    return out


@pytest.mark.parametrize("name", list(CLASS_CASES))
def test_get_loss_matches_reference_class(name, tmp_path, golden_dir):
    net, case, cfg = _net(name, tmp_path)
    g = np.load("%s/%s.npz" % (golden_dir, name))
    sp = _sp(case)
    t, noise = torch.from_numpy(g["loss_t"]), torch.from_numpy(g["loss_noise"])
    with torch.no_grad():                                    # CUDA-library value (validate_on_batch path)
        loss, ld = net.get_loss(sp, t=t, noise=noise)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=RTOL, atol=ATOL)
    keys = [k[3:] for k in g.files if k.startswith("ld.")]
    assert sorted(ld.keys()) == sorted(keys)
    for k in keys:
        np.testing.assert_allclose(float(ld[k]), float(g["ld." + k]), rtol=RTOL, atol=ATOL, err_msg=k)
    with torch.enable_grad():                                # training path: same value, gradients exist
        loss2, ld2 = net.get_loss(sp, t=t, noise=noise)
    np.testing.assert_allclose(float(loss2), float(g["loss"]), rtol=RTOL, atol=ATOL)
    assert sorted(ld2.keys()) == sorted(keys)


@pytest.mark.parametrize("name", list(CLASS_CASES))
def test_sample_matches_reference_class(name, tmp_path, golden_dir):
    net, case, cfg = _net(name, tmp_path)
    g = np.load("%s/%s.npz" % (golden_dir, name))
    B, N = case["B"], case["N"]
    room = torch.zeros(B, 1, 64, 64, device="cuda")
    draws = torch.from_numpy(g["sample_draws"])
    kw = dict(clip_denoised=True, x_init=draws[0], noise=draws[1:])
    if case.get("text"):
        kw["text"] = fake_bert_hidden(class_batch(case)["description"]).cuda()
    if case["kind"] == "arrange":
        out = net.sample(room, N, 62, batch_size=B, input_boxes=torch.from_numpy(g["input_boxes"]).cuda(), **kw)
    elif case["kind"] == "partial":
        eng_kw = dict(kw)
        out = net.sample(room, N, 62, batch_size=B, partial_boxes=torch.from_numpy(g["partial_boxes"]).cuda(),
                         partial_noise=torch.from_numpy(g["sample_draws_partial"]), **eng_kw)
    else:
        out = net.sample(room, N, 62, batch_size=B, **kw)
    ref = g["sample"]
    # error growth over the T-step loop stays inside the single-forward gate
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL, atol=5 * ATOL)
    if case["kind"] == "uncond":
        cl = slice(8, 8 + 21)
        assert np.array_equal(out.cpu().numpy()[..., cl].argmax(-1), ref[..., cl].argmax(-1))      # integer class argmax
        # generate_layout = sample + delete_empty_from_network_samples (batch 1, like the reference's script)
        d1 = torch.from_numpy(g["layout_draws"])
        kw1 = dict(kw, x_init=d1[0], noise=d1[1:])
        if case.get("text"):
            kw1["text"] = kw["text"][:1]
        s1 = net.sample(room[:1], N, 62, batch_size=1, **kw1)
        boxes = net.delete_empty_from_network_samples(s1, device="cpu")
        for k in [k[7:] for k in g.files if k.startswith("layout.")]:
            assert boxes[k].shape == g["layout." + k].shape, k
            np.testing.assert_allclose(boxes[k].numpy(), g["layout." + k], rtol=RTOL, atol=5 * ATOL, err_msg=k)


def test_trajectory_matches_reference_class(tmp_path, golden_dir):
    net, case, cfg = _net("cls_bed", tmp_path)
    g = np.load("%s/cls_bed.npz" % golden_dir)
    # the trajectory run of the generator used its own recorded stream (seed + 3): same generator formula here
    from tests.cases import noise_stream
    nz = noise_stream(case["seed"] + 3)
    nz((2, 12, 62))                                          # the shape-only draw at the top of sample()
    x_T = nz((2, 12, 62))
    noise = torch.stack([nz((2, 12, 62)) for _ in range(case["T"])])
    room = torch.zeros(2, 1, 64, 64, device="cuda")
    traj = net.sample(room, 12, 62, batch_size=2, clip_denoised=True, ret_traj=True, freq=2, x_init=x_T, noise=noise)
    ref = g["traj"]
    assert len(traj) == ref.shape[0]
    for i in range(1, len(traj)):
        np.testing.assert_allclose(traj[i].cpu().numpy(), ref[i], rtol=RTOL, atol=5 * ATOL, err_msg=str(i))
