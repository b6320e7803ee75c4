"""CPU: host-side logic of the drop-in model class against goldens of the UNMODIFIED reference class
(tests/golden/cls_*.npz, written by tests/golden/make_golden_class.py): post-processing index work is bit-exact,
including an emptiness score of exactly 0.0; reference checkpoints of the text config (with `bertmodel.*` keys)
load; the state-dict key set does not depend on whether the frozen encoder has been loaded."""
import numpy as np
import pytest
import torch

from tests.cases_class import CLASS_CASES, class_config, class_state_dict, crafted_samples


def _net(name, tmp_path):
    from scene_synthesis.networks import build_network
    case = CLASS_CASES[name]
    cfg = class_config(case, str(tmp_path / "stats.json"))
    net, _, _ = build_network(30, case["n_classes"], cfg, None, device="cpu", precision="fp32")
    net.load_state_dict(class_state_dict(net.state_dict(), case["seed"]), strict=True)
    return net, case, cfg


@pytest.mark.parametrize("name", ["cls_bed", "cls_part", "cls_text"])
def test_delete_empty_matches_reference_bit_exact(name, tmp_path, golden_dir):
    net, case, _ = _net(name, tmp_path)
    g = np.load("%s/%s.npz" % (golden_dir, name))
    cs = crafted_samples(case)[:1]
    assert np.array_equal(cs.numpy(), g["crafted"])
    for keep in (False, True):
        boxes = net.delete_empty_from_network_samples(cs, device="cpu", keep_empty=keep)
        keys = [k[5:] for k in g.files if k.startswith("del%d." % int(keep))]
        assert sorted(boxes.keys()) == sorted(keys)
        for k in keys:
            assert np.array_equal(boxes[k].numpy(), g["del%d.%s" % (int(keep), k)]), (keep, k)
    # exact zeros (+0.0 and -0.0) count as empty: the reference compares `>= 0` (:360) and then `> 0` on the bool (:379)
    assert g["del0.translations"].shape[1] == int((cs[0, :, 8 + 21] < 0).sum())
    sdict = {"translations": cs[:, :, :3], "sizes": cs[:, :, 3:6], "angles": cs[:, :, 6:8], "class_labels": cs[:, :, 8:30],
             "objfeats": cs[:, :, 30:]}
    boxes = net.delete_empty_boxes(sdict, device="cpu")
    for k in [k[5:] for k in g.files if k.startswith("delb.")]:
        assert np.array_equal(boxes[k].numpy(), g["delb." + k]), k
    # batched per-scene variant: scene 0 agrees with the reference rule, class_index is the integer argmax
    per = net.delete_empty_batched(crafted_samples(case))
    assert np.array_equal(per[0]["translations"].numpy(), g["del0.translations"][0])
    assert torch.equal(per[0]["class_index"], per[0]["class_labels"].argmax(-1))


def test_reference_text_checkpoint_loads_and_state_dict_is_stable(tmp_path):
    net, case, _ = _net("cls_text", tmp_path)
    sd = net.state_dict()
    assert not any(k.startswith("bertmodel.") for k in sd)
    ref_like = dict(sd)
    ref_like["bertmodel.embeddings.word_embeddings.weight"] = torch.zeros(4, 4)      # what a reference checkpoint carries
    ref_like["bertmodel.encoder.layer.0.attention.self.query.bias"] = torch.zeros(4)
    net.load_state_dict(ref_like, strict=True)                                       # must not raise
    # a frozen encoder attached later does not change the key set
    object.__setattr__(net, "bertmodel", torch.nn.Linear(2, 2))
    assert set(net.state_dict().keys()) == set(sd.keys())
    with pytest.raises(RuntimeError):
        bad = dict(sd)
        bad.pop("fc_text_f.weight")
        net.load_state_dict(bad, strict=True)


def test_get_loss_pack_order_and_conditions_on_cpu(tmp_path, golden_dir):
    """The host-side part of get_loss / sample that needs no GPU: attribute pack order and the condition builders
    reproduce what the reference class feeds its denoiser (checked through the reference's own golden t / noise by
    the GPU tests; here: shapes and the shared / per-scene decision)."""
    from tests.cases_class import class_batch
    net, case, _ = _net("cls_part", tmp_path)
    sp = class_batch(case)
    target = net._pack_target(sp)
    assert target.shape == (2, 12, 62)
    assert torch.equal(target[..., :3], sp["translations"]) and torch.equal(target[..., 8:30], sp["class_labels"])
    cond, shared = net._condition(2, torch.device("cpu"), layout_target=target)
    assert cond.shape == (2, 12, 192) and not shared
    net, case, _ = _net("cls_bed", tmp_path)
    cond, shared = net._condition(2, torch.device("cpu"), layout_target=target)
    assert cond.shape == (12, 128) and shared
