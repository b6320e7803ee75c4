"""CPU: the batched attribute scaling (diffuscene_b200/postprocess.py) against a per-scene numpy restatement of the
reference's decorator (threed_front_dataset.py:483-535)."""
import numpy as np
import torch

from diffuscene_b200.postprocess import encode_batch, post_process_batch


def _np_scale(x, lo, hi):
    x = np.clip(x.astype(np.float32), lo, hi)
    return 2 * ((x - lo) / (hi - lo)) - 1


def _np_descale(x, lo, hi):
    return (x + 1) / 2 * (hi - lo) + lo


def _bounds(rng):
    lo3, hi3 = rng.uniform(-3, -1, 3).astype(np.float32), rng.uniform(1, 3, 3).astype(np.float32)
    return {"translations": (lo3, hi3), "sizes": (np.float32([0.05, 0.05, 0.05]), np.float32([2.5, 1.5, 2.0])),
            "angles": (np.float32([-np.pi]), np.float32([np.pi])),
            "objfeats_32": (np.float32([1.0]), np.float32([-4.0]), np.float32([4.0]))}


def test_encode_then_post_process_round_trips_and_matches_numpy():
    rng = np.random.default_rng(0)
    B, N = 5, 12
    bounds = _bounds(rng)
    world = {"translations": rng.uniform(-4, 4, (B, N, 3)).astype(np.float32),      # some values outside the bounds
             "sizes": rng.uniform(0.1, 2.0, (B, N, 3)).astype(np.float32),
             "angles": rng.uniform(-np.pi, np.pi, (B, N, 1)).astype(np.float32),
             "objfeats_32": rng.uniform(-3, 3, (B, N, 32)).astype(np.float32),
             "class_labels": rng.uniform(-1, 1, (B, N, 22)).astype(np.float32)}
    enc = encode_batch({k: torch.from_numpy(v) for k, v in world.items()}, bounds)
    for b in range(B):                                                                # the reference works per scene
        np.testing.assert_allclose(enc["translations"][b].numpy(), _np_scale(world["translations"][b], *bounds["translations"]), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(enc["sizes"][b].numpy(), _np_scale(world["sizes"][b], *bounds["sizes"]), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(enc["objfeats_32"][b].numpy(), _np_scale(world["objfeats_32"][b], bounds["objfeats_32"][1], bounds["objfeats_32"][2]), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(enc["angles"][b].numpy(), np.concatenate([np.cos(world["angles"][b]), np.sin(world["angles"][b])], -1), rtol=1e-6, atol=1e-6)
    assert enc["angles"].shape == (B, N, 2) and torch.equal(enc["class_labels"], torch.from_numpy(world["class_labels"]))
    assert enc["translations"].abs().max() <= 1.0 + 1e-6                             # clipped into [-1, 1]
    dec = post_process_batch(enc, bounds)
    np.testing.assert_allclose(dec["angles"].numpy(), world["angles"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dec["sizes"].numpy(), np.clip(world["sizes"], *bounds["sizes"]), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dec["objfeats_32"].numpy(), world["objfeats_32"], rtol=1e-5, atol=1e-5)
    clipped = np.clip(world["translations"], *bounds["translations"])
    np.testing.assert_allclose(dec["translations"].numpy(), clipped, rtol=1e-5, atol=1e-5)
    for b in range(B):
        np.testing.assert_allclose(dec["sizes"][b].numpy(), _np_descale(enc["sizes"][b].numpy(), *bounds["sizes"]), rtol=1e-6, atol=1e-6)
    assert torch.equal(dec["class_labels"], enc["class_labels"])                      # passes through
