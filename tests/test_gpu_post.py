"""GPU: the batched front / back ends of the sampler on DEVICE tensors: catalogue retrieval
kernel (bit-exact indices against the numpy oracle of threed_future_dataset.py:28-77), attribute encode / decode and
the training collate against their CPU results."""
import numpy as np
import pytest
import torch

from diffuscene_b200.collate import collate_scenes
from diffuscene_b200.postprocess import ObjectCatalog, encode_batch, post_process_batch
from oracle import retrieval_ref as R
from tests.test_retrieval_cpu import make_catalog

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F", [32, 64, 5])
def test_retrieval_kernel_indices_are_bit_exact(F):
    rng = np.random.default_rng(3)
    labels, feats, sizes = make_catalog(rng, M=2000, n_classes=22, F=F)
    cat = ObjectCatalog(labels, feats, sizes, n_classes=22)
    B, N = 37, 12
    ql = rng.integers(0, 22, (B, N))
    qf = rng.normal(size=(B, N, F)).astype(np.float32)
    qs = rng.uniform(0.1, 2.0, (B, N, 3)).astype(np.float32)
    ql[0, :8] = labels[10]
    qs[0, :8] = sizes[10]                      # exact size ties: the feature key decides
    qf[0, :4] = feats[20]                      # exact (size, feature) ties: the catalogue order decides
    for mode, m in (("objfeats_and_size", 0), ("objfeats", 1), ("box", 2)):
        got = cat.retrieve(torch.from_numpy(ql).cuda(), torch.from_numpy(qf).cuda(), torch.from_numpy(qs).cuda(), mode)
        want = R.retrieve_batch(labels, feats, sizes, ql.reshape(-1), qf.reshape(-1, F), qs.reshape(-1, 3), m)
        assert got.dtype == torch.int64 and got.shape == (B, N)
        assert np.array_equal(got.cpu().numpy().reshape(-1), want), mode
    assert (got.cpu().numpy()[ql == 21] == -1).all()                 # class without catalogue entries


def test_retrieval_full_batch_properties():
    """BASELINE batch (4096 x 12 queries): every answer has the query's class and no other entry of that class is
    strictly better (size-independent optimality check on a sample of queries)."""
    rng = np.random.default_rng(4)
    M = 16000
    labels = rng.integers(0, 21, M)
    feats = rng.normal(size=(M, 32)).astype(np.float32)
    sizes = rng.uniform(0.1, 2.0, (M, 3)).astype(np.float32)
    cat = ObjectCatalog(labels, feats, sizes, n_classes=22)
    ql = torch.from_numpy(rng.integers(0, 21, (4096, 12))).cuda()
    qf = torch.randn(4096, 12, 32, device="cuda")
    qs = torch.rand(4096, 12, 3, device="cuda") * 2
    idx = cat.retrieve(ql, qf, qs, "objfeats_and_size")
    assert (idx >= 0).all()
    assert np.array_equal(labels[idx.cpu().numpy()], ql.cpu().numpy())
    sel = rng.integers(0, 4096 * 12, 50)
    want = R.retrieve_batch(labels, feats, sizes, ql.cpu().numpy().reshape(-1)[sel], qf.cpu().numpy().reshape(-1, 32)[sel],
                            qs.cpu().numpy().reshape(-1, 3)[sel], 0)
    assert np.array_equal(idx.cpu().numpy().reshape(-1)[sel], want)


def test_postprocess_on_device_matches_cpu():
    g = torch.Generator().manual_seed(0)
    B, N = 64, 12
    bounds = {"translations": (np.float32([-2.7, 0.04, -2.75]), np.float32([2.8, 3.6, 2.9])),
              "sizes": (np.float32([0.04, 0.02, 0.01]), np.float32([2.9, 1.8, 2.6])),
              "angles": (np.float32([-np.pi]), np.float32([np.pi])),
              "objfeats_32": (np.float32([1.0]), np.float32([-4.0]), np.float32([4.0]))}
    world = {"translations": torch.rand(B, N, 3, generator=g) * 8 - 4, "sizes": torch.rand(B, N, 3, generator=g) * 2,
             "angles": torch.rand(B, N, 1, generator=g) * 6 - 3, "objfeats_32": torch.rand(B, N, 32, generator=g) * 6 - 3,
             "class_labels": torch.rand(B, N, 22, generator=g) * 2 - 1}
    enc_c = encode_batch(world, bounds)
    enc_g = encode_batch({k: v.cuda() for k, v in world.items()}, bounds)
    for k in enc_c:
        assert enc_g[k].is_cuda
        # scale / descale are exact fp32 ops on both sides; cos / sin / atan2 differ by device libm ulps
        tol = dict(rtol=0, atol=0) if k not in ("angles",) else dict(rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(enc_g[k].cpu().numpy(), enc_c[k].numpy(), **tol)
    dec_c = post_process_batch(enc_c, bounds)
    dec_g = post_process_batch(enc_g, bounds)
    for k in dec_c:
        tol = dict(rtol=0, atol=0) if k != "angles" else dict(rtol=1e-6, atol=2e-6)
        np.testing.assert_allclose(dec_g[k].cpu().numpy(), dec_c[k].numpy(), **tol)


def test_collate_on_device_matches_cpu():
    rng = np.random.default_rng(1)

    def scene(L, C=23):
        cls = np.eye(C + 1, dtype=np.float32)[rng.integers(0, C - 1, L)]
        return {"class_labels": cls, "translations": rng.normal(size=(L, 3)).astype(np.float32),
                "sizes": rng.uniform(0.1, 1, (L, 3)).astype(np.float32), "angles": rng.normal(size=(L, 2)).astype(np.float32),
                "objfeats_32": rng.normal(size=(L, 32)).astype(np.float32)}

    scenes = [scene(int(L)) for L in rng.integers(1, 13, 257)]
    cpu = collate_scenes(scenes, max_length=12)
    dev = collate_scenes(scenes, max_length=12, device="cuda")
    for k, v in cpu.items():
        assert dev[k].is_cuda and torch.equal(dev[k].cpu(), v), k
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    pc = collate_scenes(scenes, max_length=12, permute=True, generator=g1)
    pd = collate_scenes(scenes, max_length=12, permute=True, generator=g2, device="cuda")
    for k, v in pc.items():
        assert torch.equal(pd[k].cpu(), v), k
