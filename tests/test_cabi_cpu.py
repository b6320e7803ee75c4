"""CPU: the C-ABI shared library builds, loads, exports every symbol declared in include/diffuscene_b200.h,
and its host-side plan agrees with the Python parameter inventory.  No compute calls (no GPU here)."""
import ctypes as C
import math
import os
import re

import pytest
import torch

from diffuscene_b200 import capi
from diffuscene_b200.weights import NetSpec, count_params, unet1d_param_specs
from tests.cases import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    hdr = open(os.path.join(ROOT, "include", "diffuscene_b200.h")).read()
    declared = re.findall(r"^DS_API [a-z0-9_ \*]+?(ds_[a-z0-9_]+)\(", hdr, flags=re.M)
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(capi.EXPORTED)
    assert b"sm_100a" in lib.ds_version()


@pytest.mark.parametrize("name", list(CASES))
def test_plan_inventory_matches_python_inventory(name):
    case = CASES[name]
    spec = NetSpec.from_net_kwargs(case["net_kwargs"])
    cfg = capi.make_config(spec, case["N"], case["diffusion_kwargs"]["time_num"])
    got = dict(capi.plan_expected_weights(cfg))
    want = {n[len("diffusion.model."):]: int(math.prod(shp)) for (n, shp, _) in unet1d_param_specs(spec)}
    assert got == want
    txt = capi.plan_describe(cfg)
    assert "GEMM encoder" in txt or "GEMM init_conv" in txt
    assert txt.count(" GN ") == 2 * (3 * spec.n_stages * 2 + 3 + 1)    # two Blocks per ResnetBlock


def test_bedroom_param_count():
    spec = NetSpec.from_net_kwargs(CASES["bed62"]["net_kwargs"])
    assert count_params(unet1d_param_specs(spec)) == 77676094


def test_invalid_config_is_rejected():
    spec = NetSpec.from_net_kwargs(CASES["bed62"]["net_kwargs"])
    cfg = capi.make_config(spec, 12, 1000)
    cfg.dim = 500
    with pytest.raises(capi.DsError):
        capi.plan_describe(cfg)
    with pytest.raises(NotImplementedError):
        NetSpec.from_net_kwargs(dict(CASES["bed62"]["net_kwargs"], dim_mults=[1, 2, 4, 8]))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = capi.load()
    spec = NetSpec.from_net_kwargs(CASES["bed62"]["net_kwargs"])
    cfg = capi.make_config(spec, 12, 1000)
    h = C.c_void_p()
    rc = lib.ds_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value                      # DS_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.ds_last_error()
    from diffuscene_b200.engine import DenoiserEngine
    with pytest.raises(RuntimeError):
        DenoiserEngine(spec, 12, 1000)


def test_channels_on_lanes_weight_row_order():
    """ds_gnt_weight_row (host-only): the row permutation ds_commit_weights applies to the convs of the fused GroupNorm
    blocks.  Checked against a model of the data path it exists for: TMEM lane l of a 32-lane quadrant belongs to
    thread l of an epilogue warp; the warp packs (token j, token j+1) pairs and stores them with stmatrix.trans, where
    thread l contributes column l // 4 of rows 2 (l % 4) + {0, 1} of an 8 x 8 b16 matrix whose row 2 c + e is the
    16-byte chunk c (8 channels) of token j + e in a [token][32 channel] block.  The block is channel-contiguous iff
    the thread on lane l holds channel 8 (l % 4) + l // 4."""
    lib = capi.load()
    rows = [lib.ds_gnt_weight_row(r) for r in range(256)]
    for b in range(0, 256, 32):
        blk = rows[b:b + 32]
        assert sorted(blk) == list(range(b, b + 32))                  # a permutation inside every block of 32
    block = {}                                                          # (token e, channel slot) -> channel
    for l in range(32):
        chunk, col = l % 4, l // 4                                      # stmatrix.trans: rows 2 chunk + e, column col
        for e in (0, 1):
            block[(e, 8 * chunk + col)] = rows[l]                       # slot = position inside the 64-byte token row
    for e in (0, 1):
        assert [block[(e, s)] for s in range(32)] == list(range(32))   # channels land in order: coalesced rows
