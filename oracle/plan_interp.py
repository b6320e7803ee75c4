"""ORACLE (test infrastructure): CPU interpreter of the engine's step program.

Executes the op list exported by `ds_plan_export_json` (diffuscene_b200/csrc/plan.cpp) with plain torch
ops, assembling the packed weight matrices from the same recipes the CUDA engine uses.  Comparing its
output with the golden vectors verifies -- without a GPU -- that the engine's *program* (op order, buffer
wiring, hoisted FiLM tables, block-structured encoder / decoder GEMMs, folded weight standardisation) is
a faithful restatement of Unet1D.forward (reference denoise_net.py:507-593).  With emulate_bf16=True,
weights and every op output are rounded to bf16, predicting the error of the throughput mode.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .unet1d_ref import sinusoidal_embedding

OP_PACK, OP_GEMM, OP_GN, OP_LN, OP_LINATTN, OP_ATTN, OP_XATTN, OP_GEMM_GN, OP_GEMM_LN, OP_LN_QKV_ATTN, OP_ACT = range(11)


def _r(x, on):
    return x.to(torch.bfloat16).to(torch.float32) if on else x


def assemble_wmat(r, sd):
    W = torch.zeros(r["N"], r["K"], dtype=torch.float32)
    for pc in r["pieces"]:
        src = sd[pc["name"]].reshape(pc["rows"], pc["cols"]).to(torch.float32)
        if r["ws"]:
            mean = src.mean(dim=1, keepdim=True)
            var = src.var(dim=1, unbiased=False, keepdim=True)
            src = (src - mean) * torch.rsqrt(var + 1e-5)
        W[pc["row_off"]:pc["row_off"] + pc["rows"], pc["col_off"]:pc["col_off"] + pc["cols"]] = src
    if r.get("scale_k"):       # a LayerNorm gain folded into the columns (fused LN + to_qkv op)
        W = W * sd[r["scale_k"]].reshape(1, -1).to(torch.float32)
    return W


def assemble_vec(r, sd):
    v = torch.zeros(r["n"], dtype=torch.float32)
    for pc in r["pieces"]:
        v[pc["off"]:pc["off"] + pc["n"]] += sd[pc["name"]].reshape(-1).to(torch.float32)
    return v


def run_plan(plan: dict, sd: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor, context: torch.Tensor,
             context_cross: Optional[torch.Tensor] = None, emulate_bf16: bool = False, taps: Optional[dict] = None):
    """sd keys are reference names without the 'diffusion.model.' prefix."""
    B, N, d = x.shape
    M, C = B * N, plan["C"]
    bf = emulate_bf16
    bufs = [torch.zeros(M, w) for w in plan["buf_width"]]
    wm = [_r(assemble_wmat(r, sd), bf) for r in plan["wmats"]]
    vecs = [assemble_vec(r, sd) for r in plan["vecs"]]
    # hoisted FiLM tables (always fp32 in the engine)
    e = sinusoidal_embedding(t, C)
    e = F.gelu(F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    e = F.silu(F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"]))
    film_t = [F.linear(e, sd[n + ".mlp.1.weight"], sd[n + ".mlp.1.bias"]) for n in plan["time_blocks"]]   # [B, 2C]
    cs = F.silu(context.reshape(M, -1))
    film_c = [F.linear(cs, sd[n + ".mlp.1.weight"], sd[n + ".mlp.1.bias"]) for n in plan["ctx_blocks"]]  # [M, 2C]
    xctx = []
    for n in plan["xattn_layers"]:
        kv = F.linear(context_cross, sd[n + ".fn.fn.to_kv.weight"].reshape(256, -1))     # [B, L, 256]
        k, v = kv[..., :128], kv[..., 128:]
        L = kv.shape[1]
        k = k.reshape(B, L, 4, 32).softmax(dim=1)
        v = v.reshape(B, L, 4, 32)
        xctx.append(torch.einsum("blhd,blhe->bhde", k, v))

    def heads(z):
        return z.reshape(B, N, 4, 32)

    for op in plan["ops"]:
        k = op["kind"]
        if k == OP_PACK:
            out = torch.zeros(M, plan["kin_pad"])
            out[:, :d] = x.reshape(M, d)
            bufs[op["out"]] = _r(out, bf)
            continue
        i0 = op["in0"]
        if k == OP_GEMM or k == OP_GEMM_GN or k == OP_GEMM_LN:
            a = bufs[i0["buf"]][:, i0["col"]:i0["col"] + i0["k"]]
            if op["in1"]["buf"] >= 0:
                i1 = op["in1"]
                a = torch.cat([a, bufs[i1["buf"]][:, i1["col"]:i1["col"] + i1["k"]]], dim=1)
            y = a @ wm[op["w"]].t()
            if op["b"] >= 0:
                y = y + vecs[op["b"]]
        if k == OP_GEMM_GN:      # the GEMM result stays in fp32 (TMEM) through the GroupNorm epilogue
            h = y.reshape(B, N, 8, C // 8)
            mean = h.mean(dim=(1, 3), keepdim=True)
            var = h.var(dim=(1, 3), unbiased=False, keepdim=True)
            y = ((h - mean) * torch.rsqrt(var + 1e-5)).reshape(B, N, C) * vecs[op["gamma"]] + vecs[op["beta"]]
            if op["film"] == 1:
                f = film_t[op["film_blk"]][:, None, :]
                y = y * (f[..., :C] + 1) + f[..., C:]
            elif op["film"] == 2:
                f = film_c[op["film_blk"]].reshape(B, N, 2 * C)
                y = y * (f[..., :C] + 1) + f[..., C:]
            y = F.silu(y).reshape(M, C)
            if op["res"] >= 0:
                y = y + bufs[op["res"]]
            bufs[op["out"]] = _r(y, bf)
        elif k == OP_GEMM_LN:    # the projection stays in fp32 (TMEM) through the LayerNorm epilogue
            mean = y.mean(dim=1, keepdim=True)
            var = y.var(dim=1, unbiased=False, keepdim=True)
            y = (y - mean) * torch.rsqrt(var + 1e-5) * vecs[op["gamma"]]
            if op["res"] >= 0:
                y = y + bufs[op["res"]]
            bufs[op["out"]] = _r(y, bf)
        elif k == OP_GEMM:
            if op["act"] == 1:
                y = F.gelu(y)
            elif op["act"] == 2:
                y = F.silu(y)
            if op["res"] >= 0:
                y = y + bufs[op["res"]]
            bufs[op["out"]] = bufs[op["out"]].clone()
            bufs[op["out"]][:, op["out_col"]:op["out_col"] + op["N"]] = _r(y, bf)
        elif k == OP_GN:
            h = bufs[i0["buf"]].reshape(B, N, 8, C // 8)
            mean = h.mean(dim=(1, 3), keepdim=True)
            var = h.var(dim=(1, 3), unbiased=False, keepdim=True)
            y = ((h - mean) * torch.rsqrt(var + 1e-5)).reshape(B, N, C) * vecs[op["gamma"]] + vecs[op["beta"]]
            if op["film"] == 1:
                f = film_t[op["film_blk"]][:, None, :]
                y = y * (f[..., :C] + 1) + f[..., C:]
            elif op["film"] == 2:
                f = film_c[op["film_blk"]].reshape(B, N, 2 * C)
                y = y * (f[..., :C] + 1) + f[..., C:]
            y = F.silu(y).reshape(M, C)
            if op["res"] >= 0:
                y = y + bufs[op["res"]]
            bufs[op["out"]] = _r(y, bf)
        elif k == OP_LN:
            h = bufs[i0["buf"]]
            mean = h.mean(dim=1, keepdim=True)
            var = h.var(dim=1, unbiased=False, keepdim=True)
            y = (h - mean) * torch.rsqrt(var + 1e-5) * vecs[op["b"]]
            if op["res"] >= 0:
                y = y + bufs[op["res"]]
            bufs[op["out"]] = _r(y, bf)
        elif k == OP_LN_QKV_ATTN:      # LayerNorm (gain folded into the packed weights) + to_qkv + linear-attention core
            hh = bufs[i0["buf"]]
            mean = hh.mean(dim=1, keepdim=True)
            var = hh.var(dim=1, unbiased=False, keepdim=True)
            qkv = ((hh - mean) * torch.rsqrt(var + 1e-5)) @ wm[op["w"]].t()
            q, kk, v = heads(qkv[:, :128]), heads(qkv[:, 128:256]), heads(qkv[:, 256:384])
            q = q.softmax(dim=-1) * (32 ** -0.5)
            kk = kk.softmax(dim=1)
            ctx = torch.einsum("bnhd,bnhe->bhde", kk, v)
            o = torch.einsum("bhde,bnhd->bnhe", ctx, q)
            bufs[op["out"]] = _r(o.reshape(M, 128), bf)
        elif k == OP_LINATTN:
            qkv = bufs[i0["buf"]]
            q, kk, v = heads(qkv[:, :128]), heads(qkv[:, 128:256]), heads(qkv[:, 256:384])
            q = q.softmax(dim=-1) * (32 ** -0.5)
            kk = kk.softmax(dim=1)
            ctx = torch.einsum("bnhd,bnhe->bhde", kk, v)
            o = torch.einsum("bhde,bnhd->bnhe", ctx, q)
            bufs[op["out"]] = _r(o.reshape(M, 128), bf)
        elif k == OP_ATTN:
            qkv = bufs[i0["buf"]]
            q, kk, v = heads(qkv[:, :128]), heads(qkv[:, 128:256]), heads(qkv[:, 256:384])
            sim = torch.einsum("bihd,bjhd->bhij", q * (32 ** -0.5), kk)
            o = torch.einsum("bhij,bjhd->bihd", sim.softmax(dim=-1), v)
            bufs[op["out"]] = _r(o.reshape(M, 128), bf)
        elif k == OP_XATTN:
            q = heads(bufs[i0["buf"]]).softmax(dim=-1) * (32 ** -0.5)
            o = torch.einsum("bhde,bnhd->bnhe", xctx[op["xlayer"]], q)
            bufs[op["out"]] = _r(o.reshape(M, 128), bf)
        if taps is not None:
            taps[op["name"]] = bufs[op["out"]].clone()
    return bufs[plan["out_buf"]][:, :d].reshape(B, N, d)
