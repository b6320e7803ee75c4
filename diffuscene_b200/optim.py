"""Adam on the native training path: one fused kernel over the model's flat parameter / gradient buffers.

`optimizer_factory` (networks/__init__.py, reference scene_synthesis/networks/__init__.py:15-34) returns this class
for `optimizer: Adam`.  It IS a torch.optim.Adam (param_groups, state_dict, lr schedules keep working); `step()` runs
`ds_adam_step` once over the flat buffer that backs the denoiser's parameters when their gradients came from the
native backward pass, and torch's own update for everything else (the handful of small condition parameters, or the
whole model when the autograd fallback produced the gradients).
"""
from __future__ import annotations

import ctypes as C

import torch


class NativeAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._flat_state = {}      # id(model) -> dict(exp_avg, exp_avg_sq, step)

    # ------------------------------------------------------------------------------------------------------
    def _flat_owners(self):
        owners = {}
        for group in self.param_groups:
            for p in group["params"]:
                o = getattr(p, "_ds_flat_owner", None)
                if o is not None and getattr(o, "_native_grads_ready", False):
                    owners[id(o)] = (o, group)
        return owners

    @torch.no_grad()
    def step(self, closure=None, sumsq=None, max_norm: float = 0.0):
        """sumsq (device scalar, squared global gradient norm) + max_norm: clip_grad_norm_ folded into the update."""
        loss = closure() if closure is not None else None
        owners = self._flat_owners()
        if not owners:
            if sumsq is not None and max_norm > 0:
                coef = torch.clamp(max_norm / (sumsq.sqrt() + 1e-6), max=1.0)
                for group in self.param_groups:
                    for p in group["params"]:
                        if p.grad is not None:
                            p.grad.mul_(coef)
            super().step()
            return loss
        from . import capi
        lib = capi.load()
        coef = None
        if sumsq is not None and max_norm > 0:
            coef = torch.clamp(max_norm / (sumsq.sqrt() + 1e-6), max=1.0)
        for o, group in owners.values():
            if group.get("weight_decay", 0.0) != 0.0:
                raise NotImplementedError("NativeAdam: weight decay is pinned to 0 like the reference's optimizer_factory")
            st = self._flat_state.setdefault(id(o), None)
            if st is None or st["exp_avg"].numel() != o._flat.numel() or st["exp_avg"].device != o._flat.device:
                st = {"exp_avg": torch.zeros_like(o._flat), "exp_avg_sq": torch.zeros_like(o._flat), "step": 0}
                self._flat_state[id(o)] = st
            st["step"] += 1
            b1, b2 = group["betas"]
            stream = C.c_void_p(torch.cuda.current_stream(o._flat.device).cuda_stream)
            ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            capi.check(lib.ds_adam_step(ptr(o._flat), ptr(o._flat_grads), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]),
                                        o._flat.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                        st["step"], ptr(sumsq) if (sumsq is not None and max_norm > 0) else None,
                                        float(max_norm), stream))
        # everything that does not live in a flat buffer: plain Adam, same formula
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None or getattr(p, "_ds_flat_owner", None) is not None and id(p._ds_flat_owner) in owners:
                    continue
                g = p.grad if coef is None else p.grad * coef
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                k = float(st["step"])
                st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
                st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (st["exp_avg_sq"].sqrt() / (1 - b2 ** k) ** 0.5).add_(group["eps"])
                p.addcdiv_(st["exp_avg"], denom, value=-group["lr"] / (1 - b1 ** k))
        return loss

    @torch.no_grad()
    def step_clipped(self, model, max_norm: float):
        """clip_grad_norm_(model.parameters(), max_norm) + step() without a host synchronisation: the squared global
        norm is accumulated on the device (ds_sumsq over the flat buffer + the small remaining gradients) and the
        clip coefficient is applied inside the update kernels.  Returns the total gradient norm (device scalar)."""
        from . import capi
        lib = capi.load()
        dev = model._flat_grads.device
        sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        capi.check(lib.ds_sumsq(C.c_void_p(model._flat_grads.data_ptr()), model._flat_grads.numel(),
                                C.c_void_p(sumsq.data_ptr()), stream))
        for p in model.parameters():
            if p.grad is not None and getattr(p, "_ds_flat_owner", None) is None:
                sumsq += p.grad.float().pow(2).sum()
        self.step(sumsq=sumsq, max_norm=float(max_norm))
        return sumsq.sqrt().reshape(())
