#!/usr/bin/env python
"""Benchmark of the DiffuScene denoising hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full T=1000 ancestral DDPM sample of the per-GPU batch (BASELINE.json configs[1]:
unconditional bedroom, N=12 objects, d=62 attributes, bf16): T denoiser forwards + T posterior updates,
nothing skipped, per-step noise from the in-kernel Philox generator.  `value` = scenes/s with inputs
resident in HBM; `e2e` = the same through the public sampling API with host (pinned) x_T / context in and
host result out.  Multi-GPU: scenes shard over ranks, no data-path collective (sampling is embarrassingly
parallel), weak scaling, max-over-ranks timing.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_SCENE_MFLOP = {"bedroom": 870.3, "living": 1455.7}       # forward FLOPs / scene / step
# the two bedroom shapes SURVEY 8(d) / BASELINE.md 3 name: the reference's real one (d = 62, the default) and
# BASELINE.json's synthetic one (d = 97: angle_dim 4, class_dim 23, objfeat_dim 64; 871.2 MFLOP / scene / step)
SHAPES = {"real62": (dict(), 62, 870.3),
          "synth97": (dict(channels=97, class_dim=23, angle_dim=4, objfeat_dim=64), 97, 871.2)}


def bed_kwargs(shape):
    kw = dict(BED)
    kw.update(SHAPES[shape][0])
    return kw
BED = dict(dim=512, dim_mults=[1, 1, 1, 1], channels=62, objectness_dim=0, class_dim=22, angle_dim=2,
           objfeat_dim=32, context_dim=0, instanclass_dim=128, seperate_all=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_scenes_per_sec(batch: int, steps: int, threads: int, shape="real62"):
    """CPU baseline: the oracle restatement of the reference path (fp32 torch on the host cores), timed on
    a bounded sample (`steps` diffusion steps of `batch` scenes) and scaled to a 1000-step sample."""
    from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs
    from oracle import diffusion_ref as D
    from oracle.unet1d_ref import unet1d_forward
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    spec = NetSpec.from_net_kwargs(bed_kwargs(shape))
    sd = seeded_state_dict(unet1d_param_specs(spec), seed=0)
    sched = D.make_schedule(D.make_betas("linear", 1e-4, 0.02, 1000), "v", "fixedsmall")
    ctx = torch.randn(12, 128)[None].expand(batch, 12, 128).contiguous()
    x = torch.randn(batch, 12, SHAPES[shape][1])
    den = lambda xx, tt: unet1d_forward(sd, spec, xx, tt, ctx, None)
    t0 = None
    for i, step in enumerate(reversed(range(1000 - steps - 1, 1000))):
        if i == 1:
            t0 = time.perf_counter()       # first iteration is warm-up
        t = torch.full((batch,), step, dtype=torch.int64)
        x, _ = D.p_sample_step(sched, den, x, t, torch.randn_like(x), True)
    dt = (time.perf_counter() - t0) / steps
    return batch / (dt * 1000.0), dt


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation is Python and cannot travel to the GPU
    box (no /root/reference there), so the oracle port of it is timed on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 32)      # torch CPU ops of this size stop scaling (and oversubscribe) beyond ~32 threads
    batch, dsteps = 32, 2
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = oracle_scenes_per_sec(batch, dsteps, threads, args.shape)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    sample = "%d diffusion steps x %d scenes per bench step, scaled to a 1000-step sample" % (dsteps, batch)
    print(json.dumps({
        "impl": "reference", "metric": "scenes/sec full 1000-step DDPM sample", "value": v, "unit": "scenes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * batch / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "uncond bedroom N=12 d=%d T=1000 DDPM sampling (BASELINE configs[1])" % SHAPES[args.shape][1]},
        "cpu_baseline": {"value": v, "unit": "scenes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="scenes per GPU")
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--backend", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="scenes per L2-resident sub-batch (0: whole batch at once)")
    ap.add_argument("--shape", default="real62", choices=sorted(SHAPES), help="bedroom attribute layout: the reference's "
                    "real config (d=62) or BASELINE.json's synthetic d=97")
    ap.add_argument("--fuse", type=int, default=None, help="fuse_level override (1: row-major fused GroupNorm GEMM, "
                                                            "2: channels-on-lanes variant); default: the engine's")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-op device time table to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from diffuscene_b200.engine import DenoiserEngine
    from diffuscene_b200.schedule import get_betas, make_tables
    from diffuscene_b200.weights import NetSpec, seeded_state_dict, unet1d_param_specs

    spec = NetSpec.from_net_kwargs(bed_kwargs(args.shape))
    D_ATTR, F_SCENE = SHAPES[args.shape][1], SHAPES[args.shape][2]
    N_OBJ, T, B = 12, args.timesteps, args.batch
    eng = DenoiserEngine(spec, N_OBJ, T, precision=args.precision, gemm_backend=args.backend, device=local,
                         fuse_level=args.fuse)
    eng.load_state_dict(seeded_state_dict(unet1d_param_specs(spec), seed=0))      # random-init weights
    eng.set_schedule(make_tables(get_betas("linear", 1e-4, 0.02, T), "v", "fixedsmall"))
    g = torch.Generator().manual_seed(1)
    pos_emb = torch.randn(N_OBJ, 128, generator=g).pin_memory()                      # positional_embedding
    x_T_host = torch.randn(B, N_OBJ, D_ATTR, generator=g).pin_memory()
    eng.set_context(pos_emb.to(dev), shared=True)
    x_T_dev = x_T_host.to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def resident_step(i):
        return eng.sample(B, clip_denoised=True, x_init=x_T_dev, seed=100 + i, scene_offset=rank * B, chunk_scenes=args.chunk)

    def e2e_step(i):
        # public sampling call with host buffers: H2D of this step's x_T and condition, D2H of the result
        eng.set_context(pos_emb.to(dev, non_blocking=True), shared=True)
        x0 = eng.sample(B, clip_denoised=True, x_init=x_T_host.to(dev, non_blocking=True), seed=100 + i,
                        scene_offset=rank * B, host_output=True, chunk_scenes=args.chunk)
        return x0

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record()
        for i in range(steps):
            out = fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), eng.launch_count() - l0, out

    clocks = ClockSampler(local)
    clocks.start()
    ms, launches, out = timed(resident_step, args.steps, args.warmup)
    clk = clocks.stop()
    assert torch.isfinite(out).all()
    ms_e2e, _, _ = timed(e2e_step, max(1, min(args.steps, 2)), 1)
    n_e2e = max(1, min(args.steps, 2))

    if rank == 0:
        total_scenes = B * world
        value = total_scenes * args.steps / (ms / 1000.0)
        e2e_value = total_scenes * n_e2e / (ms_e2e / 1000.0)
        ms_per_step = ms / args.steps
        us_per_dstep = ms_per_step * 1000.0 / T
        peak_tf, peak_gbs, which = measured_peaks()
        flops_per_launch = B * F_SCENE * 1e6            # one diffusion step of one GPU's batch
        achieved_tf = flops_per_launch / (us_per_dstep * 1e-6) / 1e12
        res = {
            "metric": "scenes/sec full 1000-step DDPM sample", "value": value, "unit": "scenes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "uncond bedroom N=12 d=%d T=%d DDPM sampling, batch=%d scenes/GPU "
                                   "(BASELINE configs[1]), random-init weights" % (D_ATTR, T, B),
                       "scenes_per_gpu": B, "global_batch": total_scenes, "parallelism": "scene-shard x%d" % world,
                       "l2": "per-step working set (>=20 x 50 MB activation buffers + 63 MB weights) exceeds the "
                             "126 MB L2; no explicit flush"},
            "denoiser_fwd_us_per_step": us_per_dstep,
            "gpu_launches": int(launches),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "scenes/s", "h2d_bytes_per_step": int(x_T_host.numel() * 4 + pos_emb.numel() * 4),
                    "d2h_bytes_per_step": int(B * N_OBJ * D_ATTR * 4)},
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": None, "peak_source": which + " bf16 sustained",
                         "launch": "one diffusion step (CUDA graph of the step program) over %d scenes; "
                                   "algorithmic %.1f MFLOP/scene/step" % (B, F_SCENE)},
        }
        # dominant kernel: k_gemm_gnt<12>, the channels-on-lanes tcgen05 GEMM that carries the 56 fused conv + GroupNorm
        # + FiLM + SiLU blocks and the epilogue-bound plain GEMMs (to_qkv, to_out) -- 76 of the 125 launches of a step.
        # Timed live with CUDA events around each op of one eager pass of the step program (launching stream);
        # DRAM traffic from the committed ncu capture of the same command.
        from diffuscene_b200 import capi
        ops = eng.profile_ops(B)
        us_by_name = dict(ops)

        def on_gnt(o):      # mirrors gemm_variant() in csrc/engine.cu
            if eng.fuse_level < 2 or N_OBJ != 12 or o["name"] not in us_by_name:
                return False
            if o["kind"] == 7:
                return True
            k = o["in0"]["k"] + o["in1"]["k"]
            return o["kind"] == 1 and eng.fuse_level >= 3 and o["N"] % 128 == 0 and (k <= 128 or o["N"] % 256 != 0)

        dom = [o for o in capi.plan_export(eng.cfg)["ops"] if on_gnt(o)]
        if dom:
            dom_us = sum(us_by_name[o["name"]] for o in dom)
            dom_flop = sum(2.0 * B * N_OBJ * o["N"] * (o["in0"]["k"] + o["in1"]["k"]) for o in dom)
            res["roofline"]["dominant_kernel"] = {
                "name": "k_gemm_gnt<12> (tcgen05 GEMM, output channels on the TMEM lanes; conv+GroupNorm+FiLM+SiLU "
                        "epilogue for 56 launches, bias-only for 20)", "launches_per_step": len(dom),
                "avg_us": dom_us / len(dom), "achieved": dom_flop / (dom_us * 1e-6) / 1e12, "unit": "TFLOP/s",
                "frac": dom_flop / (dom_us * 1e-6) / 1e12 / peak_tf,
                "share_of_step": dom_us / max(1e-9, sum(u for _, u in ops))}
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "round1_v8_dram_traffic.json")
        if B == 4096 and args.precision == "bf16" and eng.fuse_level >= 3 and args.shape == "real62" and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            res["roofline"]["traffic"] = tj["step_dram_bytes"]
            res["roofline"]["traffic_source"] = "profiles/round1_v8_dram_traffic.json (ncu dram__bytes_read+write, one step)"
            k = tj["kernels"].get("k_gemm_gnt<12>")
            if k and dom:
                res["roofline"]["dominant_kernel"]["traffic"] = (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["launches"]
                res["roofline"]["dominant_kernel"]["ncu_share_of_step"] = k["time_ns"] / tj["step_time_ns"]
        if args.profile_ops:
            tot = sum(u for _, u in ops)
            gemm = sum(u for n, u in ops if "proj" in n or "conv" in n or n.startswith(("enc", "dec", "out", "init"))
                       or "to_" in n or n.endswith(".5") or n == "mid_attn")
            sys.stderr.write("per-op device time (us), total %.1f, gemm-ish %.1f\n" % (tot, gemm))
            for n, u in ops:
                sys.stderr.write("  %-40s %8.1f\n" % (n, u))
        if not args.no_cpu_baseline:
            threads = min(os.cpu_count() or 1, 32)
            v, dt = oracle_scenes_per_sec(16, 3, threads, args.shape)
            res["cpu_baseline"] = {"value": v, "unit": "scenes/s", "cores": threads, "kind": "port",
                                   "sample": "3 diffusion steps x 16 scenes (oracle, fp32 torch CPU), scaled to 1000 steps"}
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
