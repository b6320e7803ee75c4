#!/usr/bin/env python
"""Generation entry point (reference scripts/generate_diffusion.py:47-468, hot call :314-323).

Builds the network from the same YAML, loads a reference-format checkpoint, and samples scenes.  Rendering,
CAD retrieval and mesh export are outside the hot path (their dependencies are absent here); results are written
as `.npz` dictionaries after the reference's `post_process` (descale + atan2).  Unlike the reference loop
(one scene per call), `--batch_size` scenes are sampled per call and emptiness is decided per scene.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from training_utils import load_config  # noqa: E402

from scene_synthesis.networks import build_network  # noqa: E402
from diffuscene_b200.synthetic import SyntheticScenes  # noqa: E402


def main(argv):
    p = argparse.ArgumentParser(description="Generate scenes using a previously trained model")
    p.add_argument("config_file")
    p.add_argument("output_directory")
    p.add_argument("--weight_file", default=None)
    p.add_argument("--n_sequences", default=10, type=int)
    p.add_argument("--batch_size", default=1, type=int)
    p.add_argument("--clip_denoised", action="store_true")
    p.add_argument("--ddim", action="store_true")
    p.add_argument("--ddim_steps", type=int, default=50)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = p.parse_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("generate_diffusion.py needs a CUDA device (B200); there is no CPU path")
    device = torch.device("cuda:0")
    os.makedirs(args.output_directory, exist_ok=True)
    torch.manual_seed(args.seed)
    config = load_config(args.config_file)
    config["network"]["diffusion_kwargs"]["loss_iou"] = False        # no dataset stats needed to sample
    ds = SyntheticScenes(config["network"], 1)
    network, _, _ = build_network(ds.feature_size, ds.n_classes, config, args.weight_file, device=device,
                                  precision=args.precision)
    network.eval()
    N, d = config["network"]["sample_num_points"], config["network"]["point_dim"]
    done = 0
    while done < args.n_sequences:
        bs = min(args.batch_size, args.n_sequences - done)
        room_mask = torch.zeros(bs, 1, 64, 64, device=device)
        samples = network.sample(room_mask, N, d, batch_size=bs, clip_denoised=args.clip_denoised, ddim=args.ddim,
                                 ddim_steps=args.ddim_steps, seed=args.seed * 100003 + done)
        for i, scene in enumerate(network.delete_empty_batched(samples)):
            boxes = {k: v[None].numpy() for k, v in scene.items() if k != "class_index"}
            out = ds.post_process(boxes)
            out["class_index"] = scene["class_index"].numpy()
            np.savez(os.path.join(args.output_directory, "scene_{:05d}.npz".format(done + i)), **out)
        done += bs
    print("generated {} scenes in {}".format(done, args.output_directory))


if __name__ == "__main__":
    main(sys.argv[1:])
