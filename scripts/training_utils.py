"""Config / checkpoint helpers with the reference's file formats (scripts/training_utils.py:15-96):
YAML experiment configs, `model_{epoch:05d}` / `opt_{epoch:05d}` checkpoint pairs, newest-pair resume."""
import json
import os
import random
import string

import torch
import yaml


def load_config(config_file):
    with open(config_file, "r") as f:
        return yaml.safe_load(f.read().replace("\r", ""))      # the reference's YAMLs carry CRLF line endings


def id_generator(size=6, chars=string.ascii_uppercase + string.digits):
    return "".join(random.choice(chars) for _ in range(size))


def save_experiment_params(args, experiment_tag, directory):
    params = {k: (None if str(v) == "" else str(v)) for k, v in vars(args).items()}
    params["experiment_tag"] = experiment_tag
    if hasattr(args, "config_file"):
        params.update(load_config(args.config_file))
    with open(os.path.join(directory, "params.json"), "w") as f:
        json.dump(params, f, indent=4)


def load_checkpoints(model, optimizer, experiment_directory, args, device):
    ids = [int(f[6:]) for f in os.listdir(experiment_directory) if f.startswith("model_")]
    if not ids:
        return
    max_id = max(ids)
    model_path = os.path.join(experiment_directory, "model_{:05d}".format(max_id))
    opt_path = os.path.join(experiment_directory, "opt_{:05d}".format(max_id))
    if not (os.path.exists(model_path) and os.path.exists(opt_path)):
        return
    print("Loading model checkpoint from {}".format(model_path))
    model.load_state_dict(torch.load(model_path, map_location=device))
    print("Loading optimizer checkpoint from {}".format(opt_path))
    optimizer.load_state_dict(torch.load(opt_path, map_location=device))
    args.continue_from_epoch = max_id + 1


def save_checkpoints(epoch, model, optimizer, experiment_directory):
    torch.save(model.state_dict(), os.path.join(experiment_directory, "model_{:05d}".format(epoch)))
    torch.save(optimizer.state_dict(), os.path.join(experiment_directory, "opt_{:05d}".format(epoch)))
