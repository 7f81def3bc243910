"""Shared plumbing of the trainers in bin/: config loading, distributed init, checkpoints."""
import json
import os
import sys

import torch as th
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pykaldi2_b200 import compat  # noqa: E402

compat.install()

from pykaldi2_b200 import dist as pkdist  # noqa: E402


def str2bool(v):
    """The reference declares boolean flags with type=bool (any non-empty string is truthy,
    bin/train_ce.py:56,62); accept that and the usual spellings of false."""
    if isinstance(v, bool):
        return v
    return str(v).lower() not in ("", "0", "false", "no", "off", "none")


def load_config(train_config, data_config=None):
    with open(train_config) as f:
        config = yaml.safe_load(f)
    if data_config:
        with open(data_config) as f:
            data = yaml.safe_load(f)
        config["source_paths"] = [j for _, j in data.get("clean_source", {}).items()]
        if "dir_noise" in data:
            config["dir_noise_paths"] = [j for _, j in data["dir_noise"].items()]
        if "rir" in data:
            config["rir_paths"] = [j for _, j in data["rir"].items()]
    return config


def init_distributed(flag):
    """-hvd is kept as the reference's switch for multi-GPU; the backend is NCCL via torchrun."""
    if flag or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        rank, world, local = pkdist.init()
        print("Run experiments with world size {}".format(world))
        return rank, world, local
    if th.cuda.is_available():
        th.cuda.set_device(0)
    return 0, 1, 0


def save_checkpoint(path, model, optimizer, epoch=None):
    ckpt = {"model": model.state_dict(), "optimizer": optimizer.state_dict()}
    if epoch is not None:
        ckpt["epoch"] = epoch
    th.save(ckpt, path)


def load_model_state(model, path, strip_module=True):
    ckpt = th.load(path, map_location="cpu")
    state = ckpt["model"] if "model" in ckpt else ckpt
    if strip_module:                     # bin/train_chain.py:149-159
        state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    model.load_state_dict(state)
    return ckpt


def dump_config(config):
    print("Experiment starts with config {}".format(json.dumps(config, sort_keys=True, indent=4)))
