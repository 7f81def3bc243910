"""Import aliases so code written against the reference's top-level packages runs unchanged:

    import pykaldi2_b200.compat; pykaldi2_b200.compat.install()
    from ops import ops                  # -> pykaldi2_b200.ops.ops       (reference ops/ops.py)
    from models import lstm              # -> pykaldi2_b200.models.lstm   (reference models/lstm.py)
    from models import transformer       # -> pykaldi2_b200.models.transformer (reference models/transformer.py)
    from reader.preprocess import GlobalMeanVarianceNormalization         (unpickles transform.pkl)
    from utils import utils
    from data import ChunkDataloader, SeqDataloader
"""
import importlib
import sys

_ALIASES = {
    "ops": "pykaldi2_b200.ops", "ops.ops": "pykaldi2_b200.ops.ops",
    "models": "pykaldi2_b200.models", "models.lstm": "pykaldi2_b200.models.lstm",
    "models.transformer": "pykaldi2_b200.models.transformer",
    "reader": "pykaldi2_b200.reader", "reader.preprocess": "pykaldi2_b200.reader.preprocess",
    "utils": "pykaldi2_b200.utils", "utils.utils": "pykaldi2_b200.utils.utils",
    "data": "pykaldi2_b200.data", "data.dataloader": "pykaldi2_b200.data.dataloader",
}


def install(force=False):
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not force:
            continue
        sys.modules[alias] = importlib.import_module(target)
    data = sys.modules["data"]
    dl = sys.modules["data.dataloader"]
    for name in ("ChunkDataloader", "SeqDataloader", "SyntheticWaveDataset", "WaveDataloader"):
        setattr(data, name, getattr(dl, name))
    data.SpeechDataset = importlib.import_module("pykaldi2_b200.data.speech_dataset").SpeechDataset
