"""Import the UNMODIFIED reference modules from /root/reference in the build container.

TEST INFRASTRUCTURE (used only by oracle/make_golden.py and the container-only validation tests;
/root/reference does not exist on the GPU box).  The reference's model code imports a few
packages that are absent here for reasons unrelated to the hot path; they are stubbed:

  unidecode, inflect          text/cleaners.py, text/numbers.py (pulled in by transformer/Models.py:7)
  matplotlib(.pyplot)         utils/tools.py:7-12
  pytorch_lightning           LightningModule := nn.Module (+ freeze), fastspeech2.py:7,16
  resemblyzer                 VoiceEncoder dummy, speaker_encoder.py:7
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

REF = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "transformer"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch.nn as nn

    if "unidecode" not in sys.modules:
        _stub("unidecode", unidecode=lambda s: s)
    if "inflect" not in sys.modules:
        class _Engine:
            def number_to_words(self, *a, **k):
                return ""
        _stub("inflect", engine=lambda: _Engine())
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        plt = _stub("matplotlib.pyplot")
        mpl.pyplot = plt
    if "pytorch_lightning" not in sys.modules:
        class LightningModule(nn.Module):
            def freeze(self):
                for p in self.parameters():
                    p.requires_grad = False
                self.eval()
        _stub("pytorch_lightning", LightningModule=LightningModule)
    if "resemblyzer" not in sys.modules:
        class VoiceEncoder(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
        _stub("resemblyzer", VoiceEncoder=VoiceEncoder)


def make_preprocessed_dir(stats, n_speaker: int) -> str:
    d = tempfile.mkdtemp(prefix="mtts_pre_")
    with open(os.path.join(d, "stats.json"), "w") as f:
        json.dump(stats, f)
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"spk{i}": i for i in range(n_speaker)}, f)
    return d


def preprocess_config(pre_dir: str):
    return {
        "path": {"preprocessed_path": pre_dir},
        "preprocessing": {
            "pitch": {"feature": "phoneme_level", "normalization": True},
            "energy": {"feature": "phoneme_level", "normalization": True},
            "mel": {"n_mel_channels": 80},
        },
    }


ALGORITHM_CONFIG = {"adapt": {"type": "spk", "speaker_emb": "table",
                              "modules": ["speaker_emb", "variance_adaptor", "decoder", "mel_linear", "postnet"],
                              "task": {"lr": 0.001}}}


def load_reference():
    """Returns the reference's (FastSpeech2, FastSpeech2Loss, transformer module) classes."""
    if not reference_available():
        raise RuntimeError("/root/reference is not available (GPU box?): use the committed goldens")
    install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import transformer  # noqa: F401  (reference package)
    from lightning.model import FastSpeech2, FastSpeech2Loss  # reference classes
    return FastSpeech2, FastSpeech2Loss, transformer


def neutralise_dropout():
    """Dropout -> identity while modules stay in train() mode (BatchNorm keeps batch statistics):
    covers SubLayers.py:54,90; modules.py:223,235; Layers.py:133-134."""
    import torch.nn as nn
    import torch.nn.functional as F

    F.dropout = lambda x, p=0.5, training=True, inplace=False: x
    nn.Dropout.forward = lambda self, x: x
