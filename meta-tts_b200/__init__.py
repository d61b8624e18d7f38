"""meta-tts_b200 — B200-native Meta-TTS meta-training hot path (hand-written sm_100a CUDA behind a
C ABI, Python host mirroring the reference's nn.Module / LightningModule API).

The directory name carries a hyphen (task layout); import it as `meta_tts_b200` (shim package at
the repo root that points its __path__ here).
"""
__version__ = "0.1.0"
