"""Diagnostic (GPU): config-2 full-size second-order task step with the Hessian-vector pass in single-pass bf16
(MTTS_HVP_SPLIT=1) against the fp32 oracle — prints total / per-tensor outer-gradient errors."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MTTS_HVP_SPLIT", "1")
import torch  # noqa: E402

from oracle import fs2_oracle as O  # noqa: E402
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("teg", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "test_engine_gpu.py"))
_teg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_teg)
_check_task, _engine = _teg._check_task, _teg._engine

cfg = O.BASE_MODEL_CONFIG
P = O.init_params(seed=0)
m = _engine(P, cfg, K=1)
print("hvp_split =", m.engine.hvp_split)
sup, qry = O.synth_task(task=0, shots=4, queries=4, L=128, T=864)
try:
    _check_task(m, P, cfg, sup, qry, 1, False, 1e-3, "config2 full size, HVP single-pass bf16")
except AssertionError as e:
    print("ASSERT:", e)
