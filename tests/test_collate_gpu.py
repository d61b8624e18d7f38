"""Row f2 on the GPU: `mtts_pack_rows` (ragged -> padded pack through the C ABI) is bit-exact against the reference's
padded 12-tuple for every field / dtype / row width (8-byte ids, 4-byte floats, 320-byte mel rows), including empty tails
and zero-length-free ragged batches; and training_step gives the same result from a collate-staged batch as from a plain
12-tuple."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import collate as B  # noqa: E402
from meta_tts_b200.ops import CudaOps  # noqa: E402
from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem  # noqa: E402
from oracle import collate_oracle as C  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "collate_golden.npz"), allow_pickle=False)


@pytest.mark.parametrize("n,seed", [(9, 0), (1, 3), (32, 7)])
def test_pack_rows_bit_exact(cuda_device, n, seed):
    data = C.synth_dataset(n=n, seed=seed, lmin=1, lmax=40)
    rag = B.reprocess_ragged(data, np.arange(n))
    out = B.pack_on_device(CudaOps(split=3), rag, cuda_device)
    torch.cuda.synchronize()
    ref = C.reprocess(data, np.arange(n))
    for k, i in (("texts", 3), ("mels", 6), ("pitches", 9), ("energies", 10), ("durations", 11), ("src_lens", 4), ("mel_lens", 7)):
        assert out[k].dtype == ref[i].dtype and torch.equal(out[k].cpu(), ref[i]), k
    if n == 9 and seed == 0:                                   # ... and against the REAL reference's output (golden dataset bounds)
        gdata = C.synth_dataset(n=9, seed=0)
        gout = B.pack_on_device(CudaOps(split=3), B.reprocess_ragged(gdata, np.arange(9)), cuda_device)
        for k in ("mels", "texts", "pitches", "energies", "durations"):
            assert np.array_equal(gout[k].cpu().numpy(), G[f"plain_{k}"]), k
    # padding to a larger static shape (CUDA-graph buffers are sized for the shape key, not for this batch)
    L2, T2 = int(ref[5]) + 5, int(ref[8]) + 13
    out2 = B.pack_on_device(CudaOps(split=3), rag, cuda_device, L=L2, T=T2)
    assert torch.equal(out2["mels"][:, :int(ref[8])].cpu(), ref[6]) and float(out2["mels"][:, int(ref[8]):].abs().max()) == 0.0
    assert torch.equal(out2["durations"][:, :int(ref[5])].cpu(), ref[11]) and int(out2["durations"][:, int(ref[5]):].abs().max()) == 0


def test_training_step_from_collate_staged_batch(cuda_device):
    """The producer's pinned staging buffer goes to the device in one copy; results equal the plain 12-tuple path."""
    cfg = O.small_model_config(1, 1)
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 1
    algo["adapt"]["test"]["steps"] = 1
    data = C.synth_dataset(n=5, seed=11, lmin=4, lmax=9)
    for d in data:
        d["speaker"] = 3
    sup, qry = B.SpeakerTaskCollate().get_meta_collate(shots=3, queries=2)(data)
    assert sup[0].staged is not None
    losses = []
    for batch in ([(sup, qry)], [([tuple(sup[0])], [tuple(qry[0])])]):
        sysm = MetaSystem(None, cfg, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cuda:0", dropout=True, seed=5)
        sysm.load_state_dict(O.init_params(seed=0, model_config=cfg))
        out = sysm.training_step(batch, 0)
        losses.append(torch.stack([out["losses"][i] for i in range(6)]).clone())
    assert torch.allclose(losses[0], losses[1], rtol=1e-5), (losses[0], losses[1])
