"""The drop-in nn.Module classes on the B200 (no backend injected: CudaOps through the C ABI): FastSpeech2 teacher forced and free
running (train / eval), Decoder beyond max_seq_len in eval mode, PostNet eval statistics, stand-alone VarianceAdaptor — against the
oracle, which is pinned to the real reference modules."""
import copy
import json
import os
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import modules as M  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402

CFG = O.small_model_config(1, 1)
ALGO = {"adapt": {"type": "spk", "speaker_emb": "table"}}


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def pre_cfg(cuda_device):
    d = tempfile.mkdtemp(prefix="mtts_pre_")
    json.dump(O.DEFAULT_STATS, open(os.path.join(d, "stats.json"), "w"))
    json.dump({f"s{i}": i for i in range(16)}, open(os.path.join(d, "speakers.json"), "w"))
    return {"path": {"preprocessed_path": d},
            "preprocessing": {"pitch": {"feature": "phoneme_level"}, "energy": {"feature": "phoneme_level"}, "mel": {"n_mel_channels": 80}}}


def test_fastspeech2_teacher_forced_and_free_running(pre_cfg):
    torch.manual_seed(0)
    model = M.FastSpeech2(pre_cfg, CFG, ALGO)
    sd = model.state_dict()
    sd["variance_adaptor.duration_predictor.linear_layer.bias"] = sd["variance_adaptor.duration_predictor.linear_layer.bias"] + 1.3
    model.load_state_dict(sd)
    b12 = O.synth_batch(2, 7, 20, seed=4, speaker=1, ragged=True)
    for train in (True, False):
        model.train(train)
        P = {k: v.detach().clone() for k, v in model.state_dict().items()}
        with torch.no_grad():
            ref_t = O.fs2_forward(copy.deepcopy(P), CFG, *b12[2:], training=train)
            ref_f = O.fs2_forward(copy.deepcopy(P), CFG, *b12[2:6], d_control=1.2, training=train)
        out_t = model(*b12[2:])
        P2 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        if train:                                                      # the train-mode forward above advanced the running statistics
            with torch.no_grad():
                ref_f = O.fs2_forward(copy.deepcopy(P2), CFG, *b12[2:6], d_control=1.2, training=train)
        out_f = model(*b12[2:6], d_control=1.2)
        torch.cuda.synchronize()
        for out, ref in ((out_t, ref_t), (out_f, ref_f)):
            for i in range(5):
                assert out[i].shape == ref[i].shape and _rel(out[i], ref[i]) < 1e-3, (train, i, _rel(out[i], ref[i]))
            assert torch.equal(out[6].cpu(), ref[6]) and torch.equal(out[7].cpu(), ref[7]) and torch.equal(out[9].cpu(), ref[9])
        assert torch.equal(out_f[5].cpu(), ref_f[5]) and int(ref_f[9].max()) > 7


def test_decoder_postnet_variance_adaptor_modules(pre_cfg):
    cfg = copy.deepcopy(CFG)
    cfg["max_seq_len"] = 24
    torch.manual_seed(0)
    dec = M.Decoder(cfg)
    Pd = {"decoder." + k: v.detach().clone() for k, v in dec.state_dict().items()}
    x = torch.randn(2, 40, 256, generator=torch.Generator().manual_seed(3))
    mask = O.get_mask_from_lengths(torch.tensor([40, 31]), 40)
    for train in (True, False):
        dec.train(train)
        out, m = dec(x, mask)
        ref, mref = O.decoder(Pd, cfg, x, mask, training=train)
        assert out.shape == ref.shape and torch.equal(m.cpu(), mref) and _rel(out, ref) < 1e-3, train
    post = M.PostNet()
    xm = torch.randn(2, 30, 80, generator=torch.Generator().manual_seed(2))
    post.train()
    post(xm)
    Pp = {"postnet." + k: v.detach().clone() for k, v in post.state_dict().items()}
    post.eval()
    assert _rel(post(xm), O.postnet(Pp, xm, training=False)) < 1e-3
    va = M.VarianceAdaptor(pre_cfg, CFG)
    sd = va.state_dict()
    sd["duration_predictor.linear_layer.bias"] = sd["duration_predictor.linear_layer.bias"] + 1.3
    va.load_state_dict(sd)
    Pv = {"variance_adaptor." + k: v.detach().clone() for k, v in va.state_dict().items()}
    xe = torch.randn(2, 7, 256, generator=torch.Generator().manual_seed(8))
    src_mask = O.get_mask_from_lengths(torch.tensor([7, 5]), 7)
    with torch.no_grad():
        ref = O.variance_adaptor(Pv, xe, src_mask, None, None, None, None, None, 1.1, 0.9, 1.4)
    out = va(xe, src_mask, p_control=1.1, e_control=0.9, d_control=1.4)
    assert torch.equal(out[4].cpu(), ref[4]) and torch.equal(out[5].cpu(), ref[5]) and torch.equal(out[6].cpu(), ref[6])
    assert out[0].shape == ref[0].shape and _rel(out[0], ref[0]) < 1e-3 and _rel(out[1], ref[1]) < 1e-3
