"""Drop-in module classes: identical state_dict keys / seeded init as the reference (via the pinned
oracle.init_params) and forward parity of Encoder / Decoder / PostNet / VariancePredictor / FastSpeech2 /
LengthRegulator semantics, with the CPU restatement of the op set injected as backend."""
import json
import os
import tempfile

import pytest
import torch

from meta_tts_b200 import modules as M
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps

CFG = O.small_model_config(1, 1)


@pytest.fixture(scope="module")
def pre_cfg():
    d = tempfile.mkdtemp(prefix="mtts_pre_")
    json.dump(O.DEFAULT_STATS, open(os.path.join(d, "stats.json"), "w"))
    json.dump({f"s{i}": i for i in range(16)}, open(os.path.join(d, "speakers.json"), "w"))
    return {"path": {"preprocessed_path": d},
            "preprocessing": {"pitch": {"feature": "phoneme_level"}, "energy": {"feature": "phoneme_level"},
                              "mel": {"n_mel_channels": 80}}}


ALGO = {"adapt": {"type": "spk", "speaker_emb": "table"}}


def test_state_dict_keys_and_seeded_init_match_reference(pre_cfg):
    torch.manual_seed(0)
    model = M.FastSpeech2(pre_cfg, O.BASE_MODEL_CONFIG, ALGO)
    P = O.init_params(seed=0)          # bit-identical to the real reference (tests/test_oracle_golden.py)
    sd = model.state_dict()
    assert sorted(sd.keys()) == sorted(P.keys())
    for k in P:
        assert torch.equal(sd[k], P[k].detach()), k


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_forward_parity_with_injected_backend(pre_cfg):
    be = RefOps(split=3)
    M._B200Module._backend = be
    try:
        torch.manual_seed(0)
        model = M.FastSpeech2(pre_cfg, CFG, ALGO).train()
        P = O.init_params(seed=0, model_config=CFG)
        b12 = O.synth_batch(2, 7, 20, seed=4, speaker=1, ragged=True)
        with torch.no_grad():
            ref = O.fs2_forward({k: v.detach().clone() for k, v in P.items()}, CFG, *b12[2:])
        out = model(*b12[2:])
        for i in range(5):
            assert _rel(out[i].reshape(ref[i].shape), ref[i]) < 2e-5, i
        assert torch.equal(out[6], ref[6]) and torch.equal(out[7], ref[7]) and torch.equal(out[9], ref[9])
        # stand-alone Encoder / Decoder keep the reference signatures
        src_mask, mel_mask = ref[6], ref[7]
        enc_o = model.encoder(b12[3], src_mask)
        assert _rel(enc_o, O.encoder({k: v.detach() for k, v in P.items()}, CFG, b12[3], src_mask)) < 2e-5
        x = torch.randn(2, 20, 256, generator=torch.Generator().manual_seed(1))
        dec_o, m2 = model.decoder(x, mel_mask)
        dref, mref = O.decoder({k: v.detach() for k, v in P.items()}, CFG, x, mel_mask)
        assert _rel(dec_o, dref) < 2e-5 and torch.equal(m2, mref)
        post = model.postnet(ref[0])
        pref = O.postnet({k: v.detach().clone() for k, v in P.items()}, ref[0], True)
        assert _rel(post, pref) < 2e-5
        vp = model.variance_adaptor.duration_predictor(enc_o, src_mask)
        vref = O.variance_predictor({k: v.detach() for k, v in P.items()}, "variance_adaptor.duration_predictor", enc_o, src_mask)
        assert _rel(vp, vref) < 5e-5
    finally:
        M._B200Module._backend = None


def test_product_modules_refuse_to_run_without_gpu(pre_cfg):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    enc = M.Encoder(CFG)
    with pytest.raises(Exception):
        enc(torch.ones(1, 4, dtype=torch.long), torch.zeros(1, 4, dtype=torch.bool))


def test_free_running_forward_matches_oracle(pre_cfg):
    """`FastSpeech2.forward(spk, texts, src_lens, max_src_len)` without targets (fastspeech2.py:73-92): eval and train mode."""
    be = RefOps(split=3)
    M._B200Module._backend = be
    try:
        torch.manual_seed(0)
        model = M.FastSpeech2(pre_cfg, CFG, ALGO)
        sd = model.state_dict()
        sd["variance_adaptor.duration_predictor.linear_layer.bias"] = sd["variance_adaptor.duration_predictor.linear_layer.bias"] + 1.3
        model.load_state_dict(sd)
        b12 = O.synth_batch(2, 7, 20, seed=4, speaker=1, ragged=True)
        for train in (False, True):
            model.train(train)
            P = {k: v.detach().clone() for k, v in model.state_dict().items()}
            with torch.no_grad():
                ref = O.fs2_forward(P, CFG, *b12[2:6], d_control=1.2, training=train)
            out = model(*b12[2:6], d_control=1.2)
            assert torch.equal(out[5], ref[5]) and torch.equal(out[9], ref[9]) and int(ref[9].max()) > 7
            assert torch.equal(out[6], ref[6]) and torch.equal(out[7], ref[7])
            for i in range(5):
                assert out[i].shape == ref[i].shape and _rel(out[i], ref[i]) < 5e-5, (train, i)
    finally:
        M._B200Module._backend = None


def test_decoder_beyond_max_seq_len(pre_cfg):
    """Decoder.forward on a sequence longer than max_seq_len: train mode truncates (Models.py:161-166), eval mode keeps the length
    with a recomputed sinusoid table (Models.py:148-156)."""
    import copy
    cfg = copy.deepcopy(CFG)
    cfg["max_seq_len"] = 24
    M._B200Module._backend = RefOps(split=3)
    try:
        torch.manual_seed(0)
        dec = M.Decoder(cfg)
        P = {"decoder." + k: v.detach().clone() for k, v in dec.state_dict().items()}
        x = torch.randn(2, 40, 256, generator=torch.Generator().manual_seed(3))
        mask = O.get_mask_from_lengths(torch.tensor([40, 31]), 40)
        for train in (True, False):
            dec.train(train)
            out, m = dec(x, mask)
            ref, mref = O.decoder(P, cfg, x, mask, training=train)
            assert out.shape == ref.shape and out.shape[1] == (24 if train else 40) and torch.equal(m, mref)
            assert _rel(out, ref) < 2e-5, train
    finally:
        M._B200Module._backend = None


def test_postnet_eval_mode_uses_running_statistics(pre_cfg):
    M._B200Module._backend = RefOps(split=3)
    try:
        torch.manual_seed(0)
        post = M.PostNet()
        x = torch.randn(2, 30, 80, generator=torch.Generator().manual_seed(2))
        post.train()
        post(x)                                                             # advances the running statistics once
        P = {"postnet." + k: v.detach().clone() for k, v in post.state_dict().items()}
        assert float(P["postnet.convolutions.0.1.running_mean"].abs().max()) > 0 and int(P["postnet.convolutions.0.1.num_batches_tracked"]) == 1
        post.eval()
        out = post(x)
        ref = O.postnet(P, x, training=False)
        assert _rel(out, ref) < 2e-5
        assert _rel(out, O.postnet({k: v.clone() for k, v in P.items()}, x, training=True)) > 1e-2     # and differs from batch statistics
    finally:
        M._B200Module._backend = None


def test_variance_adaptor_standalone_forward(pre_cfg):
    """VarianceAdaptor.forward alone (modules.py:102-158): teacher forced and free running (with controls)."""
    M._B200Module._backend = RefOps(split=3)
    try:
        torch.manual_seed(0)
        va = M.VarianceAdaptor(pre_cfg, CFG)
        sd = va.state_dict()
        sd["duration_predictor.linear_layer.bias"] = sd["duration_predictor.linear_layer.bias"] + 1.3
        va.load_state_dict(sd)
        P = {"variance_adaptor." + k: v.detach().clone() for k, v in va.state_dict().items()}
        b12 = O.synth_batch(2, 7, 20, seed=4, speaker=1, ragged=True)
        x = torch.randn(2, 7, 256, generator=torch.Generator().manual_seed(8))
        src_mask = O.get_mask_from_lengths(b12[4], 7)
        mel_mask = O.get_mask_from_lengths(b12[7], 20)
        with torch.no_grad():
            ref_t = O.variance_adaptor(P, x, src_mask, mel_mask, 20, b12[9], b12[10], b12[11])
            ref_f = O.variance_adaptor(P, x, src_mask, None, None, None, None, None, 1.1, 0.9, 1.4)
        out_t = va(x, src_mask, mel_mask, 20, b12[9], b12[10], b12[11])
        out_f = va(x, src_mask, p_control=1.1, e_control=0.9, d_control=1.4)
        for out, ref in ((out_t, ref_t), (out_f, ref_f)):
            assert len(out) == 7 and out[0].shape == ref[0].shape
            for i in range(4):
                assert _rel(out[i], ref[i]) < 5e-5, i
            assert torch.equal(torch.as_tensor(out[4]).float(), ref[4].float()) and torch.equal(out[5], ref[5]) and torch.equal(out[6], ref[6])
        assert int(ref_f[5].max()) > 7
    finally:
        M._B200Module._backend = None


def test_train_mode_forward_advances_the_module_bn_buffers(pre_cfg):
    """After a train-mode FastSpeech2 forward the module's own BatchNorm buffers (state_dict / checkpoints) hold the advanced
    running statistics, exactly as the reference's nn.BatchNorm1d would, and a following eval() forward uses them."""
    M._B200Module._backend = RefOps(split=3)
    try:
        torch.manual_seed(0)
        model = M.FastSpeech2(pre_cfg, CFG, ALGO).train()
        P = {k: v.detach().clone() for k, v in model.state_dict().items()}
        b12 = O.synth_batch(2, 7, 20, seed=4, speaker=1, ragged=True)
        model(*b12[2:])
        with torch.no_grad():
            O.fs2_forward(P, CFG, *b12[2:], training=True)                    # the oracle updates P's running statistics in place
        sd = model.state_dict()
        for i in range(5):
            for k in ("running_mean", "running_var"):
                assert _rel(sd[f"postnet.convolutions.{i}.1.{k}"], P[f"postnet.convolutions.{i}.1.{k}"]) < 1e-4, (i, k)
            assert int(sd[f"postnet.convolutions.{i}.1.num_batches_tracked"]) == 1
        model.eval()
        out = model(*b12[2:])
        with torch.no_grad():
            ref = O.fs2_forward({k: v.detach().clone() for k, v in sd.items()}, CFG, *b12[2:], training=False)
        assert _rel(out[1], ref[1]) < 2e-5
    finally:
        M._B200Module._backend = None
