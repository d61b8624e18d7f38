"""BASELINE configs[4] timing (GPU): few-shot adaptation inference = 20 first-order inner steps on a 16-shot support set
(128 phonemes -> 864 frames) + free-running synthesis of the query + 60-iteration Griffin-Lim decode.  Prints one JSON line
(a profiles/ record, not the driver's bench line).  Warm numbers: the first call allocates the tapes."""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from meta_tts_b200 import audio as PA  # noqa: E402
from meta_tts_b200 import ops as _ops  # noqa: E402
from meta_tts_b200.maml import batch_from_tuple  # noqa: E402
from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem  # noqa: E402
from meta_tts_b200 import synthetic as O  # noqa: E402  (synthetic task generator + random-init weights)


def timed(fn, reps=3):
    best = 1e30
    out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 5
    algo["adapt"]["test"] = {"steps": 20, "saving_steps": [20]}
    sysm = MetaSystem(None, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cuda:0", use_cuda_graph=False, dropout=True, seed=0)
    P = O.init_state_dict(DEFAULT_MODEL_CONFIG, seed=0)
    P["variance_adaptor.duration_predictor.linear_layer.bias"] = P["variance_adaptor.duration_predictor.linear_layer.bias"] + 2.0   # ~6.4 frames / phoneme
    sysm.load_state_dict(P)
    sup, qry = O.synth_task(task=9, shots=16, queries=1, L=128, T=864)
    m, dev = sysm.maml, sysm.device
    bs = batch_from_tuple(sup, dev)
    bq = batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True, targets=False)
    from meta_tts_b200.systems import _set_salt
    _set_salt(sysm, 1)
    tape = m.engine.new_tape()
    n0 = _ops.launch_count
    m.adapt_rolling(bs, 1, tape, fresh=True, drop_base=0)                 # allocation + warm-up
    per_step_launches = _ops.launch_count - n0
    t_adapt, _ = timed(lambda: m.adapt_rolling(bs, 20, tape, fresh=True, drop_base=0), reps=2)
    t_synth, out = timed(lambda: m.predict(bq, adapted=True, free_running=True, eval_mode=False, drop_pass=20020))
    T = int(out["T"])
    stft = PA.TacotronSTFT(1024, 256, 1024, 80, 22050, 0, 8000, device="cuda:0")
    mel = out["postnet"][0].t().contiguous()
    PA.inv_mel_spec(mel, None, stft, 60)                                    # warm-up
    t_gl, wav = timed(lambda: PA.inv_mel_spec(mel, None, stft, 60))
    t_gl_eager, _ = timed(lambda: PA.griffin_lim_fm(stft.spec_from_mel_fm(torch.nn.functional.pad(mel.t()[None], (0, 0)).contiguous())[:, :T - 1].contiguous(),
                                                    stft.stft_fn, 60, stft.stft_fn._to_fm(torch.zeros(1, 513, T - 1)), use_graph=False))
    sup_frames = 16 * 864
    rec = {"workload": "BASELINE configs[4]: 20 first-order inner steps on a 16-shot support set (128 phonemes -> 864 frames), free-running "
                       "synthesis of 1 query, Griffin-Lim x60 (n_fft 1024, hop 256)",
           "adapt_20_steps_ms": t_adapt, "adapt_ms_per_step": t_adapt / 20, "support_frame_passes_per_s": 20 * sup_frames / (t_adapt * 1e-3),
           "launches_per_inner_step": per_step_launches, "synthesis_forward_ms": t_synth, "synthesised_frames": T,
           "griffin_lim_60_ms": t_gl, "griffin_lim_60_eager_ms": t_gl_eager, "waveform_samples": int(wav.shape[0]),
           "total_ms": t_adapt + t_synth + t_gl, "dtype": "bf16x3", "dropout": "train mode (adapted learner), counter-hash masks"}
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
