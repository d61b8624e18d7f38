"""Helpers shared by tests/test_adaptation_{cpu,gpu}.py."""
import torch

from oracle import fs2_oracle as O


def talkative_params(cfg, seed=0):
    """Seeded init + a duration-predictor bias that makes free-running synthesis produce several frames per phoneme
    (the random-init predictor outputs log_d ~ 0 => zero-length output)."""
    P = O.init_params(seed=seed, model_config=cfg)
    g = torch.Generator().manual_seed(99)
    P["variance_adaptor.duration_predictor.linear_layer.bias"] = P["variance_adaptor.duration_predictor.linear_layer.bias"] + 1.3
    w = P["variance_adaptor.duration_predictor.linear_layer.weight"]
    P["variance_adaptor.duration_predictor.linear_layer.weight"] = w + 0.02 * torch.randn(w.shape, generator=g)
    for k in ("pitch", "energy"):                       # spread the predictions over several quantisation bins
        w = P[f"variance_adaptor.{k}_predictor.linear_layer.weight"]
        P[f"variance_adaptor.{k}_predictor.linear_layer.weight"] = w + 0.15 * torch.randn(w.shape, generator=g)
    return P


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_outputs(got, ref, tol=1e-3, loss_tol=1e-3, verbose=None):
    assert [k for k in got if k != "_batch"] == list(ref.keys())
    for step, r in ref.items():
        g = got[step]
        assert set(g) == set(r), step
        for kind in r:
            go, ro = g[kind]["output"], r[kind]["output"]
            assert len(go) == 10
            # integer / index path: rounded durations, masks, lengths — exact
            assert torch.equal(go[5].cpu().to(ro[5].dtype), ro[5]), (step, kind, "d_rounded", go[5], ro[5])
            assert torch.equal(go[6].cpu(), ro[6]) and torch.equal(go[7].cpu(), ro[7]), (step, kind, "masks")
            assert torch.equal(go[8].cpu(), ro[8]) and torch.equal(go[9].cpu(), ro[9]), (step, kind, "lens")
            for i in range(5):
                assert go[i].shape == ro[i].shape, (step, kind, i, go[i].shape, ro[i].shape)
                assert rel(go[i], ro[i]) < tol, (step, kind, i, rel(go[i], ro[i]))
            if kind == "recon":
                lg, lr_ = torch.stack([x.cpu() for x in g[kind]["losses"]]), torch.stack(list(r[kind]["losses"]))
                assert rel(lg, lr_) < loss_tol, (step, lg, lr_)
            if verbose:
                print(f"[adapt] {verbose} {step}/{kind}: T={go[1].shape[1]} postnet rel {rel(go[1], ro[1]):.2e} "
                      f"pitch {rel(go[2], ro[2]):.2e} logd {rel(go[4], ro[4]):.2e}")


