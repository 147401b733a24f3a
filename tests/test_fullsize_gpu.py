"""GPU: numeric parity AT THE BENCHMARKED WIDTHS (BASELINE config 2: Swin-B, embed 128, heads 4/8/16/32, depths
2/2/18/2; and the VIOLET widths 96, heads 3/6/12/24), one clip of 8x224^2, forward + all 327 parameter gradients,
through the default kernel selection (CTA-pair tcgen05 GEMMs, tcgen05 window attention for bf16/fp16).

Checked against (a) the CPU oracle run on the box's host cores on the same seeded inputs and (b) the statistics the
unmodified reference recorded in tests/golden/fullsize.pt (tests/golden/make_golden_fullsize.py).

Tolerances: fp32 rel-L2 1e-4 on the output and on every gradient tensor.  bf16 / fp16: output 2e-2; gradients
PER TENSOR  max(2e-2, 1.5 x the reference's own bf16-autocast-vs-fp32 rel-L2 for that tensor)  -- the reference-side
noise was measured once by the generating script (worst: layers.0.blocks.0.norm1.bias 0.33; median 1.4e-2).
"""
import json
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
OUT = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")

_ORACLE_CACHE = {}


def _setup(vsw, oracle, name):
    fx = torch.load(os.path.join(GOLD, "fullsize.pt"), weights_only=False)[name]
    kw = fx["kwargs"]
    cfg = oracle.SwinCfg(embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]), num_heads=tuple(kw["num_heads"]),
                         window_size=tuple(kw["window_size"]))
    sd = oracle.make_state_dict(cfg, seed=fx["sd_seed"], ln_jitter=fx["ln_jitter"])
    torch.manual_seed(fx["x_seed"])
    x = torch.randn(*fx["x_shape"])
    torch.manual_seed(fx["R_seed"])
    R = torch.randn(*fx["y"].shape) * fx["R_scale"]
    if name not in _ORACLE_CACHE:
        torch.set_num_threads(os.cpu_count() or 1)
        _ORACLE_CACHE[name] = oracle.forward_backward(sd, x, cfg, R)
    return fx, kw, sd, x, R, _ORACLE_CACHE[name]


def _dump(tag, errs, bounds=None):
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, f"fullsize_{tag}.json"), "w") as f:
            json.dump({"errs": errs, "bounds": bounds}, f, indent=0)
    except OSError:
        pass


@pytest.mark.parametrize("name", ["swin_b", "violet"])
def test_fullsize_fp32_vs_oracle_and_reference_stats(vsw, oracle, name):
    fx, kw, sd, x, R, (yo, go) = _setup(vsw, oracle, name)
    m = vsw.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **kw)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    y = m(x.cuda())
    assert rel_l2(y, fx["y"]) < 1e-4 and rel_l2(y, yo) < 1e-4
    (y * R.cuda()).sum().backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert len(grads) == 327
    g = torch.Generator().manual_seed(99)
    errs = {}
    for k, (s, n, p) in fx["grad_stats"].items():      # the reference's own gradient statistics
        r = torch.randn(grads[k].shape, generator=g, dtype=torch.float64)
        v = grads[k].double().cpu()
        assert abs(float(v.norm()) - n) <= 1e-4 * n + 1e-9, k
        assert abs(float((v * r).sum()) - p) <= 1e-4 * n * float(r.norm()) + 1e-7, k
        errs[k] = rel_l2(grads[k], go[k])
    _dump(f"{name}_fp32", errs)
    for k, e in errs.items():
        assert e < 1e-4, (k, e)


@pytest.mark.parametrize("name,mode", [("swin_b", "bf16_model"), ("swin_b", "bf16_autocast"), ("swin_b", "fp16_autocast"),
                                       ("violet", "bf16_model")])
def test_fullsize_low_precision_per_tensor_bounds(vsw, oracle, name, mode):
    fx, kw, sd, x, R, (yo, go) = _setup(vsw, oracle, name)
    m = vsw.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **kw)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    if mode == "bf16_model":
        m = m.bfloat16()
        y = m(x.cuda().bfloat16())
        assert y.dtype == torch.bfloat16
    else:
        dt = torch.bfloat16 if mode == "bf16_autocast" else torch.float16
        with torch.autocast("cuda", dtype=dt):
            y = m(x.cuda())
        assert y.dtype == torch.float32
    e_out = rel_l2(y, yo)
    (y.float() * R.cuda()).sum().backward()
    errs = {k: rel_l2(p.grad, go[k]) for k, p in m.named_parameters()}
    noise = fx["grad_bf16_noise"]
    bounds = {k: max(2e-2, 1.5 * noise[k]) for k in errs}
    _dump(f"{name}_{mode}", dict(errs, __out__=e_out), bounds)
    assert e_out < 2e-2, e_out
    bad = {k: (e, bounds[k]) for k, e in errs.items() if not e < bounds[k]}
    assert not bad, bad
    assert sorted(errs.values())[len(errs) // 2] < 2e-2


def test_use_checkpoint_matches_plain(vsw, oracle):
    """use_checkpoint=True (video_swin.py:293-295) recomputes each block in backward: same output and gradients,
    bit for bit (the kernels are deterministic), in train mode with drop-path rate 0."""
    kw = dict(embed_dim=64, depths=[2, 2], num_heads=[2, 4], window_size=(8, 7, 7))
    cfg = oracle.SwinCfg(embed_dim=64, depths=(2, 2), num_heads=(2, 4), window_size=(8, 7, 7))
    sd = oracle.make_state_dict(cfg, seed=5, ln_jitter=0.1)
    torch.manual_seed(0)
    x = torch.randn(2, 3, 8, 56, 56, device="cuda")
    outs = []
    for ck in (False, True):
        m = vsw.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, use_checkpoint=ck, **kw)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().bfloat16().train()
        y = m(x.bfloat16())
        (y.float() ** 2).sum().backward()
        outs.append((y.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}))
    assert torch.equal(outs[0][0], outs[1][0])
    for k in outs[0][1]:
        assert torch.equal(outs[0][1][k], outs[1][1][k]), k
