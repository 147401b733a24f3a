"""Reference-side bf16 noise of the tiny fixtures (build container only).

    python tests/golden/make_golden_tiny_noise.py

For tiny_a / tiny_b / tiny_d: the UNMODIFIED reference module under torch.autocast(bf16) on the CPU against its own fp32 run,
rel-L2 per gradient tensor (and for the output) -> tests/golden/tiny_bf16_noise.json.  tests/test_model_gpu.py derives its
bf16 bounds from these numbers, bound[k] = max(3e-2, 2 x noise[k]), instead of a blanket tolerance (the tiny models are far
noisier than the real widths: few tokens, LayerNorm over 32 channels).  Inputs and weights are those of make_golden.py.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import TINY, import_reference  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    from oracle import swin3d_oracle as O
    vs = import_reference()
    out = {}
    for name in ("tiny_a", "tiny_b", "tiny_d"):
        kw, xshape = TINY[name]
        cfg = O.SwinCfg(embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]), num_heads=tuple(kw["num_heads"]),
                        window_size=tuple(kw["window_size"]))
        sd = O.make_state_dict(cfg, seed=1234, ln_jitter=0.1)
        torch.manual_seed(7)
        x = torch.randn(*xshape)
        runs = []
        for autocast in (False, True):
            m = vs.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **kw)
            m.load_state_dict(sd, strict=True)
            m.eval()
            with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
                y = m(x)
            if not runs:
                torch.manual_seed(11)
                R = torch.randn(*y.shape) / 64.0
            (y.float() * R).sum().backward()
            runs.append((y.detach().float(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}))
        (y32, g32), (y16, g16) = runs
        noise = {k: rel(g16[k], g32[k]) for k in g32}
        out[name] = {"out": rel(y16, y32), "grads": noise}
        srt = sorted(noise.items(), key=lambda kv: -kv[1])
        print(f"[{name}] out {out[name]['out']:.3e}; grads median {sorted(noise.values())[len(noise) // 2]:.3e}; worst: "
              + ", ".join(f"{k}={v:.3f}" for k, v in srt[:4]))
    with open(os.path.join(HERE, "tiny_bf16_noise.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
