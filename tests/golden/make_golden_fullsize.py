"""Full-size golden statistics from the UNMODIFIED reference module (build container only).

    python tests/golden/make_golden_fullsize.py

For the two real 224^2 configurations (Swin-B widths of BASELINE config 2 and the VIOLET widths of config 1/3), one
clip of 8x224^2, loss = sum(y * R):
  * fp32 reference run: output + per-parameter gradient statistics (sum, norm, one random projection) -- the GPU
    test holds the product path against these at the benchmarked widths (heads 16/32, C = 512/1024, depth 18);
  * the SAME reference under torch.autocast(bf16) on the CPU: per-tensor rel-L2 of its gradients against its own
    fp32 gradients.  That is the reference-side bf16 noise the bf16 parity bound is derived from
    (bound[k] = max(2e-2, 1.5 * ref_noise[k])) instead of a blanket tolerance.
Only statistics are stored (fullsize.pt, < 1 MB); inputs and weights are re-created from seeds through the oracle's
make_state_dict, exactly like the tiny fixtures.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference, sha16  # noqa: E402

CONFIGS = {
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=(8, 7, 7)),
    "violet": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=(8, 7, 7)),
}
X_SHAPE = (1, 3, 8, 224, 224)
SD_SEED, LN_JITTER, X_SEED, R_SEED, R_SCALE = 4321, 0.1, 17, 19, 1.0 / 64.0


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run(vs, kw, sd, x, R, autocast):
    m = vs.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **kw)
    m.load_state_dict(sd, strict=True)
    m.eval()
    t0 = time.time()
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        y = m(x)
    (y.float() * R).sum().backward()
    print(f"   reference {'bf16-autocast' if autocast else 'fp32'} fwd+bwd {time.time() - t0:.1f} s", flush=True)
    return y.detach().float(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}


def main():
    from oracle import swin3d_oracle as O
    vs = import_reference()
    torch.set_num_threads(os.cpu_count())
    out = {}
    for name, kw in CONFIGS.items():
        cfg = O.SwinCfg(embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]), num_heads=tuple(kw["num_heads"]),
                        window_size=tuple(kw["window_size"]))
        sd = O.make_state_dict(cfg, seed=SD_SEED, ln_jitter=LN_JITTER)
        torch.manual_seed(X_SEED)
        x = torch.randn(*X_SHAPE)
        torch.manual_seed(R_SEED)
        R = torch.randn(1, 8 * kw["embed_dim"], 8, 7, 7) * R_SCALE
        y32, g32 = run(vs, kw, sd, x, R, False)
        assert y32.shape == R.shape
        y16, g16 = run(vs, kw, sd, x, R, True)
        t0 = time.time()
        yo, go = O.forward_backward(sd, x, cfg, R)
        print(f"   oracle fwd+bwd {time.time() - t0:.1f} s; oracle-vs-reference out {rel(yo, y32):.2e}, worst grad "
              f"{max(rel(go[k], g32[k]) for k in g32):.2e}", flush=True)
        assert rel(yo, y32) < 1e-5 and max(rel(go[k], g32[k]) for k in g32) < 1e-4
        g = torch.Generator().manual_seed(99)
        stats, noise = {}, {}
        for k, v in g32.items():
            r = torch.randn(v.shape, generator=g, dtype=torch.float64)
            v64 = v.double()
            stats[k] = [float(v64.sum()), float(v64.norm()), float((v64 * r).sum())]
            noise[k] = rel(g16[k], v)
        srt = sorted(noise.items(), key=lambda kv: -kv[1])
        print(f"[{name}] reference bf16-autocast vs fp32: out {rel(y16, y32):.3e}; grads median "
              f"{sorted(noise.values())[len(noise) // 2]:.3e}, worst 5: " + ", ".join(f"{k}={v:.3f}" for k, v in srt[:5]))
        out[name] = dict(kwargs=kw, x_shape=list(X_SHAPE), sd_seed=SD_SEED, ln_jitter=LN_JITTER, x_seed=X_SEED,
                         R_seed=R_SEED, R_scale=R_SCALE, x_sha=sha16(x),
                         sd_sha=sha16(torch.cat([v.flatten().double() for v in sd.values()])),
                         y=y32.contiguous().clone(), y_bf16_noise=rel(y16, y32), grad_stats=stats,
                         grad_bf16_noise=noise, torch_version=torch.__version__)
    torch.save(out, os.path.join(HERE, "fullsize.pt"))
    print("written", os.path.join(HERE, "fullsize.pt"))


if __name__ == "__main__":
    main()
