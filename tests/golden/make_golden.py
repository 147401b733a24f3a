"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference module.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (tsujuifu/pytorch_empirical-mvm, visbackbone/video_swin.py) has no tests or golden
vectors of its own, so these fixtures are outputs of the reference itself:
  * index.json   -- SHA-256 + summary statistics of relative_position_index, compute_mask, the
                    roll+window_partition gather map and get_window_size for every geometry the
                    BASELINE configs (and the edge cases of SURVEY 8c) exercise;
  * tiny_*.pt    -- for small SwinTransformer3D configurations: state_dict seed recipe, the input
                    clip, the forward output, and per-parameter gradient statistics + a few full
                    gradient tensors for loss = sum(y * R).
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    """stub addict / yapf, then import visbackbone.video_swin."""
    if "addict" not in sys.modules:
        addict = types.ModuleType("addict")

        class Dict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            __setattr__ = dict.__setitem__

        addict.Dict = Dict
        sys.modules["addict"] = addict
    for name in ("yapf", "yapf.yapflib", "yapf.yapflib.yapf_api"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["yapf.yapflib.yapf_api"].FormatCode = lambda s, **kw: (s, False)
    sys.path.insert(0, "/root/reference")
    import warnings
    warnings.filterwarnings("ignore")
    import visbackbone.video_swin as vs
    return vs


def sha16(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


MASK_GEOMS = [  # (padded grid, effective window, effective shift)
    ((8, 56, 56), (8, 7, 7), (0, 3, 3)),
    ((8, 28, 28), (8, 7, 7), (0, 3, 3)),
    ((8, 14, 14), (8, 7, 7), (0, 3, 3)),
    ((8, 7, 7), (8, 7, 7), (0, 0, 0)),
    ((8, 96, 96), (8, 12, 12), (0, 6, 6)),
    ((16, 56, 56), (8, 7, 7), (4, 3, 3)),
    ((4, 56, 56), (4, 7, 7), (0, 3, 3)),
    ((16, 14, 21), (8, 7, 7), (4, 3, 3)),
    ((2, 4, 4), (2, 4, 4), (0, 0, 0)),
    ((4, 6, 6), (2, 3, 3), (1, 1, 1)),
]
RPI_WINDOWS = [(8, 7, 7), (8, 12, 12), (16, 7, 7), (2, 7, 7), (2, 3, 3), (4, 4, 4)]
WS_CASES = [((8, 56, 56), (8, 7, 7), (4, 3, 3)), ((8, 7, 7), (8, 7, 7), (4, 3, 3)),
            ((16, 56, 56), (8, 7, 7), (4, 3, 3)), ((4, 56, 56), (8, 7, 7), (4, 3, 3)),
            ((8, 4, 4), (8, 7, 7), (4, 3, 3)), ((8, 96, 96), (8, 12, 12), (4, 6, 6)),
            ((8, 12, 12), (8, 12, 12), (4, 6, 6)), ((3, 5, 9), (2, 7, 7), (1, 3, 3))]


def index_golden(vs):
    out = {"relative_position_index": {}, "compute_mask": {}, "gather_map": {}, "get_window_size": []}
    for w in RPI_WINDOWS:
        idx = vs.WindowAttention3D(32, w, 1).relative_position_index
        out["relative_position_index"]["x".join(map(str, w))] = dict(
            shape=list(idx.shape), sum=int(idx.sum()), min=int(idx.min()), max=int(idx.max()),
            c00=int(idx[0, 0]), c01=int(idx[0, 1]), c0m1=int(idx[0, -1]), cm10=int(idx[-1, 0]),
            sha=sha16(idx.to(torch.int64)))
    for pg, ws, ss in MASK_GEOMS:
        m = vs.compute_mask.__wrapped__(pg[0], pg[1], pg[2], ws, ss, torch.device("cpu"))
        key = "g%s_w%s_s%s" % ("x".join(map(str, pg)), "x".join(map(str, ws)), "x".join(map(str, ss)))
        out["compute_mask"][key] = dict(
            grid=list(pg), window=list(ws), shift=list(ss), shape=list(m.shape),
            nonzero=int((m != 0).sum()), masked_windows=int(((m != 0).flatten(1).any(1)).sum()),
            sha=sha16(m.float()))
        D, H, W = pg
        ar = torch.arange(D * H * W).view(1, D, H, W, 1)
        rolled = torch.roll(ar, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3)) if any(ss) else ar
        gm = vs.window_partition(rolled, ws).squeeze(-1)
        out["gather_map"][key] = dict(shape=list(gm.shape), first5=gm[0, :5].tolist(),
                                      last3=gm[-1, -3:].tolist(), sha=sha16(gm.to(torch.int64)))
    for grid, w, s in WS_CASES:
        ws, ss = vs.get_window_size(grid, w, s)
        out["get_window_size"].append(dict(grid=list(grid), window=list(w), shift=list(s),
                                           ws=list(ws), ss=list(ss)))
    return out


TINY = {
    # name: (ctor kwargs, input shape).  head_dim is 32 everywhere like the real models.
    # tiny_a: window divides the grid at stage 0/1 (shifted windows + mask), clamps at stage 2/3
    #         (window (8,4,4)->N=128 and (8,2,2)->N=32: exercises the index[:N,:N] slice quirk).
    "tiny_a": (dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], window_size=(8, 7, 7)),
               (2, 3, 8, 112, 112)),
    # tiny_b: grid not a multiple of the window (post-norm zero padding), odd H/W in PatchMerging,
    #         input not a multiple of the patch, D=4 < window (temporal clamp, the shipped size_frame=4 case)
    "tiny_b": (dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=(8, 7, 7)),
               (2, 3, 4, 74, 90)),
    # tiny_c: D=16 > window depth -> real temporal shift of 4 (27-region mask)
    "tiny_c": (dict(embed_dim=32, depths=[2], num_heads=[1], window_size=(8, 7, 7)),
               (1, 3, 16, 56, 56)),
    # tiny_d: violet-like head count 3 (C=96) and Swin-L-like window (8,12,12) at small scale
    "tiny_d": (dict(embed_dim=96, depths=[2, 2], num_heads=[3, 6], window_size=(4, 6, 6)),
               (1, 3, 8, 96, 96)),
}


def model_golden(vs, name):
    from oracle import swin3d_oracle as O
    kw, xshape = TINY[name]
    cfg = O.SwinCfg(embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]), num_heads=tuple(kw["num_heads"]),
                    window_size=tuple(kw["window_size"]))
    sd = O.make_state_dict(cfg, seed=1234, ln_jitter=0.1)
    torch.manual_seed(7)
    x = torch.randn(*xshape)
    ref = vs.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **kw)
    missing = ref.load_state_dict(sd, strict=True)
    ref.eval()
    y = ref(x)
    torch.manual_seed(11)
    R = torch.randn(*y.shape) / 64.0  # explicit recipe: contiguous (B,C,D,H,W) order
    (y * R).sum().backward()
    grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    # oracle vs reference, checked at generation time as well
    yo, go = O.forward_backward(sd, x, cfg, R)
    err_y = ((yo - y).norm() / y.norm()).item()
    err_g = max(((go[k] - grads[k]).norm() / (grads[k].norm() + 1e-30)).item() for k in grads)
    print(f"[{name}] oracle-vs-reference rel-L2: out {err_y:.2e}, worst grad {err_g:.2e}")
    assert err_y < 1e-5 and err_g < 1e-4, (err_y, err_g)
    g = torch.Generator().manual_seed(99)
    stats = {}
    for k, v in grads.items():
        r = torch.randn(v.shape, generator=g, dtype=torch.float64)
        v64 = v.double()
        stats[k] = [float(v64.sum()), float(v64.norm()), float((v64 * r).sum())]
    keep_full = [k for k in grads if ("relative_position_bias_table" in k or "norm" in k or k.endswith(".bias"))]
    fixture = dict(
        name=name, kwargs=kw, x_shape=list(xshape), sd_seed=1234, ln_jitter=0.1, x_seed=7, R_seed=11,
        R_scale=1.0 / 64.0,
        x_sha=sha16(x), sd_sha=sha16(torch.cat([v.flatten().double() for v in sd.values()])),
        y=y.detach().contiguous().clone(), grad_stats=stats,
        grad_full={k: grads[k] for k in keep_full}, state_keys=list(ref.state_dict().keys()),
        torch_version=torch.__version__)
    torch.save(fixture, os.path.join(HERE, f"{name}.pt"))
    return err_y, err_g


def keys_golden(vs):
    """state_dict keys / shapes / dtypes of the three real configurations (no tensors stored)."""
    out = {}
    for name, kw in (("swin_b", dict(embed_dim=128, num_heads=[4, 8, 16, 32])),
                     ("violet", dict(embed_dim=96, num_heads=[3, 6, 12, 24])),
                     ("swin_l_384", dict(embed_dim=192, num_heads=[6, 12, 24, 48], window_size=(8, 12, 12)))):
        m = vs.SwinTransformer3D(pretrained=None, depths=[2, 2, 18, 2], **kw)
        sd = m.state_dict()
        out[name] = dict(n_entries=len(sd), n_params=sum(p.numel() for p in m.parameters()),
                         n_param_tensors=len(list(m.parameters())),
                         entries=[[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
                         drop_path=[(b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0)
                                    for l in m.layers for b in l.blocks])
    return out


def main():
    vs = import_reference()
    torch.set_num_threads(os.cpu_count())
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index_golden(vs), f, indent=1)
    with open(os.path.join(HERE, "state_keys.json"), "w") as f:
        json.dump(keys_golden(vs), f)
    for name in TINY:
        model_golden(vs, name)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
