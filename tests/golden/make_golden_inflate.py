"""Golden vectors for SwinTransformer3D.inflate_weights (video_swin.py:484-535) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_inflate.py

A synthetic 2-D Swin checkpoint (``{'model': partial state_dict}`` with 2-D conv weight, a few ordinary tensors, (2w-1)^2-row bias tables,
``relative_position_index`` / ``attn_mask`` entries that must be dropped) is inflated by the reference for two target
windows: one whose tables match (7x7) and one that needs the bicubic resize (2-D window 5 -> 3-D window (4,7,7)).
Stored: the synthetic checkpoint and the resulting 3-D state_dict entries the inflation touches.
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

KW = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], patch_size=(2, 4, 4), drop_path_rate=0.0)
KEEP = ("patch_embed.proj.bias", "layers.0.blocks.0.attn.qkv.weight", "layers.1.blocks.1.mlp.fc2.bias", "norm.weight")
CASES = {"match": dict(win2d=7, window_size=(2, 7, 7)), "bicubic": dict(win2d=5, window_size=(4, 7, 7))}


def synthetic_2d_checkpoint(model3d, win2d, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model3d.state_dict().items():
        if k == "patch_embed.proj.weight":
            sd[k] = torch.randn(v.shape[0], v.shape[1], v.shape[3], v.shape[4], generator=g)
        elif "relative_position_bias_table" in k:
            sd[k] = torch.randn((2 * win2d - 1) ** 2, v.shape[1], generator=g)
        elif "relative_position_index" in k:
            sd[k] = torch.zeros(win2d * win2d, win2d * win2d, dtype=torch.long)
        elif k in KEEP:
            sd[k] = torch.randn(v.shape, generator=g)
    sd["layers.0.blocks.1.attn_mask"] = torch.zeros(4, win2d * win2d, win2d * win2d)
    return {"model": sd}


def main():
    vs = import_reference()
    out = {"kw": KW, "cases": {}}
    for i, (name, c) in enumerate(CASES.items()):
        probe = vs.SwinTransformer3D(pretrained=None, window_size=c["window_size"], **KW)
        ckpt = synthetic_2d_checkpoint(probe, c["win2d"], seed=100 + i)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "swin2d.pth")
            torch.save(ckpt, path)
            m = vs.SwinTransformer3D(pretrained=path, pretrained2d=True, window_size=c["window_size"], **KW)
            m.init_weights()
        sd3 = {k: v.clone() for k, v in m.state_dict().items() if k in ckpt["model"] and "index" not in k}
        out["cases"][name] = dict(window_size=c["window_size"], ckpt=ckpt, result=sd3)
        print(name, len(sd3), "entries")
    torch.save(out, os.path.join(HERE, "inflate.pt"))


if __name__ == "__main__":
    main()
