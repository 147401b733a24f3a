"""CPU: host-side logic of the drop-in modules (no kernels run)."""
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def keys_gold():
    with open(os.path.join(GOLD, "state_keys.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name,kw", [
    ("swin_b", dict(embed_dim=128, num_heads=[4, 8, 16, 32])),
    ("violet", dict(embed_dim=96, num_heads=[3, 6, 12, 24])),
    ("swin_l_384", dict(embed_dim=192, num_heads=[6, 12, 24, 48], window_size=(8, 12, 12))),
])
def test_state_dict_surface_matches_reference(vsw, keys_gold, name, kw):
    """351 entries, same key order, shapes and dtypes as the reference module."""
    m = vsw.SwinTransformer3D(pretrained=None, depths=[2, 2, 18, 2], **kw)
    sd = m.state_dict()
    g = keys_gold[name]
    assert len(sd) == g["n_entries"] == 351
    assert sum(p.numel() for p in m.parameters()) == g["n_params"]
    assert len(list(m.parameters())) == g["n_param_tensors"] == 327
    assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == g["entries"]
    dp = [(b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0) for l in m.layers for b in l.blocks]
    assert dp == pytest.approx(g["drop_path"])
    assert m.norm.normalized_shape[0] == m.num_features  # model.py:13 reads this


def test_relative_position_index_closed_form(vsw, oracle):
    import numpy as np
    for w in [(8, 7, 7), (8, 12, 12), (2, 3, 3), (16, 7, 7)]:
        a = vsw.WindowAttention3D(32, w, 1).relative_position_index
        assert a.dtype == torch.int64
        assert np.array_equal(a.numpy(), oracle.relative_position_index(w))


def test_get_window_size(vsw):
    with open(os.path.join(GOLD, "index.json")) as f:
        cases = json.load(f)["get_window_size"]
    for c in cases:
        ws, ss = vsw.get_window_size(tuple(c["grid"]), tuple(c["window"]), tuple(c["shift"]))
        assert list(ws) == c["ws"] and list(ss) == c["ss"]
    assert vsw.get_window_size((8, 4, 4), (8, 7, 7)) == (8, 4, 4)


def test_load_state_dict_roundtrip_with_oracle_layout(vsw, oracle):
    cfg = oracle.SwinCfg(embed_dim=32, depths=(2, 2), num_heads=(1, 2))
    sd = oracle.make_state_dict(cfg, seed=3)
    m = vsw.SwinTransformer3D(embed_dim=32, depths=[2, 2], num_heads=[1, 2])
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_config_loader_presets_and_inheritance(vsw, tmp_path):
    from importlib import import_module
    cl = import_module("pytorch_empirical-mvm_b200.config_loader")
    bb = cl.load_backbone_cfg("videoswin/swin_violet_patch244_window877.py")  # reference's broken path
    assert bb["embed_dim"] == 96 and bb["num_heads"] == [3, 6, 12, 24] and tuple(bb["patch_size"]) == (2, 4, 4)
    base = tmp_path / "base.py"
    base.write_text("model = dict(backbone=dict(embed_dim=96, depths=[2,2,6,2], num_heads=[3,6,12,24], "
                    "patch_size=(4,4,4), window_size=(8,7,7), patch_norm=True))\n")
    child = tmp_path / "child.py"
    child.write_text("_base_ = ['./base.py']\nmodel = dict(backbone=dict(patch_size=(2,4,4), embed_dim=128))\n")
    bb = cl.load_backbone_cfg(str(child))
    assert bb["embed_dim"] == 128 and tuple(bb["patch_size"]) == (2, 4, 4) and bb["depths"] == [2, 2, 6, 2]
    with pytest.raises(FileNotFoundError):
        cl.load_backbone_cfg("nope/does_not_exist.py")


def test_factory_side_effects(vsw):
    from types import SimpleNamespace as NS
    a = NS(size_img=224, vis_backbone_size="violet", vis_backbone_init="2d", kinetics=400)
    m = vsw.get_vidswin_model(a)
    assert a.vis_backbone_init == "random" and a.vis_backbone_pretrained_weight is None  # video_swin.py:603,615
    assert m.embed_dim == 96 and len(m.state_dict()) == 351


def test_cpu_input_fails_loudly(vsw):
    m = vsw.SwinTransformer3D(embed_dim=32, depths=[2], num_heads=[1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 2, 8, 8))


def test_unsupported_options_fail_loudly(vsw):
    with pytest.raises(NotImplementedError):
        vsw.SwinTransformer3D(embed_dim=32, depths=[2], num_heads=[1], drop_rate=0.1)
    with pytest.raises(NotImplementedError):
        vsw.SwinTransformer3D(embed_dim=32, depths=[2], num_heads=[1], attn_drop_rate=0.1)


def test_init_weights_statistics(vsw):
    torch.manual_seed(0)
    m = vsw.SwinTransformer3D(embed_dim=32, depths=[2, 2], num_heads=[1, 2])
    m.init_weights()
    w = m.layers[1].blocks[0].mlp.fc1.weight
    assert float(w.abs().max()) <= 2.0 and 0.015 < float(w.std()) < 0.025
    assert float(m.layers[0].blocks[0].attn.qkv.bias.abs().max()) == 0.0
    assert torch.all(m.norm.weight == 1) and torch.all(m.norm.bias == 0)
    with pytest.raises(TypeError):
        m.pretrained = 3
        m.init_weights()


def test_window_dims_hint_packing(vsw):
    """layout hint of the attention ABI: effective depth | configured height << 8 | configured width << 16"""
    VF = vsw.functional
    assert VF.window_dims() == 0
    assert VF.window_dims(8) == 8
    assert VF.window_dims(8, (8, 7, 7)) == 8 | (7 << 8) | (7 << 16)
    assert VF.window_dims(0, (8, 12, 12)) == (12 << 8) | (12 << 16)
    # every attention module hands its configured window to the kernels
    att = vsw.WindowAttention3D(64, (4, 6, 6), 2)
    assert tuple(att.window_size) == (4, 6, 6)


def test_bias_codes_are_dense_toeplitz(vsw, oracle):
    """the codes the kernels validate in-kernel before padding their on-chip table: rowcode[n] - rowcode[0] ==
    colcode[0] - colcode[n] == d R1 + h R2 + w and rowcode[0] + colcode[0] == (L - 1) / 2 (centre of the table)"""
    VF = vsw.functional
    for wd, wh, ww in [(8, 7, 7), (4, 6, 6), (2, 3, 5)]:
        idx = torch.from_numpy(oracle.relative_position_index((wd, wh, ww)))
        N = wd * wh * ww
        rc, cc = VF.bias_codes(idx, N)
        R2, R1 = 2 * ww - 1, (2 * wh - 1) * (2 * ww - 1)
        L = (2 * wd - 1) * R1
        n = torch.arange(N)
        e = (n // (wh * ww)) * R1 + ((n // ww) % wh) * R2 + n % ww
        assert torch.equal(rc.cpu().long() - int(rc[0]), e) and torch.equal(int(cc[0]) - cc.cpu().long(), e)
        assert int(rc[0]) + int(cc[0]) == (L - 1) // 2


@pytest.mark.parametrize("case", ["match", "bicubic"])
def test_inflate_weights_vs_reference_golden(vsw, tmp_path, case):
    """2-D -> 3-D checkpoint inflation (video_swin.py:484-535): same tensors as the unmodified reference produced
    (tests/golden/make_golden_inflate.py) -- conv weight repeated over time / pd, bias tables tiled 2wd-1 times, bicubic
    resize when the 2-D window differs, relative_position_index / attn_mask entries dropped."""
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "inflate.pt"), weights_only=False)
    c = gold["cases"][case]
    path = str(tmp_path / "swin2d.pth")
    torch.save(c["ckpt"], path)
    m = vsw.SwinTransformer3D(pretrained=path, pretrained2d=True, window_size=tuple(c["window_size"]), **gold["kw"])
    m.init_weights()
    sd = m.state_dict()
    assert len(c["result"]) >= 9
    for k, ref in c["result"].items():
        assert sd[k].shape == ref.shape and torch.equal(sd[k], ref), k      # same torch ops on the CPU: bit-exact
    wd = c["window_size"][0]
    assert sd["layers.0.blocks.0.attn.relative_position_bias_table"].shape[0] == (2 * wd - 1) * 13 * 13
    assert sd["layers.0.blocks.0.attn.relative_position_index"].dtype == torch.int64     # buffer re-initialised, not loaded


def test_enc_video_state_dict_matches_reference(vsw):
    """EncVideo's own parameters (everything outside `swin.*`): same names, shapes and dtypes as the reference class
    produced (tests/golden/enc_video.pt, case "odr": latent 32 -> hidden 24, max_size_frame 6, max_size_patch 14), so
    `enc_img.*` VIOLET checkpoints load; constructing the module needs no GPU."""
    import types
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "enc_video.pt"), weights_only=False)["odr"]["params"]
    swin = vsw.SwinTransformer3D(embed_dim=4, depths=[1, 1, 1, 1], num_heads=[1, 1, 1, 1])      # latent 8 * 4 = 32
    m = vsw.EncVideo(types.SimpleNamespace(), 24, swin=swin)
    own = {k: v for k, v in m.state_dict().items() if not k.startswith("swin.")}
    assert list(own) == list(gold)                                       # same names in the same order
    for k in gold:
        assert own[k].shape == gold[k].shape and own[k].dtype == gold[k].dtype, k
    assert m.load_state_dict({**m.state_dict(), **gold}, strict=True)
    assert (m.latent_feat_size, m.img_feature_dim, m.max_size_frame, m.max_size_patch) == (32, 24, 6, 14)
    with pytest.raises(NotImplementedError):
        vsw.EncVideo(types.SimpleNamespace(swinbert=True), 24, swin=swin)
    with pytest.raises(vsw._lib.VswError):                               # CPU tensors: no fallback
        m(torch.zeros(1, 2, 3, 32, 32))


def test_block_sampler_edge_cases(vsw):
    """main_pretrain.py:312-318: a single frame always yields depth-1 blocks; grids too small for the reference's
    `randint(1, h*2//3)` fail the same way (numpy ValueError) instead of inventing a block size; a private RandomState
    reproduces the global-stream result."""
    import numpy as np
    from oracle import mvm_oracle as MO
    np.random.seed(11)
    a = vsw.mvm.sample_block_masks(3, 1, 7, 7)
    np.random.seed(11)
    b = np.stack([MO.cover_grid(MO.block_cells(1, 7, 7), 1, 7, 7).numpy() for _ in range(3)]).astype(np.uint8)
    assert a.shape == (3, 1, 7, 7) and np.array_equal(a, b) and 0 < a.sum() <= 3 * 3 * 3   # one block of at most 3x3 per sample
    with pytest.raises(ValueError):
        vsw.mvm.sample_block_masks(1, 4, 1, 7)          # h*2//3 == 0 -> randint(1, 0)
    np.random.seed(5)
    g = vsw.mvm.sample_block_masks(2, 8, 7, 7)
    assert np.array_equal(g, vsw.mvm.sample_block_masks(2, 8, 7, 7, rng=np.random.RandomState(5)))
    assert g.max() == 1 and g.dtype == np.uint8


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line on stdout with the contract's keys; in the build
    container it times the unmodified reference (kind "reference"), elsewhere the oracle port (kind "port")."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--model", "violet", "--steps", "1",
                        "--warmup", "1", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "clips/s"
    assert d["cpu_baseline"]["kind"] == ("reference" if os.path.isdir("/root/reference/visbackbone") else "port")
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
