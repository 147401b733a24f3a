"""CPU: pin the oracle (oracle/swin3d_oracle.py) against the golden vectors that
tests/golden/make_golden.py produced from the UNMODIFIED reference module."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha16(arr) -> str:
    t = torch.as_tensor(arr).contiguous()
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def index_gold():
    with open(os.path.join(GOLD, "index.json")) as f:
        return json.load(f)


def test_survey_known_answers(index_gold):
    """the SURVEY section 8c table, which was produced independently of make_golden.py"""
    r = index_gold["relative_position_index"]["8x7x7"]
    assert (r["sum"], r["max"], r["c00"], r["c01"], r["c0m1"], r["cm10"], r["sha"]) == \
           (194692288, 2534, 1267, 1266, 0, 2534, "617532ad86365ae6")
    assert index_gold["relative_position_index"]["8x12x12"]["sha"] == "779ff0bed9d27d0a"
    assert index_gold["relative_position_index"]["16x7x7"]["sha"] == "8fcb73f0949d1851"
    m = index_gold["compute_mask"]
    assert m["g8x56x56_w8x7x7_s0x3x3"]["sha"] == "6dc7dfe987078f4b" and m["g8x56x56_w8x7x7_s0x3x3"]["nonzero"] == 1167360
    assert m["g8x28x28_w8x7x7_s0x3x3"]["sha"] == "dd58cbc13cfee521"
    assert m["g8x14x14_w8x7x7_s0x3x3"]["sha"] == "e8feda2b6042e292"
    assert m["g8x7x7_w8x7x7_s0x0x0"]["sha"] == "21369b918a617d97"
    assert m["g8x96x96_w8x12x12_s0x6x6"]["sha"] == "7a80a90c974ab8e8"
    assert m["g16x56x56_w8x7x7_s4x3x3"]["sha"] == "d35f8c6cb1adbf69"
    assert m["g4x56x56_w4x7x7_s0x3x3"]["sha"] == "27587be74dc59681"
    assert index_gold["gather_map"]["g8x14x14_w8x7x7_s0x3x3"]["sha"] == "4b59eb8b47ac5cc2"


def test_oracle_rel_pos_index(oracle, index_gold):
    for key, g in index_gold["relative_position_index"].items():
        w = tuple(int(v) for v in key.split("x"))
        idx = oracle.relative_position_index(w)
        assert list(idx.shape) == g["shape"]
        assert int(idx.sum()) == g["sum"] and int(idx.max()) == g["max"] and int(idx.min()) == g["min"]
        assert sha16(idx.astype(np.int64)) == g["sha"], key


def test_oracle_mask_and_gather(oracle, index_gold):
    for key, g in index_gold["compute_mask"].items():
        pg, ws, ss = tuple(g["grid"]), tuple(g["window"]), tuple(g["shift"])
        m = oracle.shift_mask(pg, ws, ss)
        assert list(m.shape) == g["shape"]
        assert int((m != 0).sum()) == g["nonzero"]
        assert sha16(m.float()) == g["sha"], key
        gm = oracle.window_gather_map(pg, ws, ss)
        gg = index_gold["gather_map"][key]
        assert gm[0, :5].tolist() == gg["first5"] and gm[-1, -3:].tolist() == gg["last3"]
        assert sha16(gm.astype(np.int64)) == gg["sha"], key
        # reverse o roll-back is the identity: the map is a permutation of the padded grid
        assert sorted(gm.reshape(-1).tolist()) == list(range(pg[0] * pg[1] * pg[2]))


def test_oracle_get_window_size(oracle, index_gold):
    for c in index_gold["get_window_size"]:
        ws, ss = oracle.effective_window(tuple(c["grid"]), tuple(c["window"]), tuple(c["shift"]))
        assert list(ws) == c["ws"] and list(ss) == c["ss"]


@pytest.mark.parametrize("name", ["tiny_a", "tiny_b", "tiny_c", "tiny_d"])
def test_oracle_tiny_models(oracle, name):
    """forward + every parameter gradient of small SwinTransformer3D configs vs the reference's own
    outputs (fp32 round-off level: both sides are fp32 CPU)."""
    fx = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    kw = fx["kwargs"]
    cfg = oracle.SwinCfg(embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]), num_heads=tuple(kw["num_heads"]),
                         window_size=tuple(kw["window_size"]))
    sd = oracle.make_state_dict(cfg, seed=fx["sd_seed"], ln_jitter=fx["ln_jitter"])
    assert list(sd.keys()) == fx["state_keys"]
    torch.manual_seed(fx["x_seed"])
    x = torch.randn(*fx["x_shape"])
    assert sha16(x) == fx["x_sha"], "torch CPU RNG changed: regenerate goldens"
    y_ref = fx["y"]
    torch.manual_seed(fx["R_seed"])
    R = torch.randn(*y_ref.shape) * fx["R_scale"]
    y, grads = oracle.forward_backward(sd, x, cfg, R)
    assert rel_l2(y, y_ref) < 2e-6
    g = torch.Generator().manual_seed(99)
    for k, (s, n, p) in fx["grad_stats"].items():
        r = torch.randn(grads[k].shape, generator=g, dtype=torch.float64)
        v = grads[k].double()
        assert abs(float(v.norm()) - n) <= 2e-5 * n + 1e-9, k
        assert abs(float((v * r).sum()) - p) <= 5e-5 * n * float(r.norm()) / max(1.0, v.numel() ** 0.5) + 1e-7, k
    for k, gref in fx["grad_full"].items():
        assert rel_l2(grads[k], gref) < 2e-5, k


# ---------------------------------------------------------------------------------------------
# EncVideo tail (model.py:32-78): oracle/enc_video_oracle.py vs the unmodified reference class
# (tests/golden/make_golden_enc_video.py)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def enc_gold():
    return torch.load(os.path.join(GOLD, "enc_video.pt"), weights_only=False)


@pytest.mark.parametrize("case", ["plain", "odr", "nofc_vtmask"])
def test_enc_video_oracle_vs_reference(enc_gold, case):
    from oracle import enc_video_oracle as EO
    g = enc_gold[case]
    buf = g["buf"].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    f_img, m_img = EO.enc_video_tail(buf.permute(0, 4, 1, 2, 3), params, odr=g["odr"], vt_mask=g["vt_mask"])
    assert f_img.shape == g["f_img"].shape and m_img.dtype == torch.int64
    assert torch.equal(m_img, g["m_img"])                       # integer domain: bit-exact
    assert rel_l2(f_img, g["f_img"]) < 2e-6
    (f_img * g["R"]).sum().backward()
    assert rel_l2(buf.grad, g["dbuf"]) < 2e-5
    for k, ref in g["grads"].items():
        assert rel_l2(params[k].grad, ref) < 2e-5, k
    for k in params:                                            # parameters the reference left without a gradient
        if k not in g["grads"]:
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k


def test_enc_video_oracle_rejects_too_many_frames():
    from oracle import enc_video_oracle as EO
    hid = 8
    p = {"emb_cls": torch.zeros(1, 1, 1, hid), "emb_pos": torch.zeros(1, 1, 5, hid), "emb_len": torch.zeros(1, 2, 1, hid),
         "emb_odr": torch.zeros(1, 1, 1, hid), "norm.weight": torch.ones(hid), "norm.bias": torch.zeros(hid)}
    with pytest.raises(ValueError):
        EO.enc_video_tail(torch.zeros(1, hid, 3, 1, 1), p)


# ---------------------------------------------------------------------------------------------
# MVM masking + 3d_feature loss (main_pretrain.py:276-372, 508-524): oracle/mvm_oracle.py and the product's host-side
# block sampler vs the unmodified reference methods (tests/golden/make_golden_mvm.py)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mvm_gold():
    return torch.load(os.path.join(GOLD, "mvm.pt"), weights_only=False)


def test_block_sampler_and_masking_vs_reference(mvm_gold, vsw):
    from conftest import pattern_clip
    from oracle import mvm_oracle as MO
    g = mvm_gold["masking"]
    B, Tn, H, W = g["shape"]
    h, w = H // 32, W // 32
    assert g["unmask_equal"] and torch.equal(g["cover"][:, :, 0], g["cover"][:, :, 1])
    cover_ref = g["cover"][:, :, 0]
    # oracle sampler (sets + loops) and product sampler (slice assignment): same numpy stream -> same blocks, bit-exact
    np.random.seed(g["np_seed"])
    cov_o = torch.stack([MO.cover_grid(MO.block_cells(Tn, h, w), Tn, h, w) for _ in range(B)])
    np.random.seed(g["np_seed"])
    cov_p = vsw.mvm.sample_block_masks(B, Tn, h, w)
    assert torch.equal(cov_o.to(torch.uint8), cover_ref)
    assert cov_p.dtype == np.uint8 and np.array_equal(cov_p, cover_ref.numpy())
    # pixel zeroing + full-resolution mask
    img = pattern_clip(B, Tn, H, W)
    masked, mask = MO.apply_block_mask(img, cov_o, 32)
    assert torch.equal(masked[:, :, :, ::37, :], g["check_rows"])
    assert torch.equal(mask[:, :, :, ::37, :].to(torch.uint8), g["mask_rows"])
    assert float(mask.sum()) == g["mask_sum"] and abs(float(masked.double().sum()) - g["img_sum"]) < 1e-6 * abs(g["img_sum"]) + 1e-3


def test_feature_loss_oracle_vs_reference(mvm_gold):
    from oracle import mvm_oracle as MO
    g = mvm_gold["loss"]
    B, Tn, h, w, Cf = g["teacher"].shape
    P = 1 + h * w
    out_mvm = g["out_mvm"].clone().requires_grad_(True)
    fw, fb = g["fc_w"].clone().requires_grad_(True), g["fc_b"].clone().requires_grad_(True)
    non_cls = out_mvm.view(B, Tn, P, -1)[:, :, 1:]                     # drop each frame's class row (main_pretrain.py:512)
    pred = torch.nn.functional.linear(non_cls, fw, fb)
    loss = MO.feature_loss(pred, g["teacher"].permute(0, 4, 1, 2, 3), g["cov"].float(), 3)
    assert abs(float(loss) - float(g["loss"])) < 1e-6 * abs(float(g["loss"]))
    loss.backward()
    assert rel_l2(out_mvm.grad, g["d_out_mvm"]) < 1e-6
    assert rel_l2(fw.grad, g["d_fc_w"]) < 1e-6 and rel_l2(fb.grad, g["d_fc_b"]) < 1e-6
