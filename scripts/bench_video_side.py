#!/usr/bin/env python
"""Video side of one VIOLETv2 MVM `3d_feature` pretraining step (BASELINE.json config 3, the part this repo covers) on ONE
B200, every op through the C ABI:

    cov     = sample_block_masks(B, T, 7, 7)                       host, main_pretrain.py:309-321
    masked  = apply_block_mask(clips, cov)                         main_pretrain.py:355-362
    f_img   = EncVideo(masked)      student Swin-B (train, drop_path 0.2) + fc + class/pos/len embeddings + LayerNorm
    pred    = fc_mvm(non-class rows of f_img)                      (the BERT fusion encoder in between is NOT included)
    target  = teacher Swin-B(clips)  no_grad, eval                 main_pretrain.py:514-516
    loss    = masked L1(pred, target, cov) / 3 ; loss.backward()   main_pretrain.py:520-522

    python scripts/bench_video_side.py [--batch 32] [--steps 8] [--warmup 3]

Prints one JSON line (NOT the bench.py contract line: this is a widened-path measurement).  bf16 parameters and
activations, fp32 clips resident in HBM, CUDA events around the K timed steps, per-kernel-family CUDA-event split from one
extra profiled step.
"""
import argparse
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF, L = vsw.functional, vsw._lib

SWIN_B = dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=(8, 7, 7))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--hidden", type=int, default=768)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, Tn, side, hid = a.batch, 8, 224, a.hidden
    torch.manual_seed(0)
    student = vsw.SwinTransformer3D(pretrained=None, drop_path_rate=0.2, **SWIN_B)
    student.init_weights()
    teacher = vsw.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **SWIN_B)
    teacher.init_weights()
    enc = vsw.EncVideo(types.SimpleNamespace(max_size_frame=Tn, max_size_patch=14), hid, swin=student).to(dev).bfloat16().train()
    teacher = teacher.to(dev).bfloat16().eval()
    fc_mvm = torch.nn.Linear(hid, 1024).to(dev).bfloat16()
    params = list(enc.parameters()) + list(fc_mvm.parameters())
    clips = torch.randn(B, Tn, 3, side, side, device=dev)           # fp32 frames (B,T,3,H,W), as the data loader yields them
    np.random.seed(0)

    def step():
        for p in params:
            p.grad = None
        cov = vsw.mvm.sample_block_masks(B, Tn, 7, 7)
        cov_d = torch.from_numpy(cov).to(dev, non_blocking=True)
        masked, _ = vsw.mvm.apply_block_mask(clips, cov_d, 32, want_mask=False)
        f_img, _m = enc(masked)
        non_cls = f_img.view(B, Tn, 50, hid)[:, :, 1:].reshape(B * Tn * 49, hid)
        pred = VF.linear(non_cls, fc_mvm.weight, fc_mvm.bias).view(B, Tn, 49, 1024)
        with torch.no_grad():
            t_out = teacher(clips.transpose(1, 2))
        loss = vsw.mvm.mvm_3d_feature_loss(pred, t_out, cov_d, 3)
        loss.backward()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (L.launch_count() - n0) // a.steps
    VF.PROFILER = VF.KernelTimer()
    step()
    torch.cuda.synchronize()
    fam = {k: dict(launches=v["launches"], ms=round(v["ms"], 3)) for k, v in VF.PROFILER.summary().items()}
    VF.PROFILER = None
    print(json.dumps({
        "metric": "VIOLETv2 MVM 3d_feature step, video side (masking + student EncVideo fwd+bwd + teacher Swin-B fwd + loss) clips/sec",
        "value": B / ms * 1e3, "unit": "clips/s", "ms_per_step": ms, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"student+teacher Swin-B, {B} clips x 8x224^2, hidden {hid}, blockwise masks, drop_path 0.2",
                   "excluded": "BERT fusion encoder between EncVideo and fc_mvm (SURVEY 8f rank 2, not built)",
                   "l2": "inputs+activations per step >> 126 MB L2 (no explicit flush)"},
        "loss": float(loss), "vsw_launches_per_step": int(launches), "kernel_families_ms": fam,
        "mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))


if __name__ == "__main__":
    main()
