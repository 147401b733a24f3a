"""Golden vectors for the MVM masking and the 3d_feature loss (reference main_pretrain.py:276-372, 374-524), from the
UNMODIFIED reference method source.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_mvm.py

``main_pretrain.py`` cannot be imported (its star-import of utils/lib.py needs packages that are not installed), so the
``masking`` and ``calc_mvm_loss`` method definitions of ``Agent_Pretrain`` are cut out of the file with ``ast`` and
executed as plain functions against a stub ``self`` (``T = torch``, ``np``, ``random`` in the namespace).  Nothing of the
source is stored -- only seeds, input tensors and outputs.
"""
import ast
import os
import random
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/main_pretrain.py"


def load_methods(*names):
    src = open(REF).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "Agent_Pretrain")
    ns = {"T": torch, "np": np, "random": random}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(textwrap.dedent(ast.get_source_segment(src, node, padded=True)), REF, "exec"), ns)
    return [ns[n] for n in names]


def pattern_clip(B, Tn, H, W):
    """deterministic, RNG-free clip (so the fixture does not have to store it): values in [-0.5, 0.5), none exactly 0"""
    n = B * Tn * 3 * H * W
    return (((torch.arange(n, dtype=torch.int64) * 7919) % 1013).float() + 0.5).div(1013.0).sub(0.5).view(B, Tn, 3, H, W)


def main():
    masking, calc_mvm_loss = load_methods("masking", "calc_mvm_loss")
    out = {}

    # ---- masking, "bm" blocks (main_pretrain.py:309-321, 349-362) --------------------------------------------------
    B, Tn, H, W, X = 3, 4, 224, 224, 12
    seed = 5
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    img = pattern_clip(B, Tn, H, W)
    txt = torch.randint(1000, 2000, (B, X))
    txt[:, 0], txt[:, -1] = 101, 102
    stub = types.SimpleNamespace(patch_size=32, cls_token_id=101, sep_token_id=102, pad_token_id=0, mask_token_id=103,
                                 args=types.SimpleNamespace(pretrain_masks=["bm"], pretrain_tasks=["mtm", "mvm"]))
    np.random.seed(seed + 1)         # the block sampler is the only consumer of the numpy stream
    res = masking(stub, img.clone(), txt.clone(), torch.ones(B, X).long(), None, p_mask=0.15)
    out["masking"] = dict(np_seed=seed + 1, shape=(B, Tn, H, W),
                          cover=torch.nn.functional.max_pool2d(res["mvm_mask"].view(B * Tn, 3, H, W), 32)
                          .view(B, Tn, 3, 7, 7).to(torch.uint8),
                          check_rows=res["img"][:, :, :, ::37, :].clone(),          # a strided slice of the masked clip
                          mask_rows=res["mvm_mask"][:, :, :, ::37, :].to(torch.uint8),
                          unmask_equal=bool(torch.equal(res["unmask_img"], img)),
                          mask_sum=float(res["mvm_mask"].sum()), img_sum=float(res["img"].double().sum()))
    print("masking: cover mean", float(out["masking"]["cover"].float().mean()))

    # ---- 3d_feature loss (main_pretrain.py:508-524) ------------------------------------------------------------------
    B, Tn, h, w, Cf, Ch = 2, 3, 2, 2, 16, 24
    torch.manual_seed(9); np.random.seed(9)
    img = torch.randn(B, Tn, 3, 32 * h, 32 * w)
    cov = (torch.rand(B, Tn, h, w) > 0.5).float()
    mvm_mask = cov.repeat_interleave(32, -2).repeat_interleave(32, -1).unsqueeze(2).expand(-1, -1, 3, -1, -1).contiguous()
    teacher = torch.randn(B, Tn, h, w, Cf).permute(0, 4, 1, 2, 3)                   # as the Swin returns it
    fc_mvm = torch.nn.Linear(Ch, Cf)

    class _Teacher(torch.nn.Module):
        def forward(self, x):
            return teacher

    model = types.SimpleNamespace(fc_mvm=fc_mvm, feature_model=_Teacher())
    stub = types.SimpleNamespace(patch_size=32, model=model,
                                 args=types.SimpleNamespace(pretrain_tasks=["mvm"], mvm_target=["3d_feature"], deepspeed=True))
    out_mvm = torch.randn(B, Tn * (1 + h * w), Ch, requires_grad=True)
    ls = calc_mvm_loss(stub, {"unmask_img": img, "mvm_mask": mvm_mask}, out_mvm, is_train=True)
    loss = ls            # is_train: the sum over the enabled targets (here only 3d_feature), main_pretrain.py:546-550
    loss.backward()
    out["loss"] = dict(cov=cov.to(torch.uint8), teacher=teacher.permute(0, 2, 3, 4, 1).contiguous(), out_mvm=out_mvm.detach(),
                       fc_w=fc_mvm.weight.detach().clone(), fc_b=fc_mvm.bias.detach().clone(), loss=loss.detach(),
                       d_out_mvm=out_mvm.grad.clone(), d_fc_w=fc_mvm.weight.grad.clone(), d_fc_b=fc_mvm.bias.grad.clone())
    print("loss:", float(loss))
    torch.save(out, os.path.join(HERE, "mvm.pt"))


if __name__ == "__main__":
    main()
