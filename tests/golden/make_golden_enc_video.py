"""Golden vectors for the EncVideo tail (reference model.py:7-78), from the UNMODIFIED reference class source.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_enc_video.py

``model.py`` cannot be imported (``from utils.lib import *`` needs easydict / skimage / fairscale / toolz, none
installed), so the ``class EncVideo`` statement is cut out of the file with ``ast`` and executed
as-is in a namespace that provides what its body uses: ``T = torch`` and ``get_vidswin_model`` (a stand-in backbone that
returns a given feature tensor in the layout ``SwinTransformer3D.forward`` returns, video_swin.py:478-482).  The body
calls ``.cuda()`` on the all-ones mask (model.py:71); on this GPU-less container ``torch.Tensor.cuda`` is replaced by the
identity for the duration of the call.  Nothing of the class source is stored in the repo -- only tensors.
"""
import ast
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MODEL = "/root/reference/model.py"


def load_reference_encvideo():
    src = open(REF_MODEL).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "EncVideo")
    code = ast.get_source_segment(src, node)

    class _Backbone(torch.nn.Module):
        def __init__(self, latent):
            super().__init__()
            self.norm = torch.nn.LayerNorm(latent)
            self.feat = None

        def forward(self, x):            # x: (B,3,T,H,W) -> (B,latent,T,h,w) permuted view, as the Swin returns
            return self.feat

    ns = {"T": torch, "get_vidswin_model": lambda args: _Backbone(args.latent)}
    exec(compile(code, REF_MODEL, "exec"), ns)
    return ns["EncVideo"]


CASES = {
    # name: (B, T, h, w, latent, hidden, max_frame, max_patch, odr, vt_mask?)
    "plain": (2, 3, 2, 2, 16, 24, 6, 14, None, False),
    "odr": (3, 4, 2, 3, 32, 24, 6, 14, [[0, 1, 2, 3], [1, 0, 2, 3], [3, 2, 1, 0]], False),
    "nofc_vtmask": (2, 2, 1, 2, 24, 24, 2, 2, [[1, 0], [0, 1]], True),
}


def main():
    EncVideo = load_reference_encvideo()
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    try:
        for name, (B, Tn, h, w, latent, hidden, mf, mp, odr, use_vt) in CASES.items():
            torch.manual_seed(len(name))
            args = types.SimpleNamespace(latent=latent, max_size_frame=mf, max_size_patch=mp)
            m = EncVideo(args, hidden)
            with torch.no_grad():   # make every parameter matter (norm defaults are 1/0)
                m.norm.weight.uniform_(0.5, 1.5)
                m.norm.bias.uniform_(-0.5, 0.5)
            buf = torch.randn(B, Tn, h, w, latent, requires_grad=True)          # channels-last buffer
            m.swin.feat = buf.permute(0, 4, 1, 2, 3)
            img = torch.zeros(B, Tn, 3, 32 * h, 32 * w)
            vt = None
            if use_vt:
                vt = (torch.rand(B, Tn, 1 + h * w) > 0.3).long()
            f_img, m_img = m(img, odr=odr, vt_mask=vt)
            R = torch.randn_like(f_img)
            (f_img * R).sum().backward()
            params = {k: v.detach().clone() for k, v in m.state_dict().items() if not k.startswith("swin.")}
            grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()
                     if not k.startswith("swin.") and p.grad is not None}
            out[name] = dict(buf=buf.detach().clone(), odr=odr, vt_mask=vt, R=R, params=params,
                             f_img=f_img.detach().clone(), m_img=m_img.detach().clone(),
                             dbuf=buf.grad.detach().clone(), grads=grads)
            print(name, tuple(f_img.shape), tuple(m_img.shape), sorted(grads))
    finally:
        torch.Tensor.cuda = orig_cuda
    torch.save(out, os.path.join(HERE, "enc_video.pt"))


if __name__ == "__main__":
    main()
