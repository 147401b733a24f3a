"""CPU: the C-ABI library builds/loads and exports every symbol include/vsw.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vsw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|size_t|long long)\s+(vsw_\w+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("vsw_window_maps", "vsw_ln_fwd", "vsw_ln_bwd", "vsw_linear_fwd", "vsw_linear_dgrad",
                 "vsw_linear_wgrad", "vsw_window_attn_fwd", "vsw_window_attn_bwd", "vsw_patch_im2col",
                 "vsw_merge_ln_fwd", "vsw_merge_ln_bwd", "vsw_version", "vsw_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(vsw):
    lib = ctypes.CDLL(vsw._lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"libvsw_b200.so does not export {s}"


def test_binding_covers_every_declared_symbol(vsw):
    assert sorted(vsw._lib.SIGNATURES) == declared_symbols()
    assert vsw._lib.lib().vsw_version() >= 100


def test_error_path_without_gpu(vsw):
    """argument validation happens before any CUDA call, so it is testable on the CPU box"""
    L = vsw._lib
    rc = L.lib().vsw_window_maps(8, 14, 14, 8, 7, 7, 0, 9, 3, None, None, None)  # shift >= window
    assert rc == -1
    assert "bad geometry" in L.last_error()
    rc = L.lib().vsw_linear_fwd(None, None, None, None, 4, 4, 4, 0, None, None, None, None, 0, 0, 0, None)
    assert rc == -1 and "vsw_linear_fwd" in L.last_error()
    # EncVideo tail: hidden size / embedding-table sizes are validated before any launch
    one = 1   # any non-NULL pointer value: validation fails before it is dereferenced
    tail = lambda C, pos_rows, len_rows: L.lib().vsw_enc_video_tail_fwd(
        one, one, one, one, one, None, one, one, None, one, None, None, None, 2, 3, 4, C, pos_rows, len_rows, 1e-5, 0, 0, None)
    assert tail(770, 197, 6) == -3 and "multiple of 4" in L.last_error()          # VSW_ERR_UNSUPPORTED
    assert tail(2048, 197, 6) == -3
    assert tail(768, 4, 6) == -1 and "emb_pos" in L.last_error()                   # needs 1 + h*w = 5 rows
    assert tail(768, 197, 2) == -1 and "emb_len" in L.last_error()                 # 3 frames, table holds 2
    assert L.lib().vsw_enc_video_tail_bwd_workspace(2, 3, 4, 768) == (2 * 3 * 5 * 768 + 2 * 3 * 768 + 296 * 2 * 768) * 4


def test_no_oracle_on_product_path():
    """the product package must never import the oracle (it is test infrastructure)"""
    pkg = os.path.join(ROOT, "pytorch_empirical-mvm_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "oracle" not in txt.replace("no oracle", ""), f
