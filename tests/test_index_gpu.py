"""GPU: index maps produced by the CUDA kernels are BIT-EXACT with the oracle and with the golden
vectors recorded from the reference (window/shift gather map, region-id mask, relative position
index, PatchMerging map)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha16(t):
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "index.json")) as f:
        return json.load(f)


def test_rel_pos_index_kernel(vsw, oracle, gold):
    VF = vsw.functional
    for key, g in gold["relative_position_index"].items():
        w = tuple(int(v) for v in key.split("x"))
        idx = VF.rel_pos_index(w, "cuda")
        assert sha16(idx) == g["sha"], key
        assert np.array_equal(idx.cpu().numpy(), oracle.relative_position_index(w))


def test_window_maps_and_mask_kernels(vsw, oracle, gold):
    VF = vsw.functional
    for key, g in gold["compute_mask"].items():
        pg, ws, ss = tuple(g["grid"]), tuple(g["window"]), tuple(g["shift"])
        plan = VF.window_plan(pg, ws, ss, "cuda")
        assert (plan.nW, plan.N) == tuple(g["shape"][:2])
        gm = plan.gather.view(plan.nW, plan.N).cpu().long()
        assert sha16(gm) == gold["gather_map"][key]["sha"], key
        m = vsw.compute_mask(pg[0], pg[1], pg[2], ws, ss, "cuda")
        assert m.dtype == torch.float32 and list(m.shape) == g["shape"]
        assert sha16(m) == g["sha"], key
        if any(ss):
            assert np.array_equal(plan.region.view(plan.nW, plan.N).cpu().numpy(),
                                  oracle.window_region_ids(pg, ws, ss).astype(np.uint8))


@pytest.mark.parametrize("grid,window,shift", [((4, 11, 13), (8, 7, 7), (4, 3, 3)), ((3, 5, 9), (2, 7, 7), (1, 3, 3)),
                                               ((16, 10, 10), (8, 7, 7), (4, 3, 3)), ((1, 1, 1), (8, 7, 7), (4, 3, 3))])
def test_window_maps_with_padding(vsw, oracle, grid, window, shift):
    """grid not a multiple of the window: slots that fall into the post-norm zero padding are -1"""
    VF = vsw.functional
    ws, ss = oracle.effective_window(grid, window, shift)
    pg = oracle.padded_grid(grid, ws)
    plan = VF.window_plan(grid, window, shift, "cuda")
    assert plan.ws == ws and plan.ss == ss and plan.pgrid == pg
    ref = oracle.window_gather_map(pg, ws, ss)  # flat over the PADDED grid
    d, h, w = ref // (pg[1] * pg[2]), (ref // pg[2]) % pg[1], ref % pg[2]
    inside = (d < grid[0]) & (h < grid[1]) & (w < grid[2])
    unp = np.where(inside, (d * grid[1] + h) * grid[2] + w, -1)
    got = plan.gather.view(plan.nW, plan.N).cpu().numpy()
    assert np.array_equal(got, unp)
    # every real token appears exactly once
    assert sorted(got[got >= 0].tolist()) == list(range(grid[0] * grid[1] * grid[2]))


@pytest.mark.parametrize("grid", [(8, 56, 56), (4, 19, 23), (2, 1, 1), (3, 7, 8)])
def test_merge_map_kernel(vsw, oracle, grid):
    got = vsw.functional.merge_map(grid, "cuda").view(-1, 4).cpu().numpy()
    assert np.array_equal(got, oracle.merge_gather_map(grid))


def test_window_partition_reverse_roundtrip(vsw):
    x = torch.randn(2, 8, 14, 14, 16, device="cuda")
    w = vsw.window_partition(x, (8, 7, 7))
    assert w.shape == (8, 392, 16)
    assert torch.equal(vsw.window_reverse(w, (8, 7, 7), 2, 8, 14, 14), x)
    ref = x.view(2, 1, 8, 2, 7, 2, 7, 16).permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, 392, 16)
    assert torch.equal(w, ref)
