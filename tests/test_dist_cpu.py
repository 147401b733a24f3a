"""CPU, world_size 2, gloo: the data-parallel contract of the hot path -- clips shard by rank, weights are
replicated, and after the gradient all-reduce every rank holds the gradients of the single-process run on the
concatenated batch (the loss is a mean over clips).  The compute here is the CPU oracle; the GPU product path
obeys the same contract through DDP in bench.py."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import swin3d_oracle as O
        dp = importlib.import_module("pytorch_empirical-mvm_b200.dp")
        torch.manual_seed(0)  # identical weights on every rank
        cfg = O.SwinCfg(embed_dim=32, depths=(2,), num_heads=(1,), window_size=(2, 7, 7))
        sd = O.make_state_dict(cfg, seed=5, ln_jitter=0.1)
        params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items() if v.is_floating_point()}
        full = dict(sd)
        full.update(params)
        n_clips = 5  # deliberately not a multiple of the world size
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n_clips, 3, 2, 28, 28, generator=g)
        R = torch.randn(n_clips, 32, 2, 7, 7, generator=g)
        sl = dp.clip_shard(n_clips, rank, world)
        # per-rank loss = sum over local clips / n_clips * world, so that the AVERAGE over ranks is the global mean loss
        y = O.swin_forward(full, x[sl], cfg)
        loss = (y * R[sl]).sum() / n_clips * world
        loss.backward()
        local = {k: p.grad.clone() for k, p in params.items()}
        ncoll = dp.all_reduce_gradients(params.values(), bucket_bytes=64 << 10)
        assert ncoll >= 2  # several buckets at this bucket size
        # the copy-free, in-place form of the same exchange step (what bench.py uses) must give the same averages
        twins = {k: torch.nn.Parameter(torch.zeros_like(v)) for k, v in params.items()}
        for k, t in twins.items():
            t.grad = local[k].clone()
        assert dp.all_reduce_gradients_coalesced(twins.values()) == len(twins)
        for k in twins:
            assert torch.allclose(twins[k].grad, params[k].grad, rtol=1e-6, atol=1e-7), k
        if rank == 0:
            ref = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items() if v.is_floating_point()}
            fr = dict(sd)
            fr.update(ref)
            ((O.swin_forward(fr, x, cfg) * R).sum() / n_clips).backward()
            worst = max(float((params[k].grad - ref[k].grad).norm() / (ref[k].grad.norm() + 1e-30)) for k in ref)
            assert worst < 1e-5, worst
        # every rank ends with identical gradients
        flat = torch.cat([p.grad.reshape(-1) for p in params.values()])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        assert torch.equal(other[0], other[1])
        with open(os.path.join(tmpdir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def _worker_overlap(rank, world, port, tmpdir):
    """OverlappedGradReducer: hooks launch one reduction per block during backward; result == full-batch gradients"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dp = importlib.import_module("pytorch_empirical-mvm_b200.dp")

        class Block(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a, self.b = torch.nn.Linear(8, 8), torch.nn.Linear(8, 8)

            def forward(self, x):
                return x + self.b(torch.tanh(self.a(x)))

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.embed = torch.nn.Linear(4, 8)
                self.layers = torch.nn.ModuleList([torch.nn.ModuleList([Block(), Block()]), torch.nn.ModuleList([Block()])])
                self.unused = torch.nn.Linear(3, 3)   # never receives a gradient: finish() must cope
                self.head = torch.nn.Linear(8, 2)

            def forward(self, x):
                x = self.embed(x)
                for stage in self.layers:
                    for blk in stage:
                        x = blk(x)
                return self.head(x)

        torch.manual_seed(0)
        net, ref = Net().double(), Net().double()
        ref.load_state_dict(net.state_dict())
        red = dp.OverlappedGradReducer(net)
        assert len(red.groups) == 6   # embed, 3 blocks, unused, head
        g = torch.Generator().manual_seed(3)
        x, t = torch.randn(6, 4, generator=g).double(), torch.randn(6, 2, generator=g).double()
        sl = dp.clip_shard(6, rank, world)
        for _ in range(2):   # two steps: the reducer re-arms itself
            net.zero_grad(set_to_none=True)
            ((net(x[sl]) - t[sl]) ** 2).sum().mul(world / 6).backward()
            red.finish()
        ref.zero_grad(set_to_none=True)
        ((ref(x) - t) ** 2).sum().div(6).backward()
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            if q.grad is None:
                assert p.grad is None, n
            else:
                assert torch.allclose(p.grad, q.grad, rtol=1e-9, atol=1e-12), n
        with open(os.path.join(tmpdir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_overlapped_grad_reducer_gloo(tmp_path):
    world, port = 2, 29000 + os.getpid() % 1000 + 1
    mp.spawn(_worker_overlap, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def _worker_flat(rank, world, port, tmpdir):
    """FlatGradReducer: gradients live in one flat buffer, a few fixed contiguous ranges are all-reduced as they fill;
    a parameter that gets a gradient on ONE rank only must not change the collective's shape (zeros travel instead)"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dp = importlib.import_module("pytorch_empirical-mvm_b200.dp")

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.embed = torch.nn.Linear(4, 8)
                self.blocks = torch.nn.ModuleList([torch.nn.Linear(8, 8) for _ in range(5)])
                self.unused = torch.nn.Linear(3, 3)      # never receives a gradient
                self.sometimes = torch.nn.Linear(8, 1)   # receives a gradient on rank 0 only
                self.head = torch.nn.Linear(8, 2)

            def forward(self, x, extra):
                x = self.embed(x)
                for b in self.blocks:
                    x = x + torch.tanh(b(x))
                y = self.head(x)
                return y + self.sometimes(x).sum() * 0.01 if extra else y

        torch.manual_seed(0)
        net, ref = Net().double(), Net().double()
        ref.load_state_dict(net.state_dict())
        red = dp.FlatGradReducer(net, n_chunks=3, install_sink=False, unused="none")
        assert len(red.chunks) == 3 and red.chunks[0][1] == 0 and red.chunks[-1][2] == red.flat[torch.float64].numel()
        g = torch.Generator().manual_seed(3)
        x, t = torch.randn(6, 4, generator=g).double(), torch.randn(6, 2, generator=g).double()
        sl = dp.clip_shard(6, rank, world)
        for _ in range(2):   # two steps: zero_grad() re-arms the reducer
            red.zero_grad()
            ((net(x[sl], rank == 0) - t[sl]) ** 2).sum().mul(world / 6).backward()
            assert red.finish() == 3
        # reference: the same two shards on one process, gradients averaged by hand
        tot = {}
        for r in range(world):
            ref.zero_grad(set_to_none=True)
            s2 = dp.clip_shard(6, r, world)
            ((ref(x[s2], r == 0) - t[s2]) ** 2).sum().mul(world / 6).backward()
            for n, q in ref.named_parameters():
                if q.grad is not None:
                    tot[n] = tot.get(n, 0) + q.grad / world
        flat = red.flat[torch.float64]
        for n, p in net.named_parameters():
            if n.startswith("unused"):
                assert p.grad is None, n
                continue
            # (`sometimes` has no local gradient on rank 1: finish() hands it the averaged slot, so both replicas step alike)
            assert p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr(), n   # a view of the flat buffer
            assert torch.allclose(p.grad, tot[n], rtol=1e-9, atol=1e-12), n
        red.remove()
        with open(os.path.join(tmpdir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_flat_grad_reducer_gloo(tmp_path):
    world, port = 2, 29000 + os.getpid() % 1000 + 2
    mp.spawn(_worker_flat, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_clip_shard_covers_batch():
    dp = importlib.import_module("pytorch_empirical-mvm_b200.dp")
    for n in (1, 5, 32, 33):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s = dp.clip_shard(n, r, world)
                seen += list(range(n))[s]
            assert seen == list(range(n))


@pytest.mark.timeout(180)
def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_flat_grad_reducer_layout_and_single_process():
    """no process group: slots are aligned, chunks tile every flat buffer exactly, mixed dtypes get one buffer each, gradients
    that arrive the ordinary way are copied into their slots and equal plain autograd's"""
    dp = importlib.import_module("pytorch_empirical-mvm_b200.dp")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3).double(), torch.nn.Linear(3, 1).double())
    ref = [p.detach().clone().requires_grad_(True) for p in net.parameters()]
    red = dp.FlatGradReducer(net, n_chunks=8, install_sink=False)     # more chunks than parameters per dtype: clamped
    assert set(red.flat) == {torch.float32, torch.float64}
    for dt, flat in red.flat.items():
        ch = sorted((a, b) for d, a, b in red.chunks if d == dt)
        assert ch[0][0] == 0 and ch[-1][1] == flat.numel() and all(x[1] == y[0] for x, y in zip(ch, ch[1:]))
    for p in red.params:
        _, off, n = red.slot[id(p)]
        assert off % dp.FlatGradReducer.ALIGN == 0 and n == p.numel()
    x = torch.randn(4, 5)

    def loss(ps):
        h = torch.tanh(x @ ps[0].t() + ps[1]).double()
        return ((h @ ps[2].t() + ps[3]) @ ps[4].t() + ps[5]).pow(2).sum()
    for _ in range(2):
        red.zero_grad()
        loss(list(net.parameters())).backward()
        assert red.finish() == 0
    loss(ref).backward()
    for p, q in zip(net.parameters(), ref):
        assert torch.allclose(p.grad, q.grad, rtol=1e-12, atol=0)
        assert p.grad.untyped_storage().data_ptr() == red.flat[p.dtype].untyped_storage().data_ptr()
    red.remove()
