#!/usr/bin/env python
"""bench.py -- Video-Swin-B fwd+bwd clips/sec on N B200s (BASELINE.json metric), one JSON line.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference algorithm on the host CPU (oracle port)

A step = forward + backward (all 327 parameter gradients) of SwinTransformer3D (Swin-B widths,
embed_dim 128, heads 4/8/16/32, depths 2/2/18/2, drop_path_rate 0.2, train mode) over one batch of
synthetic 8x224^2 clips; bf16 parameters and activations, fp32 softmax/LN statistics.  Data parallel:
every rank owns `--batch` clips (weak scaling); gradients are all-reduced by DDP/NCCL inside backward.

  value : clips/s with the clips already resident in HBM.
  e2e   : same step called through the public module API with fp32 clips in PINNED HOST memory:
          H2D copy of the batch + D2H read of the scalar loss inside the timed region.
  roofline : the kernel family with the largest share of the timed step, timed live with CUDA events.
  cpu_baseline : the oracle port of the reference on the host cores, bounded sample.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "pytorch_empirical-mvm_b200"

MODELS = {
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=(8, 7, 7)),
    "violet": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=(8, 7, 7)),
    "swin_l_384": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48], window_size=(8, 12, 12)),
}
# kernel symbol behind each profiler family in the default configuration (roofline.traffic is only reported for a capture of it)
CURRENT_KERNEL = {"window_attn_bwd": "attn2_bwd_kernel", "window_attn_fwd": "attn2_fwd_kernel"}
FWD_BWD_GFLOP_PER_CLIP = {"swin_b": 844.0, "violet": 497.0, "swin_l_384": 6319.1}  # SURVEY 8d (3 x fwd)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def reference_available():
    return os.path.isdir("/root/reference/visbackbone")


def cpu_true_reference_run(model_name, batch, steps, warmup, threads):
    """the UNMODIFIED reference module (visbackbone/video_swin.py, imported where it lies) forward+backward on the host cores.
    Only possible where /root/reference exists (the build container); the GPU box has no copy of it."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    vs = import_reference()
    torch.set_num_threads(threads)
    kw = MODELS[model_name]
    torch.manual_seed(0)
    m = vs.SwinTransformer3D(pretrained=None, drop_path_rate=0.0, **kw)
    m.train()
    side = 384 if model_name == "swin_l_384" else 224
    x = torch.randn(batch, 3, 8, side, side)
    with torch.no_grad():
        y0 = m(x[:1])
    R = torch.randn(batch, *y0.shape[1:]) / 1024
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in m.parameters():
            p.grad = None
        (m(x) * R).sum().backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), sum(times) / len(times)


def cpu_reference_run(model_name, batch, steps, warmup, threads):
    """the reference algorithm (oracle port, fp32) forward+backward on the host cores -> clips/s"""
    from oracle import swin3d_oracle as O
    torch.set_num_threads(threads)
    kw = MODELS[model_name]
    cfg = O.SwinCfg(embed_dim=kw["embed_dim"], depths=tuple(kw["depths"]), num_heads=tuple(kw["num_heads"]),
                    window_size=tuple(kw["window_size"]))
    sd = O.make_state_dict(cfg, seed=0)
    side = 384 if model_name == "swin_l_384" else 224
    torch.manual_seed(0)
    x = torch.randn(batch, 3, 8, side, side)
    y0 = O.swin_forward(sd, x[:1], cfg)
    R = torch.randn(batch, *y0.shape[1:]) / 1024
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward_backward(sd, x, cfg, R)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), sum(times) / len(times)


def attn_sweep(args):
    """BASELINE config 5: WindowAttention3D core (QK^T + relative-position bias + shift mask + softmax + PV, forward and
    backward) over window 8x7x7 vs 8x12x12, shifted vs unshifted, heads 4-32, next to the reference algorithm on the same
    tensors on the host cores (video_swin.py:152-169 restated with torch ops; bounded sample of windows).  One JSON line."""
    vsw = importlib.import_module(PKG)
    VF, L = vsw.functional, vsw._lib
    from oracle import swin3d_oracle as O   # host-side checker only (mask + index of the CPU reference leg)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    hd, clips = 32, args.batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    w0 = torch.randn(4, 4, 392, 32, requires_grad=True)     # warm the host thread pool / allocator before the first timed case
    (torch.softmax(w0 @ w0.transpose(-1, -2), -1) @ w0).sum().backward()
    cases = []
    for window, grid in (((8, 7, 7), (8, 14, 14)), ((8, 12, 12), (8, 24, 24))):
        for shifted in (False, True):
            for nH in (4, 8, 16, 32):
                shift = tuple(w // 2 for w in window) if shifted else (0, 0, 0)
                plan = VF.window_plan(grid, window, shift, dev)
                nW, N = plan.nW, plan.N
                B_, C = clips * nW, nH * hd
                Lt = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
                g = torch.Generator().manual_seed(nH * 7 + N + shifted)
                qkv_h = torch.randn(B_ * N, 3 * C, generator=g)
                table_h = torch.randn(Lt, nH, generator=g) * 0.1
                dout_h = torch.randn(B_ * N, C, generator=g)
                qkv, table, dout = (t.to(dev).bfloat16() for t in (qkv_h, table_h, dout_h))
                rc, cc = VF.bias_codes(VF.rel_pos_index(window, dev), N)   # product path: no oracle
                region = plan.region if plan.shifted else None
                scale = hd ** -0.5

                def fwd():
                    return VF.attn_fwd(qkv, table, rc, cc, region, None, B_, nW, N, nH, hd, scale, window=window)
                out, lse = fwd()

                def bwd():
                    return VF.attn_bwd(qkv, out, dout, lse, table, rc, cc, region, None, B_, nW, N, nH, hd, scale,
                                       planes=plan.ws[0], window=window)
                ms = []
                for fn in (fwd, bwd):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.steps):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1) / args.steps)
                # ---- the same tensors through the reference algorithm on the host (first clip's windows only)
                ns = nW   # windows of one clip
                q3 = qkv_h[: ns * N].view(ns, N, 3, nH, hd).clone().requires_grad_(True)
                tb = table_h.clone().requires_grad_(True)
                mask = O.shift_mask(plan.pgrid, plan.ws, plan.ss) if plan.shifted else None
                idx = torch.from_numpy(O.relative_position_index(window))[:N, :N].reshape(-1)
                t0 = time.perf_counter()
                q, k, v = q3[:, :, 0].transpose(1, 2) * scale, q3[:, :, 1].transpose(1, 2), q3[:, :, 2].transpose(1, 2)
                sc = q @ k.transpose(-1, -2) + tb[idx].view(N, N, nH).permute(2, 0, 1)[None]
                if mask is not None:
                    sc = sc + mask[:, None]
                o_ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(ns * N, C)
                o_ref.backward(dout_h[: ns * N])
                cpu_s = time.perf_counter() - t0
                err = float((out[: ns * N].float().cpu() - o_ref.detach()).norm() / o_ref.detach().norm())
                fl = 4.0 * B_ * nH * N * N * hd
                cases.append(dict(window=list(window), N=N, heads=nH, shifted=shifted, windows=B_,
                                  fwd_ms=ms[0], bwd_ms=ms[1], windows_per_s=B_ / ((ms[0] + ms[1]) * 1e-3),
                                  fwd_tflops=fl / ms[0] / 1e9, bwd_tflops=2 * fl / ms[1] / 1e9,
                                  path="tcgen05" if N <= 448 else "cuda-core (no tcgen05 kernel for 1152-token windows)",
                                  cpu_reference_windows_per_s=ns / cpu_s, cpu_sample_windows=ns,
                                  fwd_rel_l2_vs_cpu_reference=err))
    print(json.dumps({"metric": "WindowAttention3D core fwd+bwd windows/sec (BASELINE config 5 sweep)", "unit": "windows/s",
                      "n_gpus": 1, "steps": args.steps, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "window 8x7x7 vs 8x12x12, shifted / unshifted, heads 4-32, head_dim 32, "
                                 f"{clips} clips x 4 windows", "cpu": f"{threads} threads, {cpu_model_name()}, fp32, "
                                 "video_swin.py:152-169 restated with torch ops on the first clip's windows"},
                      "cases": cases}), flush=True)


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="swin_b", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU (BASELINE config 2: 32)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="clips per CPU-baseline step (BASELINE config 1: 2)")
    ap.add_argument("--backend", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"], help="16-bit compute type (both run on tcgen05)")
    ap.add_argument("--mode", default="step", choices=["step", "attn_sweep"],
                    help="step: the fwd+bwd encoder step (the contract line); attn_sweep: BASELINE config 5 microbench sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl != "reference" and args.warmup < 3:
        args.warmup = 3   # timing rules: at least 3 warm-up steps (the JSON line reports the value actually used)

    if args.mode == "attn_sweep":
        if int(os.environ.get("RANK", 0)) == 0:
            attn_sweep(args)
        return
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    side = 384 if args.model == "swin_l_384" else 224
    workload = f"Video-{args.model} fwd+bwd, {args.batch} clips/GPU x 8x{side}^2, {args.dtype}, drop_path 0.2 train mode"
    tdtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16

    # ------------------------------------------------------------------ reference arm (host CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        steps = max(1, min(args.steps, 3))
        kind = "reference" if reference_available() else "port"
        run = cpu_true_reference_run if kind == "reference" else cpu_reference_run
        cps, sec = run(args.model, args.cpu_batch, steps, 1, threads)
        what = ("the unmodified visbackbone/video_swin.py imported from /root/reference" if kind == "reference"
                else "oracle port of visbackbone/video_swin.py (the reference itself does not exist on this box)")
        sample = (f"{steps} timed steps (1 warm-up) of fwd+bwd over {args.cpu_batch} clips of 8x{side}^2, fp32, "
                  f"{what}, torch {torch.__version__} CPU, {threads} threads, {cpu_model_name()}")
        print(json.dumps({
            "impl": "reference", "metric": "Video-Swin-B fwd+bwd clips/sec", "value": cps, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "cpu_sample_clips_per_step": args.cpu_batch},
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm (B200)
    vsw = importlib.import_module(PKG)
    VF, L = vsw.functional, vsw._lib
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    saved_stdout = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout; the contract is ONE JSON line there, so park stdout on stderr
        # until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    L.set_gemm_backend({"auto": L.GEMM_AUTO, "simt": L.GEMM_SIMT, "tcgen05": L.GEMM_TCGEN05}[args.backend])

    torch.manual_seed(0)  # identical weights on every rank
    model = vsw.SwinTransformer3D(pretrained=None, drop_path_rate=0.2, **MODELS[args.model])
    model.init_weights()
    model = model.to(dev).to(tdtype).train()
    net = model
    # gradient exchange (the one collective of the path): "flat" = gradients written into one flat buffer, 4 contiguous ranges
    # all-reduced while backward runs (dp.FlatGradReducer); "coalesced" / "overlap" / "ddp" are the earlier forms, kept for A/B
    dp_mode = os.environ.get("VSW_DP_MODE", "flat") if world > 1 else os.environ.get("VSW_DP_MODE", "none")
    reducer = vsw.dp.OverlappedGradReducer(model) if dp_mode == "overlap" else None
    flat_red = vsw.dp.FlatGradReducer(model, n_chunks=int(os.environ.get("VSW_DP_CHUNKS", "4"))) if dp_mode == "flat" else None
    if world > 1 and dp_mode == "ddp":
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[local_rank], gradient_as_bucket_view=True,
                  bucket_cap_mb=int(os.environ.get("VSW_DDP_BUCKET_MB", "50")))
    B = args.batch
    torch.manual_seed(1 + rank)  # per-rank clips
    x_host = torch.randn(B, 3, 8, side, side).pin_memory()  # fp32 frames as the data loader yields them
    x_dev = x_host.to(dev, non_blocking=True)
    with torch.no_grad():
        y0 = model(x_dev[:1])
    Rm = (torch.randn(B, *y0.shape[1:], device=dev) / 1024).to(tdtype)
    # inputs (154 MB fp32) + saved activations (GBs) exceed the 126 MB L2 every step: no explicit flush needed
    l2_note = "inputs+activations per step >> 126 MB L2 (no explicit flush)"

    def step(x, module=None):
        if flat_red is not None and module is None:
            flat_red.zero_grad()
        else:
            for p in model.parameters():
                p.grad = None
        y = (module or net)(x)
        loss = (y * Rm).sum(dtype=torch.float32)
        loss.backward()
        if flat_red is not None and module is None:
            flat_red.finish()
        elif dp_mode == "coalesced" and module is None:
            vsw.dp.all_reduce_gradients_coalesced(model.parameters())   # the one exchange step (NCCL, in place, averaged)
        elif dp_mode == "overlap" and module is None:
            reducer.finish()   # per-block coalesced all-reduces were launched from autograd hooks during backward
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step(x_dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = L.launch_count()
    ms_total = timed(lambda: step(x_dev), args.steps)
    launches = L.launch_count() - n0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- e2e: pinned host clips -> H2D -> module API -> D2H loss
    e2e = None
    if not args.no_e2e:
        # Loader-style pipeline: the H2D copy of step i+1 is issued on a copy stream before the compute of step i is
        # launched (double-buffered device clips), so every step still copies its own inputs from pinned host memory but
        # the PCIe transfer hides under the previous step; the loss is read back (D2H) every step.
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [torch.empty_like(x_dev) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0}

        def prefetch(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])          # the buffer's previous step has finished with it
                bufs[b].copy_(x_host, non_blocking=True)     # fp32 clips, pinned host memory -> HBM
                ready[b].record(copy_stream)

        def e2e_step():
            i = state["i"]
            b = i & 1
            torch.cuda.current_stream().wait_event(ready[b])
            prefetch(i + 1)
            loss = step(bufs[b])
            consumed[b].record()
            state["i"] = i + 1
            return float(loss.item())  # D2H read of the result
        prefetch(0)
        e2e_step()
        ms_e2e = timed(e2e_step, args.steps) / args.steps
        e2e = {"value": world * B / (ms_e2e / 1e3), "unit": "clips/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": world * x_host.numel() * x_host.element_size(), "d2h_bytes_per_step": 4 * world,
               "h2d_bytes_per_step_per_gpu": x_host.numel() * x_host.element_size(),
               "pipeline": "H2D of step i+1 on a copy stream overlaps the compute of step i (double-buffered)"}

    if reducer is not None:
        reducer.remove()   # the next leg runs on rank 0 alone: no collective may be issued from its backward
    if flat_red is not None:
        flat_red.remove()
    # ---- roofline of the dominant kernel family: per-launch CUDA events on the launching stream
    roofline, families = None, None
    if rank == 0:
        VF.PROFILER = VF.KernelTimer()
        torch.cuda.synchronize()
        psteps = 2
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(psteps):
            step(x_dev, model)   # the bare module: rank 0 alone runs this leg, so no collective may be issued
        pe1.record()
        torch.cuda.synchronize()
        summ = VF.PROFILER.summary()
        VF.PROFILER = None
        pk = peaks()
        step_ms = pe0.elapsed_time(pe1)
        families = {k: dict(launches=v["launches"] // psteps, ms_per_step=v["ms"] / psteps,
                            share_of_step=v["ms"] / step_ms,
                            tflops=(v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0),
                            gbs=(v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0))
                    for k, v in summ.items()}
        top = max(summ, key=lambda k: summ[k]["ms"])
        v = summ[top]
        tensor_bound = v["flops"] > 0
        ach = (v["flops"] / (v["ms"] * 1e-3) / 1e12) if tensor_bound else (v["bytes"] / (v["ms"] * 1e-3) / 1e9)
        peak = pk["tf_sustained"] if tensor_bound else pk["hbm"]
        # dram__bytes per launch from an `ncu --set full` capture of THIS workload and THIS kernel generation
        # (profiles/r02_ncu_traffic.json names the kernel symbol and the command); anything else reads null
        traffic, traffic_source = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        if os.path.exists(tpath) and args.model == "swin_b" and B == 32 and args.dtype == "bf16":
            with open(tpath) as f:
                ent = json.load(f).get(top)
            if ent and ent.get("kernel") == CURRENT_KERNEL.get(top) and os.environ.get("VSW_ATTN_TC2") is None:
                traffic, traffic_source = ent.get("avg_dram_bytes_per_launch"), ent.get("source")
        roofline = {"kernel": top, "bound": "tensor" if tensor_bound else "hbm", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": traffic_source,
                    "algorithmic_bytes_per_launch": v["bytes"] / max(1, v["launches"]),
                    "peak_source": pk["source"] + (" (sustained bf16)" if tensor_bound else ""),
                    "launches_per_step": v["launches"] // psteps,
                    "avg_launch_ms": v["ms"] / max(1, v["launches"])}
    # ---- CPU baseline (rank 0, N=1 only): oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        kind = "reference" if reference_available() else "port"
        run = cpu_true_reference_run if kind == "reference" else cpu_reference_run
        cps, sec = run(args.model, args.cpu_batch, 2, 1, threads)
        cpu = {"value": cps, "unit": "clips/s", "cores": threads, "kind": kind,
               "sample": f"2 timed steps (1 warm-up) of fwd+bwd over {args.cpu_batch} clips of 8x{side}^2, fp32, "
                         f"{sec:.2f} s/step, {cpu_model_name()}"
                         + ("" if kind == "reference" else "; oracle port (no /root/reference on this box)")}

    if rank == 0:
        pk = peaks()
        step_tflops = world * B * FWD_BWD_GFLOP_PER_CLIP[args.model] / ms_step  # GF/ms == TF/s
        out = {
            "metric": "Video-Swin-B fwd+bwd clips/sec", "value": value, "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "variant": args.model, "clips_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"dp{world}", "grad_exchange": dp_mode, "gemm_backend": args.backend, "l2": l2_note,
                       "optimizer": "excluded (metric is encoder fwd+bwd, SURVEY 8d)"},
            "clocks": sampler.result(), "e2e": e2e, "gpu_launches": int(launches),
            "step_tflops_per_gpu": step_tflops / world,
            "step_frac_of_bf16_sustained": step_tflops / world / pk["tf_sustained"],
            "roofline": roofline, "kernel_families": families, "cpu_baseline": cpu,
        }
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
