import importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a): print(f"[r{rank} {time.time()%1000:.1f}]", *a, flush=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
log("init pg")
dist.init_process_group("nccl", device_id=dev)
log("pg ok"); t = torch.ones(4, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); log("allreduce ok", t[0].item())
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
torch.manual_seed(0)
m = vsw.SwinTransformer3D(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], drop_path_rate=0.2)
m.init_weights(); m = m.to(dev).bfloat16().train()
x = torch.randn(4, 3, 8, 224, 224, device=dev)
log("single-rank step"); y = m(x); y.float().sum().backward(); torch.cuda.synchronize(); log("single ok")
from torch.nn.parallel import DistributedDataParallel as DDP
net = DDP(m, device_ids=[lr], gradient_as_bucket_view=True, bucket_cap_mb=50)
log("ddp wrapped")
for i in range(3):
    for p in m.parameters(): p.grad = None
    y = net(x); log("fwd", i); y.float().sum().backward(); log("bwd queued", i); torch.cuda.synchronize(); log("step ok", i)
dist.barrier(); log("done"); dist.destroy_process_group()
