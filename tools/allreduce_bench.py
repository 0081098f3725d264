"""All-reduce of the flat gradient buffer alone: NCCL vs this library's NVLink kernel (multimem / peer), several sizes.
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/allreduce_bench.py"""
import json, os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from hier_slam_b200.mapping import FlatParams, SymmetricAllReduce
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def t(fn, n=30):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)
for mb in (12, 48, 352):
    n = mb * 1024 * 1024 // 4
    row = dict(world=world, mbytes=mb)
    x = torch.ones(n, device=dev)
    row["nccl_ms"] = round(t(lambda: dist.all_reduce(x)), 4)
    for name, mc in (("multimem", True), ("p2p", False)):
        for blocks in (64, 128):
            p = FlatParams({"g": torch.zeros(n, device=dev)}, direct_grads=False)
            try:
                ar = SymmetricAllReduce(p, blocks=blocks, multicast=mc)
                if mc and not ar.multicast_ptr:
                    row[name] = "no multicast"; break
                p.flat_grad.fill_(float(rank + 1)); ar(); torch.cuda.synchronize()
                ok = bool((p.flat_grad == world * (world + 1) / 2).all())
                row[f"{name}_b{blocks}_ms"] = round(t(ar), 4); row[f"{name}_ok"] = ok
            except Exception as ex:
                row[name] = repr(ex)[:200]; break
            del ar, p
    if rank == 0:
        for k in list(row):
            if k.endswith("_ms"): row[k.replace("_ms", "_busGBps")] = round(2 * (world - 1) / world * mb * 1.048576 / row[k], 1)
        print(json.dumps(row), flush=True)
dist.destroy_process_group()
