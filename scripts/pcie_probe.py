"""Pinned-memory copy peaks of this box (plain cudaMemcpyAsync through torch): the denominator of
the end-to-end roofline.  One process per GPU under torchrun measures the aggregate at N ranks.
Usage: python scripts/pcie_probe.py            (1 GPU)
       torchrun --nproc-per-node N scripts/pcie_probe.py"""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 2 << 30
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
out = {}
for name, dst, src in (("d2h", h, d), ("h2d", d, h)):
    best = 0.0
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        gbs = nbytes / (a.elapsed_time(b) * 1e-3) / 1e9
        if rep:
            best = max(best, gbs)
    t = torch.tensor([best], device="cuda", dtype=torch.float64)
    if world > 1:
        mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(t)
        out[name + "_gbs_min_rank"] = float(mn.item())
    out[name + "_gbs_aggregate"] = float(t.item())
if rank == 0:
    out.update(n_gpus=world, bytes_per_copy=nbytes, how="torch copy_ pinned<->device, 2 GiB, best of 3 after a warm-up, all ranks concurrently")
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
