"""PCIe probe with every rank of a box copying at once (torchrun): per-rank and aggregate H2D / D2H GB/s from pinned host
memory.  Explains the multi-GPU e2e floor: GPUs that share a PCIe switch / root port share its bandwidth."""
import json, os, sys, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else (1 << 30)
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
def timed(fn):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t_h2d = timed(lambda: d.copy_(h, non_blocking=True)); t_d2h = timed(lambda: h.copy_(d, non_blocking=True))
v = torch.tensor([n / t_h2d / 1e6, n / t_d2h / 1e6], dtype=torch.float64, device="cuda")
allv = [torch.empty_like(v) for _ in range(world)]
dist.all_gather(allv, v)
if rank == 0:
    print(json.dumps({"ranks_copying_at_once": world, "GiB_each": n / (1 << 30), "h2d_GBps_per_rank": [round(float(x[0]), 1) for x in allv],
                      "d2h_GBps_per_rank": [round(float(x[1]), 1) for x in allv], "h2d_GBps_aggregate": round(float(sum(x[0] for x in allv)), 1),
                      "d2h_GBps_aggregate": round(float(sum(x[1] for x in allv)), 1)}))
dist.destroy_process_group()
