"""What bounds the host->device copies when all GPUs of the box copy at once (the end-to-end arm at N = 8)?
torchrun --nproc-per-node N scripts/h2d_wall.py : per rank, a 1 GiB pinned buffer -> its GPU, (a) one rank at a time, (b) all ranks together,
and (c) all ranks together doing a plain host memcpy of the same size (host DRAM bandwidth under the same concurrency)."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); dist.barrier(); torch.cuda.synchronize()
sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
N = 1 << 30
src = torch.empty(N, dtype=torch.uint8).pin_memory(); src.fill_(1)
dst = torch.empty(N, dtype=torch.uint8, device="cuda")
host2 = np.empty(N, dtype=np.uint8)

def h2d(reps=6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    return reps * N / (a.elapsed_time(b) * 1e-3) / 1e9

def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]

alone = 0.0
for r in range(world):
    dist.barrier()
    if r == rank:
        alone = h2d()
dist.barrier()
together = h2d()
dist.barrier()
s = src.numpy()
t0 = time.perf_counter()
for _ in range(3):
    np.copyto(host2, s)
hostcpy = 3 * N / (time.perf_counter() - t0) / 1e9
res = {"alone_gbs": gather(alone), "together_gbs": gather(together), "host_memcpy_together_gbs": gather(hostcpy)}
if rank == 0:
    res["aggregate_together_gbs"] = sum(res["together_gbs"]); res["aggregate_host_memcpy_gbs"] = sum(res["host_memcpy_together_gbs"])
    res["cpus"] = os.cpu_count()
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
