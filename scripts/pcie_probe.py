"""Host<->device bandwidth probe, alone and with all ranks at once (diagnostic for the e2e path)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 256 * 1024 * 1024
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
def bw(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return reps * n * 4 / (time.perf_counter() - t) / 1e9
def both():
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
h2 = torch.empty(n, dtype=torch.float32, pin_memory=True); d2 = torch.ones(n, dtype=torch.float32, device="cuda")
for phase in ("alone", "together"):
    if world > 1: dist.barrier()
    if phase == "alone" and rank != 0:
        if world > 1: dist.barrier()
        continue
    r = (bw(lambda: d.copy_(h, non_blocking=True)), bw(lambda: h.copy_(d, non_blocking=True)), 2 * bw(both))
    print(f"rank {rank} {phase}: H2D {r[0]:.1f} GB/s  D2H {r[1]:.1f} GB/s  bidirectional total {r[2]:.1f} GB/s", flush=True)
    if phase == "alone" and world > 1: dist.barrier()
t = time.perf_counter(); x = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda"); del x; torch.cuda.synchronize()
if world > 1: dist.destroy_process_group()
