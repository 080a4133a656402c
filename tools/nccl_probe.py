"""all-reduce bandwidth probe (torchrun): prints algbw for a few sizes; run with NCCL_DEBUG=INFO to see the transport."""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
for mb in (16, 64, 298):
    x = torch.ones(mb * 1024 * 1024 // 4, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        dist.all_reduce(x)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    if rank == 0:
        print(f"all_reduce {mb} MB x{world}: {ms:.3f} ms  algbw {mb / 1024 / (ms / 1e3):.1f} GB/s", flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1) if world > 1 else None)
dist.destroy_process_group()
