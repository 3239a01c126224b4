import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = sm.empty(1 << 22, dtype=torch.float64, device=torch.device("cuda", local))
h = sm.rendezvous(t, dist.group.WORLD)
print(rank, "has_multicast", h.has_multicast_support if hasattr(h, "has_multicast_support") else None,
      "mc_ptr", hex(h.multicast_ptr) if hasattr(h, "multicast_ptr") else None, "bufs", [hex(p) for p in h.buffer_ptrs], "sigpads", len(h.signal_pad_ptrs),
      [n for n in dir(h) if not n.startswith("_")], flush=True)
t.fill_(rank + 1)
h.barrier()
peer = h.get_buffer((rank + 1) % world, (8,), torch.float64)
print(rank, "peer value", peer[:2].tolist(), flush=True)
h.barrier()
dist.destroy_process_group()
