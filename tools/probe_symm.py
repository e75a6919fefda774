import os, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1024, dtype=torch.int64, device=torch.device("cuda", local))
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2],
      "mc", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "pad size", hdl.signal_pad_size, flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.int64)
print(rank, "peer value", int(peer[0]), flush=True)
dist.barrier(); dist.destroy_process_group()
