"""Probe: does torch symmetric memory (peer pointers over NVLink) work on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, 'ptrs', [hex(p) for p in hdl.buffer_ptrs], 'pads', [hex(p) for p in hdl.signal_pad_ptrs],
      'mc', hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, 'pad bytes', hdl.signal_pad_size, flush=True)
hdl.barrier()
peer = (rank + 1) % world
rt = hdl.get_buffer(peer, (1 << 20,), torch.float32)
s = rt.sum().item()
print(rank, 'peer sum', s, 'expected', (peer + 1) * (1 << 20), flush=True)
hdl.barrier()
dist.destroy_process_group()
