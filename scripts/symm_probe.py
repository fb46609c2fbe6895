"""Probe (2+ GPUs, torchrun): torch symmetric memory -- peer buffers, direct copies, signals."""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 32 << 20
t = symm_mem.empty(n, dtype=torch.uint8, device=torch.device("cuda", local))
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "rendezvous ok", hdl.rank, hdl.world_size, hdl.buffer_size, hdl.signal_pad_size, flush=True)
lo, hi = (rank - 1) % world, (rank + 1) % world
src = torch.full((n // 2,), rank + 1, dtype=torch.uint8, device="cuda")
peer_hi = hdl.get_buffer(hi, (n // 2,), torch.uint8, 0)        # upward message -> hi's first half
peer_lo = hdl.get_buffer(lo, (n // 2,), torch.uint8, n // 2)   # downward message -> lo's second half
torch.cuda.synchronize()
dist.barrier()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    peer_hi.copy_(src)
    peer_lo.copy_(src)
    hdl.put_signal(hi, 1, 10000)
    hdl.put_signal(lo, 0, 10000)
    hdl.wait_signal(lo, 1, 10000)
    hdl.wait_signal(hi, 0, 10000)
    e1.record()
    torch.cuda.synchronize()
    ok = bool((t[: n // 2] == lo + 1).all()) and bool((t[n // 2:] == hi + 1).all())
    print(rank, "iter", it, "ok", ok, "ms", round(e0.elapsed_time(e1), 3), "GB/s per dir",
          round(n / 2 / e0.elapsed_time(e1) / 1e6, 1), flush=True)
# small-message latency: 64 KB both ways
small = src[:65536]
ph, pl = peer_hi[:65536], peer_lo[:65536]
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ph.copy_(small); pl.copy_(small)
    hdl.put_signal(hi, 1, 10000); hdl.put_signal(lo, 0, 10000)
    hdl.wait_signal(lo, 1, 10000); hdl.wait_signal(hi, 0, 10000)
e1.record(); torch.cuda.synchronize()
print(rank, "64 KB exchange", round(e0.elapsed_time(e1) / 50 * 1000, 1), "us", flush=True)
# the same through NCCL batch_isend_irecv
r0, r1 = torch.empty_like(small), torch.empty_like(small)
def nccl_x():
    ops = [dist.P2POp(dist.isend, small, lo, tag=1), dist.P2POp(dist.isend, small, hi, tag=0),
           dist.P2POp(dist.irecv, r0, hi, tag=1), dist.P2POp(dist.irecv, r1, lo, tag=0)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
for _ in range(5): nccl_x()
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(50): nccl_x()
e1.record(); torch.cuda.synchronize()
print(rank, "64 KB NCCL exchange", round(e0.elapsed_time(e1) / 50 * 1000, 1), "us", flush=True)
w = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(5): dist.all_reduce(w, op=dist.ReduceOp.MAX)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(50): dist.all_reduce(w, op=dist.ReduceOp.MAX)
e1.record(); torch.cuda.synchronize()
print(rank, "4 B all_reduce", round(e0.elapsed_time(e1) / 50 * 1000, 1), "us", flush=True)
e0.record()
for _ in range(50): hdl.barrier(2, 10000)
e1.record(); torch.cuda.synchronize()
print(rank, "symm barrier", round(e0.elapsed_time(e1) / 50 * 1000, 1), "us", flush=True)
dist.destroy_process_group()
