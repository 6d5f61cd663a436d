"""Feasibility probe (2 GPUs, torchrun): symmetric-memory allocation, peer pointers, a kernel of this library storing into the
PEER's memory over NVLink, and how large a symmetric buffer the pool accepts."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
from henbun_b200 import _lib
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
lib = _lib.load()
nbytes = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else (1 << 30)
t = sm.empty(nbytes // 4, dtype=torch.float32, device="cuda")
h = sm.rendezvous(t, dist.group.WORLD)
ptrs = list(h.buffer_ptrs)
print(f"rank {rank}: {nbytes / 2**30:.1f} GiB symmetric buffer, ptrs {[hex(p) for p in ptrs]}, signal pads {len(h.signal_pad_ptrs)}", flush=True)
t.zero_(); torch.cuda.synchronize(); dist.barrier()
# this library's gather kernel writes 1 MiB of local data into the PEER's buffer (dst = peer pointer)
n = 1 << 18
src = torch.full((n,), float(rank + 1), device="cuda")
idx = torch.arange(n // 64, device="cuda", dtype=torch.int64)
peer = ptrs[(rank + 1) % world]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rc = lib.hb_gather_rows(C.c_void_p(peer), _lib.ptr(src.view(n // 64, 64)), _lib.ptr(idx), n // 64, 64, _lib.stream())
torch.cuda.synchronize(); dist.barrier()
got = t[:n]
ok = bool(torch.all(got == float((rank - 1) % world + 1)))
print(f"rank {rank}: rc={rc}, peer store visible: {ok}", flush=True)
# bandwidth of a big peer copy driven by a kernel (torch copy into a tensor view of the peer pointer is not available: use cudaMemcpyPeer via torch)
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dist.barrier()
dist.destroy_process_group()
