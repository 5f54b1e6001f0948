"""dev: 2-rank check of the cudaIpc peer regions (torchrun --nproc-per-node 2 scripts/peer_probe.py)."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import _lib  # noqa: E402
from easyfea_b200 import device as dv  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 1 << 16
p = ctypes.c_void_p()
_lib.call("efb_peer_alloc", n * 8, ctypes.byref(p))
mine = dv.view_f64(p.value, n)
print(rank, "view ptr", hex(mine.data_ptr()), "alloc ptr", hex(p.value), "aliases", mine.data_ptr() == p.value, flush=True)
h = (ctypes.c_char * 64)()
_lib.call("efb_peer_export", p, h)
got = [None] * world
dist.all_gather_object(got, bytes(h.raw))
peers = {}
for q in range(world):
    if q != rank:
        pp = ctypes.c_void_p()
        _lib.call("efb_peer_open", (ctypes.c_char * 64).from_buffer_copy(got[q]), ctypes.byref(pp))
        peers[q] = dv.view_f64(pp.value, n)
dist.barrier()
for q, t in peers.items():
    t[rank * 8:(rank + 1) * 8] = float(rank + 1)  # torch kernel storing into the peer's region
torch.cuda.synchronize()
dist.barrier()
print(rank, "my region after the peers' stores:", mine[:16].tolist(), flush=True)
dist.barrier()
dist.destroy_process_group()
