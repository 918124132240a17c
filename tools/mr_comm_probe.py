import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from nekstab_b200 import lib
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
ids = [lib.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
print(rank, "id len", len(ids[0]), ids[0][:16].hex(), flush=True)
with open("/proc/self/maps") as f:
    print(rank, sorted({l.split()[-1] for l in f if "nccl" in l}), flush=True)
L = lib.load_library()
rc = L.nsb_comm_init(rank, world, ids[0], lr)
print(rank, "comm_init rc", rc, L.nsb_last_error(), flush=True)
dist.barrier()
