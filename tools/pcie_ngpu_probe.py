"""Copy-only PCIe ceiling with N ranks at once: every rank moves one 1080p frame up and mask +
background down per iteration from page-locked memory (bgsb_copy_probe), all ranks between the same barriers.

    python tools/pcie_ngpu_probe.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/pcie_ngpu_probe.py
One JSON line per run (rank 0): per-GPU and aggregate figures, slowest rank."""
import ctypes as C, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tracking_b200 import capi
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    fd = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NPX = 1920 * 1080
out = {}
for name, bind in (("unbound", False), ("bound", True)):
    if bind:
        cores = sorted(os.sched_getaffinity(0)); per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
    up, dn, both = C.c_double(0), C.c_double(0), C.c_double(0)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    capi.check(capi.lib().bgsb_copy_probe(local, NPX * 3, NPX * 4, 300, C.byref(up), C.byref(dn), C.byref(both)))
    t = torch.tensor([up.value, dn.value, both.value], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u, d, b = t.tolist()
    out[name] = dict(h2d_us=u * 1e6, d2h_us=d * 1e6, duplex_us=b * 1e6, h2d_gbs_per_gpu=NPX * 3 / u / 1e9, d2h_gbs_per_gpu=NPX * 4 / d / 1e9,
                     duplex_gbs_per_gpu=NPX * 7 / b / 1e9, aggregate_duplex_gbs=world * NPX * 7 / b / 1e9, ceiling_mpixel_s=world * NPX / b / 1e6)
if rank == 0:
    line = json.dumps(dict(probe="pcie_ngpu", n_gpus=world, host_cores=os.cpu_count(), **out))
    if world > 1: os.write(fd, (line + "\n").encode())
    else: print(line)
if world > 1: dist.destroy_process_group()
