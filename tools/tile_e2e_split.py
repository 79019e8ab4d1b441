"""What the end-to-end tile loop pays on top of the device-timed pass: wall clock per tile of the pipelined loop of bench.py
(two pinned result buffers) with / without the host upload and with / without the result copy, max over ranks.
torchrun --nproc-per-node N tools/tile_e2e_split.py [stride ...]"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from instageo_b200.model import infer_utils as IU  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model, pinned = bench.tile_setup(dev)
d_tile = pinned.to(dev)
H = W = bench.TILE_HW
steps = 12


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for stride in [int(a) for a in sys.argv[1:]] or [112, 224]:
    kw = bench.tile_kw(stride, 256)
    for src, sname in ((d_tile, "device tile"), (pinned, "host tile  ")):
        for d2h in (False, True):
            res = [torch.empty((H, W), dtype=torch.int8).pin_memory() for _ in range(2)]
            done = [torch.cuda.Event(), torch.cuda.Event()]
            for i in range(3):
                IU.sliding_window_inference_sharded(src, model, rank, world, copy=False, out_host=res[i & 1] if d2h else None, **kw)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                k = i & 1
                if i >= 2:
                    done[k].synchronize()
                IU.sliding_window_inference_sharded(src, model, rank, world, copy=False, out_host=res[k] if d2h else None, **kw)
                if d2h:
                    done[k] = IU.tile_result_event()
                else:
                    done[k].record()
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / steps * 1e3], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"stride {stride} {world} GPU(s) {sname} {'+ result to host' if d2h else '                '}: {dt.item():7.3f} ms per tile")
if world > 1:
    dist.destroy_process_group()
