"""Where one sharded tile pass spends its device time, phase by phase (CUDA events on the compute stream, max and mean
over ranks): torchrun --nproc-per-node N tools/tile_phases.py [stride ...].  The NCCL work of the exchange and of the
gather runs on NCCL's stream; the marks are taken after the compute stream has been made to wait for it."""
import os
import sys

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
for stride in [int(a) for a in sys.argv[1:]] or [112, 224]:
    kw = bench.tile_kw(stride, 256)
    for src, label in ((d_tile, "device-resident tile"), (pinned, "pinned host tile")):
        for _ in range(3):
            IU.sliding_window_inference_sharded(src, model, rank, world, copy=False, **kw)
        eng = next(e for e in IU._ENGINES.values() if e.stride == stride)
        eng.timing = True
        acc = {}
        reps = 5
        for _ in range(reps):
            IU.sliding_window_inference_sharded(src, model, rank, world, copy=False, **kw)
            for k, v in eng.phase_ms().items():
                acc[k] = acc.get(k, 0.0) + v / reps
        eng.timing = False
        keys = list(acc)
        t = torch.tensor([acc[k] for k in keys], device=dev)
        mx, mean = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(mean)
            mean /= world
        if rank == 0:
            print(f"== stride {stride}, {world} GPU(s), {label}: {eng.n_calls} model call(s) of <= {eng.per_call} windows per rank, "
                  f"{len(eng.sends)} sends / {len(eng.recvs)} receives on rank 0")
            for k, a, b in zip(keys, mx.tolist(), mean.tolist()):
                print(f"   {k:40s} max {a:7.3f} ms   mean {b:7.3f} ms")
            print(f"   {'sum of the per-phase maxima':40s}     {sum(mx.tolist()):7.3f} ms")
if world > 1:
    dist.destroy_process_group()
