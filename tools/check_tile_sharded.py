"""torchrun --nproc-per-node N tools/check_tile_sharded.py : the sharded sliding window (window-row exchange and
halo recompute) against the single-GPU result computed on every rank -- must be bit-identical."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200.model import PrithviSeg  # noqa: E402
from instageo_b200.model import infer_utils as IU  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mean = [0.14245495, 0.13921481, 0.12434631, 0.31420089, 0.20743526, 0.12046503]
std = [0.04036231, 0.04186983, 0.05267646, 0.0822221, 0.06834774, 0.05294205]
from bench import calibrate_head_bias, stress_init  # noqa: E402  (randomised BatchNorm statistics, zero-mean logits)
from instageo_b200 import ops  # noqa: E402
torch.manual_seed(0)
model = PrithviSeg(temporal_step=1, num_classes=2, load_pretrained_weights=False, variant="prithvi_eo_tiny", depth=2)
stress_init(model, seed=3)
model = model.to(dev).eval()
cal = torch.randint(0, 10001, (8, 6, 224, 224), generator=torch.Generator().manual_seed(999), dtype=torch.int16).to(dev)
spec = ops.PreprocessSpec([m * 1e4 for m in mean], [s * 1e4 for s in std], 1, None, 1.0, -9999, dev)
calibrate_head_bias(model, ops.preprocess(cal, spec, want_f32=False, want_patches=True)["patches"])  # both classes appear
ok = True
for (H, W, stride) in ((1500, 1100, 112), (1030, 900, 224), (3660, 3660, 112)):
    g = torch.Generator().manual_seed(1042)
    tile = torch.randint(0, 10001, (6, H, W), generator=g, dtype=torch.int16)
    tile[:, :200, :300] = -9999
    kw = dict(window_size=(224, 224), stride=stride, batch_size=256, mean=[m * 1e4 for m in mean], std=[s * 1e4 for s in std],
              constant_multiplier=1.0, no_data_value=-9999)
    d_tile = tile.to(dev)
    full = IU.sliding_window_inference(d_tile, model, return_tensor=True, **kw)
    a = IU.sliding_window_inference_sharded(d_tile, model, rank, world, **kw)
    b = IU.sliding_window_inference_sharded(d_tile, model, rank, world, halo_recompute=True, **kw)
    c = IU.sliding_window_inference_sharded(tile.pin_memory(), model, rank, world, **kw)   # host tile: partial upload
    hist = [int((full == k).sum()) for k in (-1, 0, 1)]
    same = torch.equal(a, full) and torch.equal(b, full) and torch.equal(c, full) and min(hist) > 0
    ok = ok and same
    print(f"rank {rank}: {H}x{W} stride {stride}: exchange == single {torch.equal(a, full)}, halo == single "
          f"{torch.equal(b, full)}, host tile == single {torch.equal(c, full)}, class_hist (nodata, 0, 1) {hist}", flush=True)
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED TILE CHECK", "PASS" if int(t) == 1 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
