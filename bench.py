#!/usr/bin/env python
"""Headline benchmark: 224-px 6-band chips/sec through the chip-inference hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Prithvi-V1-100M PrithviSeg, multi-temporal crop
segmentation shape (6 bands x 3 timesteps x 224 x 224, 13 classes), batch 64 per GPU, random-init
weights, synthetic int16 chips.  One step = raw int16 chips -> fused normalise/mask (kernel 1) ->
PrithviSeg (kernels 2-4) -> argmax int8 (fused).  N > 1: one process per GPU (torchrun), every rank
its own 64 chips (weak scaling), one NCCL all-gather of the int8 masks per step -- no other collective.

`--impl reference` times the reference's own CPU algorithm (the in-repo oracle port: the reference
is pure Python/PyTorch and is not present on the GPU box) on the host cores, same config and metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VARIANT, T, NC, BATCH = "prithvi_eo_v1_100", 3, 13, 64
# --workload: the default is BASELINE.json configs[1] (the bench line the driver records); the others are
# extra measurements of configs[2], configs[3] (tile_3660) and configs[4] (chipset_100k) with the same JSON schema
# (profiles/r01_bench_*.json).
WORKLOADS = {
    "chips_v1_100m_t3": ("prithvi_eo_v1_100", 3, 13, 64),
    "chips_v2_300m_t3": ("prithvi_eo_v2_300", 3, 13, 128),
}
CROP_MEAN = [494.905781, 815.239594, 924.335066, 2968.881459, 2634.621962, 1739.579917]
CROP_STD = [284.925432, 357.84876, 575.566823, 896.601013, 951.900334, 921.407808]
METRIC = "224px 6-band chips/sec/box (device-timed)"
WORKLOAD = "prithvi_v1_100m_T3_nc13_b64: raw int16 chips -> normalise/mask -> PrithviSeg -> argmax int8"
CPU_SAMPLE_CHIPS = 8      # chips per CPU step (reference arm and cpu_baseline leg): a bounded sample of the batch of 64
PARITY_CHIPS = 2          # chips of the TIMED GPU batch that are re-computed by the oracle (bench "parity" key)
TOL = 2e-2                # north_star: logits within 2e-2 max-abs (bf16 engine vs fp32 reference)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1436.2), d.get("hbm_gbs", 6456.2), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def gemm_traffic_from_profile():
    """Average DRAM bytes (read + write) per gemm_kernel launch of one bench step, from the committed ncu launch
    list of this same command (profiles/r01_launches_bench_step.csv: --metrics gpu__time_duration.sum,
    dram__bytes_read.sum,dram__bytes_write.sum).  None when the list is absent or has no DRAM columns."""
    path = os.path.join(ROOT, "profiles", "r02_launches_bench_step.csv")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_launches_bench_step.csv")
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import launch_summary
        seq = launch_summary.load(path)
        idx = [i for i, x in enumerate(seq) if "preprocess" in x["k"]]
        step = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
        g = [x for x in step if "gemm_kernel" in x["k"] and "dram__bytes_read.sum" in x]
        if not g:
            return None
        return sum(x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"] for x in g) / len(g)
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def stress_init(model, seed: int = 0) -> None:
    """Random-init weights of the reference's init law (the module's own constructor) plus randomised BatchNorm
    statistics / affine terms and head biases, so that the class maps are not one constant class (default BatchNorm
    statistics + zero biases predict almost a single class everywhere: a parity or bit-identity check on such maps
    proves nothing).  Plain torch, no oracle code: the oracle later runs on THIS model's state_dict."""
    import torch
    g = torch.Generator().manual_seed(1000 + seed)
    with torch.no_grad():
        for mod in model.segmentation_head.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            elif isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)) and mod.bias is not None:
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
        for blk in model.prithvi_encoder.blocks:
            for lin in (blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2):
                lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.02)


def calibrate_head_bias(model, patches) -> None:
    """Subtract the per-class mean logit of one batch from the final 1x1 convolution's bias.  Synthetic chips are
    white noise, so a random-init head's logits are a per-class constant plus small per-pixel fluctuations and the
    argmax is one class almost everywhere; with zero-mean logits the class map follows the pixels and every class
    appears.  Weights stay random-init (this is initialisation, done once before any timing; the oracle later runs on
    the model's final state_dict)."""
    import torch
    with torch.no_grad():
        logits = model.forward_patches(patches, want_logits=True)[0]
        model.segmentation_head[-1].bias.sub_(logits.mean(dim=(0, 2, 3)))


def cpu_oracle_logits(sd, raw, heads, t=None, mean=None, std=None):
    """Reference CPU path for a chip batch: normalise/mask -> PrithviSeg fp32 -> logits (torch f32)."""
    import numpy as np
    import torch
    from oracle import preprocess as OP
    from oracle import prithvi as P
    t = T if t is None else t
    x = np.stack([OP.preprocess_chip(r, None, 1.0, mean or CROP_MEAN, std or CROP_STD, t, None)[0] for r in raw])
    return P.prithvi_seg_forward(torch.from_numpy(x), sd, heads, t)


def cpu_oracle_step(sd, raw, heads):
    """... -> argmax int8 (instageo/model/infer_utils.py:99-101)."""
    from oracle import prithvi as P
    return P.argmax_int8(cpu_oracle_logits(sd, raw, heads))


def cpu_setup(model=None, raw=None):
    """Weights and chips of the CPU leg.  Inside the GPU bench they are the timed model's own state_dict and the
    first chips of a timed batch; the stand-alone reference arm draws random-init weights of the same architecture
    from the oracle's own initialiser."""
    import torch
    from oracle import prithvi as P
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if model is None:   # stand-alone reference arm: nothing of the B200 package on this path
        sd = P.make_state_dict(VARIANT, T, NC, seed=0, stress=True)
    else:
        sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items() if v.is_floating_point()}
    if raw is None:
        g = torch.Generator(device="cpu").manual_seed(1042)
        raw = torch.randint(0, 10001, (CPU_SAMPLE_CHIPS, T * 6, 224, 224), generator=g, dtype=torch.int16).numpy()
    return sd, raw, P.VARIANTS[VARIANT][2], torch.get_num_threads()


def parity_report(model, sd, heads, raw_np, logits_gpu, amax_gpu, timed_identical):
    """The oracle against the GPU result for the first chips of a TIMED batch (B = 64 forward_patches path)."""
    import torch
    ref = cpu_oracle_logits(sd, raw_np, heads)
    got = logits_gpu[: len(raw_np)].float().cpu()
    max_abs = (got - ref).abs().max().item()
    top2 = ref.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * max(max_abs, 1e-6)     # ties: margin <= 2 eps
    agree = (amax_gpu[: len(raw_np)].cpu().long() == ref.argmax(1))[safe].float().mean().item() if safe.any() else 0.0
    hist = torch.bincount(amax_gpu[: len(raw_np)].cpu().long().flatten(), minlength=NC).tolist()
    ok = bool(max_abs < TOL and agree == 1.0 and safe.float().mean().item() > 0.5 and timed_identical)
    return {"chips": len(raw_np), "of_timed_batch": BATCH, "max_abs": max_abs, "tol": TOL,
            "argmax_agree_outside_ties": agree, "excluded": 1.0 - safe.float().mean().item(),
            "timed_batch_argmax_identical": bool(timed_identical), "class_hist": hist,
            "oracle": "oracle/ (fp32 CPU restatement of the reference) on the timed model's own state_dict", "ok": ok}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd, raw, heads, cores = cpu_setup()
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_oracle_step(sd, raw, heads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(sd, raw, heads)
    dt = time.perf_counter() - t0
    v = CPU_SAMPLE_CHIPS * args.steps / dt
    sample = f"{CPU_SAMPLE_CHIPS} chips/step of the same workload (of {BATCH}), oracle port of the reference CPU path, fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "chips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "chips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "chips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


FLOOD_MEAN = [0.14245495, 0.13921481, 0.12434631, 0.31420089, 0.20743526, 0.12046503]
FLOOD_STD = [0.04036231, 0.04186983, 0.05267646, 0.0822221, 0.06834774, 0.05294205]
TILE_HW = 3660


def tile_setup(dev):
    """BASELINE.json configs[3] inputs: Prithvi-V1-100M T=1 flood head (2 classes, stress-initialised so that the
    class map is not constant) and a synthetic 3660 x 3660 x 6 int16 HLS tile with a diagonal nodata wedge."""
    import torch
    from instageo_b200.model import PrithviSeg
    torch.manual_seed(0)
    model = PrithviSeg(temporal_step=1, num_classes=2, load_pretrained_weights=False, variant="prithvi_eo_v1_100")
    stress_init(model, seed=1)
    model = model.to(dev).eval()
    H = W = TILE_HW
    g = torch.Generator().manual_seed(1042)
    tile = torch.randint(0, 10001, (6, H, W), generator=g, dtype=torch.int16)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    tile[:, (yy + xx) < 700] = -9999  # diagonal nodata wedge, like the corner of an HLS tile
    from instageo_b200 import ops
    spec = ops.PreprocessSpec([m * 1e4 for m in FLOOD_MEAN], [s * 1e4 for s in FLOOD_STD], 1, None, 1.0, -9999, dev)
    wins = torch.tensor([(0, 1800 + 224 * i, 1500) for i in range(8)], dtype=torch.int32, device=dev)
    pre = ops.preprocess(tile[:, :, :].to(dev).unsqueeze(0), spec, windows=wins, win=224, want_f32=False, want_patches=True)
    calibrate_head_bias(model, pre["patches"])
    return model, tile.pin_memory()


def tile_kw(stride, tile_batch):
    # the nodata comparison happens AFTER the constant multiplier (dataloader.py:741, 899 -- SURVEY F10), so the
    # tile is normalised in raw DN units (statistics x 1e4, multiplier 1.0) to keep -9999 recognisable
    return dict(window_size=(224, 224), stride=stride, batch_size=tile_batch, mean=[m * 1e4 for m in FLOOD_MEAN],
                std=[s * 1e4 for s in FLOOD_STD], constant_multiplier=1.0, no_data_value=-9999)


def tile_measure(model, pinned_tile, d_tile, rank, world, dev, stride, tile_batch, steps, warmup, families=False):
    """Time `steps` whole-tile passes (device-resident tile: CUDA events, max over ranks; host tile in / host map out:
    wall clock, max over ranks).  Returns a dict; `out` = the gathered class map of the last pass (device)."""
    import torch
    import torch.distributed as dist
    from instageo_b200 import _lib, ops
    from instageo_b200.model import infer_utils as IU
    kw = tile_kw(stride, tile_batch)
    H = W = TILE_HW
    n_win = len(ops.window_origins(H, 224, stride, True)) ** 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return IU.sliding_window_inference_sharded(d_tile, model, rank, world, copy=False, **kw)

    for _ in range(warmup):
        out = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item() / steps
    fam = None
    if families:
        _lib.profile_enable(True)
        _lib.profile_report()
        step()
        torch.cuda.synchronize()
        fam = _lib.profile_report()
        _lib.profile_enable(False)
    # end to end: pinned host tile in (each rank uploads only the raster rows it touches, on a copy stream), host class
    # map out (every rank brings the whole map back).  A stream of tiles: two pinned result buffers alternate and the
    # host waits for the map of tile i-1 only after tile i has been enqueued, like ChipPipeline does for chip batches.
    res = [torch.empty((H, W), dtype=torch.int8).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    for i in range(max(2, warmup)):   # both upload slots and the result stream exist before the clock starts (cudaMalloc synchronises)
        IU.sliding_window_inference_sharded(pinned_tile, model, rank, world, copy=False, out_host=res[i & 1], **kw)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        k = i & 1
        if i >= 2:
            done[k].synchronize()      # the buffer's previous map has reached the host (a consumer would read it here)
        IU.sliding_window_inference_sharded(pinned_tile, model, rank, world, copy=False, out_host=res[k], **kw)
        done[k] = IU.tile_result_event()   # the map travels on the engine's result stream while the next tile is computed
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    res = res[(steps - 1) & 1]
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return {"ms": ms, "windows": n_win, "windows_per_s": n_win / (ms / 1e3),
            "e2e_windows_per_s": n_win * steps / dt.item(), "families": fam, "out": out,
            "same_as_host_path": bool(torch.equal(out.cpu(), res))}


def tile_report(model, pinned_tile, rank, world, dev, steps, warmup, tile_batch=256):
    """The `tile` key of the default bench line: configs[3] at this world size -- stride 112 and 224 -- plus, for
    N > 1, the bit-identity of the sharded result (window exchange AND halo recompute) against the single-GPU map
    computed on every rank (the class map has both classes and the nodata wedge: `class_hist`)."""
    import torch
    import torch.distributed as dist
    from instageo_b200.model import infer_utils as IU
    d_tile = pinned_tile.to(dev)
    rep = {"workload": "tile_3660x3660x6_int16: sliding windows -> normalise/mask -> PrithviSeg V1-100M T=1 nc=2 "
                       "(stress init) -> overlap-average stitch -> int8 map; strong scaling over window / stripe shards",
           "unit": "windows/s", "steps": steps, "warmup": warmup}
    ok = True
    for stride in (112, 224):
        r = tile_measure(model, pinned_tile, d_tile, rank, world, dev, stride, tile_batch, steps, warmup)
        out = r.pop("out")
        r.pop("families")
        entry = {k: r[k] for k in ("ms", "windows", "windows_per_s", "e2e_windows_per_s", "same_as_host_path")}
        entry["class_hist"] = [int((out == k).sum()) for k in (-1, 0, 1)]
        if world > 1:
            kw = tile_kw(stride, tile_batch)
            single = IU.sliding_window_inference(d_tile, model, return_tensor=True, **kw)
            halo = IU.sliding_window_inference_sharded(d_tile, model, rank, world, halo_recompute=True, **kw)
            same = torch.tensor([int(torch.equal(out, single)), int(torch.equal(halo, single))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            entry["bit_identical_to_single_gpu"] = {"window_exchange": bool(same[0]), "halo_recompute": bool(same[1])}
            ok = ok and bool(same.min())
        ok = ok and entry["same_as_host_path"] and min(entry["class_hist"]) > 0
        rep[f"stride{stride}"] = entry
    rep["ok"] = ok
    return rep


def v2_300m_report(rank, world, dev, steps, warmup):
    """The `chips_v2_300m` key of the default bench line: BASELINE.json configs[2] (Prithvi-V2-300M, T = 3, 13 classes,
    batch 128 per GPU, bf16) through the same step as the headline workload -- device-timed chips/s over all ranks --
    and the first chip of the last timed batch against the fp32 CPU oracle on the model's own weights."""
    import torch
    import torch.distributed as dist
    from instageo_b200 import ops
    from instageo_b200.model import PrithviSeg
    variant, t, nc, batch = "prithvi_eo_v2_300", 3, 13, 128
    torch.manual_seed(0)
    model = PrithviSeg(temporal_step=t, num_classes=nc, load_pretrained_weights=False, variant=variant)
    stress_init(model, seed=2)
    model = model.to(dev).eval()
    spec = ops.PreprocessSpec(CROP_MEAN, CROP_STD, t, None, 1.0, None, dev)
    cal = torch.randint(0, 10001, (8, t * 6, 224, 224), generator=torch.Generator().manual_seed(999), dtype=torch.int16)
    calibrate_head_bias(model, ops.preprocess(cal.to(dev), spec, want_f32=False, want_patches=True)["patches"])
    g = torch.Generator().manual_seed(2042 + rank)
    raws = [torch.randint(0, 10001, (batch, t * 6, 224, 224), generator=g, dtype=torch.int16) for _ in range(2)]
    d_raws = [r.to(dev) for r in raws]   # 2 x 231 MB of int16 + > 3 GB of activations per step: larger than L2

    def step(i):
        pre = ops.preprocess(d_raws[i & 1], spec, want_f32=False, want_patches=True)
        return model.forward_patches(pre["patches"], want_logits=False, want_argmax=True)[1]

    for i in range(warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        amax = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item() / steps
    rep = {"workload": "prithvi_v2_300m_T3_nc13_b128 per GPU: raw int16 chips -> normalise/mask -> PrithviSeg -> argmax int8",
           "value": world * batch / (ms / 1e3), "unit": "chips/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "forward_path": model.graph_status()}
    if rank == 0:
        from oracle import prithvi as P
        idx = (steps - 1) & 1
        pre = ops.preprocess(d_raws[idx], spec, want_f32=False, want_patches=True)
        lg, am = model.forward_patches(pre["patches"], want_logits=True, want_argmax=True)
        sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items() if v.is_floating_point()}
        ref = cpu_oracle_logits(sd, raws[idx][:1].numpy(), P.VARIANTS[variant][2], t)
        max_abs = (lg[:1].cpu() - ref).abs().max().item()
        top2 = ref.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 2 * max(max_abs, 1e-6)
        agree = (am[:1].cpu().long() == ref.argmax(1))[safe].float().mean().item() if safe.any() else 0.0
        rep["parity"] = {"chips": 1, "max_abs": max_abs, "tol": TOL, "argmax_agree_outside_ties": agree,
                         "excluded": 1.0 - safe.float().mean().item(),
                         "timed_batch_argmax_identical": bool(torch.equal(am, amax)),
                         "ok": bool(max_abs < TOL and agree == 1.0 and torch.equal(am, amax))}
    return rep


def run_tile(args):
    """BASELINE.json configs[3]: sliding-window inference over a synthetic 3660 x 3660 x 6 int16 HLS tile with a
    nodata wedge, Prithvi-V1-100M T=1 flood head (2 classes); ranks split the windows and the output row stripes,
    exchange the window logits a stripe needs from other ranks (NCCL send/recv) and all-gather the int8 stripes in
    place.  One step = the whole tile; value = windows (224-px chips) per second over all ranks."""
    import torch
    import torch.distributed as dist

    import instageo_b200
    from instageo_b200 import _lib, ops
    from instageo_b200.model.model import flops_per_chip

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    model, pinned = tile_setup(dev)
    d_tile = pinned.to(dev)
    H = W = TILE_HW
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = tile_measure(model, pinned, d_tile, rank, world, dev, args.stride, args.tile_batch, args.steps, args.warmup,
                     families=True)
    clocks = sampler.stop() if rank == 0 else None
    fam, out, ms, n_win_total = r["families"], r["out"], r["ms"], r["windows"]
    peak_tf, peak_gbs, peak_src = peaks()
    # dominant family = the tcgen05 GEMMs (encoder linears + head convs); FLOPs of the windows THIS rank computed
    enc = model.prithvi_encoder
    fl = flops_per_chip(enc.embed_dim, len(enc.blocks), 1, 2)
    attn_fl = len(enc.blocks) * 4 * 197 * 197 * enc.embed_dim
    n_local = -(-n_win_total // world)
    gemm_ms = fam["gemm_linear"][0] + fam["gemm_conv"][0]
    gemm_n = fam["gemm_linear"][1] + fam["gemm_conv"][1]
    achieved = (fl["total"] - attn_fl) * n_local / (gemm_ms / 1e3) / 1e12 if gemm_ms else None
    st_ms, st_n = fam["stitch"]
    stitch_bytes = (n_win_total // world) * 2 * 224 * 224 * 4 + 2 * H * W // world
    fams = {k: {"ms_per_step": v[0], "launches_per_step": v[1]} for k, v in fam.items()}
    if st_ms:
        fams["stitch"]["achieved_gbs"] = stitch_bytes / (st_ms / max(1, st_n) / 1e3) / 1e9
        fams["stitch"]["frac_of_hbm_peak"] = fams["stitch"]["achieved_gbs"] / peak_gbs
    out_j = {"metric": METRIC, "value": r["windows_per_s"], "unit": "chips/s", "n_gpus": world,
             "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
             "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
             "config": {"workload": f"tile_3660x3660x6_int16_stride{args.stride}: sliding windows -> normalise/mask -> "
                                    "PrithviSeg V1-100M T=1 nc=2 -> overlap-average stitch -> int8 map",
                        "windows": n_win_total, "windows_per_call": args.tile_batch,
                        "parallelism": (f"windows and output row stripes x{world}: window logits a stripe needs from other ranks "
                                        "exchanged by NCCL send/recv (no recompute), in-place int8 stripe all-gather; "
                                        "bit-identical to 1 GPU") if world > 1 else "single GPU",
                        "l2": "tile 80 MB + window logits 116-411 MB + >1 GB activations per step, larger than L2"},
             "clocks": clocks,
             "e2e": {"value": r["e2e_windows_per_s"], "unit": "chips/s",
                     "h2d_bytes_per_step": 6 * H * W * 2 // world, "d2h_bytes_per_step": H * W,
                     "api": "instageo_b200.model.infer_utils.sliding_window_inference_sharded (pinned host tile in, host map out)"},
             "gpu_launches": int(sum(v[1] for v in fam.values())) * args.steps,
             "roofline": {"kernel": "gemm_kernel<EPI> (tcgen05 GEMM: encoder linears + head implicit-GEMM convs, T=1 shapes)",
                          "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                          "frac": achieved / peak_tf if achieved else None, "traffic": None, "peak_source": peak_src,
                          "avg_launch_ms": gemm_ms / max(1, gemm_n)},
             "kernel_families": fams,
             "same_as_host_path": r["same_as_host_path"],
             "class_hist": [int((out == k).sum()) for k in (-1, 0, 1)]}
    if rank == 0:
        print(json.dumps(out_j))
    if world > 1:
        dist.destroy_process_group()


def run_chipset(args):
    """BASELINE.json configs[4]: a large chip set (100 000 chips of 6 bands x 3 timesteps at 8 GPUs = 12 500 per GPU,
    22.6 GB of int16 resident in each GPU's HBM) through fused normalise/mask -> Prithvi-V2-300M -> argmax, rank r
    owning the contiguous chip block partition(n, world, r), then ONE NCCL all-gather of the int8 masks into dataset
    order.  One step = one pass over the whole set; value = chips/s over all ranks (weak scaling: 12 500 chips/GPU)."""
    import torch
    import torch.distributed as dist

    import instageo_b200
    from instageo_b200 import _lib, ops
    from instageo_b200.model import PrithviSeg
    from instageo_b200.model import infer_utils as IU
    from instageo_b200.model.model import flops_per_chip

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    variant, t, nc, batch = "prithvi_eo_v2_300", 3, 13, 128
    n_total = args.chips if args.chips else 12500 * world
    lo, hi = IU.partition(n_total, world, rank)
    n_local = hi - lo
    torch.manual_seed(0)
    model = PrithviSeg(temporal_step=t, num_classes=nc, load_pretrained_weights=False, variant=variant).to(dev).eval()
    spec = ops.PreprocessSpec(CROP_MEAN, CROP_STD, t, None, 1.0, None, dev)
    # the shard is generated on the device, chunk by chunk, from a seed per chunk's first chip index
    d_raw = torch.empty((n_local, t * 6, 224, 224), dtype=torch.int16, device=dev)
    for c0 in range(0, n_local, 500):
        g = torch.Generator(device=dev).manual_seed(1042 + lo + c0)
        c1 = min(n_local, c0 + 500)
        d_raw[c0:c1] = torch.randint(0, 10001, (c1 - c0, t * 6, 224, 224), generator=g, dtype=torch.int16, device=dev)
    masks = torch.empty((n_local, 224, 224), dtype=torch.int8, device=dev)

    def one_batch(i):
        pre = ops.preprocess(d_raw[i:i + batch], spec, want_f32=False, want_patches=True)
        masks[i:i + batch] = model.forward_patches(pre["patches"], want_logits=False, want_argmax=True)[1]

    def step():
        for i in range(0, n_local, batch):
            one_batch(i)
        return IU.gather_chip_masks(masks, n_total, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):          # warm-up: macro-batches, not whole passes (a pass is ~8 s)
        one_batch((i * batch) % max(1, n_local - batch))
    if world > 1:
        IU.gather_chip_masks(masks, n_total, world)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        full = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    clocks = sampler.stop() if rank == 0 else None
    assert full.shape[0] == n_total
    # gather alone (the only collective of the path)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record()
    IU.gather_chip_masks(masks, n_total, world)
    g1.record()
    barrier()
    gather_ms = g0.elapsed_time(g1)
    del full
    # roofline of the GEMM family over 8 instrumented macro-batches
    n_prof = min(8, max(1, n_local // batch))
    _lib.profile_enable(True)
    _lib.profile_report()
    for i in range(n_prof):
        one_batch(i * batch)
    torch.cuda.synchronize()
    fam = _lib.profile_report()
    _lib.profile_enable(False)
    enc = model.prithvi_encoder
    fl = flops_per_chip(enc.embed_dim, len(enc.blocks), t, nc)
    n_tok = t * 196 + 1
    gemm_fl = (fl["total"] - len(enc.blocks) * 4 * n_tok * n_tok * enc.embed_dim) * batch * n_prof
    gemm_ms = fam["gemm_linear"][0] + fam["gemm_conv"][0]
    gemm_n = fam["gemm_linear"][1] + fam["gemm_conv"][1]
    peak_tf, peak_gbs, peak_src = peaks()
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms else 0.0
    all_ms = sum(v[0] for v in fam.values())
    # end to end: the same number of macro-batches streamed from pinned HOST memory (4 rotating host batches --
    # pinning the whole 22.6 GB shard per rank would only measure the allocator), masks back on the host
    pipe = IU.ChipPipeline(model, spec, batch, dev)
    g = torch.Generator().manual_seed(7 + rank)
    pinned = [torch.randint(0, 10001, (batch, t * 6, 224, 224), generator=g, dtype=torch.int16).pin_memory() for _ in range(4)]
    n_b = (n_local + batch - 1) // batch
    sink = []
    pipe.run([pinned[i % 4] for i in range(3)], consume=lambda a: None)
    barrier()
    t0 = time.perf_counter()
    pipe.run([pinned[i % 4] for i in range(n_b)], consume=lambda a: sink.append(int(a[0, 0, 0])))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out = {"metric": METRIC, "value": n_total * args.steps / (ms / 1e3), "unit": "chips/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": f"chipset_{n_total}_chips_prithvi_v2_300m_T3_nc13: resident int16 shard -> normalise/mask -> "
                                  "PrithviSeg -> argmax int8 -> NCCL all-gather of the masks",
                      "chips_per_gpu": n_local, "shard_bytes": n_local * t * 6 * 224 * 224 * 2, "macro_batch": batch,
                      "parallelism": f"contiguous chip blocks x{world}, one int8 all-gather per pass",
                      "warmup_unit": "macro-batches of 128 chips (one pass is one step)",
                      "l2": "22.6 GB shard streamed once per pass, far larger than L2"},
           "clocks": clocks,
           "e2e": {"value": world * n_b * batch / dt.item(), "unit": "chips/s", "h2d_bytes_per_step": pipe.h2d_bytes * n_b,
                   "d2h_bytes_per_step": pipe.d2h_bytes * n_b,
                   "api": "instageo_b200.model.infer_utils.ChipPipeline.run over the shard's macro-batches (pinned host int16 in, int8 masks out)"},
           "gpu_launches": (model.launches_per_forward()) * n_b * args.steps,
           "gather": {"ms": gather_ms, "bytes_out_per_rank": n_total * 224 * 224,
                      "gbs_per_rank": n_total * 224 * 224 / (gather_ms / 1e3) / 1e9 if world > 1 else None},
           "roofline": {"kernel": "gemm_kernel<EPI> (tcgen05 GEMM: encoder linears + head implicit-GEMM convs)", "bound": "tensor",
                        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": None,
                        "peak_source": peak_src, "avg_launch_ms": gemm_ms / max(1, gemm_n),
                        "share_of_step": gemm_ms / all_ms if all_ms else None,
                        "timing": f"cuda events around every launch of {n_prof} macro-batches after the timed region"},
           "kernel_families": {k: {"ms_per_batch": v[0] / n_prof, "launches_per_batch": v[1] / n_prof} for k, v in fam.items()}}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 20 (chipset_100k: 1 pass)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--chips", type=int, default=0, help="chipset_100k: total chips (default 12500 per GPU)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tile", action="store_true", help="default workload: skip the configs[3] tile section")
    ap.add_argument("--workload", default="chips_v1_100m_t3", choices=sorted(WORKLOADS) + ["tile_3660", "chipset_100k"])
    ap.add_argument("--stride", type=int, default=224, help="tile_3660: sliding-window stride")
    ap.add_argument("--tile-batch", type=int, default=256, help="tile_3660: windows per model call (at most)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 1 if args.workload == "chipset_100k" else 20
    if args.workload == "tile_3660":
        return run_tile(args)
    if args.workload == "chipset_100k":
        return run_chipset(args)
    global VARIANT, T, NC, BATCH, WORKLOAD
    VARIANT, T, NC, BATCH = WORKLOADS[args.workload]
    if args.workload != "chips_v1_100m_t3":
        WORKLOAD = f"{VARIANT}_T{T}_nc{NC}_b{BATCH}: raw int16 chips -> normalise/mask -> PrithviSeg -> argmax int8"
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import instageo_b200
    from instageo_b200 import _lib, ops
    from instageo_b200.model import PrithviSeg
    from instageo_b200.model.infer_utils import ChipPipeline
    from instageo_b200.model.model import flops_per_chip

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    torch.manual_seed(0)
    model = PrithviSeg(temporal_step=T, num_classes=NC, load_pretrained_weights=False, variant=VARIANT)
    stress_init(model)
    model = model.to(dev).eval()
    spec = ops.PreprocessSpec(CROP_MEAN, CROP_STD, T, None, 1.0, None, dev)
    g = torch.Generator(device="cpu").manual_seed(1042 + rank)
    n_rot = 4  # rotating input batches: 4 x 115.6 MB > 126 MB L2
    raws = [torch.randint(0, 10001, (BATCH, T * 6, 224, 224), generator=g, dtype=torch.int16) for _ in range(n_rot)]
    d_raws = [r.to(dev) for r in raws]
    cal = torch.randint(0, 10001, (8, T * 6, 224, 224), generator=torch.Generator().manual_seed(999), dtype=torch.int16)
    calibrate_head_bias(model, ops.preprocess(cal.to(dev), spec, want_f32=False, want_patches=True)["patches"])  # same on every rank
    # N > 1: the step's only collective, one all-gather of the int8 masks [world * 64, 224, 224], runs on a side
    # stream so that it overlaps the next step's compute ("pipelined per macro-batch on a side stream");
    # two result buffers alternate, and the timed region ends only when the last gather has finished.
    gathered = [torch.empty((world * BATCH, 224, 224), dtype=torch.int8, device=dev) for _ in range(2)] if world > 1 else None
    gather_stream = torch.cuda.Stream(dev) if world > 1 else None

    def step(i):
        pre = ops.preprocess(d_raws[i % n_rot], spec, want_f32=False, want_patches=True)
        amax = model.forward_patches(pre["patches"], want_logits=False, want_argmax=True)[1]
        if world > 1:
            gather_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(gather_stream):
                amax.record_stream(gather_stream)
                dist.all_gather_into_tensor(gathered[i & 1], amax)
        return amax

    def barrier():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(gather_stream)
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    launches_per_step = 1 + model.launches_per_forward() - 1  # preprocess + forward (no patchify: bf16 rows in)
    graph = model.graph_status()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        last_amax = step(i)
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(gather_stream)  # the last gather is inside the timed region
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 GEMM: encoder linears + head implicit-GEMM convs):
    # CUDA events around every launch on its own stream, over `steps` instrumented steps of the same loop.
    _lib.profile_enable(True)
    _lib.profile_report()
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    fam = _lib.profile_report()
    _lib.profile_enable(False)
    enc = model.prithvi_encoder
    fl = flops_per_chip(enc.embed_dim, len(enc.blocks), T, NC)
    n_tok = T * 196 + 1
    attn_fl = len(enc.blocks) * 4 * n_tok * n_tok * enc.embed_dim
    gemm_fl_step = (fl["total"] - attn_fl) * BATCH
    gemm_ms = fam["gemm_linear"][0] + fam["gemm_conv"][0]
    gemm_launches = fam["gemm_linear"][1] + fam["gemm_conv"][1]
    peak_tf, peak_gbs, peak_src = peaks()
    achieved = gemm_fl_step * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    all_ms = sum(v[0] for v in fam.values())
    roofline = {"kernel": "gemm_kernel<EPI> (tcgen05 GEMM: encoder linears + head implicit-GEMM convs)",
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": gemm_traffic_from_profile() if args.workload == "chips_v1_100m_t3" else None,
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the "
                                "57 gemm_kernel launches of one step, profiles/r02_launches_bench_step.csv)",
                "peak_source": peak_src,
                "algorithmic_flops_per_launch": gemm_fl_step * args.steps / max(1, gemm_launches),
                "avg_launch_ms": gemm_ms / max(1, gemm_launches), "share_of_step": gemm_ms / all_ms if all_ms else None,
                "timing": "cuda events around every launch, %d instrumented steps right after the timed region" % args.steps}
    families = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in fam.items()}
    pre_bytes = BATCH * (T * 6 * 224 * 224 * 4)  # int16 in + bf16 out
    if fam["preprocess"][0] > 0:
        families["preprocess"]["achieved_gbs"] = pre_bytes * args.steps / (fam["preprocess"][0] / 1e3) / 1e9
        families["preprocess"]["frac_of_hbm_peak"] = families["preprocess"]["achieved_gbs"] / peak_gbs

    # ---- end to end through the public host API (host buffers, H2D + D2H inside the timed region)
    pipe = ChipPipeline(model, spec, BATCH, dev)
    pinned = [r.pin_memory() for r in raws]
    sink = []
    pipe.run([pinned[i % n_rot] for i in range(args.warmup)], consume=lambda a: None)
    barrier()
    t0 = time.perf_counter()
    pipe.run([pinned[i % n_rot] for i in range(args.steps)], consume=lambda a: sink.append(int(a[0, 0, 0])))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": world * BATCH * args.steps / dt.item(), "unit": "chips/s",
           "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
           "api": "instageo_b200.model.infer_utils.ChipPipeline.run (pinned host int16 in, int8 masks out)"}

    out = {"metric": METRIC, "value": value, "unit": "chips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic",
           "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                      "parallelism": f"chip-sharded x{world}, int8 mask all-gather" if world > 1 else "single GPU",
                      "l2": "4 rotating input batches (462 MB int16) + >1 GB of activations per step, larger than the 126 MB L2",
                      "weights": "random init (reference init law) + randomised BatchNorm statistics / biases (bench.stress_init)"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roofline,
           "kernel_families": families,
           "forward_path": {"cuda_graph_replay": bool(graph["last_forward_was_graph"]), "kernels_per_graph": graph["kernels_in_graph"],
                            "note": graph["note"]}}

    # ---- parity of the TIMED configuration: the argmax of the last timed step's batch (B = 64 forward_patches path)
    # must be reproduced by one more call on the same batch (which also returns the logits), and the first chips of
    # that batch are re-computed by the fp32 CPU oracle on this model's own weights
    idx = (args.steps - 1) % n_rot
    pre = ops.preprocess(d_raws[idx], spec, want_f32=False, want_patches=True)
    l_par, a_par = model.forward_patches(pre["patches"], want_logits=True, want_argmax=True)
    timed_identical = bool(torch.equal(a_par, last_amax))

    # ---- configs[3] (sliding-window tile) at this world size: throughput + N-GPU bit-identity
    tile = None
    if args.workload == "chips_v1_100m_t3" and not args.no_tile:
        del pipe, pinned
        t_model, t_pinned = tile_setup(dev)
        tile = tile_report(t_model, t_pinned, rank, world, dev, steps=max(2, min(10, args.steps)), warmup=2)
        out["tile"] = tile
        del t_model, t_pinned
        torch.cuda.empty_cache()
        out["chips_v2_300m"] = v2_300m_report(rank, world, dev, steps=max(2, min(5, args.steps)), warmup=3)

    ok = True
    if rank == 0:
        sd, raw_cpu, heads, cores = cpu_setup(model, raws[idx][:CPU_SAMPLE_CHIPS].numpy())
        out["parity"] = parity_report(model, sd, heads, raw_cpu[:PARITY_CHIPS], l_par, a_par, timed_identical)
        ok = out["parity"]["ok"] and (tile is None or tile["ok"])
        if "chips_v2_300m" in out:
            ok = ok and out["chips_v2_300m"]["parity"]["ok"]
        if world == 1 and not args.no_cpu_baseline:
            cpu_oracle_step(sd, raw_cpu[:1], heads)
            t0, n = time.perf_counter(), 0
            while n < 3 and (n == 0 or time.perf_counter() - t0 < 12):
                cpu_oracle_step(sd, raw_cpu, heads)
                n += 1
            v = CPU_SAMPLE_CHIPS * n / (time.perf_counter() - t0)
            out["cpu_baseline"] = {"value": v, "unit": "chips/s", "cores": cores, "kind": "port",
                                   "sample": f"{n} x {CPU_SAMPLE_CHIPS} chips of the timed batch (of {BATCH}) through oracle/ "
                                             "(reference CPU path, fp32, all host threads)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
