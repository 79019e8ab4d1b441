"""CPU: the C-ABI library loads and exports every symbol include/instageo_b200.h declares;
the host mirror refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
import instageo_b200
from instageo_b200 import _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "instageo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ig_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.ig_version() >= 100
    assert isinstance(lib.ig_last_error(), bytes)


def test_no_cpu_fallback():
    from instageo_b200.model import PrithviSeg
    m = PrithviSeg(temporal_step=1, num_classes=2, load_pretrained_weights=False, variant="prithvi_eo_tiny", depth=1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 6, 1, 224, 224))
    from instageo_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.stitch(torch.zeros(1, 2, 8, 8), [0], [0], 8, 8)
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(1, 6, 1, 224, 224))


def test_state_dict_is_drop_in():
    """Same keys/shapes as the reference PrithviSeg (188 tensors for V1-100M; model.py:292-390)."""
    from instageo_b200.model import PrithviSeg
    from oracle import prithvi as P
    m = PrithviSeg(temporal_step=3, num_classes=13, load_pretrained_weights=False, variant="prithvi_eo_v1_100", depth=2)
    sd = m.state_dict()
    ref = P.make_state_dict("prithvi_eo_v1_100", 3, 13, depth=2)
    assert set(sd) == set(ref)
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in sd)
    m.load_state_dict(ref, strict=True)
    full = PrithviSeg(temporal_step=1, num_classes=2, load_pretrained_weights=False, variant="prithvi_eo_tiny")
    assert len(full.state_dict()) == 4 + 12 * 4 + 2 + 38
    assert not any(p.requires_grad for p in full.prithvi_encoder.parameters())  # freeze_backbone default
    assert torch.equal(m.prithvi_encoder.pos_embed, ref["prithvi_encoder.pos_embed"])
    tl = PrithviSeg(temporal_step=1, load_pretrained_weights=False, variant="prithvi_eo_v2_300_tl", depth=0)
    assert "prithvi_encoder.temporal_embed_enc.scale" in tl.state_dict()
    with pytest.raises(NotImplementedError):
        PrithviSeg(load_pretrained_weights=False, variant="prithvi_eo_v2_600")


def test_new_host_mirrors_refuse_cpu():
    """§8(f) mirrors (metrics, chip masking): same no-fallback rule -- without CUDA they raise, they never compute on the host."""
    import numpy as np
    from instageo_b200.data import apply_mask, create_chip, mask_segmentation_map
    from instageo_b200.model.metrics import RunningAUC, RunningConfusionMatrix, RunningRegressionMetrics
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    for cls, args in ((RunningConfusionMatrix, (3,)), (RunningAUC, (3,)), (RunningRegressionMetrics, ())):
        with pytest.raises(RuntimeError, match="no CPU path"):
            cls(*args)
    chip = np.zeros((6, 8, 8), dtype=np.int16)
    for fn in (lambda: create_chip(chip, np.zeros((1, 8, 8), np.uint8)), lambda: apply_mask(chip, np.zeros((1, 8, 8), np.uint8), 0),
               lambda: mask_segmentation_map(chip, np.zeros((8, 8), np.int8), 0)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            fn()


def test_header_cites_reference_for_every_entry_point():
    """every compute entry point of the C ABI names the reference interface it replaces (file:line)."""
    src = open(os.path.join(ROOT, "include", "instageo_b200.h")).read()
    for ref in ("instageo/model/dataloader.py:495-524", "instageo/model/model.py:292-419", "instageo/model/infer_utils.py:96-101",
                "instageo/model/metrics.py:86-108", "instageo/model/metrics.py:209-244", "instageo/model/segmentation.py:117-156",
                "instageo/model/segmentation.py:202-213", "instageo/data/data_pipeline.py:229-267", "instageo/data/hls_utils.py:359-403"):
        assert ref in src, ref


def test_committed_bench_lines_follow_the_contract():
    """the bench lines kept under profiles/ carry every key the driver's contract names (bench.py docstring / DESIGN §6)"""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0[12]_bench_*.json")))
    assert files, "no bench lines committed"
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (f, k)
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert d["gpu_launches"] > 0 and d["value"] > 0 and d["vs_baseline"] is None
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert d["n_gpus"] > 1 or not (bad & set(d["clocks"]["reasons"])), (f, d["clocks"])
    for name in ("r01_bench_chips_v1.json", "r02_bench_chips_v1.json"):
        head = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert {"value", "unit", "cores", "kind", "sample"} <= set(head["cpu_baseline"])   # the driver's N = 1 line
        assert head["roofline"]["bound"] == "tensor" and 0 < head["roofline"]["frac"] < 1
    # round 2: parity of the timed batch and the tile section travel in the same line
    for name in ("r02_bench_chips_v1.json", "r02_bench_chips_v1_2gpu.json", "r02_bench_chips_v1_4gpu.json",
                 "r02_bench_chips_v1_8gpu.json"):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert d["parity"]["ok"] and d["parity"]["max_abs"] < d["parity"]["tol"] and d["parity"]["timed_batch_argmax_identical"]
        assert d["tile"]["ok"] and {"stride112", "stride224"} <= set(d["tile"])
        assert min(d["tile"]["stride112"]["class_hist"]) > 0                     # nodata, class 0 and class 1 all present
        if d["n_gpus"] > 1:
            assert d["tile"]["stride112"]["bit_identical_to_single_gpu"] == {"window_exchange": True, "halo_recompute": True}
            assert d["tile"]["stride224"]["bit_identical_to_single_gpu"] == {"window_exchange": True, "halo_recompute": True}
        assert d["chips_v2_300m"]["value"] > 0 and d["forward_path"]["cuda_graph_replay"]


def test_roofline_traffic_comes_from_the_committed_launch_list():
    """bench.py's roofline.traffic is read from profiles/r02_launches_bench_step.csv (the ncu launch list of the same
    command): the file is there, parses, and gives a per-launch DRAM figure of the right order (100 MB - 1 GB)."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    assert os.path.exists(os.path.join(ROOT, "profiles", "r02_launches_bench_step.csv"))
    t = bench.gemm_traffic_from_profile()
    assert t is not None and 1e8 < t < 1e9
