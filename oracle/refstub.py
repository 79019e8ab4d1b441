"""Dev-container-only recipe for importing the REAL reference modules (TEST INFRASTRUCTURE).

``/root/reference`` is read-only and does not exist on the GPU box, so nothing that runs
there may import this.  It is used by ``oracle/gen_golden.py`` (to freeze golden vectors
under ``tests/golden/``) and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent).

The reference's ``instageo/model/{model,pritvhi,dataloader}.py`` run unmodified once the
third-party imports that are absent from this image are stubbed in ``sys.modules``
(SURVEY.md Appendix C).  The only stub with arithmetic is timm's ``Block`` -- timm==1.0.20
is a pinned dependency that is not vendored in the reference; its published definition is
restated below with timm's attribute names so reference checkpoints keep their keys.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("INSTAGEO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "instageo", "model"))


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    """timm.models.vision_transformer.Block at the arguments pritvhi.py:445-457 passes."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, norm_layer=nn.LayerNorm,
                 drop_path=0.0, **_):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def install() -> None:
    """Seed sys.modules and put the reference on sys.path (idempotent)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    _stub("timm")
    _stub("timm.layers", to_2tuple=to_2tuple)
    _stub("timm.models")
    _stub("timm.models.vision_transformer", Block=Block)
    _stub("codecarbon", EmissionsTracker=object)
    _stub("codecarbon.output", EmissionsData=object)
    _stub("neptune", Run=object)
    _stub("ptflops", get_model_complexity_info=lambda *a, **k: (0, 0))
    _stub("pytorch_lightning", LightningModule=nn.Module, Trainer=object, Callback=object)
    _stub("pytorch_lightning.callbacks", Callback=object)
    _stub("rasterio")
    _stub("xarray", Dataset=object, DataArray=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_prithvi_seg(**kwargs):
    install()
    from instageo.model.model import PrithviSeg  # type: ignore

    return PrithviSeg(load_pretrained_weights=False, **kwargs).eval()


def reference_dataloader():
    install()
    import instageo.model.dataloader as dl  # type: ignore

    return dl
