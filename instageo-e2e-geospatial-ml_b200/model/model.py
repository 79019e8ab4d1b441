"""Drop-in ``PrithviSeg`` whose forward runs on hand-written sm_100a kernels.

Host-side mirror of instageo/model/model.py:292-419 (``PrithviSeg``) and
instageo/model/pritvhi.py:370-530 (``PrithviViT``): same constructor signature, same
attribute names (``prithvi_encoder``, ``segmentation_head``, ``model_args``), same
``state_dict()`` keys and shapes -- so ``load_state_dict`` of an InstaGeo checkpoint, and
``patch("instageo.model.base.PrithviSeg", ...)`` as in the reference's own tests
(tests/model_tests/test_run.py:97), both work.  PyTorch is used only to own the parameter
tensors; ``forward`` hands device pointers to the C ABI (``ig_model_forward``).

Inference only (eval semantics, model.py:369,376,388).  There is no CPU path: calling
``forward`` with a non-CUDA tensor, in training mode, or without the built extension raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

# variant table: instageo/model/model.py:128-168 (embed_dim, depth, heads, patch, coords)
PRITHVI_VARIANTS = {
    "prithvi_eo_tiny": dict(embed_dim=256, depth=4, num_heads=4, patch=16, coords=()),
    "prithvi_eo_v1_100": dict(embed_dim=768, depth=12, num_heads=12, patch=16, coords=()),
    "prithvi_eo_v2_100": dict(embed_dim=768, depth=12, num_heads=12, patch=16, coords=()),
    "prithvi_eo_v2_300": dict(embed_dim=1024, depth=24, num_heads=16, patch=16, coords=()),
    "prithvi_eo_v2_300_tl": dict(embed_dim=1024, depth=24, num_heads=16, patch=16, coords=("time", "location")),
    "prithvi_eo_v2_600": dict(embed_dim=1280, depth=32, num_heads=16, patch=14, coords=()),
    "prithvi_eo_v2_600_tl": dict(embed_dim=1280, depth=32, num_heads=16, patch=14, coords=("time", "location")),
}
_NUM_FRAMES_DEFAULT = {"prithvi_eo_tiny": 1, "prithvi_eo_v1_100": 3}
IN_CHANS = 6


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(dim // 2, dtype=np.float32)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_3d_sincos_pos_embed(embed_dim: int, grid_size, cls_token: bool = False) -> np.ndarray:
    """Fixed 3-D sin-cos table laid out [w | h | t] (instageo/model/pritvhi.py:92-127)."""
    assert embed_dim % 16 == 0
    t, h, w = grid_size
    we = np.tile(_sincos_1d(embed_dim // 16 * 6, np.arange(w)), (t * h, 1))
    he = np.tile(np.repeat(_sincos_1d(embed_dim // 16 * 6, np.arange(h)), w, axis=0), (t, 1))
    te = np.repeat(_sincos_1d(embed_dim // 16 * 4, np.arange(t)), h * w, axis=0)
    pe = np.concatenate((we, he, te), axis=1)
    if cls_token:
        pe = np.concatenate([np.zeros([1, embed_dim]), pe], axis=0)
    return pe


class _Attention(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    """Parameter container with timm 1.0.20 ``Block`` attribute names (no forward)."""

    def __init__(self, dim: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _CoordEncoder(nn.Module):
    """``scale`` parameter of Temporal/LocationEncoder (pritvhi.py:273-367); never read by forward."""

    def __init__(self, learn: bool):
        super().__init__()
        if learn:
            self.scale = nn.Parameter(torch.zeros(1))
        else:
            self.register_buffer("scale", torch.ones(1))


class PatchEmbed(nn.Module):
    """Parameter container of the 3-D tubelet embedding (pritvhi.py:206-270)."""

    def __init__(self, input_size, patch_size, in_chans: int, embed_dim: int):
        super().__init__()
        self.input_size = tuple(input_size)
        self.patch_size = tuple(patch_size)
        self.grid_size = tuple(s // p for s, p in zip(self.input_size, self.patch_size))
        self.num_patches = self.grid_size[0] * self.grid_size[1] * self.grid_size[2]
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size, bias=True)


class PrithviViT(nn.Module):
    """Encoder parameters with the reference's names; compute lives in the CUDA engine."""

    def __init__(self, img_size=224, patch_size=(1, 16, 16), num_frames=1, in_chans=6, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, coords_encoding=None, coords_scale_learn=False, **_: Any):
        super().__init__()
        self.in_chans, self.num_frames, self.embed_dim, self.num_heads = in_chans, num_frames, embed_dim, num_heads
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.patch_embed = PatchEmbed((num_frames,) + self.img_size, patch_size, in_chans, embed_dim)
        coords_encoding = coords_encoding or []
        self.temporal_encoding = "time" in coords_encoding
        self.location_encoding = "location" in coords_encoding
        if self.temporal_encoding:
            self.temporal_embed_enc = _CoordEncoder(coords_scale_learn)
        if self.location_encoding:
            self.location_embed_enc = _CoordEncoder(coords_scale_learn)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.register_buffer("pos_embed", torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        self.initialize_weights()

    def initialize_weights(self) -> None:
        """Same init law as pritvhi.py:463-477 (xavier Linear / patch-embed, N(0,0.02) cls)."""
        pe = get_3d_sincos_pos_embed(self.pos_embed.shape[-1], self.patch_embed.grid_size, cls_token=True)
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.normal_(self.cls_token, std=0.02)
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                if mod.bias is not None:
                    mod.bias.data.zero_()
            elif isinstance(mod, nn.LayerNorm):
                mod.bias.data.zero_()
                mod.weight.data.fill_(1.0)


class PrithviSeg(nn.Module):
    """Prithvi segmentation model -- B200 engine behind InstaGeo's interface (model.py:292-419)."""

    def __init__(self, temporal_step: int = 1, image_size: int = 224, num_classes: int = 2,
                 load_pretrained_weights: bool = True, freeze_backbone: bool = True,
                 model_bands: list[int] = list(range(6)), variant: str = "prithvi_eo_v1_100",
                 embed_dims: list[int] | None = None, depth: int = -1, **kwargs: Any) -> None:
        super().__init__()
        if variant not in PRITHVI_VARIANTS:
            raise KeyError(f"unknown Prithvi variant {variant!r}")
        v = PRITHVI_VARIANTS[variant]
        if v["patch"] != 16:
            raise NotImplementedError(
                f"{variant}: patch 14 / head kernel sizes [5,5,5,7] (model.py:154-176) are outside the "
                "B200 engine's scope; supported: tiny, v1_100, v2_100, v2_300, v2_300_tl")
        in_chans = IN_CHANS * (len(model_bands) // IN_CHANS)
        if in_chans != IN_CHANS:
            raise NotImplementedError("the engine supports exactly 6 input bands per timestep (model.py:330)")
        D = int(kwargs.pop("embed_dim", v["embed_dim"]))
        L = v["depth"] if depth == -1 else depth
        heads = int(kwargs.pop("num_heads", v["num_heads"]))
        self.prithvi_encoder = PrithviViT(img_size=image_size, patch_size=(1, 16, 16), num_frames=temporal_step,
                                          in_chans=in_chans, embed_dim=D, depth=L, num_heads=heads,
                                          coords_encoding=list(v["coords"]),
                                          coords_scale_learn=bool(v["coords"]), **kwargs)
        if load_pretrained_weights:
            self._load_pretrained(variant, L)
        if freeze_backbone:
            for p in self.prithvi_encoder.parameters():
                p.requires_grad = False
        self.model_args = dict(img_size=image_size, num_frames=temporal_step, patch_size=[1, 16, 16], in_chans=in_chans,
                               embed_dim=D, depth=v["depth"], num_heads=heads, mlp_ratio=4,
                               coords_encoding=list(v["coords"]), coords_scale_learn=bool(v["coords"]))
        if embed_dims is None:
            embed_dims = [(D * temporal_step) // (2 ** i) for i in range(5)]
        if len(embed_dims) != 5:
            raise ValueError("embed_dims must list 5 channel widths")
        self.embed_dims = [int(e) for e in embed_dims]

        def upscaling_block(cin: int, cout: int) -> nn.Module:
            return nn.Sequential(
                nn.ConvTranspose2d(cin, cout, kernel_size=3, stride=2, padding=1, output_padding=1),
                nn.Dropout(0.1),
                nn.Conv2d(cout, cout, kernel_size=3, padding=1),
                nn.BatchNorm2d(cout),
                nn.ReLU(),
            )

        self.segmentation_head = nn.Sequential(
            *[upscaling_block(self.embed_dims[i], self.embed_dims[i + 1]) for i in range(4)],
            nn.Dropout(0.1),
            nn.Conv2d(self.embed_dims[-1], num_classes, kernel_size=1),
        )
        self.num_classes = num_classes
        self.temporal_step = temporal_step
        self.image_size = image_size
        self._engine: Optional[int] = None
        self._engine_sig = None
        self._engine_device = None
        self._workspaces: dict = {}
        self._sig_tensors: Optional[list] = None
        self._tap_buf: Optional[torch.Tensor] = None
        self.eval()

    # ------------------------------------------------------------------ weights
    def _load_pretrained(self, variant: str, depth: int) -> None:
        """HF download + key filtering like model.py:220-247 / utils.py:271-315 (needs network)."""
        hub = {"prithvi_eo_v1_100": ("ibm-nasa-geospatial/Prithvi-EO-1.0-100M", "Prithvi_EO_V1_100M.pt"),
               "prithvi_eo_v2_300": ("ibm-nasa-geospatial/Prithvi-EO-2.0-300M", "Prithvi_EO_V2_300M.pt"),
               "prithvi_eo_v2_300_tl": ("ibm-nasa-geospatial/Prithvi-EO-2.0-300M-TL", "Prithvi_EO_V2_300M_TL.pt")}
        if variant not in hub:
            raise AssertionError(f"No pre-trained model found for variant {variant}")
        from huggingface_hub import hf_hub_download

        path = hf_hub_download(repo_id=hub[variant][0], filename=hub[variant][1])
        sd = torch.load(path, map_location="cpu", weights_only=True)
        enc = self.prithvi_encoder
        clean = {}
        for k, val in sd.items():
            k = k.replace("_timm_module.", "")
            if "decoder" in k or "_dec" in k or k == "mask_token":
                continue
            if "pos_embed" in k:
                val = enc.pos_embed
            if not enc.temporal_encoding and "temporal_embed" in k:
                continue
            if not enc.location_encoding and "location_embed" in k:
                continue
            k = k[len("encoder."):] if k.startswith("encoder.") else k
            if k.startswith("blocks.") and int(k.split(".")[1]) >= depth:
                continue
            clean[k] = val
        enc.load_state_dict(clean, strict=True)

    def train(self, mode: bool = True):
        if mode:
            # keep nn.Module bookkeeping usable (Lightning toggles it) but refuse to run forward
            return super().train(True)
        return super().train(False)

    # ------------------------------------------------------------------ engine
    def _signature(self):
        """(address, version) of every parameter / buffer.  The tensor list is cached: building ``state_dict()`` on
        every forward cost 188-332 prefix-joined dictionary entries of host work per step.  ``_apply`` (``.to()``,
        ``.cuda()``, ``.half()``) and ``load_state_dict`` drop the cache; in-place edits bump ``_version``."""
        if self._sig_tensors is None:
            self._sig_tensors = list(self.state_dict(keep_vars=True).values())
        return tuple((t.data_ptr(), t._version) for t in self._sig_tensors)

    def _apply(self, fn, *args, **kwargs):
        self._sig_tensors = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._sig_tensors = None
        out = super().load_state_dict(*args, **kwargs)
        self._sig_tensors = None
        return out

    def _destroy_engine(self) -> None:
        if getattr(self, "_engine", None):
            try:
                _lib.load().ig_model_destroy(self._engine)
            except Exception:
                pass
        self._engine = None
        self._workspaces = {}
        self._tap_buf = None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:  # interpreter shutdown
            pass

    def _sync_engine(self, device: torch.device) -> None:
        sig = self._signature()
        if self._engine is not None and sig == self._engine_sig and device == self._engine_device:
            return
        lib = _lib.load()
        enc = self.prithvi_encoder
        if self._engine is None or device != self._engine_device:
            self._destroy_engine()
            cfg = _lib.ModelCfg(enc.embed_dim, len(enc.blocks), enc.num_heads, self.temporal_step, self.num_classes,
                                self.image_size, 16, IN_CHANS, (C.c_int * 5)(*self.embed_dims))
            handle = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(lib.ig_model_create(C.byref(cfg), C.byref(handle)))
            self._engine = handle.value
            self._engine_device = device
        with torch.cuda.device(device):
            stream = _lib.current_stream(device)
            for key, t in self.state_dict().items():
                if not t.is_floating_point():
                    continue  # num_batches_tracked
                if t.device != device:
                    raise RuntimeError(f"parameter {key} is on {t.device}, expected {device}: call model.to(device)")
                tf = t.detach().to(torch.float32).contiguous()
                shape = (C.c_int64 * max(1, tf.dim()))(*tf.shape)
                _lib.check(lib.ig_model_load_weight(self._engine, key.encode(), tf.data_ptr(), shape, tf.dim(), stream))
            _lib.check(lib.ig_model_finalize(self._engine, stream))
        torch.cuda.current_stream(device).synchronize()  # temporaries (tf) may be freed after this
        self._engine_sig = sig

    def refresh_engine(self) -> None:
        """Force a re-pack of the weights (call after in-place edits made through ``.data``)."""
        self._engine_sig = None

    def _workspace(self, batch: int, device: torch.device) -> torch.Tensor:
        need = _lib.load().ig_model_workspace_bytes(self._engine, batch) + 1024
        ws = self._workspaces.get(device)
        if ws is None or ws.numel() < need:
            # the engine caches plans / graphs / cleared borders per workspace address: the old block is about to be
            # freed (and its address may be handed to someone else), so that cache goes with it
            _lib.check(_lib.load().ig_model_reset_cache(self._engine))
            self._workspaces = {}
            ws = None
            ws = torch.zeros(need, dtype=torch.uint8, device=device)
            self._workspaces = {device: ws}
        off = (-ws.data_ptr()) % 1024
        return ws[off:]

    def _run(self, x: torch.Tensor, x_dtype: int, batch: int, want_logits: bool, want_argmax: bool,
             want_feats: bool, out_logits: Optional[torch.Tensor] = None, out_argmax: Optional[torch.Tensor] = None):
        if self.training:
            raise RuntimeError("instageo_b200.PrithviSeg is inference-only: call model.eval() first")
        if not x.is_cuda:
            raise RuntimeError("PrithviSeg.forward needs a CUDA tensor: there is no CPU path (use model.to('cuda'))")
        dev = x.device
        self._sync_engine(dev)
        lib = _lib.load()
        S, nc, T = self.image_size, self.num_classes, self.temporal_step
        ws = self._workspace(batch, dev)
        logits = amax = None
        if want_logits:
            logits = out_logits if out_logits is not None else torch.empty((batch, nc, S, S), dtype=torch.float32, device=dev)
            if logits.dtype != torch.float32 or not logits.is_contiguous() or tuple(logits.shape) != (batch, nc, S, S) \
                    or logits.device != dev:
                raise ValueError(f"out_logits must be a contiguous float32 [{batch}, {nc}, {S}, {S}] tensor on {dev}")
        if want_argmax and nc > 1:
            amax = out_argmax if out_argmax is not None else torch.empty((batch, S, S), dtype=torch.int8, device=dev)
            if amax.dtype != torch.int8 or not amax.is_contiguous() or tuple(amax.shape) != (batch, S, S) \
                    or amax.device != dev:
                raise ValueError(f"out_argmax must be a contiguous int8 [{batch}, {S}, {S}] tensor on {dev}")
        feats = (torch.empty((batch, self.embed_dims[0], S // 16, S // 16), dtype=torch.float32, device=dev)
                 if want_feats else None)
        with torch.cuda.device(dev):
            self._attach_taps(batch, dev)
            _lib.check(lib.ig_model_forward(self._engine, x.data_ptr(), x_dtype, batch, _lib.ptr(logits),
                                            _lib.ptr(amax), _lib.ptr(feats), ws.data_ptr(), ws.numel(),
                                            _lib.current_stream(dev)))
        return logits, amax, feats

    # ------------------------------------------------------------------ parity taps
    def enable_taps(self, on: bool = True) -> None:
        """Record the encoder taps of every following forward ('embed', 'block<i>', 'tokens'; test use: the forward
        then runs kernel by kernel with one device copy per block instead of replaying its CUDA graph)."""
        self._taps_on = bool(on)
        if not on:
            self._tap_buf = None
            if self._engine:
                _lib.check(_lib.load().ig_model_set_tap_buffer(self._engine, None, 0))

    def _attach_taps(self, batch: int, dev: torch.device) -> None:
        if not getattr(self, "_taps_on", False):
            return
        enc = self.prithvi_encoder
        n = (len(enc.blocks) + 2) * batch * (1 + self.temporal_step * (self.image_size // 16) ** 2) * enc.embed_dim
        if self._tap_buf is None or self._tap_buf.numel() < n or self._tap_buf.device != dev:
            self._tap_buf = torch.empty(n, dtype=torch.float32, device=dev)
        _lib.check(_lib.load().ig_model_set_tap_buffer(self._engine, self._tap_buf.data_ptr(), self._tap_buf.numel()))

    def graph_status(self) -> dict:
        """{'enabled', 'last_forward_was_graph', 'kernels_in_graph', 'note'} of the engine's CUDA-graph replay."""
        if not self._engine:
            return {"enabled": False, "last_forward_was_graph": False, "kernels_in_graph": 0, "note": "no engine yet"}
        last, nk, note = C.c_int(0), C.c_int(0), C.create_string_buffer(256)
        en = _lib.load().ig_model_graph_status(self._engine, C.byref(last), C.byref(nk), note, 256)
        return {"enabled": bool(en), "last_forward_was_graph": bool(last.value), "kernels_in_graph": nk.value,
                "note": note.value.decode("utf-8", "replace")}

    def _check_img(self, img: torch.Tensor) -> torch.Tensor:
        T, S = self.temporal_step, self.image_size
        if img.dim() == 4 and T == 1:
            img = img.unsqueeze(2)  # pritvhi.py:507-509
        if img.dim() != 5 or tuple(img.shape[1:]) != (IN_CHANS, T, S, S):
            raise ValueError(
                f"expected input [B, {IN_CHANS}, {T}, {S}, {S}], got {tuple(img.shape)}; pos-embed interpolation "
                "for other sizes (pritvhi.py:182-203) is not supported by the engine")
        return img.to(torch.float32).contiguous()

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def forward(self, img: torch.Tensor, return_features: bool = False):
        """logits [B, nc, H, W] float32 (and the pre-head features [B, D*T, 14, 14])."""
        img = self._check_img(img)
        logits, _, feats = self._run(img, _lib.IG_F32, img.shape[0], True, False, return_features)
        return (logits, feats) if return_features else logits

    @torch.no_grad()
    def predict(self, img: torch.Tensor) -> torch.Tensor:
        """Fused ``argmax(dim=1)`` -> int8 [B, H, W] (instageo/model/infer_utils.py:99-101);
        logits never leave the SM."""
        if self.num_classes == 1:
            raise RuntimeError("predict() is for classification heads; use forward() for regression")
        img = self._check_img(img)
        return self._run(img, _lib.IG_F32, img.shape[0], False, True, False)[1]

    @torch.no_grad()
    def predict_proba(self, img: torch.Tensor) -> torch.Tensor:
        """``PrithviSegmentationModule.predict_step`` (instageo/model/segmentation.py:202-213):
        ``softmax(logits, dim=1)[:, 1]`` float32 [B, H, W], fused into the head epilogue."""
        if self.num_classes < 2:
            raise RuntimeError("predict_proba() needs a classification head; a regression head's "
                               "predict_step is forward(img).squeeze(1) (regression.py:338-339)")
        img = self._check_img(img)
        if self.training:
            raise RuntimeError("instageo_b200.PrithviSeg is inference-only: call model.eval() first")
        if not img.is_cuda:
            raise RuntimeError("PrithviSeg.predict_proba needs a CUDA tensor: there is no CPU path")
        dev, B, S = img.device, img.shape[0], self.image_size
        self._sync_engine(dev)
        ws = self._workspace(B, dev)
        prob = torch.empty((B, S, S), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ig_model_predict_proba(self._engine, img.data_ptr(), _lib.IG_F32, B,
                                                          prob.data_ptr(), ws.data_ptr(), ws.numel(),
                                                          _lib.current_stream(dev)))
        return prob

    @torch.no_grad()
    def forward_patches(self, patches: torch.Tensor, want_logits: bool = True, want_argmax: bool = False,
                        out_logits: Optional[torch.Tensor] = None, out_argmax: Optional[torch.Tensor] = None):
        """Production entry: tubelet rows written by the fused preprocessing kernel
        (``ops.preprocess(..., want_patches=True)``), bf16 [B*T*196, 1536].  ``out_logits`` / ``out_argmax``: write
        into the caller's buffers (e.g. a slice of a tile's window-logit array) instead of new tensors."""
        rows = self.temporal_step * (self.image_size // 16) ** 2
        if patches.dtype != torch.bfloat16 or patches.dim() != 2 or patches.shape[1] != IN_CHANS * 256 \
                or patches.shape[0] % rows:
            raise ValueError(f"patches must be bf16 [B*{rows}, {IN_CHANS * 256}]")
        logits, amax, _ = self._run(patches.contiguous(), _lib.IG_BF16, patches.shape[0] // rows, want_logits,
                                    want_argmax, False, out_logits, out_argmax)
        return logits, amax

    def launches_per_forward(self) -> int:
        return int(_lib.load().ig_model_launches_per_forward(self._engine)) if self._engine else 0

    def debug_tap(self, name: str, batch: int, shape) -> torch.Tensor:
        """Intermediate activation of the last forward (parity tests): 'x', 'feat', 'convt<i>', 'stage<i>', and --
        after ``enable_taps()`` -- 'embed', 'block<i>', 'tokens' ([B, N, D])."""
        dev = self._engine_device
        ws = self._workspace(batch, dev)
        out = torch.empty(tuple(shape), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ig_model_debug_tap(self._engine, name.encode(), batch, ws.data_ptr(),
                                                      out.data_ptr(), out.numel(), _lib.current_stream(dev)))
        return out


def flops_per_chip(embed_dim: int, depth: int, temporal: int, num_classes: int, head_dims=None) -> dict:
    """2*MAC count per chip (SURVEY.md §8d formulas)."""
    n = temporal * 196 + 1
    dims = head_dims or [(embed_dim * temporal) // (2 ** i) for i in range(5)]
    patch = 2 * temporal * 196 * 1536 * embed_dim
    enc = depth * (24 * n * embed_dim ** 2 + 4 * n * n * embed_dim)
    head, hw = 0, 14
    for i in range(4):
        head += 2 * hw * hw * 9 * dims[i] * dims[i + 1] + 2 * (2 * hw) ** 2 * 9 * dims[i + 1] ** 2
        hw *= 2
    head += 2 * 224 * 224 * dims[4] * num_classes
    return {"patch": patch, "encoder": enc, "head": head, "total": patch + enc + head}
