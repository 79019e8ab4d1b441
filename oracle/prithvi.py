"""Torch-CPU fp32 oracle of ``PrithviSeg`` (TEST INFRASTRUCTURE ONLY).

Functional restatement over a plain ``state_dict`` (same keys / shapes as the reference
module so the very same tensors feed the reference, this oracle and the CUDA engine):

* patch embed, pos-embed, cls token, block loop, final norm:
  instageo/model/pritvhi.py:206-270 (PatchEmbed), :92-127 (3-D sin-cos), :498-530 (forward)
* transformer block: third-party **timm==1.0.20** ``Block`` (pinned in the reference's
  uv.lock; constructed at instageo/model/pritvhi.py:445-457 as
  ``Block(D, heads, mlp_ratio, qkv_bias=True, norm_layer=nn.LayerNorm)``).  Published
  definition: pre-norm, ``x + proj(SDPA(split(qkv(LN1 x))))``, ``x + fc2(GELU_erf(fc1(LN2 x)))``,
  LayerNorm eps 1e-5, softmax scale head_dim**-0.5, no LayerScale / DropPath / qk-norm.
  PARITY UNPINNED at this boundary: no reference test holds a Block golden vector.
* token -> image reshape and segmentation head: instageo/model/model.py:349-419
  (4 x [ConvTranspose2d k3 s2 p1 op1 -> Conv2d k3 p1 -> BatchNorm2d(eval) -> ReLU] -> Conv2d 1x1).
* argmax -> int8: instageo/model/infer_utils.py:96-101.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# variant -> (embed_dim, depth, heads, default num_frames)   instageo/model/model.py:128-168
VARIANTS = {
    "prithvi_eo_tiny": (256, 4, 4, 1),
    "prithvi_eo_v1_100": (768, 12, 12, 3),
    "prithvi_eo_v2_100": (768, 12, 12, 4),
    "prithvi_eo_v2_300": (1024, 24, 16, 4),
    "prithvi_eo_v2_300_tl": (1024, 24, 16, 4),
}
PATCH = 16
IN_CHANS = 6


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    # pritvhi.py:68-89: omega built in float32, outer product promotes to float64
    omega = np.arange(dim // 2, dtype=np.float32)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_3d(embed_dim: int, grid) -> np.ndarray:
    """pritvhi.py:92-127 -- [1 + t*h*w, D] float64, cls row zeros, layout [w | h | t]."""
    t, h, w = grid
    wd = hd = embed_dim // 16 * 6
    td = embed_dim // 16 * 4
    we = np.tile(_sincos_1d(wd, np.arange(w)), (t * h, 1))
    he = np.tile(np.repeat(_sincos_1d(hd, np.arange(h)), w, axis=0), (t, 1))
    te = np.repeat(_sincos_1d(td, np.arange(t)), h * w, axis=0)
    pe = np.concatenate((we, he, te), axis=1)
    return np.concatenate([np.zeros([1, embed_dim]), pe], axis=0)


def head_dims(embed_dim: int, temporal: int):
    return [(embed_dim * temporal) // (2 ** i) for i in range(5)]  # model.py:380-383


def make_state_dict(variant="prithvi_eo_v1_100", temporal=1, num_classes=2, depth=-1,
                    image_size=224, seed=0, stress=False) -> dict:
    """Random-init weights with the reference's init law (pritvhi.py:130-146, :463-477).

    ``stress=True`` also randomises BatchNorm statistics/affine and the final bias so that
    BN folding bugs and argmax ties are visible.
    """
    D, L, heads, _ = VARIANTS[variant]
    if depth != -1:
        L = depth
    g = torch.Generator().manual_seed(seed)

    def xavier(*shape, fan=None):
        fo, fi = fan if fan else (shape[0], int(np.prod(shape[1:])))
        a = math.sqrt(6.0 / (fi + fo))
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def uni(bound, *shape):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    grid = (temporal, image_size // PATCH, image_size // PATCH)
    sd = {}
    p = "prithvi_encoder."
    sd[p + "cls_token"] = torch.randn(1, 1, D, generator=g) * 0.02
    sd[p + "pos_embed"] = torch.from_numpy(sincos_pos_embed_3d(D, grid)).float().unsqueeze(0)
    K = IN_CHANS * PATCH * PATCH
    sd[p + "patch_embed.proj.weight"] = xavier(D, IN_CHANS, 1, PATCH, PATCH)
    sd[p + "patch_embed.proj.bias"] = uni(1 / math.sqrt(K), D)
    for i in range(L):
        b = f"{p}blocks.{i}."
        sd[b + "norm1.weight"] = torch.ones(D)
        sd[b + "norm1.bias"] = torch.zeros(D)
        sd[b + "attn.qkv.weight"] = xavier(3 * D, D)
        sd[b + "attn.qkv.bias"] = torch.zeros(3 * D)
        sd[b + "attn.proj.weight"] = xavier(D, D)
        sd[b + "attn.proj.bias"] = torch.zeros(D)
        sd[b + "norm2.weight"] = torch.ones(D)
        sd[b + "norm2.bias"] = torch.zeros(D)
        sd[b + "mlp.fc1.weight"] = xavier(4 * D, D)
        sd[b + "mlp.fc1.bias"] = torch.zeros(4 * D)
        sd[b + "mlp.fc2.weight"] = xavier(D, 4 * D)
        sd[b + "mlp.fc2.bias"] = torch.zeros(D)
    sd[p + "norm.weight"] = torch.ones(D)
    sd[p + "norm.bias"] = torch.zeros(D)
    dims = head_dims(D, temporal)
    h = "segmentation_head."
    for i in range(4):
        ci, co = dims[i], dims[i + 1]
        bt = 1 / math.sqrt(co * 9)  # torch fan_in of a ConvTranspose2d weight [Cin,Cout,3,3]
        sd[f"{h}{i}.0.weight"] = uni(bt, ci, co, 3, 3)
        sd[f"{h}{i}.0.bias"] = uni(bt, co)
        bc = 1 / math.sqrt(co * 9)
        sd[f"{h}{i}.2.weight"] = uni(bc, co, co, 3, 3)
        sd[f"{h}{i}.2.bias"] = uni(bc, co)
        if stress:
            sd[f"{h}{i}.3.weight"] = torch.rand(co, generator=g) + 0.5
            sd[f"{h}{i}.3.bias"] = torch.randn(co, generator=g) * 0.1
            sd[f"{h}{i}.3.running_mean"] = torch.randn(co, generator=g) * 0.1
            sd[f"{h}{i}.3.running_var"] = torch.rand(co, generator=g) + 0.5
        else:
            sd[f"{h}{i}.3.weight"] = torch.ones(co)
            sd[f"{h}{i}.3.bias"] = torch.zeros(co)
            sd[f"{h}{i}.3.running_mean"] = torch.zeros(co)
            sd[f"{h}{i}.3.running_var"] = torch.ones(co)
        sd[f"{h}{i}.3.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    bf = 1 / math.sqrt(dims[4])
    sd[h + "5.weight"] = uni(bf, num_classes, dims[4], 1, 1)
    sd[h + "5.bias"] = torch.randn(num_classes, generator=g) * 0.1 if stress else uni(bf, num_classes)
    if stress:
        # trained-like scale so that logits are O(1) and argmax margins are not all ties
        sd[h + "5.weight"] = sd[h + "5.weight"] * 8.0
    return sd


def _depth_of(sd) -> int:
    n = 0
    while f"prithvi_encoder.blocks.{n}.norm1.weight" in sd:
        n += 1
    return n


def block(x: torch.Tensor, sd: dict, pre: str, heads: int) -> torch.Tensor:
    """timm 1.0.20 ``Block.forward`` (fused-attention branch).  timm is not installed here (parity unpinned against
    timm itself); tests/test_oracle_block_vs_hf.py checks this function against transformers' ViTMAELayer."""
    B, N, D = x.shape
    hd = D // heads
    h = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    q, k, v = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    a = F.scaled_dot_product_attention(q, k, v)  # scale = hd ** -0.5
    a = a.transpose(1, 2).reshape(B, N, D)
    x = x + F.linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    h = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))  # exact erf
    return x + F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def encoder(img: torch.Tensor, sd: dict, heads: int, taps: dict | None = None) -> torch.Tensor:
    """``PrithviViT.forward`` (pritvhi.py:498-530) -> [B, 1+T*196, D]."""
    p = "prithvi_encoder."
    if img.dim() == 4:
        img = img.unsqueeze(2)
    x = F.conv3d(img, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"],
                 stride=(1, PATCH, PATCH))
    x = x.flatten(2).transpose(1, 2)
    pos = sd[p + "pos_embed"]
    x = x + pos[:, 1:, :]
    cls = (sd[p + "cls_token"] + pos[:, :1, :]).expand(x.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1)
    if taps is not None:
        taps["embed"] = x.clone()
    for i in range(_depth_of(sd)):
        x = block(x, sd, f"{p}blocks.{i}.", heads)
        if taps is not None:
            taps[f"block{i}"] = x.clone()
    return F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


def tokens_to_image(tokens: torch.Tensor, temporal: int) -> torch.Tensor:
    """model.py:405-413 -- drop cls, [B,T*196,D] -> [B, D*T, 14, 14] (channel = d*T + t)."""
    t = tokens[:, 1:, :]
    side = int(np.sqrt(t.shape[1] // temporal))
    return t.permute(0, 2, 1).reshape(tokens.shape[0], -1, side, side)


def seg_head(feat: torch.Tensor, sd: dict, taps: dict | None = None) -> torch.Tensor:
    """model.py:349-390 in eval mode (Dropout = identity, BatchNorm uses running stats)."""
    h = "segmentation_head."
    x = feat
    for i in range(4):
        x = F.conv_transpose2d(x, sd[f"{h}{i}.0.weight"], sd[f"{h}{i}.0.bias"], stride=2,
                               padding=1, output_padding=1)
        if taps is not None:
            taps[f"convt{i}"] = x.clone()
        x = F.conv2d(x, sd[f"{h}{i}.2.weight"], sd[f"{h}{i}.2.bias"], padding=1)
        x = F.batch_norm(x, sd[f"{h}{i}.3.running_mean"], sd[f"{h}{i}.3.running_var"],
                         sd[f"{h}{i}.3.weight"], sd[f"{h}{i}.3.bias"], False, 0.1, 1e-5)
        x = F.relu(x)
        if taps is not None:
            taps[f"stage{i}"] = x.clone()
    return F.conv2d(x, sd[h + "5.weight"], sd[h + "5.bias"])


@torch.no_grad()
def prithvi_seg_forward(img: torch.Tensor, sd: dict, heads: int, temporal: int,
                        return_features: bool = False, taps: dict | None = None):
    """``PrithviSeg.forward`` (model.py:392-419): logits [B, nc, 224, 224] float32."""
    tok = encoder(img.float(), sd, heads, taps)
    if taps is not None:
        taps["tokens"] = tok.clone()
    feat = tokens_to_image(tok, temporal)
    out = seg_head(feat, sd, taps)
    return (out, feat) if return_features else out


def argmax_int8(logits: torch.Tensor) -> np.ndarray:
    """infer_utils.py:99-101."""
    return torch.argmax(logits, dim=1).cpu().numpy().astype(np.int8)


def flops_per_chip(variant: str, temporal: int, num_classes: int, depth: int = -1) -> dict:
    """2*MAC flop count per chip (SURVEY.md §8d formulas)."""
    D, L, _, _ = VARIANTS[variant]
    if depth != -1:
        L = depth
    n = temporal * 196 + 1
    patch = 2 * temporal * 196 * 1536 * D
    enc = L * (24 * n * D * D + 4 * n * n * D)
    dims = head_dims(D, temporal)
    head, hw = 0, 14
    for i in range(4):
        head += 2 * hw * hw * 9 * dims[i] * dims[i + 1] + 2 * (2 * hw) ** 2 * 9 * dims[i + 1] ** 2
        hw *= 2
    head += 2 * 224 * 224 * dims[4] * num_classes
    return {"patch": patch, "encoder": enc, "head": head, "total": patch + enc + head}
