"""CPU: the oracle's transformer block (a restatement of timm==1.0.20 ``Block``, which is absent from the reference
tree and from this image -- "parity unpinned" in DESIGN.md §3) cross-checked against an INDEPENDENT implementation of
the same MAE/ViT block that IS installed: transformers' ``ViTMAELayer`` (pre-norm MHSA with bias, scale head_dim**-0.5,
exact-erf GELU MLP, LayerNorm eps 1e-5 as timm's default).  Prithvi's encoder is the MAE ViT, so the two must agree to
fp32 round-off once timm's fused qkv weight is split into the q / k / v linears.  This does not pin timm itself, but it
does pin the block arithmetic against a second published implementation."""
import pytest
import torch

from oracle import prithvi as P


@pytest.mark.parametrize("variant,T", [("prithvi_eo_tiny", 1), ("prithvi_eo_v1_100", 3)])
def test_oracle_block_matches_transformers_vit_mae_layer(variant, T):
    M = pytest.importorskip("transformers.models.vit_mae.modeling_vit_mae")
    from transformers import ViTMAEConfig
    D, _, heads = P.VARIANTS[variant][:3]
    sd = P.make_state_dict(variant, T, 2, depth=1, seed=11, stress=True)
    pre = "prithvi_encoder.blocks.0."
    cfg = ViTMAEConfig(hidden_size=D, num_attention_heads=heads, intermediate_size=4 * D, hidden_act="gelu",
                       layer_norm_eps=1e-5, qkv_bias=True, attention_probs_dropout_prob=0.0, hidden_dropout_prob=0.0)
    cfg._attn_implementation = "eager"
    layer = M.ViTMAELayer(cfg).eval()
    qw, qb = sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]       # timm: rows (q | k | v) x (head, 64)
    hf = {"attention.attention.query.weight": qw[:D], "attention.attention.key.weight": qw[D:2 * D],
          "attention.attention.value.weight": qw[2 * D:], "attention.attention.query.bias": qb[:D],
          "attention.attention.key.bias": qb[D:2 * D], "attention.attention.value.bias": qb[2 * D:],
          "attention.output.dense.weight": sd[pre + "attn.proj.weight"], "attention.output.dense.bias": sd[pre + "attn.proj.bias"],
          "intermediate.dense.weight": sd[pre + "mlp.fc1.weight"], "intermediate.dense.bias": sd[pre + "mlp.fc1.bias"],
          "output.dense.weight": sd[pre + "mlp.fc2.weight"], "output.dense.bias": sd[pre + "mlp.fc2.bias"],
          "layernorm_before.weight": sd[pre + "norm1.weight"], "layernorm_before.bias": sd[pre + "norm1.bias"],
          "layernorm_after.weight": sd[pre + "norm2.weight"], "layernorm_after.bias": sd[pre + "norm2.bias"]}
    layer.load_state_dict(hf, strict=True)
    x = torch.randn(2, 1 + T * 196, D, generator=torch.Generator().manual_seed(5)) * 2.0
    with torch.no_grad():
        want = layer(x)
        want = want[0] if isinstance(want, (tuple, list)) else want
        got = P.block(x, sd, pre, heads)
    err = (got - want).abs().max().item()
    assert err < 5e-5 * max(1.0, want.abs().max().item()), err   # fp32 CPU, different op order only
