"""Oracle restatement of openai-CLIP ``VisionTransformer.forward`` (clip==1.0, clip/model.py:
``VisionTransformer``, ``Transformer``, ``ResidualAttentionBlock``, ``QuickGELU``,
``LayerNorm``) for ViT-B/32 -- plain torch float32 on CPU, written out op by op.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).  PARITY UNPINNED: no
pretrained CLIP weights or clip package offline; the architecture is cross-checked against
``transformers.CLIPVisionModelWithProjection`` with shared random weights in
tests/test_oracle.py.

Reference call sites: hub/compressor.py:39-40,93 (``clip.load("ViT-B/32")`` ->
``model.visual`` -> ``self.clip(X)``), lossyless/architectures.py:351-354,367-376.
"""
import math

import torch
import torch.nn.functional as F

WIDTH, LAYERS, HEADS, PATCH, RES, OUT = 768, 12, 12, 32, 224, 512


def vit_config(width=WIDTH, layers=LAYERS, heads=HEADS, patch=PATCH, res=RES, out=OUT):
    return dict(width=width, layers=layers, heads=heads, patch=patch, res=res, out=out)


def init_state_dict(seed=0, cfg=None, dtype=torch.float32):
    """Deterministic random weights with the key names of ``clip.model.VisionTransformer``
    and the init scales of ``CLIP.initialize_parameters`` (no pretrained weights offline).
    LayerNorm gains/biases and linear biases are perturbed so that a test cannot pass with
    them ignored."""
    cfg = cfg or vit_config()
    g = torch.Generator().manual_seed(seed)
    W, L, P = cfg["width"], cfg["layers"], cfg["patch"]
    n_tok = (cfg["res"] // P) ** 2 + 1
    scale = W ** -0.5
    proj_std = (W ** -0.5) * ((2 * L) ** -0.5)
    attn_std = W ** -0.5
    fc_std = (2 * W) ** -0.5
    rn = lambda *s: torch.randn(*s, generator=g)
    sd = {
        "conv1.weight": rn(W, 3, P, P) * (3 * P * P) ** -0.5,
        "class_embedding": scale * rn(W),
        "positional_embedding": scale * rn(n_tok, W),
        "ln_pre.weight": 1 + 0.1 * rn(W), "ln_pre.bias": 0.1 * rn(W),
        "ln_post.weight": 1 + 0.1 * rn(W), "ln_post.bias": 0.1 * rn(W),
        "proj": scale * rn(W, cfg["out"]),
    }
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        sd[p + "attn.in_proj_weight"] = attn_std * rn(3 * W, W)
        sd[p + "attn.in_proj_bias"] = 0.02 * rn(3 * W)
        sd[p + "attn.out_proj.weight"] = proj_std * rn(W, W)
        sd[p + "attn.out_proj.bias"] = 0.02 * rn(W)
        sd[p + "ln_1.weight"] = 1 + 0.1 * rn(W)
        sd[p + "ln_1.bias"] = 0.1 * rn(W)
        sd[p + "mlp.c_fc.weight"] = fc_std * rn(4 * W, W)
        sd[p + "mlp.c_fc.bias"] = 0.02 * rn(4 * W)
        sd[p + "mlp.c_proj.weight"] = proj_std * rn(W, 4 * W)
        sd[p + "mlp.c_proj.bias"] = 0.02 * rn(W)
        sd[p + "ln_2.weight"] = 1 + 0.1 * rn(W)
        sd[p + "ln_2.bias"] = 0.1 * rn(W)
    return {k: v.to(dtype) for k, v in sd.items()}


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)


def forward(sd, x, cfg=None, return_tokens=False):
    """x: [B,3,res,res] -> z [B,out], float32 throughout."""
    cfg = cfg or vit_config()
    W, L, H, P = cfg["width"], cfg["layers"], cfg["heads"], cfg["patch"]
    f = lambda k: sd[k].float()
    x = x.float()
    B = x.shape[0]
    t = F.conv2d(x, f("conv1.weight"), stride=P)  # [B,W,g,g]
    t = t.reshape(B, W, -1).permute(0, 2, 1)  # [B,g*g,W]
    cls = f("class_embedding")[None, None, :].expand(B, 1, W)
    t = torch.cat([cls, t], dim=1) + f("positional_embedding")[None]
    t = _ln(t, f("ln_pre.weight"), f("ln_pre.bias"))
    S, d = t.shape[1], W // H
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        h = _ln(t, f(p + "ln_1.weight"), f(p + "ln_1.bias"))
        qkv = h @ f(p + "attn.in_proj_weight").t() + f(p + "attn.in_proj_bias")
        q, k, v = qkv.split(W, dim=-1)
        q = q.reshape(B, S, H, d).transpose(1, 2)
        k = k.reshape(B, S, H, d).transpose(1, 2)
        v = v.reshape(B, S, H, d).transpose(1, 2)
        att = torch.softmax((q * (1.0 / math.sqrt(d))) @ k.transpose(-1, -2), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, W)
        t = t + o @ f(p + "attn.out_proj.weight").t() + f(p + "attn.out_proj.bias")
        h = _ln(t, f(p + "ln_2.weight"), f(p + "ln_2.bias"))
        h = h @ f(p + "mlp.c_fc.weight").t() + f(p + "mlp.c_fc.bias")
        h = h * torch.sigmoid(1.702 * h)  # QuickGELU
        t = t + h @ f(p + "mlp.c_proj.weight").t() + f(p + "mlp.c_proj.bias")
    if return_tokens:
        return t
    z = _ln(t[:, 0, :], f("ln_post.weight"), f("ln_post.bias"))
    return z @ f("proj")


def to_hf_state_dict(sd, cfg=None):
    """Map the CLIP key names onto transformers' CLIPVisionModelWithProjection (used only by
    the architecture cross-check test)."""
    cfg = cfg or vit_config()
    W = cfg["width"]
    out = {
        "vision_model.embeddings.class_embedding": sd["class_embedding"],
        "vision_model.embeddings.patch_embedding.weight": sd["conv1.weight"],
        "vision_model.embeddings.position_embedding.weight": sd["positional_embedding"],
        "vision_model.pre_layrnorm.weight": sd["ln_pre.weight"],
        "vision_model.pre_layrnorm.bias": sd["ln_pre.bias"],
        "vision_model.post_layernorm.weight": sd["ln_post.weight"],
        "vision_model.post_layernorm.bias": sd["ln_post.bias"],
        "visual_projection.weight": sd["proj"].t().contiguous(),
    }
    for i in range(cfg["layers"]):
        p = f"transformer.resblocks.{i}."
        h = f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[p + "attn.in_proj_weight"].split(W, dim=0)
        bq, bk, bv = sd[p + "attn.in_proj_bias"].split(W, dim=0)
        out.update({
            h + "self_attn.q_proj.weight": wq, h + "self_attn.q_proj.bias": bq,
            h + "self_attn.k_proj.weight": wk, h + "self_attn.k_proj.bias": bk,
            h + "self_attn.v_proj.weight": wv, h + "self_attn.v_proj.bias": bv,
            h + "self_attn.out_proj.weight": sd[p + "attn.out_proj.weight"],
            h + "self_attn.out_proj.bias": sd[p + "attn.out_proj.bias"],
            h + "layer_norm1.weight": sd[p + "ln_1.weight"], h + "layer_norm1.bias": sd[p + "ln_1.bias"],
            h + "layer_norm2.weight": sd[p + "ln_2.weight"], h + "layer_norm2.bias": sd[p + "ln_2.bias"],
            h + "mlp.fc1.weight": sd[p + "mlp.c_fc.weight"], h + "mlp.fc1.bias": sd[p + "mlp.c_fc.bias"],
            h + "mlp.fc2.weight": sd[p + "mlp.c_proj.weight"], h + "mlp.fc2.bias": sd[p + "mlp.c_proj.bias"],
        })
    return out
