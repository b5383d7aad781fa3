"""ORACLE -- test infrastructure only.  Plain-PyTorch fp32 restatement of the VGGT aggregator path
(vggt/vggt/models/aggregator.py, layers/{block,attention,rope,mlp,layer_scale,patch_embed,vision_transformer}.py),
functional style with the reference's state_dict keys.  Pinned by tests/golden/vggt_mini.npz, which
oracle/make_golden.py generates from the reference package itself."""
import math

import torch
import torch.nn.functional as F

_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
_STD = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)


def rope_1d(x, pos, base):
    """rope.py:127-188 for one spatial direction; x [B,H,N,d], pos [B,N] integer."""
    d = x.shape[-1]
    inv = 1.0 / (base ** (torch.arange(0, d, 2, device=x.device).float() / d))
    ang = pos[..., None].float() * inv
    ang = torch.cat((ang, ang), -1)[:, None]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    return x * ang.cos() + torch.cat((-x2, x1), -1) * ang.sin()


def rope_2d(x, pos, base):
    v, h = x.chunk(2, -1)
    return torch.cat((rope_1d(v, pos[..., 0], base), rope_1d(h, pos[..., 1], base)), -1)


def block(sd, p, x, heads, ln_eps, pos=None, rope_base=0.0):
    """block.py:77-98 (eval path): x + ls1(attn(norm1 x)); x + ls2(mlp(norm2 x))."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ln_eps)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, heads, C // heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
    if p + "attn.q_norm.weight" in sd:
        q = F.layer_norm(q, (C // heads,), sd[p + "attn.q_norm.weight"], sd[p + "attn.q_norm.bias"], 1e-5)
        k = F.layer_norm(k, (C // heads,), sd[p + "attn.k_norm.weight"], sd[p + "attn.k_norm.bias"], 1e-5)
    if rope_base > 0 and pos is not None:
        q, k = rope_2d(q, pos, rope_base), rope_2d(k, pos, rope_base)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C)
    o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = x + (o * sd[p + "ls1.gamma"] if p + "ls1.gamma" in sd else o)
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ln_eps)
    h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                 sd[p + "mlp.fc2.bias"])
    return x + (h * sd[p + "ls2.gamma"] if p + "ls2.gamma" in sd else h)


def dino_patch_tokens(sd, p, images, depth, heads, patch, nreg):
    """vision_transformer.py forward_features -> x_norm_patchtokens."""
    x = F.conv2d(images, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=patch)
    hp, wp = x.shape[-2:]
    x = x.flatten(2).transpose(1, 2)
    C = x.shape[-1]
    x = torch.cat((sd[p + "cls_token"].expand(x.shape[0], -1, -1), x), 1)
    pos = sd[p + "pos_embed"]
    N = pos.shape[1] - 1
    if not (hp * wp == N and images.shape[-2] == images.shape[-1]):
        M = int(math.sqrt(N))
        pp = F.interpolate(pos[:, 1:].reshape(1, M, M, C).permute(0, 3, 1, 2), size=(hp, wp), mode="bicubic",
                           antialias=True).permute(0, 2, 3, 1).reshape(1, -1, C)
        pos = torch.cat((pos[:, :1], pp), 1)
    x = x + pos
    if nreg:
        x = torch.cat((x[:, :1], sd[p + "register_tokens"].expand(x.shape[0], -1, -1), x[:, 1:]), 1)
    for i in range(depth):
        x = block(sd, f"{p}blocks.{i}.", x, heads, 1e-6)
    x = F.layer_norm(x, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    return x[:, 1 + nreg:]


def aggregator(sd, images, depth, heads, patch=14, nreg=4, rope_base=100.0, vit_depth=0):
    """aggregator.py:184-258.  images [B,S,3,H,W] in [0,1]; returns the list of `depth` [B,S,P,2C] tensors."""
    B, S, _, H, W = images.shape
    dev = images.device
    x = ((images - _MEAN.to(dev)[None]) / _STD.to(dev)[None]).view(B * S, 3, H, W)
    if vit_depth:
        pt = dino_patch_tokens(sd, "patch_embed.", x, vit_depth, heads, patch, nreg)
    else:
        pt = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    C = pt.shape[-1]

    def special(tok):
        return torch.cat([tok[:, 0:1].expand(B, 1, *tok.shape[2:]), tok[:, 1:].expand(B, S - 1, *tok.shape[2:])],
                         1).reshape(B * S, *tok.shape[2:])

    tokens = torch.cat([special(sd["camera_token"]), special(sd["register_token"]), pt], 1)
    hp, wp = H // patch, W // patch
    ys, xs = torch.meshgrid(torch.arange(hp, device=dev), torch.arange(wp, device=dev), indexing="ij")
    pos = torch.stack((ys, xs), -1).reshape(1, hp * wp, 2).expand(B * S, -1, -1) + 1
    pos = torch.cat([torch.zeros(B * S, 1 + nreg, 2, dtype=pos.dtype, device=dev), pos], 1)
    P = tokens.shape[1]
    out = []
    for i in range(depth):
        tokens = block(sd, f"frame_blocks.{i}.", tokens.view(B * S, P, C), heads, 1e-5, pos.view(B * S, P, 2), rope_base)
        fi = tokens.view(B, S, P, C)
        tokens = block(sd, f"global_blocks.{i}.", tokens.view(B, S * P, C), heads, 1e-5, pos.view(B, S * P, 2), rope_base)
        out.append(torch.cat([fi, tokens.view(B, S, P, C)], -1))
    return out
