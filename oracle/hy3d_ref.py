"""ORACLE -- test infrastructure only.  Nothing under 3d-re-gen_b200/ imports this file.

Plain-PyTorch fp32 restatements (functional style, weights passed as the reference's own state_dict) of the
floating-point part of the hot path.  Each function names the reference lines it restates; all paths are
relative to /root/reference/Hunyuan3D-2/hy3dgen/shapegen/ unless they start with vggt/.

Pinned by tests/test_oracle_vs_reference.py against the reference's own modules imported from
/root/reference (when that checkout is present, i.e. in the build container) and against the committed
fixtures tests/golden/*.npz that oracle/make_golden.py generated from those modules.
Device-agnostic: the GPU parity tests run these same functions in fp32 (TF32 off) as the numerical
reference for the fp16 tensor-core kernels.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- DiT
def timestep_embedding(t, dim=256, max_period=10000, time_factor=1000.0):
    """models/denoisers/hunyuan3ddit.py:39-60 -- cat(cos, sin); `time_factor * t` in t's dtype."""
    t = time_factor * t
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return emb.to(t.dtype)


def _lin(sd, name, x):
    b = sd.get(name + ".bias")
    return F.linear(x, sd[name + ".weight"], b)


def _rms(x, scale):
    """hunyuan3ddit.py:83-92 (RMSNorm over the head dim, eps 1e-6)."""
    xf = x.float()
    rrms = torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + 1e-6)
    return (xf * rrms).to(x.dtype) * scale


def _split_heads_khd(qkv, heads):
    """rearrange "B L (K H D) -> K B H L D" (hunyuan3ddit.py:196)."""
    B, L, _ = qkv.shape
    return qkv.view(B, L, 3, heads, -1).permute(2, 0, 3, 1, 4)


def _sdpa(q, k, v):
    """hunyuan3ddit.py:33-36: SDPA then "B H L D -> B L (H D)"."""
    o = F.scaled_dot_product_attention(q, k, v)
    B, H, L, D = o.shape
    return o.permute(0, 2, 1, 3).reshape(B, L, H * D)


def _modulation(sd, name, vec, n):
    """hunyuan3ddit.py:138-152."""
    out = _lin(sd, name + ".lin", F.silu(vec))[:, None, :]
    return out.chunk(n, dim=-1)


def _ln(x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


def double_block(sd, pre, img, txt, vec, heads):
    """DoubleStreamBlock.forward, hunyuan3ddit.py:189-217."""
    i_sh1, i_sc1, i_g1, i_sh2, i_sc2, i_g2 = _modulation(sd, pre + "img_mod", vec, 6)
    t_sh1, t_sc1, t_g1, t_sh2, t_sc2, t_g2 = _modulation(sd, pre + "txt_mod", vec, 6)
    img_m = (1 + i_sc1) * _ln(img) + i_sh1
    iq, ik, iv = _split_heads_khd(_lin(sd, pre + "img_attn.qkv", img_m), heads)
    iq = _rms(iq, sd[pre + "img_attn.norm.query_norm.scale"])
    ik = _rms(ik, sd[pre + "img_attn.norm.key_norm.scale"])
    txt_m = (1 + t_sc1) * _ln(txt) + t_sh1
    tq, tk, tv = _split_heads_khd(_lin(sd, pre + "txt_attn.qkv", txt_m), heads)
    tq = _rms(tq, sd[pre + "txt_attn.norm.query_norm.scale"])
    tk = _rms(tk, sd[pre + "txt_attn.norm.key_norm.scale"])
    attn = _sdpa(torch.cat((tq, iq), 2), torch.cat((tk, ik), 2), torch.cat((tv, iv), 2))
    t_attn, i_attn = attn[:, :txt.shape[1]], attn[:, txt.shape[1]:]
    img = img + i_g1 * _lin(sd, pre + "img_attn.proj", i_attn)
    h = _lin(sd, pre + "img_mlp.0", (1 + i_sc2) * _ln(img) + i_sh2)
    img = img + i_g2 * _lin(sd, pre + "img_mlp.2", F.gelu(h, approximate="tanh"))
    txt = txt + t_g1 * _lin(sd, pre + "txt_attn.proj", t_attn)
    h = _lin(sd, pre + "txt_mlp.0", (1 + t_sc2) * _ln(txt) + t_sh2)
    txt = txt + t_g2 * _lin(sd, pre + "txt_mlp.2", F.gelu(h, approximate="tanh"))
    return img, txt


def single_block(sd, pre, x, vec, heads):
    """SingleStreamBlock.forward, hunyuan3ddit.py:254-267."""
    shift, scale, gate = _modulation(sd, pre + "modulation", vec, 3)
    hidden = x.shape[-1]
    x_mod = (1 + scale) * _ln(x) + shift
    lin1 = _lin(sd, pre + "linear1", x_mod)
    qkv, mlp = lin1[..., :3 * hidden], lin1[..., 3 * hidden:]
    q, k, v = _split_heads_khd(qkv, heads)
    q = _rms(q, sd[pre + "norm.query_norm.scale"])
    k = _rms(k, sd[pre + "norm.key_norm.scale"])
    attn = _sdpa(q, k, v)
    out = _lin(sd, pre + "linear2", torch.cat((attn, F.gelu(mlp, approximate="tanh")), 2))
    return x + gate * out


def dit_forward(sd, x, t, cond, heads, depth, depth_single, taps=None):
    """Hunyuan3DDiT.forward, hunyuan3ddit.py:381-410 (guidance_embed False).
    taps (optional list) receives the hidden state after every block."""
    latent = _lin(sd, "latent_in", x)
    # reference quirk (hunyuan3ddit.py:390): timestep_embedding(t, 256, self.time_factor) passes time_factor
    # POSITIONALLY into max_period, so max_period = 1000 (not 10000) and time_factor keeps its default 1000.
    temb = timestep_embedding(t, 256, max_period=1000.0, time_factor=1000.0).to(latent.dtype)
    vec = _lin(sd, "time_in.out_layer", F.silu(_lin(sd, "time_in.in_layer", temb)))
    c = _lin(sd, "cond_in", cond)
    for i in range(depth):
        latent, c = double_block(sd, f"double_blocks.{i}.", latent, c, vec, heads)
        if taps is not None:
            taps.append(torch.cat((c, latent), 1))
    latent = torch.cat((c, latent), 1)
    for i in range(depth_single):
        latent = single_block(sd, f"single_blocks.{i}.", latent, vec, heads)
        if taps is not None:
            taps.append(latent)
    latent = latent[:, c.shape[1]:]
    shift, scale = _lin(sd, "final_layer.adaLN_modulation.1", F.silu(vec)).chunk(2, dim=1)
    latent = (1 + scale[:, None, :]) * _ln(latent) + shift[:, None, :]
    return _lin(sd, "final_layer.linear", latent)


# ----------------------------------------------------------------------------------------------- scheduler
def flow_euler_sigmas(num_inference_steps, shift=1.0):
    """pipelines.py:725-731 + schedulers.py:181-221: sigmas = linspace(0,1,N) (float64 -> float32),
    timesteps = sigmas*1000, sigmas extended by a trailing 1.0."""
    s = np.linspace(0, 1, num_inference_steps)
    s = shift * s / (1 + (shift - 1) * s)
    sig = torch.from_numpy(s).to(torch.float32)
    timesteps = sig * 1000
    return timesteps, torch.cat([sig, torch.ones(1)])


def flow_euler_step(sample, model_output, sigma, sigma_next):
    """schedulers.py:300-309."""
    prev = sample.to(torch.float32) + (sigma_next - sigma) * model_output
    return prev.to(model_output.dtype)


def denoise_loop(sd, latents, cond2, heads, depth, depth_single, steps, guidance, t_dtype=None):
    """pipelines.py:741-759 with classifier-free guidance; cond2 = cat(cond, uncond).
    t_dtype=torch.float16 quantises the timestep the way the fp16 pipeline does (`t.to(latents.dtype) / 1000`,
    pipelines.py:747-749) while the rest of the oracle stays fp32."""
    timesteps, sigmas = flow_euler_sigmas(steps)
    for i, t in enumerate(timesteps):
        x2 = torch.cat([latents] * 2)
        ts = t.expand(x2.shape[0]).to(t_dtype or latents.dtype) / 1000
        ts = ts.to(latents.dtype).to(latents.device)
        v = dit_forward(sd, x2, ts, cond2, heads, depth, depth_single)
        vc, vu = v.chunk(2)
        v = vu + guidance * (vc - vu)
        latents = flow_euler_step(latents, v, sigmas[i], sigmas[i + 1])
    return latents


# ----------------------------------------------------------------------------------------------- ShapeVAE
def _ln_aff(sd, name, x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def vae_resblock(sd, pre, x, heads):
    """ResidualAttentionBlock, models/autoencoders/attention_blocks.py:366-433 (qk_norm=True, c_qkv laid out
    (H, (q,k,v), D), :319-322)."""
    B, L, W = x.shape
    qkv = _lin(sd, pre + "attn.c_qkv", _ln_aff(sd, pre + "ln_1", x)).view(B, L, heads, -1)
    D = W // heads
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    if pre + "attn.attention.q_norm.weight" in sd:
        q = _ln_aff(sd, pre + "attn.attention.q_norm", q)
        k = _ln_aff(sd, pre + "attn.attention.k_norm", k)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    o = o.transpose(1, 2).reshape(B, L, W)
    x = x + _lin(sd, pre + "attn.c_proj", o)
    h = F.gelu(_lin(sd, pre + "mlp.c_fc", _ln_aff(sd, pre + "ln_2", x)))
    return x + _lin(sd, pre + "mlp.c_proj", h)


def vae_forward(sd, latents, heads, layers, taps=None):
    """ShapeVAE.forward, models/autoencoders/model.py:279-282."""
    x = _lin(sd, "post_kl", latents)
    for i in range(layers):
        x = vae_resblock(sd, f"transformer.resblocks.{i}.", x, heads)
        if taps is not None:
            taps.append(x)
    return x


def fourier_embed(x, num_freqs, include_pi):
    """FourierEmbedder.forward, attention_blocks.py:113-131 (include_input=True, logspace=True)."""
    freqs = 2.0 ** torch.arange(num_freqs, dtype=torch.float32, device=x.device)
    if include_pi:
        freqs = freqs * math.pi
    freqs = freqs.to(x.dtype)
    emb = (x[..., None].contiguous() * freqs).view(*x.shape[:-1], -1)
    return torch.cat((x, emb.sin(), emb.cos()), dim=-1)


def geo_decoder(sd, queries, latents, heads, num_freqs=8, include_pi=False, pre="geo_decoder."):
    """CrossAttentionDecoder.forward, attention_blocks.py:484-494, with ResidualCrossAttentionBlock :296-299 and
    MultiheadCrossAttention :250-261 (c_kv laid out (H, (k,v), D), :206-209)."""
    B, n, _ = queries.shape
    x = _lin(sd, pre + "query_proj", fourier_embed(queries, num_freqs, include_pi).to(latents.dtype))
    ca = pre + "cross_attn_decoder."
    W = x.shape[-1]
    D = W // heads
    q = _lin(sd, ca + "attn.c_q", _ln_aff(sd, ca + "ln_1", x)).view(B, n, heads, D)
    kv = _lin(sd, ca + "attn.c_kv", _ln_aff(sd, ca + "ln_2", latents)).view(B, latents.shape[1], heads, 2 * D)
    k, v = kv[..., :D], kv[..., D:]
    if ca + "attn.attention.q_norm.weight" in sd:
        q = _ln_aff(sd, ca + "attn.attention.q_norm", q)
        k = _ln_aff(sd, ca + "attn.attention.k_norm", k)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    o = o.transpose(1, 2).reshape(B, n, W)
    x = x + _lin(sd, ca + "attn.c_proj", o)
    x = x + _lin(sd, ca + "mlp.c_proj", F.gelu(_lin(sd, ca + "mlp.c_fc", _ln_aff(sd, ca + "ln_3", x))))
    if pre + "ln_post.weight" in sd:
        x = _ln_aff(sd, pre + "ln_post", x, eps=1e-5)
    return _lin(sd, pre + "output_proj", x)


def dense_grid_points(bounds, R):
    """generate_dense_grid_points, models/autoencoders/volume_decoders.py:122-138 -> [(R+1)^3, 3] float32,
    x slowest / z fastest."""
    lo, hi = np.array(bounds[0:3]), np.array(bounds[3:6])
    ax = [np.linspace(lo[i], hi[i], int(R) + 1, dtype=np.float32) for i in range(3)]
    xs, ys, zs = np.meshgrid(*ax, indexing="ij")
    return np.stack((xs, ys, zs), axis=-1).reshape(-1, 3)


def vanilla_volume_decode(sd, latents, heads, R, bounds=1.01, num_chunks=10000, num_freqs=8, include_pi=False,
                          query_dtype=torch.float16):
    """VanillaVolumeDecoder.__call__, volume_decoders.py:141-182.  The queries are quantised to the
    pipeline dtype (fp16) before the embedding (:168); everything after runs in latents.dtype."""
    if isinstance(bounds, float):
        bounds = [-bounds] * 3 + [bounds] * 3
    xyz = torch.from_numpy(dense_grid_points(bounds, R)).to(query_dtype).to(latents.dtype).to(latents.device)
    outs = []
    for s in range(0, xyz.shape[0], num_chunks):
        q = xyz[s:s + num_chunks][None].expand(latents.shape[0], -1, -1)
        outs.append(geo_decoder(sd, q, latents, heads, num_freqs, include_pi))
    g = torch.cat(outs, dim=1)
    return g.view(latents.shape[0], R + 1, R + 1, R + 1).float()


# ----------------------------------------------------------------------------------------------- VGGT geometry
def unproject_depth_map_to_point_map(depth, extrinsic, intrinsic):
    """vggt/vggt/utils/geometry.py:15-117, numpy, same dtype flow (float64 result)."""
    depth = np.asarray(depth)
    out = []
    for s in range(depth.shape[0]):
        d = depth[s]
        if d.ndim == 3:
            d = d.squeeze(-1)
        K, E = intrinsic[s], extrinsic[s]
        H, W = d.shape
        u, v = np.meshgrid(np.arange(W), np.arange(H))
        x = (u - K[0, 2]) * d / K[0, 0]
        y = (v - K[1, 2]) * d / K[1, 1]
        cam = np.stack((x, y, d), axis=-1).astype(np.float32)
        R, T = E[None, :3, :3], E[None, :3, 3:]
        Rt = np.transpose(R, (0, 2, 1))
        inv = np.tile(np.eye(4), (1, 1, 1))
        inv[:, :3, :3] = Rt
        inv[:, :3, 3:] = -np.matmul(Rt, T)
        out.append(np.dot(cam, inv[0, :3, :3].T) + inv[0, :3, 3])
    return np.stack(out, axis=0)
