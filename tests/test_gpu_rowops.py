"""GPU: row-wise operators through the C ABI against the oracle / golden fixtures."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_layernorm_plain_affine_and_modulated():
    from r3g import ops
    torch.manual_seed(0)
    for width in (128, 1024, 2048):
        x = (torch.randn(2 * 77, width, device="cuda") * 2 + 0.3).half()
        w = (1 + 0.1 * torch.randn(width, device="cuda")).half()
        b = (0.1 * torch.randn(width, device="cuda")).half()
        y = ops.layernorm(x, w, b, eps=1e-6)
        ref = F.layer_norm(x.float(), (width,), w.float(), b.float(), 1e-6)
        assert (y.float() - ref).abs().max().item() < 4e-3
        sc = (0.2 * torch.randn(2, width, device="cuda")).half()
        sh = (0.2 * torch.randn(2, width, device="cuda")).half()
        y = ops.layernorm(x, eps=1e-6, scale=sc, shift=sh, rows_per_batch=77)
        ref = (1 + sc.float()[:, None]) * F.layer_norm(x.float(), (width,), eps=1e-6).view(2, 77, width) + sh.float()[:, None]
        assert (y.float().view(2, 77, width) - ref).abs().max().item() < 8e-3


def test_layernorm_strided_rows():
    from r3g import ops
    x = torch.randn(50, 512, device="cuda").half()
    view = x[:, 128:384]
    y = ops.layernorm(view, eps=1e-5, out=torch.empty(50, 256, device="cuda", dtype=torch.float16))
    assert (y.float() - F.layer_norm(view.float(), (256,), eps=1e-5)).abs().max().item() < 4e-3


def test_qk_norm_rms_and_layernorm_modes():
    import hy3d_ref as R
    from r3g import ops
    torch.manual_seed(1)
    rows, H = 123, 4
    buf = torch.randn(rows, 3 * H * 64, device="cuda").half()
    qs = (1 + 0.2 * torch.randn(64, device="cuda")).half()
    ks = (1 + 0.2 * torch.randn(64, device="cuda")).half()
    ref = buf.float().view(rows, 3, H, 64).clone()
    ref[:, 0] = R._rms(ref[:, 0], qs.float())
    ref[:, 1] = R._rms(ref[:, 1], ks.float())
    out = ops.qk_norm_(buf.clone(), H, 0, H * 64, 64, 0, 1e-6, qs, None, ks, None)
    assert (out.float().view(rows, 3, H, 64) - ref).abs().max().item() < 6e-3
    # mode 1: LayerNorm with bias, (H, (q,k,v), D) layout of the VAE
    qb = (0.1 * torch.randn(64, device="cuda")).half()
    kb = (0.1 * torch.randn(64, device="cuda")).half()
    ref = buf.float().view(rows, H, 3, 64).clone()
    ref[:, :, 0] = F.layer_norm(ref[:, :, 0], (64,), qs.float(), qb.float(), 1e-6)
    ref[:, :, 1] = F.layer_norm(ref[:, :, 1], (64,), ks.float(), kb.float(), 1e-6)
    out = ops.qk_norm_(buf.clone(), H, 0, 64, 192, 1, 1e-6, qs, qb, ks, kb)
    assert (out.float().view(rows, H, 3, 64) - ref).abs().max().item() < 6e-3
    # q only (geo-decoder query side)
    out = ops.qk_norm_(buf.clone(), H, 0, 0, 192, 1, 1e-6, qs, qb, None, None)
    assert torch.equal(out.view(rows, H, 3, 64)[:, :, 1:], buf.view(rows, H, 3, 64)[:, :, 1:])


def test_gemv_and_timestep_embedding(golden_dir):
    from r3g import ops
    torch.manual_seed(2)
    w = (torch.randn(6144, 1024, device="cuda") * 0.03).half()
    b = (torch.randn(6144, device="cuda") * 0.1).half()
    v = torch.randn(2, 1024, device="cuda").half()
    out = ops.gemv(w, b, v, silu_in=True)
    ref = F.linear(F.silu(v.float()), w.float(), b.float())
    assert (out.float() - ref).abs().max().item() < 6e-3
    z = np.load(os.path.join(golden_dir, "dit_mini.npz"))
    t = torch.from_numpy(z["temb_t"]).cuda()
    emb = ops.timestep_embedding(t, 256, 1000.0, 1000.0)
    # fixture produced by the reference function in fp16; device sin/cos may differ by one fp16 ulp
    diff = (emb.float().cpu() - torch.from_numpy(z["temb_out"]).float()).abs().max().item()
    assert diff <= 1e-3


def test_cfg_euler_step_matches_reference_dtype_flow(golden_dir):
    import hy3d_ref as R
    from r3g import ops
    torch.manual_seed(3)
    x = torch.randn(1, 3072, 64).half()
    v = torch.randn(2, 3072, 64).half()
    g, (sig0, sig1) = 5.0, (torch.tensor(0.5102, dtype=torch.float32), torch.tensor(0.5306, dtype=torch.float32))
    vc, vu = v.chunk(2)
    mix = vu + g * (vc - vu)
    ref = R.flow_euler_step(x, mix, sig0, sig1)
    # torch's CPU kernels round the 0-dim fp32 (sigma_next - sigma) to fp16 before the product, its CUDA
    # kernels keep it in fp32 (opmath).  The kernel takes d_sigma as given, so both flows are reproducible:
    dup = torch.empty(2, 3072, 64, device="cuda", dtype=torch.float16)
    xd = x.cuda().clone()
    ops.cfg_euler_step_(xd, v.cuda(), g, float((sig1 - sig0).half()), x_dup=dup)
    assert torch.equal(xd.cpu(), ref), "CPU dtype flow (the oracle): must be bit-exact"
    assert torch.equal(dup[0], xd[0]) and torch.equal(dup[1], xd[0])
    # what torch's own CUDA kernels do with the same expression (the reference's actual execution): equal to one of
    # the two d_sigma roundings -- the pipeline uses the fp16-rounded one, which must be the matching one
    vcc, vuc = v.cuda().chunk(2)
    ref_cuda = R.flow_euler_step(x.cuda(), vuc + g * (vcc - vuc), sig0.cuda(), sig1.cuda())
    assert torch.equal(xd, ref_cuda), "torch CUDA flow differs from the CPU flow: revisit the pipeline's d_sigma"


def test_grid_fourier_against_reference_fixture(golden_dir):
    from r3g import ops
    z = np.load(os.path.join(golden_dir, "vae_mini.npz"))
    Rr = int(z["cfg_R"])
    n = (Rr + 1) ** 3
    out = torch.full((n, 64), 7.0, device="cuda", dtype=torch.float16)
    ops.grid_fourier(out, 0, n, Rr, [-1.01] * 3 + [1.01] * 3, 8, 0)
    got = out.float().cpu().numpy()
    # coordinates: bit-exact fp16 quantisation of np.linspace (volume_decoders.py:168)
    assert np.array_equal(got[:, :3], z["xyz"].astype(np.float16).astype(np.float32))
    # features: the fixture is fp32 sin/cos of fp16 products; ours is rounded to fp16
    assert np.abs(got[:, :51] - z["fourier"]).max() < 1e-3
    assert (got[:, 51:] == 0).all()
    # sub-range
    part = torch.empty(100, 64, device="cuda", dtype=torch.float16)
    ops.grid_fourier(part, 300, 100, Rr, [-1.01] * 3 + [1.01] * 3, 8, 0)
    assert torch.equal(part, out[300:400])


def test_unproject_against_reference_fixture(golden_dir):
    from r3g import ops
    z = np.load(os.path.join(golden_dir, "unproject.npz"))
    d = torch.from_numpy(z["depth"][..., 0]).cuda()
    pts = ops.unproject(d, z["extrinsic"], z["intrinsic"], torch.float64)
    np.testing.assert_allclose(pts.cpu().numpy(), z["points"], rtol=0, atol=1e-12)
    pts32 = ops.unproject(d, z["extrinsic"], z["intrinsic"], torch.float32)
    np.testing.assert_allclose(pts32.cpu().numpy(), z["points"].astype(np.float32), rtol=0, atol=1e-6)
    # full-size property: z_cam is exact, so inverting the rigid transform recovers depth
    S, H, W = 2, 518, 518
    depth = 0.5 + torch.rand(S, H, W, device="cuda")
    E = np.tile(np.eye(3, 4, dtype=np.float32), (S, 1, 1)); E[:, :, 3] = [0.1, -0.2, 0.3]
    K = np.tile(np.array([[400, 0, 259], [0, 410, 259], [0, 0, 1]], np.float32), (S, 1, 1))
    p = ops.unproject(depth, E, K, torch.float64)
    back = p[..., 2] + 0.3
    assert (back - depth.double()).abs().max().item() < 1e-6
