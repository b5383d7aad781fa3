"""GPU: marching cubes through the C ABI, bit-exact against the CPU oracle (integer face indices, cell
classification, and float32 vertex positions)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _vols():
    from test_mc_oracle import sphere, torus, random_field
    yield "sphere65", sphere(65)
    yield "torus49", torus(49)
    for s in (0, 1, 2):
        v = random_field(33, s)
        yield f"rand33_{s}", v
    rng = np.random.default_rng(7)
    yield "noise_ragged", rng.normal(size=(7, 12, 9)).astype(np.float32)
    yield "noise20", rng.normal(size=(20, 20, 20)).astype(np.float32)
    yield "tiny2", np.array([[[1, -1], [-1, -1]], [[-1, -1], [-1, 1]]], np.float32)
    # one isolated cell per cube index x 6 log-normal magnitude draws: every face-test sub-case, and Lewiner's interior
    # test on both sides for 4.1.1 / 4.1.2 and 10.1.1 / 10.1.2 (tools/dump_skimage_goldens.py: case_atlas)
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "dump_goldens", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "dump_skimage_goldens.py"))
    dg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dg)
    yield "case_atlas", dg.case_atlas(0)
    yield "noise_closed", dg.white_noise((17, 18, 19), 3)


@pytest.mark.parametrize("name,vol", list(_vols()))
def test_mc_bit_exact(name, vol):
    import mc as omc
    from r3g import ops
    for level, bounds in ((0.0, None), (0.05, [-1.01] * 3 + [1.01] * 3)):
        ov, of, oc = omc.marching_cubes(vol, level, bounds=bounds, return_cases=True)
        g = torch.from_numpy(vol).cuda()
        v, f = ops.marching_cubes(g, level, bounds=bounds)
        assert f.dtype == torch.int32 and v.dtype == torch.float32
        assert np.array_equal(f.cpu().numpy(), of), f"{name}: face indices differ"
        assert np.array_equal(v.cpu().numpy(), ov), f"{name}: vertex positions differ"
        assert np.array_equal(ops.mc_classify(g, level).cpu().numpy(), oc), f"{name}: cell cases differ"


def test_mc_unaligned_grid_pointer():
    """A grid whose base is not 16-byte aligned takes the scalar-load variant of the sign-bit pass: same mesh."""
    import mc as omc
    from r3g import ops
    from test_mc_oracle import random_field
    vol = random_field(33, 5)
    ov, of = omc.marching_cubes(vol, 0.0)[:2]
    flat = torch.empty(vol.size + 1, device="cuda")
    g = flat[1:].view(vol.shape)
    g.copy_(torch.from_numpy(vol))
    assert g.data_ptr() % 16 != 0 and g.is_contiguous()
    v, f = ops.marching_cubes(g, 0.0)
    assert np.array_equal(f.cpu().numpy(), of) and np.array_equal(v.cpu().numpy(), ov)


def test_mc_errors_match_skimage_behaviour():
    from r3g import ops
    from test_mc_oracle import sphere
    g = torch.from_numpy(sphere(9)).cuda()
    with pytest.raises(ValueError):
        ops.marching_cubes(g, 10.0)
    with pytest.raises(RuntimeError):
        ops.marching_cubes(torch.zeros(5, 5, 5, device="cuda"), 0.0)


def test_mc_full_size_properties():
    """257^3 (the metric's size): closed oriented manifold, Euler characteristic, sequential-order property."""
    from r3g import ops
    from test_mc_oracle import check_closed_manifold
    n = 257
    ax = torch.linspace(-1.01, 1.01, n, device="cuda")
    x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = 0.25 - torch.sqrt((torch.sqrt(x * x + y * y) - 0.6) ** 2 + z * z)  # torus
    v, f = ops.marching_cubes(vol.contiguous(), 0.0)
    vn, fn = v.cpu().numpy(), f.cpu().numpy()
    assert check_closed_manifold(vn, fn) == 0
    # vertex ids appear in increasing order of first use (the sequential algorithm's property)
    first = np.full(len(vn), -1, np.int64)
    flat = fn.reshape(-1)
    idx = np.arange(len(flat))
    order = np.argsort(flat, kind="stable")
    first_use = np.minimum.reduceat(idx[order], np.r_[0, np.flatnonzero(np.diff(flat[order])) + 1])
    assert (np.diff(first_use) > 0).all()
    # re-run: deterministic
    v2, f2 = ops.marching_cubes(vol.contiguous(), 0.0)
    assert torch.equal(v, v2) and torch.equal(f, f2)
