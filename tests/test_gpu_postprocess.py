"""GPU: FloaterRemover on the device (union-find components) against scipy's connected components on the same mesh."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene():
    """A big sphere, a torus and three tiny floaters (marching cubes of an analytic field), as numpy arrays."""
    import mc as omc
    n = 97
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    field = 0.6 - np.sqrt(x * x + y * y + z * z)
    # a 6 % component that stays and two specks under MeshLab's 0.5 % rule (faces ~ r^2: (0.035 / 0.6)^2 = 0.34 %)
    for cx, cy, cz, r in ((0.8, 0.8, 0.8, 0.15), (-0.85, 0.8, -0.8, 0.035), (0.85, -0.85, 0.2, 0.03)):
        field = np.maximum(field, r - np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2))
    return omc.marching_cubes(field.astype(np.float32), 0.0)


def test_components_match_scipy_and_floaters_are_removed():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from r3g import ops
    from r3g.postprocessors import DegenerateFaceRemover, FaceReducer, FloaterRemover
    from r3g.vae import Latent2MeshOutput
    v, f = _scene()
    labels = ops.mesh_components(torch.from_numpy(f).cuda(), len(v)).cpu().numpy()
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]]])
    ncomp, ref = connected_components(coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(len(v), len(v))), directed=False)
    assert ncomp == 4, ncomp
    # same partition, and the label is the smallest vertex index of the component
    for c in range(ncomp):
        members = np.flatnonzero(ref == c)
        assert (labels[members] == members.min()).all()
    out = FloaterRemover()(Latent2MeshOutput(mesh_v=v, mesh_f=f))
    sizes = np.bincount(ref[f[:, 0]])
    kept = sizes >= 0.005 * sizes.max()
    assert kept.sum() < ncomp, "the scene is meant to contain components under the 0.5 % threshold"
    assert len(out.mesh_f) == sizes[kept].sum() and len(out.mesh_v) == np.isin(ref, np.flatnonzero(kept)).sum()
    # compaction keeps order and geometry: every kept face maps to the same three points
    keep_face = kept[ref[f[:, 0]]]
    assert np.array_equal(out.mesh_v[out.mesh_f], v[f[keep_face]])
    # tensors stay on the device
    vt, ft = FloaterRemover()((torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()))
    assert vt.is_cuda and np.array_equal(ft.cpu().numpy(), out.mesh_f)
    assert DegenerateFaceRemover()(out) is out and FaceReducer()(out, max_facenum=10 ** 9) is out
    with pytest.raises(NotImplementedError):
        FaceReducer()(out, max_facenum=10)
    with pytest.raises(Exception):
        ops.mesh_components(torch.tensor([[0, 1, 99]], dtype=torch.int32).cuda(), 3)


def test_components_on_a_large_noise_mesh():
    """4 M-vertex marching-cubes output of a noise field (the bench's mesh shape): partition equals scipy's."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from r3g import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    grid = torch.randn(129, 129, 129, device="cuda", generator=g)
    v, f = ops.marching_cubes(grid, 0.0)
    labels = ops.mesh_components(f, v.shape[0]).cpu().numpy()
    fn = f.cpu().numpy()
    e = np.concatenate([fn[:, [0, 1]], fn[:, [1, 2]]])
    n = v.shape[0]
    ncomp, ref = connected_components(coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n)), directed=False)
    first = np.full(ncomp, n, dtype=np.int64)
    np.minimum.at(first, ref, np.arange(n))
    assert np.array_equal(labels, first[ref])
