"""Multi-GPU plumbing: one process per GPU, objects sharded round-robin, finished meshes gathered to rank 0.

The reference's only parallelism on this path is data-parallel over objects (src/2d_to_3d_models/run.py:176-194:
task i -> GPU i % num_devices through a process pool, results meet on disk).  Here rank r takes objects
{i : i % world == r} of the sorted list and rank 0 -- which runs the downstream scene assembly
(src/scene_reconstruction/run.py:62-76 reads <out>/3D/<name>/<name>.glb) -- receives every mesh over
NCCL (NVLink / NVSwitch): one all_gather of the (V, F) counts, then one send/recv of a packed payload per rank.
There is no compute kernel to fuse this with: the payload is a few MB per object, after the last kernel.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank=None, world=None):
    """Indices of the items this rank processes (round-robin, like task i -> GPU i % num_devices)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    return list(range(rank, n_items, world))


def _pack(meshes, device):
    parts = []
    for v, f in meshes:
        parts.append(v.reshape(-1).contiguous().view(torch.int32))
        parts.append(f.reshape(-1).to(torch.int32).contiguous())
    if not parts:
        return torch.empty(0, dtype=torch.int32, device=device)
    return torch.cat(parts)


def gather_meshes(meshes, to_host=False, max_objects=None):
    """meshes: list of (verts float32 [V,3], faces int32 [F,3]) tensors on this rank's device.
    Returns, on rank 0, the list of all ranks' meshes ordered by (rank, local index); [] elsewhere.
    Works on NCCL (CUDA tensors) and on gloo (CPU tensors; used by the world_size-2 CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [(v.cpu(), f.cpu()) if to_host else (v, f) for v, f in meshes]
    world, rank = dist.get_world_size(), dist.get_rank()
    device = meshes[0][0].device if meshes else torch.device("cuda", torch.cuda.current_device()) \
        if dist.get_backend() == "nccl" else torch.device("cpu")
    n_local = torch.tensor([len(meshes)], dtype=torch.int64, device=device)
    n_all = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(n_all, n_local)
    n_max = max(int(t.item()) for t in n_all) if max_objects is None else max_objects
    counts = torch.zeros(max(n_max, 1), 2, dtype=torch.int64, device=device)
    for i, (v, f) in enumerate(meshes):
        counts[i, 0], counts[i, 1] = v.shape[0], f.shape[0]
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    payload = _pack(meshes, device)
    out = []
    if rank == 0:
        bufs, ops_ = {}, []
        for r in range(1, world):
            n_int = int((all_counts[r][:, 0].sum() + all_counts[r][:, 1].sum()).item()) * 3
            bufs[r] = torch.empty(n_int, dtype=torch.int32, device=device)
            if n_int:
                ops_.append(dist.P2POp(dist.irecv, bufs[r], r))
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
        bufs[0] = payload
        for r in range(world):
            off = 0
            for i in range(int(n_all[r].item())):
                nv, nf = int(all_counts[r][i, 0]), int(all_counts[r][i, 1])
                v = bufs[r][off:off + 3 * nv].view(torch.float32).view(nv, 3)
                off += 3 * nv
                f = bufs[r][off:off + 3 * nf].view(nf, 3)
                off += 3 * nf
                out.append((v.cpu(), f.cpu()) if to_host else (v, f))
    else:
        if payload.numel():
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload, 0)]):
                w.wait()
    return out
