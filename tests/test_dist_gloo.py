"""CPU, world_size 2 over gloo: the N > 1 path's host logic (object sharding + variable-length mesh gather)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from r3g.dist import gather_meshes, shard_indices
    items = list(range(7))
    mine = shard_indices(len(items))
    meshes = []
    for i in mine:  # mesh i has i+1 vertices and 2i+1 faces with recognisable content
        g = torch.Generator().manual_seed(i)
        meshes.append((torch.rand(i + 1, 3, generator=g), torch.randint(0, i + 1, (2 * i + 1, 3), generator=g,
                                                                        dtype=torch.int32)))
    got = gather_meshes(meshes)
    if rank == 0:
        # plain numpy through the queue: torch tensors travel as shared-memory file descriptors that the parent has to
        # fetch from THIS process, which may already have exited (ConnectionResetError in a slow parent)
        q.put((mine, [(v.numpy().copy(), f.numpy().copy()) for v, f in got]))
    else:
        q.put((mine, len(got)))
    # streaming gather: one fixed-capacity message per rank and step, ring of depth 2, uneven object counts
    from r3g.dist import MeshStreamGatherer
    sg = MeshStreamGatherer(cap_vertices=8 + 8 * rank, cap_faces=32 - 8 * rank, device="cpu")   # ranks agree on the max
    for step in range(4):
        if step < len(meshes):
            sg.submit(*meshes[step])
        else:
            sg.submit(None, None)
    streamed = sg.finish()
    if rank == 0:
        assert len(streamed) == 7
        for (v, f), (gv, gf) in zip(streamed, got):
            assert torch.equal(v, gv) and torch.equal(f, gf)
    else:
        assert streamed == []
    # batch gather (bench.py at N > 1): staging buffer per rank, one message per peer at the end, pinned-ring landing
    from r3g.dist import MeshBatchGatherer
    bg = MeshBatchGatherer(cap_vertices=8 + 8 * rank, cap_faces=32 - 8 * rank, steps=4, device="cpu")
    for step in range(4):
        bg.submit(*(meshes[step] if step < len(meshes) else (None, None)))
    landed = []
    bufs = bg.finish(to_host=True, sink=lambda k, r, v, f: landed.append((r, k, v.clone(), f.clone())))
    if rank == 0:
        assert len(bufs) == world and len(landed) == 7
        for (r, k, v, f), (gv, gf) in zip(sorted(landed, key=lambda t: (t[0], t[1])), got):
            assert torch.equal(v, gv) and torch.equal(f, gf)
    else:
        assert bufs == [] and landed == []
    # a rank with nothing to send must not deadlock the gather
    got2 = gather_meshes(meshes if rank == 0 else [])
    if rank == 0:
        assert len(got2) == len(meshes)
    dist.barrier()
    dist.destroy_process_group()


def _run_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=120) for _ in range(2)]
    except Exception:
        res = None
    for p in procs:
        p.join(timeout=120)
    return res if res is not None and all(p.exitcode == 0 for p in procs) else None


def test_shard_and_gather_world2():
    # the TCP rendezvous on a just-released port can be reset on a loaded box: one retry on a fresh port
    res = _run_world2() or _run_world2()
    assert res is not None, "world-size-2 gloo run failed twice"
    r0 = [r for r in res if isinstance(r[1], list)][0]
    r1 = [r for r in res if not isinstance(r[1], list)][0]
    assert r0[0] == [0, 2, 4, 6] and r1[0] == [1, 3, 5] and r1[1] == 0
    order = [0, 2, 4, 6, 1, 3, 5]  # (rank, local index)
    assert len(r0[1]) == 7
    for i, (v, f) in zip(order, r0[1]):
        g = torch.Generator().manual_seed(i)
        ev = torch.rand(i + 1, 3, generator=g)
        ef = torch.randint(0, i + 1, (2 * i + 1, 3), generator=g, dtype=torch.int32)
        assert torch.equal(torch.from_numpy(v), ev) and torch.equal(torch.from_numpy(f), ef)


def test_stream_gatherer_single_process():
    sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
    from r3g.dist import MeshStreamGatherer
    seen = []
    sg = MeshStreamGatherer(8, 8, device="cpu", sink=lambda step, r, v, f: seen.append((step, r, v.shape[0], f.shape[0])))
    for i in range(5):
        sg.submit(torch.rand(i + 1, 3), torch.zeros(i, 3, dtype=torch.int32))
    assert sg.finish() == [] and seen == [(i, 0, i + 1, i) for i in range(5)]
    import pytest
    with pytest.raises(ValueError):
        MeshStreamGatherer(2, 2, device="cpu").submit(torch.rand(3, 3), torch.zeros(1, 3, dtype=torch.int32))


def test_single_process_passthrough():
    sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
    from r3g.dist import gather_meshes, shard_indices
    assert shard_indices(5, 1, 3) == [1, 4]
    m = [(torch.rand(4, 3), torch.zeros(2, 3, dtype=torch.int32))]
    assert gather_meshes(m)[0][0] is m[0][0]


def test_shard_indices_partition_property():
    """Every object index is owned by exactly one rank, in increasing order per rank, for any (n, world)."""
    sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
    from hypothesis import given, settings, strategies as st
    from r3g.dist import shard_indices

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 200), st.integers(1, 16))
    def check(n, world):
        seen = []
        for r in range(world):
            mine = shard_indices(n, r, world)
            assert mine == sorted(mine) and all(i % world == r for i in mine)
            seen += mine
        assert sorted(seen) == list(range(n))
    check()
