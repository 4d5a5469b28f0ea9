"""World_size-2 `gloo` tests (CPU) of the multi-GPU host logic in unnatural_planets_b200/sharded.py: slab ranges, vertex-base
rebasing with two's-complement relative ids, variable-length gathers, row-tile reassembly.  The per-slab COMPUTE is a GPU
call (tests/test_gpu_mesh.py checks it stitches bit-exactly on one device); here the slab results are cut out of a golden
mesh the reference produced, so the collectives and the bookkeeping are checked end to end without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "grid_sphere_lakes_40.npz")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_slabs(world):
    """cut the golden whole mesh into `world` slab-like results: contiguous vertex blocks, contiguous face blocks whose ids are
    relative to the block's first vertex (faces placed so that some refer to the previous block: negative ids)"""
    g = np.load(GOLDEN)
    v, q = g["dmc_vertices"], g["dmc_quads"].astype(np.uint32)
    nv, nq = len(v), len(q)
    vcut = [nv * k // world for k in range(world + 1)]
    # a quad goes to the block that owns its LARGEST vertex id (the creator of its newest vertex comes last in scan order)
    owner = np.searchsorted(np.asarray(vcut[1:]), q.max(axis=1), side="right")
    parts = []
    for k in range(world):
        mine = q[owner == k]
        parts.append({
            "positions": v[vcut[k] : vcut[k + 1]] * np.float32(2.0),
            "normals": v[vcut[k] : vcut[k + 1]] * np.float32(0.5),
            "dmc_vertices": v[vcut[k] : vcut[k + 1]],
            "indices": (mine[:, :3].reshape(-1) - np.uint32(vcut[k])).astype(np.uint32),
            "dmc_quads": (mine - np.uint32(vcut[k])).astype(np.uint32),
            "near_zero_samples": k + 1,
        })
    whole_quads = np.concatenate([q[owner == k] for k in range(world)])
    return parts, v, whole_quads


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from unnatural_planets_b200 import sharded

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts, v, whole_quads = _fake_slabs(world)
        mesh = sharded.assemble_mesh(parts[rank], dst=0)
        # the tensor path (device-resident chunks on the GPU box; CPU tensors under gloo): same bookkeeping, padded gathers
        import torch

        tens = {k: (torch.from_numpy(a.view(np.int32).copy()) if a.dtype == np.uint32 else torch.from_numpy(a.copy())) if isinstance(a, np.ndarray) else a for k, a in parts[rank].items()}
        mesh2 = sharded.assemble_mesh_tensors(tens, dst=0, keys=[k for k, _ in sharded.MESH_KEYS])
        # several fields in one exchange (what ShardedPlanet.field_meshes does): a second "field" with the arrays scaled / shifted
        tens_b = {k: (t * 3 if k in ("positions", "normals", "dmc_vertices") else t) for k, t in tens.items()}
        both = sharded.assemble_meshes_tensors({"land": tens, "water": tens_b}, ["land", "water"], dst=0, keys=[k for k, _ in sharded.MESH_KEYS])
        # a mesh only rank 1 holds reaches every rank (ShardedPlanet.bake(src=...)): one payload broadcast
        held = [np.arange(21, dtype=np.float32).reshape(7, 3) * np.float32(0.5), np.arange(14, dtype=np.float32).reshape(7, 2), np.array([4000000000, 1, 2, 3, 4, 5], dtype=np.uint32)]
        got_b = sharded.broadcast_arrays(held if rank == 1 else None, src=1)
        assert len(got_b) == 3 and got_b[0].shape == (7, 3) and got_b[1].shape == (7, 2) and got_b[2].shape == (6,)
        assert (got_b[0].numpy() == held[0]).all() and (got_b[1].numpy() == held[1]).all() and (got_b[2].numpy().view(np.uint32) == held[2]).all()
        # row tiles of a 37-row image (ragged split)
        h, w = 37, 5
        img = [np.arange(h * w * c, dtype=np.uint32).reshape(h, w, c).astype(np.uint8) for c in (3, 2, 1)]
        r0, r1 = sharded.split_rows(h, rank, world)
        rows = sharded.assemble_rows(tuple(i[r0:r1] for i in img), dst=0)
        # an empty contribution from one rank
        lens = sharded.gather_varlen(np.zeros(0 if rank == 1 else 4, dtype=np.float32), dst=0)
        if rank == 0:
            ok = (mesh["dmc_vertices"] == v).all() and (mesh["dmc_quads"] == whole_quads).all() and (mesh["positions"] == v * np.float32(2)).all()
            ok = ok and (mesh["indices"] == whole_quads[:, :3].reshape(-1)).all() and mesh["near_zero_samples"] == sum(range(1, world + 1))
            ok = ok and all((mesh2[k] == mesh[k]).all() and mesh2[k].dtype == mesh[k].dtype and mesh2[k].shape == mesh[k].shape for k in ("positions", "normals", "dmc_vertices", "indices", "dmc_quads")) and mesh2["near_zero_samples"] == mesh["near_zero_samples"]
            ok = ok and all((both["land"][k] == mesh[k]).all() and both["land"][k].dtype == mesh[k].dtype and both["land"][k].shape == mesh[k].shape for k in ("positions", "normals", "dmc_vertices", "indices", "dmc_quads"))
            ok = ok and (both["water"]["positions"] == mesh["positions"] * np.float32(3)).all() and (both["water"]["dmc_quads"] == mesh["dmc_quads"]).all() and both["water"]["near_zero_samples"] == mesh["near_zero_samples"]
            ok = ok and all((a == b).all() for a, b in zip(rows, img)) and [len(x) for x in lens] == [4, 0][:world]
            negative = any((p["dmc_quads"].view(np.int32) < 0).any() for p in parts[1:])
            np.save(out_path, np.array([int(ok), int(negative)]))
        else:
            assert mesh is None and mesh2 is None and both is None and rows is None and lens is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_assembly(tmp_path):
    out = str(tmp_path / "result.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok, negative = np.load(out)
    assert ok == 1
    assert negative == 1  # the fixture does exercise ids that point into the previous slab


def test_split_rows_and_slab_ranges():
    sys.path.insert(0, ROOT)
    from unnatural_planets_b200 import sharded, slab_range

    for h, world in ((8192, 8), (37, 2), (5, 8), (4096, 3)):
        cuts = [sharded.split_rows(h, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == h and all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    # dualmc's z loop covers [0, res - 2): slab ranges tile it (upg_slab_range; pure host arithmetic, no GPU needed)
    for res, n in ((1024, 8), (500, 3), (40, 64)):
        rs = [slab_range(res, k, n) for k in range(n)]
        assert rs[0][0] == 0 and rs[-1][1] == res - 2 and all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
