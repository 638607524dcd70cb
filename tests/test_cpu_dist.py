"""World-size-2 gloo test of the N>1 path's host logic (no GPU): group sharding, map broadcast and the result
gather.  The per-rank search here runs on the CPU oracle (test infrastructure standing in for the kernel), so
what is checked is the plumbing: union of the shards == the serial run, maps replicated bit for bit."""
import os
import socket

import numpy as np
import pytest

import orc
import scenarios as S

REC_DT = np.dtype([("group", "i4"), ("query", "i4"), ("success", "i4"), ("n_pops", "i4"), ("cost", "f4")])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run_group(g):
    sc = S.c1_scenario(g)
    o = orc.port(orc.make_params(grid_size=100, resolution=0.3))
    sc = dict(sc); sc["goal"] = np.array([18.0, 3.0, 0.2], np.float32)
    S.build_map(o, sc)
    r = o.find_path(3.0, [0, 0, 0])
    return o.get_map(), np.array([(g, 0, int(r["success"]), r["n_pops"], r["cost"])], REC_DT)


def _worker(rank, world, port, n_groups, out_dir):
    import torch
    import torch.distributed as dist
    from path_planning_pkg_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_groups(n_groups, rank, world)
    recs, maps = [], {}
    for g in mine:
        m, r = _run_group(g)
        maps[g] = m; recs.append(r)
    # map replication: the owner broadcasts, everyone ends with identical bytes
    crcs = []
    for g in range(n_groups):
        t = torch.from_numpy(maps[g].copy()) if g in maps else torch.zeros(100, 100)
        shard.broadcast_map(t, shard.owner_of(g, world))
        crcs.append(int(np.frombuffer(t.numpy().tobytes(), np.uint32).sum(dtype=np.uint64)))
    allrec = shard.gather_records(np.concatenate(recs) if recs else np.zeros(0, REC_DT), world)
    np.save(os.path.join(out_dir, f"rec_{rank}.npy"), allrec)
    np.save(os.path.join(out_dir, f"crc_{rank}.npy"), np.array(crcs, np.uint64))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_broadcast_gather_world2(tmp_path, built):
    import torch.multiprocessing as mp
    n_groups, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_groups, str(tmp_path)), nprocs=world, join=True)
    serial = np.concatenate([_run_group(g)[1] for g in range(n_groups)])
    for rank in range(world):
        rec = np.load(tmp_path / f"rec_{rank}.npy")
        rec = np.sort(rec, order="group")
        assert np.array_equal(rec, serial)
    assert np.array_equal(np.load(tmp_path / "crc_0.npy"), np.load(tmp_path / "crc_1.npy"))


def test_shard_groups_partition():
    from path_planning_pkg_b200 import shard
    for world in (1, 2, 4, 8):
        seen = sorted(g for r in range(world) for g in shard.shard_groups(64, r, world))
        assert seen == list(range(64))
        sizes = [len(shard.shard_groups(64, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_shard_queries_partition():
    from path_planning_pkg_b200 import shard
    import numpy as np
    for world in (1, 2, 3, 8):
        parts = [shard.shard_queries(65536 + 5, r, world) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(65536 + 5))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
