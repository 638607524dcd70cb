"""Worker of tests/test_gpu_multirank.py (one process per GPU, launched by torch.distributed.run): rank 0 rasterises the maps of a
few C4 groups, pp_broadcast_maps replicates them into every rank's context over NCCL, every rank answers its shard of the queries;
the replicas must be bit-equal to rank 0's maps and the gathered results equal to the single-rank run of the same batch."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import scenarios as S  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import path_planning_pkg_b200 as pp
    from path_planning_pkg_b200.shard import gather_records, shard_queries

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cpu_group = dist.new_group(backend="gloo")
    G, n_starts = 3, 8
    groups = [S.c4_group(g, n_starts=n_starts) for g in range(G)]
    P = pp.make_params(grid_size=512, resolution=0.2)
    ctx = pp.Context(P, num_groups=G, device=local_rank)
    for gi, sc in enumerate(groups):
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        if rank == 0:
            for _ in range(sc["rounds"]):
                ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
                ctx.decay(group=gi)
        else:
            ctx.update_apf(sc["boxes"], S.APF_ADDED_RADIUS, group=gi)      # frames + APF lists only: the map arrives by broadcast
    if rank != 0:
        assert not ctx.get_map(0).any()
    ctx.field2d(group=0, download=False)                                  # derived state that the broadcast must invalidate
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    ctx.broadcast_maps(0, G, 0)
    ctx.sync()
    maps = np.stack([ctx.get_map(g) for g in range(G)])
    # bit-equal replicas: compare every rank's maps with rank 0's
    t = torch.from_numpy(maps.view(np.int32).copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same_maps = bool(torch.equal(t, ref))
    # queries: selected from the (now replicated) maps, sharded by query
    thr = ctx.consts().log_threshold
    queries, qgroups = [], []
    for gi, sc in enumerate(groups):
        cand = sc["start_candidates"]
        st = ctx.set_start(ctx.make_queries(cand, [gi] * len(cand)))
        sel = cand[maps[gi][st["ci"], st["cj"]] < thr][:n_starts]
        queries.append(sel); qgroups += [gi] * len(sel)
    queries = np.concatenate(queries); qgroups = np.array(qgroups, np.int32)
    mine = shard_queries(len(queries), rank, world)
    out = {"rank": rank, "same_maps": same_maps}
    rec_dt = np.dtype([("q", "i4"), ("success", "i4"), ("n_pops", "i4"), ("cost", "u4"), ("hash", "u8")])
    for mode, tag in ((0, "exact"), (1, "kpop")):
        res, paths, curv, _ = ctx.find_path_batch(ctx.make_queries(queries[mine], qgroups[mine]), ctx.make_opts(path_cap=2048, mode=mode, kpop=32))
        assert (res["status"] == 0).all()
        rec = np.zeros(len(mine), rec_dt)
        rec["q"] = mine; rec["success"] = res["success"]; rec["n_pops"] = res["n_pops"]; rec["cost"] = res["cost"].view(np.uint32)
        import orc
        rec["hash"] = [orc.path_hash(paths[k, :res["n_path"][k]], curv[k, :res["n_path"][k]]) if res["success"][k] else 0 for k in range(len(mine))]
        allrec = gather_records(rec, world, group=cpu_group)
        if rank == 0:
            allrec = allrec[np.argsort(allrec["q"])]
            full, fp, fc, _ = ctx.find_path_batch(ctx.make_queries(queries, qgroups), ctx.make_opts(path_cap=2048, mode=mode, kpop=32))
            ok = (np.array_equal(allrec["q"], np.arange(len(queries))) and np.array_equal(allrec["success"], full["success"])
                  and np.array_equal(allrec["n_pops"], full["n_pops"]) and np.array_equal(allrec["cost"], full["cost"].view(np.uint32))
                  and all(int(allrec["hash"][k]) == (orc.path_hash(fp[k, :full["n_path"][k]], fc[k, :full["n_path"][k]]) if full["success"][k] else 0)
                          for k in range(len(queries))))
            out[tag + "_sharded_equals_single_rank"] = bool(ok)
            out[tag + "_expansions"] = int(full["n_pops"].sum())
    flags = [None] * world
    dist.all_gather_object(flags, out, group=cpu_group)
    if rank == 0:
        print("MULTIRANK " + json.dumps(flags), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
