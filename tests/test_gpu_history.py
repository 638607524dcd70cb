"""GPU tests of the planner-object history (pp_set_history; run with `pytest -m gpu`).

One reference HybridAStar object carries the `_visted` flags (until reset()) and the `_node_map` costs (for ever) of its 2D
heuristic from one find_path call to the next, stale against every map update in between -- that is how
src/local_planner.cpp drives it (find_path every tick, reset() only per waypoint).  With history enabled the device returns,
for every query of such a session, what the same call sequence returns on ONE unmodified reference object: expansion
sequence, cost, path, curvature, bit for bit.  Oracle: the compiled reference with pinned libm (oracle/_ref)."""
import numpy as np
import pytest

import orc
import scenarios as S
from test_gpu_parity import _bits, _ctx, _states_equal

pytestmark = pytest.mark.gpu


def _same(a, b):
    if not (a["success"] == b["success"] and a["n_pops"] == b["n_pops"] and a["cost"] == b["cost"]):
        return False
    return (_states_equal(a["pops"], b["pops"])[0] and np.array_equal(_bits(a["path"]), _bits(b["path"])) and
            np.array_equal(_bits(a["curvature"]), _bits(b["curvature"])))


@pytest.mark.parametrize("seed", [0, 3])
def test_session_on_one_planner_object_bitexact(seed):
    """11 queries: 6 ticks without reset, a bare reset, a second waypoint (update_goal relocates the non-empty map, reset)."""
    sc, ops = S.session_ops(seed, goal_changes=True)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, crm, fresh = _ctx(P), orc.ref(P), orc.ref(P)
    ctx.set_history(0, True)
    ra, rb = S.run_session(ctx, ops), S.run_session(crm, ops)
    rf = S.run_session(fresh, ops, fresh_each_query=lambda p: p.scrub())
    assert np.array_equal(_bits(ctx.get_map()), _bits(crm.get_map()))
    assert len(ra) == len(rb) == 11 and all(r["status"] == 0 for r in ra)
    usable = [k for k in range(len(rb)) if rb[k]["n_pops_bin_oob"] == 0]      # F7: undefined in the reference
    assert len(usable) >= 8
    bad = [k for k in usable if not _same(ra[k], rb[k])]
    assert not bad, [(k, ra[k]["n_pops"], rb[k]["n_pops"], float(ra[k]["cost"]), float(rb[k]["cost"])) for k in bad]
    # the history matters in this session: the reference itself answers differently on a fresh cache per query
    assert sum(1 for k in range(1, len(rb)) if rb[k]["n_pops"] != rf[k]["n_pops"]) >= 3


def test_history_off_and_batches_use_a_fresh_cache():
    """Without pp_set_history, and for every batch of more than one query, each query sees the freshly constructed cache."""
    sc, ops = S.session_ops(0, goal_changes=False, n_ticks=3)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, fresh = _ctx(P), orc.ref(P)
    ra = S.run_session(ctx, ops)
    rf = S.run_session(fresh, ops, fresh_each_query=lambda p: p.scrub())
    for a, b in zip(ra, rf):
        if b["n_pops_bin_oob"] == 0:
            assert _same(a, b)
    # history enabled, but a 2-query batch: both answers are the fresh-cache ones, and the carried cache is not touched
    ctx.set_history(0, True)
    last = [op for op in ops if op[0] == "query"][-1]
    q = ctx.make_queries([[last[2][0], last[2][1], last[2][2], float(last[1])]] * 2, [0, 0])
    res, _, _, _ = ctx.find_path_batch(q, ctx.make_opts(path_cap=4096))
    assert int(res[0]["n_pops"]) == int(res[1]["n_pops"]) == rf[-1]["n_pops"]
    one = ctx.find_path(float(last[1]), last[2])          # first single query of the carried cache = fresh state
    assert one["n_pops"] == rf[-1]["n_pops"] and one["cost"] == rf[-1]["cost"]


def test_capacity_retry_restarts_from_the_same_history():
    """A query that exhausts its pools is re-run with larger ones (the reference is unbounded); the aborted attempt must not
    leave a trace in the carried cache."""
    sc, ops = S.session_ops(0, goal_changes=False, n_ticks=5)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, crm = _ctx(P), orc.ref(P)
    ctx.set_history(0, True)
    rb = S.run_session(crm, ops)
    ra, retried = [], 0
    for op in ops:
        if op[0] == "query":
            ra.append(ctx.find_path(float(op[1]), op[2], max_expansions=256, max_open=512))
            retried += ctx.batch_retried()
        else:
            S.run_session(ctx, [op])
    assert retried >= 2, "the tiny pools were meant to overflow"
    for k, (a, b) in enumerate(zip(ra, rb)):
        if b["n_pops_bin_oob"] == 0:
            assert a["status"] == 0 and _same(a, b), (k, a["n_pops"], b["n_pops"])
