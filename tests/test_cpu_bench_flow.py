"""Dry run of bench.py's whole default flow on the CPU against a stub device context: every branch of main() -- EXACT timing,
end-to-end pass, K-POP blocks, single-query latencies, footprint side block, reference baseline (the real compiled reference on a
tiny sample) and the JSON line with every key the bench contract names -- executes without a GPU, so that an edit of bench.py
cannot break the round-end measurement with a NameError / TypeError.  The numbers the stub returns are meaningless; only the
control flow and the shape of the JSON line are checked here."""
import ctypes as C
import json
import sys
import types

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="needs the compiled reference for the cpu_baseline leg")


def _fake_pp(real):
    RESULT_DT = real._cabi.RESULT_DT

    class FakeLib:
        def __init__(self, ctx):
            self.ctx = ctx

        def pp_batch_upload(self, h, q, n, opts):
            self.ctx._n = n.value
            return 0

        def pp_batch_fetch(self, h, res, paths, curv, trace):
            r = self.ctx._results(self.ctx._n)
            C.memmove(res.value, r.ctypes.data, r.nbytes)
            return 0

        def pp_timer_begin(self, h):
            return 0

        def pp_timer_end(self, h, ms):
            ms._obj.value = 2.5
            return 0

        def pp_last_error(self):
            return b""

    class FakeContext:
        make_queries = staticmethod(real.Context.make_queries)
        make_opts = staticmethod(real.Context.make_opts)

        def __init__(self, params, num_groups=1, device=0):
            self.params, self.num_groups = params, num_groups
            self.N, self.h, self.lib, self._n, self.launches = params.grid_size, C.c_void_p(1), FakeLib(self), 0, 0

        def _chk(self, rc):
            assert rc == 0

        def _results(self, n):
            r = np.zeros(n, RESULT_DT)
            r["success"] = 1; r["cost"] = 10.0; r["n_pops"] = 100 + np.arange(n) % 7; r["n_path"] = 4
            return r

        def consts(self):
            return types.SimpleNamespace(log_threshold=0.85)

        def create_lane(self): return FakeContext(self.params, self.num_groups)
        def set_memory_budget(self, n): pass
        def close(self): pass
        def update_goal(self, *a, **k): pass
        def update_boxes(self, *a, **k): pass
        def update_boxes_2d(self, *a, **k): pass
        def update_boxes_2d_decay(self, *a, **k): pass
        def update_apf(self, *a, **k): pass
        def decay(self, *a, **k): pass
        def sync(self): pass
        def kernel_launches(self): self.launches += 1; return self.launches
        def batch_retried(self): return 0
        def get_map(self, g=0): return np.zeros((self.N, self.N), np.float32)
        def field2d(self, group=0, download=True): return None, 12, 0.7
        def field3d(self, group=0, use_h2d=True, download=True): return None, 3.9

        def set_start(self, q):
            return {"ci": np.full(len(q), 10, np.int32), "cj": np.full(len(q), 10, np.int32)}

        def batch_upload(self, q, opts=None): self._n = len(q)
        def batch_run(self): return 1.5
        def batch_run_async(self): pass
        def batch_wait(self): return 1.5

        def batch_fetch(self, want_paths=False):
            return self._results(self._n), None, None

        def find_path_batch(self, q, opts=None, want_paths=True):
            n = len(q)
            return self._results(n), np.zeros((n, 2048, 3), np.float32), np.zeros((n, 2048), np.float32), None

    fake = types.ModuleType("path_planning_pkg_b200")
    fake.Context, fake.PPError, fake.make_params, fake._cabi = FakeContext, real.PPError, real.make_params, real._cabi
    return fake


def test_default_bench_flow_prints_a_complete_line(monkeypatch, capsys):
    import torch
    import path_planning_pkg_b200 as real
    import bench

    real_tensor = torch.tensor
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (100 << 30, 180 << 30))
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda *a: types.SimpleNamespace(multi_processor_count=148))
    monkeypatch.setattr(torch, "tensor", lambda data, dtype=None, device=None: real_tensor(data, dtype=dtype))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setitem(sys.modules, "path_planning_pkg_b200", _fake_pp(real))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--groups", "2", "--starts", "4", "--steps", "3", "--warmup", "1", "--cpu-sample", "4",
                                      "--lanes", "2", "--c5-groups", "2", "--c5-steps", "1"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    out = [l for l in capsys.readouterr().out.split("\n") if l.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "kpop", "c5", "c1",
                "c2_map_update", "c3_fields", "batch_latency_ms", "retried_queries", "queries_with_bin_oob"):
        assert key in line, key
    assert line["steps"] == 3 and line["config"]["workload"].startswith("C4")
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] > 0
    for blk in ("cpu_baseline", "kpop", "c5", "c2_map_update", "c3_fields"):
        assert "error" not in line[blk], (blk, line[blk])
    assert line["cpu_baseline"]["kind"] == "reference" and "identical_to_gpu" in line["cpu_baseline"]
    assert "cost_vs_reference" in line["kpop"] and line["c5"]["scaling"] == "strong"


def test_reference_arm_prints_its_line(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--groups", "1", "--starts", "3", "--steps", "1", "--warmup", "1"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    line = json.loads([l for l in capsys.readouterr().out.split("\n") if l.startswith("{")][0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "busy_time" in line
