"""Host-side logic that needs no GPU: parameters/status mapping, C-ABI surface, sharding."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parameters_and_status_mapping():
    from sqpsolver_jl_b200.host.parameters import Parameters, moi_termination_status

    p = Parameters(max_iter=7)
    assert p.get_parameter("max_iter") == 7 and p.algorithm == "SQP-TR"
    with pytest.raises(KeyError):
        p.set_parameter("nope", 1)
    # MOI_wrapper.jl:1238-1278
    assert moi_termination_status(0) == "LOCALLY_SOLVED" and moi_termination_status(6) == "LOCALLY_SOLVED"
    assert moi_termination_status(2) == "LOCALLY_INFEASIBLE" and moi_termination_status(-1) == "ITERATION_LIMIT"
    assert moi_termination_status(4) == "NORM_LIMIT" and moi_termination_status(-5) == "MEMORY_LIMIT"  # quirk: else arm


def test_library_exports_every_declared_symbol(built_lib):
    from sqpsolver_jl_b200 import capi

    hdr = open(os.path.join(ROOT, "include", "sqpqp.h")).read()
    declared = set(re.findall(r"\b(sqpqp_[a-z_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in sqpqp.h but not exported"
    assert declared == set(capi.EXPORTS)


def test_default_options_roundtrip_without_gpu(built_lib):
    from sqpsolver_jl_b200 import capi

    o = capi.Options()
    built_lib.sqpqp_default_options(C.byref(o))
    assert o.rho0 == 0.1 and o.alpha == 1.6 and o.sigma == 1e-6 and o.check_every == 25
    assert o.warm_start == 1 and o.team == 0
    # struct layout agreement between the header and the ctypes mirror: last field intact
    assert o.threads == 0 and o.smem_kb == -1 and o.polish_cg_max > 0 and o.method == 0 and o.ipm_eps == 1e-9 and o.ipm_kappa_eps == 10.0


def test_create_fails_loudly_without_cuda(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sqpsolver_jl_b200 import capi

    with pytest.raises(capi.SqpQpError):
        capi.Engine(0)


def test_shard_ranges_cover_batch():
    from sqpsolver_jl_b200.host.batch import shard_range

    for B, W in ((1024, 8), (1024, 3), (5, 8), (1, 1)):
        seen = []
        for r in range(W):
            lo, hi = shard_range(B, r, W)
            seen += list(range(lo, hi))
        assert seen == list(range(B))
    assert shard_range(1024, 3, 8) == (384, 512)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gather_results_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from sqpsolver_jl_b200.host.batch import shard_range, pack_results, gather_results
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        B = 7
        lo, hi = shard_range(B, r, w)
        ids = np.arange(lo, hi)
        rec = gather_results(pack_results(ids % 3, ids + 10, ids * 1.5), B)
        assert rec.shape[0] == B, rec.shape
        assert np.array_equal(rec["iters"], np.arange(B) + 10) and np.array_equal(rec["objective"], np.arange(B) * 1.5)
        assert np.array_equal(rec["status"], np.arange(B) % 3)
        dist.destroy_process_group()
        print("rank", r, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", env["MASTER_PORT"], str(script)], capture_output=True, text=True,
                         env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_grouped_driver_partition_merge_and_errors(monkeypatch):
    """GroupedBatchSqpTR without a GPU: the sub-driver is replaced by a stub that fills its arrays from the instance ids it was
    given.  Checks the partition (contiguous, unequal sizes), that every group gets ITS slice of the per-instance NLP data and of
    x0, the hand-off default for groups, the merge order, batch-wide ids in the log, summed timers / stats, and that an exception
    in one group's thread reaches the caller."""
    from sqpsolver_jl_b200.host import sqp_trust_region as tr

    class Nlp:
        def __init__(self, ids):
            self.ids = np.asarray(ids)

        def subset(self, lo, hi):
            return Nlp(self.ids[lo:hi])

    made = []

    class Stub:
        def __init__(self, nlp, batch, params=None, device=0, engine_options=None, x0=None, device_evaluator=False, layout=None):
            assert len(nlp.ids) == batch
            self.ids, self.B, self.layout, self.x0 = nlp.ids, batch, layout, x0
            self.options = params
            self.mixed_phases = True
            made.append(self)

        def run(self, log=None):
            if 13 in self.ids and getattr(Stub, "fail", False):
                raise RuntimeError("group with instance 13 failed")
            ids = self.ids.astype(float)
            self.x = np.stack([ids, ids + 0.5], axis=1); self.lam = self.x.copy()
            self.mult_x_L = self.x.copy(); self.mult_x_U = self.x.copy(); self.mult_g = -self.lam
            self.obj_val = 10.0 * ids; self.f = self.obj_val.copy()
            self.status = (self.ids % 3).astype(np.int64); self.ret = self.status.copy()
            self.n_qp = self.ids + 1; self.iter = self.ids + 2
            self.prim_infeas = ids * 0.0; self.dual_infeas = ids * 0.0
            self.rounds = int(self.ids.max())
            self.timers = {"callbacks": 1.0, "device": 2.0}
            self.optimizer = type("O", (), {"stats": {"solve_ms": 3.0, "solves": 1}})()
            if log is not None:
                for k in range(self.B):
                    log.append({"b": k, "id": int(self.ids[k])})
            return self

        def close(self):
            self.closed = True

    monkeypatch.setattr(tr, "BatchSqpTR", Stub)
    B = 14
    x0 = np.arange(B * 2, dtype=float).reshape(B, 2)
    g = tr.GroupedBatchSqpTR(Nlp(np.arange(B)), B, params="P", groups=3, x0=x0, layout={"tail": 64})
    assert g.bounds == [(0, 4), (4, 9), (9, 14)]
    assert [list(s.ids) for s in made] == [list(range(0, 4)), list(range(4, 9)), list(range(9, 14))]
    assert all(s.layout == {"tail": 64, "handoff": 0} for s in made)          # groups: no hand-off to the resident launch
    assert all(np.array_equal(s.x0, x0[lo:hi]) for s, (lo, hi) in zip(made, g.bounds))
    log = []
    g.run(log)
    assert np.array_equal(g.x[:, 0], np.arange(B)) and np.array_equal(g.obj_val, 10.0 * np.arange(B))
    assert np.array_equal(g.status, np.arange(B) % 3) and np.array_equal(g.n_qp, np.arange(B) + 1)
    assert g.rounds == B - 1 and g.timers == {"callbacks": 3.0, "device": 6.0} and g.stats["solve_ms"] == 9.0
    assert sorted(e["b"] for e in log) == list(range(B)) and all(e["b"] == e["id"] for e in log)
    g.close()
    assert all(s.closed for s in made)
    # one group (or more groups than instances) degenerates gracefully; a shared 1-D x0 is passed through
    made.clear()
    g1 = tr.GroupedBatchSqpTR(Nlp(np.arange(3)), 3, groups=8, x0=np.zeros(2))
    assert g1.bounds == [(0, 1), (1, 2), (2, 3)] and all(s.x0.shape == (2,) for s in made)
    made.clear()
    g2 = tr.GroupedBatchSqpTR(Nlp(np.arange(5)), 5, groups=1, layout=None)
    assert g2.bounds == [(0, 5)] and made[0].layout is None                      # lock step: the engine's own launch rule
    Stub.fail = True
    try:
        with pytest.raises(RuntimeError, match="instance 13"):
            tr.GroupedBatchSqpTR(Nlp(np.arange(B)), B, groups=3).run()
    finally:
        Stub.fail = False
