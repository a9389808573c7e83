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
