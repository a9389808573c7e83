"""Pin the oracle: the reference's own known answers + the public case9 optimum.

  * test/runtests.jl:12-14     toy: X ~ -1, Y ~ -1 (rtol 1e-4), LOCALLY_SOLVED
  * README.md:17-20            1-variable toy: x = -1
  * SURVEY section 4 [derived] first toy QP infeasible -> feasibility-restoration LP with optimum 1, p1 = -2
  * public MATPOWER case9 polar-ACOPF optimum 5296.69 $/h (validates evaluator + driver)
and the committed golden fixtures (tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest

from oracle.sqp_tr import Parameters, SqpTROracle
from sqpsolver_jl_b200.host.parameters import moi_termination_status
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9
from sqpsolver_jl_b200.nlp.toy import ReadmeToy, ToyExample

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_toy_example_known_answer():
    trace, log = [], []
    s = SqpTROracle(ToyExample(), Parameters(max_iter=100), trace=trace).run(log)
    assert np.allclose(s.x, [-1.0, -1.0], rtol=1e-4)
    assert s.status == 0 and moi_termination_status(s.status) == "LOCALLY_SOLVED"
    # derived trace: QP 1 infeasible, QP 2 is the FR LP: p1 = -2, sum of slacks = 1
    assert trace[0]["status"] == "INFEASIBLE" and not trace[0]["fr"]
    assert trace[1]["fr"] and trace[1]["status"] == "LOCALLY_SOLVED"
    assert abs(trace[1]["p"][0] + 2.0) < 1e-8


def test_readme_toy_known_answer():
    s = SqpTROracle(ReadmeToy(), Parameters(max_iter=100)).run()
    assert s.status == 0
    assert np.allclose(s.x, [-1.0], rtol=1e-6)
    assert abs(s.obj_val) < 1e-8  # f(-1) = 0


def test_case9_public_optimum():
    s = SqpTROracle(AcopfPolar(case9()), Parameters(max_iter=100, init_mu=1e4)).run()
    assert s.status == 0
    assert abs(s.obj_val - 5296.69) < 0.01  # MATPOWER case9 ACOPF optimum
    g = np.load(os.path.join(GOLD, "case9_mu1e4.npz"))
    assert abs(s.obj_val - float(g["obj"])) <= 1e-9 * abs(float(g["obj"]))
    assert s.iter == int(g["iters"])


@pytest.mark.parametrize("name,cls", [("toy_example", ToyExample), ("readme_toy", ReadmeToy)])
def test_golden_fixture_matches(name, cls):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    s = SqpTROracle(cls(), Parameters(max_iter=100)).run()
    assert s.status == int(g["status"]) and s.iter == int(g["iters"])
    assert np.allclose(s.x, g["x"], rtol=0, atol=1e-9)


def test_default_parameters_and_constants():
    # parameters.jl:17-29 and sqp_trust_region.jl:66-71,100-101
    p = Parameters()
    assert (p.tol_direction, p.tol_residual, p.tol_infeas) == (1e-8, 1e-8, 1e-8)
    assert p.max_iter == 3000 and p.init_mu == 1.0 and p.tr_size == 10.0 and p.use_soc is False
    s = SqpTROracle(ReadmeToy(), p)
    assert s.Delta_max == 1e8 and s.phi == 1e20
