"""Regenerate tests/golden/*.npz with the CPU oracle (run from the repo root).

The reference itself (Julia + Ipopt) cannot run in the build container, so these vectors
are produced by the oracle restatement, which in turn is pinned against the reference's own
known answers (tests/test_oracle_golden.py: toy (-1,-1) LOCALLY_SOLVED, README toy -1, and
the public MATPOWER case9 optimum 5296.69 $/h).  Stored per case: the final SQP answer and
the full list of QP subproblems of the trajectory (inputs + the oracle's solution), so the
GPU parity tests can replay every subproblem through the C-ABI.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sqp_tr import Parameters, SqpTROracle  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import case9, synth_net  # noqa: E402
from sqpsolver_jl_b200.nlp.toy import ReadmeToy, ToyExample  # noqa: E402

CASES = {
    "toy_example": (ToyExample, dict(max_iter=100)),
    "readme_toy": (ReadmeToy, dict(max_iter=100)),
    "case9_mu1e4": (lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4)),
    "case9_default": (lambda: AcopfPolar(case9()), dict(max_iter=60)),
    # BASELINE configs[2]: the case118-shaped synthetic network, SQP trust-region variant -- the first subproblems only
    # (the as-coded algorithm does not converge on it within 100 iterations; a fixture of the whole run would be 20 MB)
    "case118_first": (lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=6, init_mu=1e5)),
}
STATUS_CODE = {"LOCALLY_SOLVED": 4, "INFEASIBLE": 2, "LOCALLY_INFEASIBLE": 5, "ITERATION_LIMIT": 11, "NUMERICAL_ERROR": 20}

if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, (mk, kw) in CASES.items():
        if only and name not in only:
            continue
        trace = []
        s = SqpTROracle(mk(), Parameters(**kw), trace=trace).run()
        keys = ("x", "dE", "h_val", "df", "E", "p", "lambda_qp", "mult_x_U", "mult_x_L", "lam")
        data = {f"qp_{k}": np.stack([t[k] for t in trace]) for k in keys}
        data["qp_Delta"] = np.array([t["Delta"] for t in trace])
        data["qp_fr"] = np.array([t["fr"] for t in trace])
        data["qp_iter"] = np.array([t["iter"] for t in trace])
        data["qp_status"] = np.array([STATUS_CODE[t["status"]] for t in trace])
        np.savez_compressed(os.path.join(out, name + ".npz"), x=s.x, obj=s.obj_val, status=s.status, iters=s.iter,
                            n_qp=s.n_qp, lam=s.lam, **data)
        print(name, "status", s.status, "obj", repr(s.obj_val), "iters", s.iter, "qps", len(trace))
