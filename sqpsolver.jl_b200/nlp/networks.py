"""Power-network data for the ACOPF workloads.

* :func:`case9` -- the public MATPOWER/WSCC 9-bus case (re-entered from public
  knowledge; the reference's ``examples/data/case9.m`` is git-ignored and absent,
  see examples/acopf/opf.jl:84 and .gitignore:2).  Known polar-ACOPF optimum
  ~5296.69 $/h.
* :func:`synth_net` -- seeded synthetic networks for the BASELINE.json shapes
  "case118-shaped" (118 bus / 186 branch / 54 gen) and "~2000-bus"
  (2000 / 3000 / 400): spanning ring + random chords.

All quantities are stored in per-unit on ``baseMVA`` as PowerModels does after
``make_per_unit!``; angles in radians.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Network:
    name: str
    baseMVA: float
    # buses
    nbus: int
    pd: np.ndarray
    qd: np.ndarray
    gs: np.ndarray
    bs: np.ndarray
    vmin: np.ndarray
    vmax: np.ndarray
    ref_bus: int  # 0-based
    # generators
    gen_bus: np.ndarray  # 0-based
    pmin: np.ndarray
    pmax: np.ndarray
    qmin: np.ndarray
    qmax: np.ndarray
    cost2: np.ndarray  # $/h per pu^2
    cost1: np.ndarray  # $/h per pu
    cost0: np.ndarray
    # branches
    f_bus: np.ndarray  # 0-based
    t_bus: np.ndarray
    br_r: np.ndarray
    br_x: np.ndarray
    br_b: np.ndarray
    rate_a: np.ndarray
    tap: np.ndarray
    shift: np.ndarray
    angmin: np.ndarray
    angmax: np.ndarray
    meta: dict = field(default_factory=dict)

    @property
    def ngen(self):
        return int(self.gen_bus.shape[0])

    @property
    def nbranch(self):
        return int(self.f_bus.shape[0])

    def perturbed_loads(self, batch: int, rel_sigma: float = 0.05, seed: int = 1234):
        """Per-instance loads for the batched workload (BASELINE.json configs[4]).

        Instance ``b`` uses ``pd*(1+rel_sigma*N(0,1))`` drawn from a Philox
        counter-based generator keyed by ``seed`` with the instance id as the
        stream (``Philox(key=seed, counter=[0,0,0,b])``), so any rank can
        generate exactly its own shard without touching the others'.
        Returns ``(pd[batch,nbus], qd[batch,nbus])``.
        """
        pd = np.empty((batch, self.nbus))
        qd = np.empty((batch, self.nbus))
        for b in range(batch):
            rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, b]))
            pd[b] = self.pd * (1.0 + rel_sigma * rng.standard_normal(self.nbus))
            qd[b] = self.qd * (1.0 + rel_sigma * rng.standard_normal(self.nbus))
        return pd, qd


def case9() -> Network:
    base = 100.0
    bus_pd = np.array([0, 0, 0, 0, 90, 0, 100, 0, 125], float) / base
    bus_qd = np.array([0, 0, 0, 0, 30, 0, 35, 0, 50], float) / base
    nbus = 9
    gen_bus = np.array([1, 2, 3]) - 1
    pmax = np.array([250, 300, 270], float) / base
    pmin = np.array([10, 10, 10], float) / base
    qmax = np.array([300, 300, 300], float) / base
    qmin = -qmax
    # gencost model 2, 3 coefficients ($/MW^2 h, $/MW h, $/h)
    c2 = np.array([0.11, 0.085, 0.1225])
    c1 = np.array([5.0, 1.2, 1.0])
    c0 = np.array([150.0, 600.0, 335.0])
    br = np.array(
        [
            # f, t, r, x, b, rateA
            [1, 4, 0.0, 0.0576, 0.0, 250],
            [4, 5, 0.017, 0.092, 0.158, 250],
            [5, 6, 0.039, 0.17, 0.358, 150],
            [3, 6, 0.0, 0.0586, 0.0, 300],
            [6, 7, 0.0119, 0.1008, 0.209, 150],
            [7, 8, 0.0085, 0.072, 0.149, 250],
            [8, 2, 0.0, 0.0625, 0.0, 250],
            [8, 9, 0.032, 0.161, 0.306, 250],
            [9, 4, 0.01, 0.085, 0.176, 250],
        ]
    )
    nbr = br.shape[0]
    # PowerModels clamps the +-360 degree MATPOWER defaults to +-60 degrees.
    ang = np.deg2rad(60.0)
    return Network(
        name="case9",
        baseMVA=base,
        nbus=nbus,
        pd=bus_pd,
        qd=bus_qd,
        gs=np.zeros(nbus),
        bs=np.zeros(nbus),
        vmin=np.full(nbus, 0.9),
        vmax=np.full(nbus, 1.1),
        ref_bus=0,
        gen_bus=gen_bus,
        pmin=pmin,
        pmax=pmax,
        qmin=qmin,
        qmax=qmax,
        cost2=c2 * base * base,
        cost1=c1 * base,
        cost0=c0,
        f_bus=br[:, 0].astype(int) - 1,
        t_bus=br[:, 1].astype(int) - 1,
        br_r=br[:, 2].copy(),
        br_x=br[:, 3].copy(),
        br_b=br[:, 4].copy(),
        rate_a=br[:, 5] / base,
        tap=np.ones(nbr),
        shift=np.zeros(nbr),
        angmin=np.full(nbr, -ang),
        angmax=np.full(nbr, ang),
    )


def _dc_flows(nbus, f, t, x, inj, ref):
    """DC power flow: returns branch flows (pu) for bus injections ``inj``."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    nbr = f.shape[0]
    bsus = 1.0 / x
    A = sp.coo_matrix(
        (np.concatenate([np.ones(nbr), -np.ones(nbr)]), (np.tile(np.arange(nbr), 2), np.concatenate([f, t]))),
        shape=(nbr, nbus),
    ).tocsr()
    B = (A.T @ sp.diags(bsus) @ A).tocsc()
    keep = np.setdiff1d(np.arange(nbus), [ref])
    theta = np.zeros(nbus)
    theta[keep] = spla.spsolve(B[keep][:, keep], inj[keep])
    return bsus * (A @ theta)


def synth_net(nbus: int, nbranch: int, ngen: int, seed: int, load_scale: float = 1.0) -> Network:
    """Seeded synthetic network: spanning ring + random chords (degree <= 9).

    r~U(0.005,0.05)*ls, x~U(0.03,0.3)*ls, b~U(0,0.1) where ``ls`` shortens chord
    impedances a little so the meshed network stays well inside the +-30 degree
    angle-difference limit; loads pd~U(0,1.5)*load_scale pu on 80% of the
    buses, qd=0.3 pd; 5% of buses carry a small shunt (so the balance rows are
    genuinely nonlinear); quadratic costs c2~U(0.01,0.1), c1~U(10,40) $/MW;
    vm in [0.94,1.06]; rate_a = max(1.5 |DC flow|, 0.5) pu.
    """
    assert nbranch >= nbus and ngen <= nbus
    rng = np.random.default_rng(seed)
    base = 100.0
    f = list(range(nbus))
    t = [(i + 1) % nbus for i in range(nbus)]
    deg = np.full(nbus, 2)
    have = set(zip(f, t))
    # chords: mostly local (short electrical distance) with a few long links
    while len(f) < nbranch:
        a = int(rng.integers(nbus))
        if rng.random() < 0.8:
            hop = int(rng.integers(2, max(3, min(12, nbus // 2))))
            b = (a + hop) % nbus
        else:
            b = int(rng.integers(nbus))
        if a == b or deg[a] >= 9 or deg[b] >= 9:
            continue
        if (a, b) in have or (b, a) in have:
            continue
        have.add((a, b))
        f.append(a)
        t.append(b)
        deg[a] += 1
        deg[b] += 1
    f = np.array(f)
    t = np.array(t)
    r = rng.uniform(0.005, 0.05, nbranch)
    x = rng.uniform(0.03, 0.3, nbranch)
    bc = rng.uniform(0.0, 0.1, nbranch)

    pd = np.where(rng.random(nbus) < 0.8, rng.uniform(0.0, 1.5, nbus), 0.0) * load_scale
    qd = 0.3 * pd
    gs = np.where(rng.random(nbus) < 0.05, rng.uniform(0.0, 0.05, nbus), 0.0)
    bs = np.where(gs > 0, rng.uniform(0.0, 0.2, nbus), 0.0)

    # generators spread evenly round the ring (plus jitter), one per bus
    gen_bus = np.unique((np.arange(ngen) * nbus // ngen + rng.integers(0, max(1, nbus // ngen), ngen)) % nbus)
    while gen_bus.shape[0] < ngen:
        extra = rng.integers(nbus)
        gen_bus = np.unique(np.append(gen_bus, extra))
    share = pd.sum() / ngen
    pmax = share * rng.uniform(1.6, 2.6, ngen)
    pmin = 0.1 * pmax * (rng.random(ngen) < 0.5)
    qmax = 0.75 * pmax + 0.5
    qmin = -qmax
    c2 = rng.uniform(0.01, 0.1, ngen)
    c1 = rng.uniform(10.0, 40.0, ngen)
    c0 = np.zeros(ngen)

    # size thermal limits from a proportional DC dispatch
    inj = -pd.copy()
    np.add.at(inj, gen_bus, pd.sum() * pmax / pmax.sum())
    flows = _dc_flows(nbus, f, t, x, inj, 0)
    rate = np.maximum(1.5 * np.abs(flows), 0.5)
    ang = np.deg2rad(30.0)
    return Network(
        name=f"synth{nbus}",
        baseMVA=base,
        nbus=nbus,
        pd=pd,
        qd=qd,
        gs=gs,
        bs=bs,
        vmin=np.full(nbus, 0.94),
        vmax=np.full(nbus, 1.06),
        ref_bus=0,
        gen_bus=gen_bus.astype(int),
        pmin=pmin,
        pmax=pmax,
        qmin=qmin,
        qmax=qmax,
        cost2=c2 * base * base,
        cost1=c1 * base,
        cost0=c0,
        f_bus=f,
        t_bus=t,
        br_r=r,
        br_x=x,
        br_b=bc,
        rate_a=rate,
        tap=np.ones(nbranch),
        shift=np.zeros(nbranch),
        angmin=np.full(nbranch, -ang),
        angmax=np.full(nbranch, ang),
        meta={"seed": seed, "load_scale": load_scale, "max_dc_angle": float(np.max(np.abs(flows * x)))},
    )


def case118_shaped() -> Network:
    return synth_net(118, 186, 54, seed=118)


def case2000_shaped() -> Network:
    return synth_net(2000, 3000, 400, seed=2000)
