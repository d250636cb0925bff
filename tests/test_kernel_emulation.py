"""Host emulation of the CUDA kernel SOURCE against the oracle, the reference's dumps and analytic solutions (CPU; no GPU).

tests/emul/kernel_emulation.cpp compiles mallard_b200/csrc/kernels_impl.cuh - the file nvcc compiles for sm_100a - for the host
(STRICT arithmetic: -ffp-contract=off) and runs its kernels thread by thread.  Covered: teno_recon_kernel, the generic TENO
kernel, visc_grad_kernel, face_flux_kernel (QT = 0 / 1), gather_stage_kernel; not covered: the PTX streaming kernels and the CFL
kernel.  Two uses:
  * the kernels added after this round's GPU budget was spent (generic TENO for basis_order 5..9 / other stencil factors,
    quadrilaterals under TENO, viscous terms) are checked here against the oracle / analytically, since their GPU tests
    (tests/test_gpu_parity.py, MLB_RUN_UNVERIFIED=1) have not run on a B200 yet;
  * a second execution of kernels that HAVE run on hardware: bit-exact against the oracle on triangles.
This is test infrastructure; the product has no CPU path (mallard_b200 refuses to compute without a CUDA device)."""
import os
import sys

import numpy as np
import pytest

import golden_util as gu
import mallard_b200 as mb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
from emulation import EmulatedSolver  # noqa: E402

SYM4 = [dict(name=n, type="symmetry") for n in ("left", "right", "top", "bottom")]
EXTRAP4 = [dict(name=n, type="extrapolation") for n in ("left", "right", "top", "bottom")]
CONN_KEYS = ["node_coords", "offsets_nodes_of_cell", "nodes_of_cell", "offsets_faces_of_cell", "faces_of_cell", "offsets_nodes_of_face",
             "nodes_of_face", "cells_of_face"]


def _smooth(xy, rng):
    k = rng.uniform(0.5, 2.0, 4)
    rho = 1.0 + 0.3 * np.sin(2 * np.pi * k[0] * xy[:, 0]) * np.cos(2 * np.pi * k[1] * xy[:, 1])
    u = 0.4 + 0.2 * np.cos(2 * np.pi * k[2] * xy[:, 1]); v = -0.3 + 0.2 * np.sin(2 * np.pi * k[3] * xy[:, 0])
    p = 1.0 + 0.2 * np.cos(2 * np.pi * (xy[:, 0] - xy[:, 1]))
    e = p / (0.4 * rho)
    return np.stack([rho, rho * u, rho * v, rho * (e + 0.5 * (u * u + v * v))], 1)


@pytest.mark.parametrize("order,basis,fixed,riemann", [(3, "legendre", False, "HLLC"), (3, "legendre", True, "HLL"), (1, "legendre", True, "Rusanov"),
                                                       (4, "legendre", True, "HLLC"), (2, "monomial", True, "HLLC")])
def test_emulated_teno_kernels_are_bit_exact_against_the_oracle(oracle_mod, order, basis, fixed, riemann):
    """teno_recon_kernel + face_flux_kernel + gather_stage_kernel, compiled for the host: the oracle's face values and residual,
    bit for bit (legendre; monomial goes through pow: same libm here), and the generic kernel equals the specialised one."""
    mesh = mb.Mesh.generate("cartesian_tri", 16, 14, 2.0, 1.0)
    om = oracle_mod.Mesh.generate("cartesian_tri", 16, 14, 2.0, 1.0)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="p_out", p=0.9), dict(name="top", type="symmetry"),
           dict(name="bottom", type="wall_adiabatic")]
    kw = dict(recon="TENO", riemann=riemann, integrator="SSPRK3", bcs=bcs, basis=basis, order=order, teno_fixed=fixed)
    so, se = oracle_mod.Solver(om, **kw), EmulatedSolver(mesh, **kw)
    U0 = _smooth(mesh.arrays["cell_coords"], np.random.default_rng(11))
    so.set_state(U0); se.set_state(U0)
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    Fo, Fe = so.calc_face_values()[real][:, :, 0], se.calc_face_values()[real][:, :, 0]
    ro, re = so.calc_rhs(), se.calc_rhs()
    assert np.array_equal(Fe, Fo, equal_nan=True)
    assert np.array_equal(re, ro, equal_nan=True)
    se.force_generic(True)
    assert np.array_equal(se.calc_face_values()[real][:, :, 0], Fe, equal_nan=True) and np.array_equal(se.calc_rhs(), re, equal_nan=True)


@pytest.mark.parametrize("order,factor,qc,basis", [(5, 2.0, 5, "legendre"), (6, 2.0, 5, "legendre"), (7, 2.0, 5, "monomial"), (9, 2.0, 5, "legendre"),
                                                   (2, 1.5, 0, "legendre"), (3, 3.0, 0, "monomial"), (4, 2.5, 0, "legendre")])
def test_emulated_generic_kernel_orders_5_to_9_and_other_stencil_factors_vs_oracle(oracle_mod, order, factor, qc, basis):
    """Everything the reference's TOML accepts beyond the specialised kernels (face_reconstruction.cpp:110-116,170-180)."""
    nx, ny = (12, 10) if order >= 7 else (14, 12)          # 240 cells hold the 110-cell stencils of p = 9; the host QR dominates the run time
    om = oracle_mod.Mesh.generate("cartesian_tri", nx, ny, 2.0, 1.0)
    mesh = mb.Mesh.generate("cartesian_tri", nx, ny, 2.0, 1.0)
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", bcs=SYM4, basis=basis, order=order, factor=factor, quad_cell_order=qc, teno_fixed=True)
    so, se = oracle_mod.Solver(om, **kw), EmulatedSolver(mesh, **kw)
    U0 = _smooth(mesh.arrays["cell_coords"], np.random.default_rng(23))
    so.set_state(U0); se.set_state(U0)
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    Fo, Fe = so.calc_face_values()[real][:, :, 0], se.calc_face_values()[real][:, :, 0]
    assert np.array_equal(Fe, Fo, equal_nan=True), gu.rel_err(Fe, Fo)
    assert np.array_equal(se.calc_rhs(), so.calc_rhs(), equal_nan=True)


@pytest.mark.parametrize("name", ["teno_legendre_12x10_p5", "teno_legendre_8x7_p2_f15", "teno_smooth_6x6", "teno_monomial_7x6_p3",
                                  "teno_bcs_rk4_10x8", "teno_hll_riemann_9x7"])
def test_emulated_kernels_against_dumps_of_the_unmodified_reference(name):
    meta, g = gu.load(name)
    mm = meta["mesh"]
    mesh = mb.Mesh.generate(mm["type"], mm["Nx"], mm["Ny"], mm["Lx"], mm["Ly"])
    se = EmulatedSolver(mesh, **gu.solver_kwargs(meta))
    se.set_state(g["U0"])
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    interior = real & (mesh.arrays["cells_of_face"][:, 1] >= 0)
    F = se.calc_face_values()
    # 1/(SI+eps)^6 is three multiplications here and libm pow in the reference (last ulp of a weight that is then normalised away
    # in the smooth branch): element-wise 1e-12, bit-exact in practice on these cases
    assert gu.rel_err(F[real][:, :, 0], g["F_stage1"][real][:, :, 0]) <= 1e-12
    assert gu.rel_err(F[interior][:, :, 1], g["F_stage1"][interior][:, :, 1]) <= 1e-12
    assert gu.rel_err(se.calc_rhs(), g["rhs_stage1"]) <= 1e-12


# ---- quadrilaterals and mixed meshes under TENO: k-exactness (no oracle exists: the reference throws, face_reconstruction.cpp:485-487)
_D5 = (np.array([[1 / 3, 1 / 3], [0.059715871789770, 0.470142064105115], [0.470142064105115, 0.059715871789770], [0.470142064105115, 0.470142064105115],
                 [0.797426985353087, 0.101286507323456], [0.101286507323456, 0.797426985353087], [0.101286507323456, 0.101286507323456]]),
       np.array([0.225, 0.132394152788506, 0.132394152788506, 0.132394152788506, 0.125939180544827, 0.125939180544827, 0.125939180544827]))


def _cell_averages(mesh, f):
    A = mesh.arrays
    X, onc, noc = A["node_coords"], A["offsets_nodes_of_cell"].astype(np.int64), A["nodes_of_cell"].astype(np.int64)
    out, area = np.zeros((mesh.n_cells, 4)), np.zeros(mesh.n_cells)
    xy, w = _D5
    for c in range(mesh.n_cells):
        n = noc[onc[c]:onc[c + 1]]
        for t in range(len(n) - 2):
            v0, v1, v2 = X[n[0]], X[n[t + 1]], X[n[t + 2]]
            a = 0.5 * abs((v1[0] - v0[0]) * (v2[1] - v0[1]) - (v2[0] - v0[0]) * (v1[1] - v0[1]))
            p = v0[None, :] + xy[:, :1] * (v1 - v0)[None, :] + xy[:, 1:] * (v2 - v0)[None, :]
            out[c] += a * (w[:, None] * f(p[:, 0], p[:, 1])).sum(axis=0)
            area[c] += a
    return out / area[:, None]


@pytest.mark.parametrize("order,tri_fraction", [(1, 0.5), (2, 0.5), (3, 0.5), (3, 0.0), (2, 1.0), (4, 0.6)])
def test_emulated_teno_on_quadrilateral_and_mixed_meshes_is_k_exact(order, tri_fraction):
    """An order-p reconstruction must reproduce every polynomial of degree <= p from its cell averages at every face quadrature
    point, from both sides of every interior face - on quadrilaterals, triangles and any mix, jittered."""
    from mallard_b200 import synthetic as syn
    mesh = syn.mixed_tri_quad(12, 10, 3.0, 2.0, seed=3, tri_fraction=tri_fraction)
    nn = np.diff(mesh.arrays["offsets_nodes_of_cell"])
    assert (tri_fraction == 1.0 or (nn == 4).any()) and (tri_fraction == 0.0 or (nn == 3).any())
    coef = np.random.default_rng(17).uniform(-1.0, 1.0, size=(4, order + 1, order + 1))

    def f(x, y):
        out = np.zeros(x.shape + (4,))
        for v in range(4):
            for i in range(order + 1):
                for j in range(order + 1 - i):
                    out[..., v] += coef[v, i, j] * x ** i * y ** j
        out[..., 0] += 5.0
        return out
    se = EmulatedSolver(mesh, "TENO", "HLLC", "SSPRK3", order=order, bcs=EXTRAP4, teno_fixed=True)
    se.set_state(_cell_averages(mesh, f))
    F = se.calc_face_values()
    A = mesh.arrays
    nof, cof = A["nodes_of_face"].reshape(-1, 2).astype(np.int64), A["cells_of_face"]
    xi = {1: [0.0], 2: [-0.5773502691896257, 0.5773502691896257], 3: [-0.7745966692414834, 0.0, 0.7745966692414834]}[se.n_quad]
    x0, x1 = A["node_coords"][nof[:, 0]], A["node_coords"][nof[:, 1]]
    worst = 0.0
    for q, z in enumerate(xi):
        pq = (z + 1.0) * 0.5 * (x1 - x0) + x0
        exact = f(pq[:, 0], pq[:, 1])
        worst = max(worst, np.abs(F[:, q, 0] - exact).max(), np.abs(F[cof[:, 1] >= 0][:, q, 1] - exact[cof[:, 1] >= 0]).max())
    assert worst <= 2e-9 * (10.0 ** max(0, order - 3)), worst
    # free-stream preservation through the flux and gather kernels on the mixed mesh
    se.set_state(np.tile([1.2, 0.36, -0.24, 2.6], (mesh.n_cells, 1)))
    rhs = se.calc_rhs()
    interior = np.ones(mesh.n_cells, bool)
    interior[cof[cof[:, 1] < 0, 0]] = False
    assert np.abs(rhs[interior]).max() < 1e-9


def test_emulated_first_order_on_a_mixed_mesh_is_bit_exact_against_the_oracle(oracle_mod):
    from mallard_b200 import synthetic as syn
    mesh = syn.mixed_tri_quad(14, 10, 3.0, 2.0, seed=9, tri_fraction=0.4)
    om = oracle_mod.Mesh.from_arrays({k: mesh.arrays[k] for k in CONN_KEYS}, mesh.zones)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="p_out", p=0.9), dict(name="top", type="symmetry"),
           dict(name="bottom", type="wall_adiabatic")]
    so, se = oracle_mod.Solver(om, "FO", "HLLC", "SSPRK3", bcs=bcs), EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", bcs=bcs)
    U0 = _smooth(mesh.arrays["cell_coords"] / 3.0, np.random.default_rng(2))
    so.set_state(U0); se.set_state(U0)
    assert np.array_equal(se.calc_rhs(), so.calc_rhs())


# ---- viscous terms (new: the reference is Euler only): analytic checks ------------------------------------------------------
def _gas(mu):
    return dict(gamma=1.4, p_ref=101325.0, T_ref=298.15, rho_ref=1.225, p_min=-1e20, p_max=1e20, mu=mu, Pr=0.72)


def _state_from_prim(rho, u, v, T, R):
    cv = R / 0.4
    return np.stack([rho, rho * u, rho * v, rho * (cv * T + 0.5 * (u * u + v * v))], 1)


R_GAS = 101325.0 / (298.15 * 1.225)


@pytest.mark.parametrize("mtype", ["cartesian", "cartesian_tri", "mixed"])
def test_viscous_residual_of_couette_flow(mtype):
    """Plane Couette flow u = U y / H, v = 0, uniform p and T between a fixed and a moving no-slip wall.  The least-squares
    gradients are exact (du/dy = U / H, everything else 0) on quadrilaterals, triangles and jittered mixed meshes, wall cells
    included; the VISCOUS part of the residual, rhs(mu) - rhs(0), is then exactly (0, 0, 0, mu (U / H)^2): uniform shear stress,
    viscous heating.  On the wall-aligned quadrilateral mesh the inviscid part vanishes as well (HLLC resolves the tangential
    jump across faces parallel to the flow exactly), so the full residual is the analytic one."""
    from mallard_b200 import synthetic as syn
    mu, Uw, H, L = 0.05, 3.0, 1.0, 2.0
    mesh = syn.mixed_tri_quad(12, 10, L, H, seed=5, tri_fraction=0.5) if mtype == "mixed" else mb.Mesh.generate(mtype, 12, 10, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0]),
           dict(name="top", type="wall_noslip", u=[Uw, 0.0])]
    xy = mesh.arrays["cell_coords"]
    n = mesh.n_cells
    U0 = _state_from_prim(np.full(n, 1.2), Uw * xy[:, 1] / H, np.zeros(n), np.full(n, 300.0), R_GAS)
    res = []
    for m in (mu, 0.0):
        se = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(m), bcs=bcs)
        se.set_state(U0)
        if m > 0:
            G = se.gradients()
            assert np.abs(G[:, 1] - Uw / H).max() < 1e-10 and np.abs(G[:, [0, 2, 3, 4, 5]]).max() < 1e-9
        res.append(se.calc_rhs())
    heat = mu * (Uw / H) ** 2
    dv = res[0] - res[1]
    scale = np.abs(res[1]).max() + 1.0                              # the difference of two residuals carries their rounding
    assert np.abs(dv[:, :3]).max() < 1e-12 * scale + 1e-9 and np.abs(dv[:, 3] - heat).max() < 1e-12 * scale + 1e-8 * heat
    if mtype == "cartesian":
        rhs = res[0]
        assert np.abs(rhs[:, 0]).max() < 1e-9 and np.abs(rhs[:, 1]).max() < 1e-9 and np.abs(rhs[:, 2]).max() < 1e-7
        assert np.abs(rhs[:, 3] - heat).max() < 1e-8 * heat + 1e-9


def test_viscous_operator_is_second_order_accurate():
    """(rhs with mu) - (rhs without) against div(tau) and div(tau.u - q) of a smooth field at the cell centroids, interior cells,
    two resolutions of a regular quadrilateral mesh: the error falls by ~4 when the mesh is refined by 2."""
    mu, Pr, gamma = 0.02, 0.72, 1.4
    cp = R_GAS * gamma / (gamma - 1.0)
    kappa = mu * cp / Pr
    k = np.pi

    def fields(x, y):
        u = np.sin(k * x) * np.cos(k * y); v = -0.5 * np.cos(k * x) * np.sin(k * y); T = 300.0 + 10.0 * np.sin(k * x) * np.sin(k * y)
        ux, uy = k * np.cos(k * x) * np.cos(k * y), -k * np.sin(k * x) * np.sin(k * y)
        vx, vy = 0.5 * k * np.sin(k * x) * np.sin(k * y), -0.5 * k * np.cos(k * x) * np.cos(k * y)
        uxx, uyy, uxy = -k * k * u, -k * k * u, -k * k * np.cos(k * x) * np.sin(k * y)
        vxx, vyy, vxy = -k * k * v, -k * k * v, 0.5 * k * k * np.sin(k * x) * np.cos(k * y)
        Txx_yy = -2 * k * k * 10.0 * np.sin(k * x) * np.sin(k * y)
        div = ux + vy
        txx, tyy, txy = mu * (2 * ux - 2 / 3 * div), mu * (2 * vy - 2 / 3 * div), mu * (uy + vx)
        dtxx_dx = mu * (2 * uxx - 2 / 3 * (uxx + vxy)); dtxy_dy = mu * (uyy + vxy)
        dtxy_dx = mu * (uxy + vxx); dtyy_dy = mu * (2 * vyy - 2 / 3 * (uxy + vyy))
        mx, my = dtxx_dx + dtxy_dy, dtxy_dx + dtyy_dy
        en = u * mx + v * my + txx * ux + txy * (uy + vx) + tyy * vy + kappa * Txx_yy
        return u, v, T, mx, my, en

    errs = []
    for n in (16, 32):
        mesh = mb.Mesh.generate("cartesian", n, n, 1.0, 1.0)
        xy = mesh.arrays["cell_coords"]
        u, v, T, mx, my, en = fields(xy[:, 0], xy[:, 1])
        U0 = _state_from_prim(np.full(mesh.n_cells, 1.0), u, v, T, R_GAS)
        res = []
        for m in (mu, 0.0):
            se = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", gas=dict(_gas(m), Pr=Pr), bcs=EXTRAP4)
            se.set_state(U0)
            res.append(se.calc_rhs())
        dv = res[0] - res[1]
        cof = mesh.arrays["cells_of_face"]
        inner = np.ones(mesh.n_cells, bool)
        inner[cof[cof[:, 1] < 0, 0]] = False
        for _ in range(1):                                   # one more layer: the boundary cells' Green-Gauss gradients are first order
            edge = ~inner
            nb = np.zeros(mesh.n_cells, bool)
            two = cof[:, 1] >= 0
            nb[cof[two, 0]] |= edge[cof[two, 1]]; nb[cof[two, 1]] |= edge[cof[two, 0]]
            inner &= ~nb
        assert np.abs(dv[:, 0]).max() == 0.0                 # no viscous mass flux
        errs.append(max(np.abs(dv[inner, 1] - mx[inner]).max() / np.abs(mx).max(), np.abs(dv[inner, 2] - my[inner]).max() / np.abs(my).max(),
                        np.abs(dv[inner, 3] - en[inner]).max() / np.abs(en).max()))
    assert errs[0] < 0.05 and errs[1] < errs[0] / 3.0, errs


def test_zero_viscosity_is_the_inviscid_path_bit_for_bit(oracle_mod):
    mesh = mb.Mesh.generate("wedge", 14, 8, 4.0, 1.5)
    U0 = _smooth(mesh.arrays["cell_coords"] / 4.0, np.random.default_rng(4))
    a = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", bcs=SYM4)
    b = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(0.0), bcs=SYM4)
    a.set_state(U0); b.set_state(U0)
    assert np.array_equal(a.calc_rhs(), b.calc_rhs())


@pytest.mark.parametrize("recon,mu,order,mtype", [("FO", 0.0, 3, "wedge"), ("FO", 0.02, 3, "cartesian_tri"), ("TENO", 0.0, 3, "cartesian_tri"),
                                                  ("TENO", 0.02, 2, "cartesian_tri"), ("TENO", 0.0, 5, "cartesian_tri"), ("TENO", 0.01, 2, "mixed"),
                                                  ("TENO", 0.0, 3, "cartesian_tri:graph"), ("TENO", 0.02, 2, "mixed:graph")])
def test_emulated_rank_contexts_reproduce_the_single_context_residual(recon, mu, order, mtype):
    """Partitioned contexts (owned cells + the ghost rings the preprocessor decides to hold: one ring for first order, the stencil
    reach for TENO, one more ring for the least-squares gradients of viscous runs) with every held cell filled as a completed halo
    exchange leaves it: every rank's residual of its own cells equals the single-context residual bit for bit - specialised and
    generic kernels, quadrilaterals, viscous terms."""
    from mallard_b200 import synthetic as syn
    mtype, _, partitioner = mtype.partition(":")          # ":graph" = cut by mlb_partition_graph instead of the coordinate bisection
    if mtype == "mixed":
        mesh = syn.mixed_tri_quad(14, 12, 3.0, 2.0, seed=4, tri_fraction=0.5)
    else:
        mesh = mb.Mesh.generate(mtype, 14, 12, 3.0, 2.0)
    bcs = EXTRAP4 if mtype == "mixed" else SYM4
    kw = dict(recon=recon, riemann="HLLC", integrator="SSPRK3", bcs=bcs, order=order, quad_cell_order=5 if order >= 5 else 0, teno_fixed=True, gas=_gas(mu))
    U0 = _smooth(mesh.arrays["cell_coords"] / 3.0, np.random.default_rng(8))
    one = EmulatedSolver(mesh, **kw)
    one.set_state(U0)
    ref = one.calc_rhs()
    n_ranks = 3
    part = mb.partition_graph(mesh, n_ranks) if partitioner == "graph" else mb.partition(mesh, n_ranks)
    got = np.zeros_like(ref)
    held = []
    for r in range(n_ranks):
        e = EmulatedSolver(mesh, part=part, rank=r, n_ranks=n_ranks, **kw)
        assert e.n_owned == int((part == r).sum()) and e.n_held > e.n_owned
        held.append(e.n_held - e.n_owned)
        e.set_state(U0)
        got += e.calc_rhs()                       # zeros outside the rank's own cells
    assert np.isfinite(ref).all() and np.array_equal(got, ref)
    if recon == "FO":
        inviscid = EmulatedSolver(mesh, part=part, rank=1, n_ranks=n_ranks, **dict(kw, gas=_gas(0.0)))
        assert (held[1] > inviscid.n_held - inviscid.n_owned) == (mu > 0)      # the second ghost ring exists exactly when it is needed


# ---- FAST floating-point mode of the same kernels (MLB_STREAM_KERNELS source + FMA contraction by the host compiler, see
#      tests/emul/emulation.py: FLAGS): the tolerances the gated GPU tests (tests/test_gpu_parity.py, fp = "fast") assert -------------
from emulation import fast_available  # noqa: E402

FAST_EMULATION = pytest.mark.skipif(not fast_available(), reason="this host has no FMA unit: the FAST emulation cannot be built")
TOL = 1e-12


@FAST_EMULATION
@pytest.mark.parametrize("order,factor,qc,basis,fixed", [(3, 2.0, 0, "legendre", False), (3, 2.0, 0, "monomial", True), (5, 2.0, 5, "legendre", True),
                                                         (7, 2.0, 5, "monomial", True), (2, 1.5, 0, "legendre", False), (4, 2.5, 0, "legendre", True)])
def test_fast_emulation_stays_inside_the_tolerance_of_the_gpu_tests(oracle_mod, order, factor, qc, basis, fixed):
    """FAST against the oracle on the field scale: 1e-12 up to p = 5, 1e-9 above (conditioning of the pseudo-inverse rows), face values
    and residual - specialised kernel where one exists, generic kernel otherwise and when forced."""
    nx, ny = (12, 10) if order >= 7 else (14, 12)
    om = oracle_mod.Mesh.generate("cartesian_tri", nx, ny, 2.0, 1.0)
    mesh = mb.Mesh.generate("cartesian_tri", nx, ny, 2.0, 1.0)
    kw = dict(recon="TENO", riemann="HLLC", integrator="SSPRK3", bcs=SYM4, basis=basis, order=order, factor=factor, quad_cell_order=qc, teno_fixed=fixed)
    so, sf = oracle_mod.Solver(om, **kw), EmulatedSolver(mesh, fp_mode="fast", **kw)
    U0 = _smooth(mesh.arrays["cell_coords"], np.random.default_rng(23))
    so.set_state(U0); sf.set_state(U0)
    real = gu.real_faces(mesh.arrays["cells_of_face"], mesh.arrays["nodes_of_face"])
    Fo, ro = so.calc_face_values()[real][:, :, 0], so.calc_rhs()
    tol = TOL if order <= 5 else 1e-9
    for generic in (False, True):
        sf.force_generic(generic)
        Ff, rf = sf.calc_face_values()[real][:, :, 0], sf.calc_rhs()
        assert np.array_equal(np.isfinite(Ff), np.isfinite(Fo)) and np.array_equal(np.isfinite(rf), np.isfinite(ro))
        assert gu.field_err(Ff, Fo) <= tol and gu.field_err(rf, ro) <= tol, (generic, gu.field_err(Ff, Fo), gu.field_err(rf, ro))
        assert not np.array_equal(rf, ro) or order == 0          # (it IS a different rounding: the check is not vacuous)


@FAST_EMULATION
@pytest.mark.parametrize("order,tri_fraction", [(2, 0.5), (3, 0.0), (4, 0.6)])
def test_fast_emulation_on_mixed_meshes_is_k_exact(order, tri_fraction):
    from mallard_b200 import synthetic as syn
    mesh = syn.mixed_tri_quad(12, 10, 3.0, 2.0, seed=3, tri_fraction=tri_fraction)
    coef = np.random.default_rng(17).uniform(-1.0, 1.0, size=(4, order + 1, order + 1))

    def f(x, y):
        out = np.zeros(x.shape + (4,))
        for v in range(4):
            for i in range(order + 1):
                for j in range(order + 1 - i):
                    out[..., v] += coef[v, i, j] * x ** i * y ** j
        out[..., 0] += 5.0
        return out
    U0 = _cell_averages(mesh, f)
    F = {}
    for fp in ("strict", "fast"):
        se = EmulatedSolver(mesh, "TENO", "HLLC", "SSPRK3", order=order, bcs=EXTRAP4, teno_fixed=True, fp_mode=fp)
        se.set_state(U0)
        F[fp] = se.calc_face_values()
        n_quad = se.n_quad
    A = mesh.arrays
    nof, cof = A["nodes_of_face"].reshape(-1, 2).astype(np.int64), A["cells_of_face"]
    xi = {1: [0.0], 2: [-0.5773502691896257, 0.5773502691896257], 3: [-0.7745966692414834, 0.0, 0.7745966692414834]}[n_quad]
    x0, x1 = A["node_coords"][nof[:, 0]], A["node_coords"][nof[:, 1]]
    worst = 0.0
    for q, z in enumerate(xi):
        pq = (z + 1.0) * 0.5 * (x1 - x0) + x0
        exact = f(pq[:, 0], pq[:, 1])
        worst = max(worst, np.abs(F["fast"][:, q, 0] - exact).max(), np.abs(F["fast"][cof[:, 1] >= 0][:, q, 1] - exact[cof[:, 1] >= 0]).max())
    assert worst <= 2e-9 * (10.0 ** max(0, order - 3)), worst
    assert gu.field_err(F["fast"], F["strict"]) <= 1e-10 and not np.array_equal(F["fast"], F["strict"])


@FAST_EMULATION
@pytest.mark.parametrize("mtype", ["cartesian", "cartesian_tri", "mixed"])
def test_fast_emulation_of_the_viscous_couette_residual(mtype):
    """The assertion of tests/test_gpu_parity.py::test_viscous_residual_of_couette_flow[*-fast] with the FAST source."""
    from mallard_b200 import synthetic as syn
    mu, Uw, H, L = 0.05, 3.0, 1.0, 2.0
    mesh = syn.mixed_tri_quad(24, 20, L, H, seed=5, tri_fraction=0.5) if mtype == "mixed" else mb.Mesh.generate(mtype, 24, 20, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0]),
           dict(name="top", type="wall_noslip", u=[Uw, 0.0])]
    xy, n = mesh.arrays["cell_coords"], mesh.n_cells
    U0 = _state_from_prim(np.full(n, 1.2), Uw * xy[:, 1] / H, np.zeros(n), np.full(n, 300.0), R_GAS)
    res = []
    for m in (mu, 0.0):
        se = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(m), bcs=bcs, fp_mode="fast")
        se.set_state(U0)
        res.append(se.calc_rhs())
    heat, dv = mu * (Uw / H) ** 2, res[0] - res[1]
    scale = np.abs(res[1]).max() + 1.0
    assert np.abs(dv[:, :3]).max() < 1e-11 * scale + 1e-9 and np.abs(dv[:, 3] - heat).max() < 1e-11 * scale + 1e-8 * heat
    if mtype == "cartesian":
        # STRICT: the inviscid fluxes through opposite faces cancel exactly; FAST re-rounds them (FMA, reciprocals): what is left is a few ulp
        # of the energy flux per cell height, (rho E + p) U / dy ~ 1e7 here
        flux = (np.abs(U0[:, 3]).max() + 1.2 * R_GAS * 300.0) * Uw / (H / 20)
        slack = 1e-13 * flux
        assert np.abs(res[0][:, :3]).max() < 1e-7 + slack and np.abs(res[0][:, 3] - heat).max() < 1e-8 * heat + 1e-9 + slack


def test_emulated_decaying_shear_layer_follows_the_diffusion_equation():
    """tests/test_gpu_parity.py::test_decaying_shear_layer_follows_the_diffusion_equation with the emulated kernels and SSPRK3 in numpy
    (fixed dt below the acoustic limit): u(y, t) = U erf(y / sqrt(delta0^2 + 4 nu t)); the thresholds of the GPU test hold and the
    error falls by ~4 when the mesh is refined by 2 (measured: 7.6e-3 -> 1.8e-3)."""
    from math import erf
    nu, Uw, d0, T0, rho0 = 0.02, 1.0, 0.08, 300.0, 1.0
    errs = []
    for ny in (40, 80):
        mesh = mb.Mesh.generate("cartesian", 4, ny, 0.1, 2.0)
        y = mesh.arrays["cell_coords"][:, 1] - 1.0
        bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="symmetry"),
               dict(name="top", type="symmetry")]
        s = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(nu * rho0), bcs=bcs, fp_mode="fast" if fast_available() else "strict")
        n = mesh.n_cells
        U = _state_from_prim(np.full(n, rho0), Uw * np.vectorize(erf)(y / d0), np.zeros(n), np.full(n, T0), R_GAS)
        dt, t = 3.0e-5, 0.0                                  # acoustic limit h / (c + U) = 7e-5 on both meshes (h = 0.025), viscous limit far above

        def rhs(V):
            s.set_state(V)
            return s.calc_rhs()
        while t < 0.2:
            U1 = U + dt * rhs(U)
            U2 = 0.75 * U + 0.25 * (U1 + dt * rhs(U1))
            U = U / 3.0 + (2.0 / 3.0) * (U2 + dt * rhs(U2))
            t += dt
        assert np.isfinite(U).all()
        exact = Uw * np.vectorize(erf)(y / np.sqrt(d0 * d0 + 4.0 * nu * t))
        errs.append(np.abs(U[:, 1] / U[:, 0] - exact).max())
    assert errs[0] < 0.02 and errs[1] < errs[0] / 2.5, errs


# ---- whole time steps: the multi-kernel sequence of mlb_run and the cooperative small-mesh kernel (csrc/small_step.cuh) ---------------
WEDGE_BCS = [dict(name="left", type="upt", u=[600.0, 0.0], p=101325.0, T=300.0), dict(name="right", type="p_out", p=101325.0),
             dict(name="top", type="symmetry"), dict(name="bottom", type="wall_adiabatic")]


def _wedge_state(mesh):
    R = 101325.0 / (298.15 * 1.225)
    n = mesh.n_cells
    p, T, u = np.full(n, 101325.0), np.full(n, 300.0), np.full(n, 600.0)
    rho = p / (R * T)
    return np.stack([rho, rho * u, 0.0 * rho, rho * (p / (0.4 * rho) + 0.5 * u * u)], 1)


@pytest.mark.parametrize("mtype,nx,ny,Lx,Ly,riemann,integ,bcs,cfl,n_steps", [
    ("cartesian", 200, 1, 1.0, 0.005, "HLLC", "SSPRK3", SYM4, 1.0, 40),            # examples/sod numerics, downsized
    ("wedge", 30, 10, 4.0, 1.5, "HLLC", "SSPRK3", WEDGE_BCS, 1.0, 60),             # examples/wedge numerics + a wall, downsized
    ("cartesian_tri", 14, 11, 1.0, 0.8, "HLL", "RK4", EXTRAP4, 0.6, 12),
    ("cartesian", 24, 9, 2.0, 1.0, "Rusanov", "SSPRK3", SYM4, 0.8, 15)])
def test_emulated_time_steps_and_the_cooperative_small_mesh_kernel_vs_oracle(oracle_mod, mtype, nx, ny, Lx, Ly, riemann, integ, bcs, cfl, n_steps):
    """n steps of calc_dt + take_step: the oracle, the emulated multi-kernel sequence of mlb_run (CFL kernel, face kernel, gather / RK kernel
    per stage) and the phases of the cooperative kernel on grids of 1, 3 and 7 blocks (grid-stride loops, dt published between phases):
    the same bits, the same time, the same step count."""
    mesh = mb.Mesh.generate(mtype, nx, ny, Lx, Ly)
    om = oracle_mod.Mesh.generate(mtype, nx, ny, Lx, Ly)
    if mtype == "wedge":
        U0 = _wedge_state(mesh)
    elif ny == 1:
        x = mesh.arrays["cell_coords"][:, 0]
        rho, p = np.where(x < 0.5, 1.0, 0.125), np.where(x < 0.5, 1.0, 0.1)
        U0 = np.stack([rho, 0 * rho, 0 * rho, p / 0.4], 1)
    else:
        U0 = _smooth(mesh.arrays["cell_coords"], np.random.default_rng(3))
    kw = dict(recon="FO", riemann=riemann, integrator=integ, bcs=bcs)
    so = oracle_mod.Solver(om, **kw)
    so.set_state(U0)
    t = 0.0
    for _ in range(n_steps):
        dt = so.calc_dt(cfl)
        so.take_step(dt)
        t += dt
    Uo = so.get("U")
    assert np.isfinite(Uo).all() and np.abs(Uo - U0).max() > 0
    for blocks in (0, 1, 3, 7):
        se = EmulatedSolver(mesh, **kw)
        se.set_state(U0)
        te, dte = se.run(n_steps, cfl=cfl, small_blocks=blocks)
        assert np.array_equal(se.get_state(), Uo), blocks
        assert te == t and dte == dt and se.step_count() == n_steps, blocks
    # a fixed dt (cfl <= 0: no CFL phase), continued from where the run above stopped
    a, b = EmulatedSolver(mesh, **kw), EmulatedSolver(mesh, **kw)
    a.set_state(Uo); b.set_state(Uo)
    a.run(5, dt=0.5 * dt); b.run(5, dt=0.5 * dt, small_blocks=2)
    for _ in range(5):
        so.take_step(0.5 * dt)
    assert np.array_equal(a.get_state(), so.get("U")) and np.array_equal(b.get_state(), so.get("U"))


@pytest.mark.parametrize("name", ["wedge_30x10", "wedge_wall_30x10", "sod_hll_rk4"])
def test_emulated_cooperative_kernel_against_the_unmodified_reference_after_many_steps(name):
    """The reference's own state after 100 / 200 (wedge) and 40 (sod, HLL + RK4) steps (tests/golden, dumped by the unmodified reference):
    the cooperative kernel's phases reproduce it to 1e-12 of the field scale (bit for bit wherever pow() is not involved)."""
    meta, g = gu.load(name)
    mm = meta["mesh"]
    mesh = mb.Mesh.generate(mm["type"], mm["Nx"], mm["Ny"], mm["Lx"], mm["Ly"])
    se = EmulatedSolver(mesh, **gu.solver_kwargs(meta))
    se.set_state(g["U0"])
    done = 0
    for k in sorted(int(key[4:].split(":")[0]) for key in g if key.startswith("step") and key.endswith(":U")):
        n = k + 1 - done
        se.run(n, cfl=meta["cfl"], small_blocks=3)
        done += n
        err = gu.field_err(se.get_state(), g["step%d:U" % k])
        assert err <= 1e-12, (k, err)
    assert done >= 40 and se.step_count() == done


@pytest.mark.parametrize("recon,mu", [("TENO", 1.0e-3), ("TENO", 0.0), ("FO", 1.0e-3)])
def test_emulated_rank_local_contexts_reproduce_the_single_context_residual(recon, mu):
    """What bench.py's strong-scaling records run at N > 1 (BASELINE configs[3] / [4]): every rank builds its context from ITS PART of the
    jittered, id-shuffled triangulation only (synthetic.jittered_tri_local -> mlb_create_local: ghost layers, cut faces, global order kept),
    viscous contexts with their second ghost ring.  Each rank's residual of its own cells equals the single-context residual of the global
    mesh bit for bit."""
    from mallard_b200 import synthetic as syn
    nq, world = 26, 3
    mesh = syn.jittered_tri(nq, nq, 10.0, 10.0, seed=12345)
    U0 = syn.isentropic_vortex(mesh.arrays["cell_coords"])
    kw = dict(recon=recon, riemann="HLLC", integrator="SSPRK3", order=3, bcs=EXTRAP4, teno_fixed=True, gas=_gas(mu))
    one = EmulatedSolver(mesh, **kw)
    one.set_state(U0)
    ref = one.calc_rhs()
    assert np.isfinite(ref).all()
    got, n_owned = np.zeros_like(ref), 0
    for r in range(world):
        lp = syn.jittered_tri_local(nq, nq, 10.0, 10.0, world, r, seed=12345, layers=8)
        gids = np.asarray(lp.local["global_ids"], dtype=np.int64)
        assert np.array_equal(lp.mesh.arrays["cell_coords"], mesh.arrays["cell_coords"][gids])
        e = EmulatedSolver(lp.mesh, part=lp.part_local, rank=r, n_ranks=world, local=lp.local, **kw)
        assert e.n_owned == lp.n_owned
        n_owned += e.n_owned
        e.set_state(U0[gids])                     # every held cell filled as a completed halo exchange leaves it
        rhs = e.calc_rhs()                        # local numbering, zeros outside the rank's own cells
        own = lp.part_local == r
        assert np.abs(rhs[~own]).max() == 0.0
        got[gids[own]] = rhs[own]
    assert n_owned == mesh.n_cells and np.array_equal(got, ref)


@pytest.mark.parametrize("mtype", ["cartesian", "cartesian_tri", "mixed"])
def test_steady_heat_conduction_between_isothermal_walls(mtype):
    """MLB_BC_WALL_NOSLIP with T > 0 (isothermal wall): gas at rest, uniform pressure, T linear between a cold and a hot wall.  The heat flux
    kappa dT/dy is uniform, so the viscous part of the residual vanishes in every cell, wall cells included (the mirror state puts the wall's
    temperature ON the face), and the gradient of T is exact on quadrilaterals, triangles and jittered mixed meshes."""
    from mallard_b200 import synthetic as syn
    mu, H, L, T0, T1 = 0.05, 1.0, 2.0, 280.0, 340.0
    mesh = syn.mixed_tri_quad(12, 10, L, H, seed=7, tri_fraction=0.5) if mtype == "mixed" else mb.Mesh.generate(mtype, 12, 10, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0], T=T0),
           dict(name="top", type="wall_noslip", u=[0.0, 0.0], T=T1)]
    y = mesh.arrays["cell_coords"][:, 1]
    T = T0 + (T1 - T0) * y / H
    p = 1.0e5
    U0 = _state_from_prim(p / (R_GAS * T), np.zeros_like(y), np.zeros_like(y), T, R_GAS)
    res = []
    for m in (mu, 0.0):
        se = EmulatedSolver(mesh, "FO", "HLLC", "SSPRK3", gas=_gas(m), bcs=bcs)
        se.set_state(U0)
        if m > 0:
            G = se.gradients()
            assert np.abs(G[:, 5] - (T1 - T0) / H).max() < 1e-9 * (T1 - T0) / H and np.abs(G[:, :5]).max() < 1e-9
        res.append(se.calc_rhs())
    dv = res[0] - res[1]
    kappa = mu * (R_GAS * 1.4 / 0.4) / 0.72
    flux = kappa * (T1 - T0) / H                                     # the uniform heat flux; a cell's residual is its (vanishing) divergence
    h = (mesh.arrays["cell_volume"].min()) ** 0.5
    assert np.abs(dv[:, :3]).max() == 0.0 and np.abs(dv[:, 3]).max() < 1e-9 * flux / h


@pytest.mark.parametrize("order,mtype", [(2, "cartesian_tri"), (3, "cartesian_tri"), (3, "cartesian")])
def test_couette_flow_under_teno_is_a_steady_state_up_to_viscous_heating(order, mtype):
    """The viscous terms behind the TENO face kernel (face_flux_kernel<RS, true, Q, true>): plane Couette flow is reproduced exactly by a
    k-exact reconstruction (u is linear), both sides of every face carry the same state, the inviscid fluxes of this steady Euler solution
    cancel, and what is left of the residual is the viscous heating mu (U / H)^2 in the energy equation - in every cell that does not touch
    a wall.  Regular meshes (triangles, quadrilaterals): the viscous operator works on the primitives of the cell AVERAGES (second order), and the
    average of rho E over a cell carries u's variance across that cell - the same in every cell of a regular mesh, a cell-to-cell O(h^2)
    temperature noise on a jittered one (measured: 0.12 in the momentum residual of the 14 x 12 jittered mixed mesh, first order and TENO alike)."""
    mu, Uw, H, L = 0.05, 3.0, 1.0, 2.0
    mesh = mb.Mesh.generate(mtype, 14, 12, L, H)
    bcs = [dict(name="left", type="extrapolation"), dict(name="right", type="extrapolation"), dict(name="bottom", type="wall_noslip", u=[0.0, 0.0]),
           dict(name="top", type="wall_noslip", u=[Uw, 0.0])]
    # cell AVERAGES of the conserved state: rho, rho u (linear: value at the centroid) and rho E = rho (cv T + u^2 / 2) (quadratic in y)
    # (pressure 1, not 1e5: the reconstruction's conditioning error is relative to the size of rho E, and here it is compared with a heating of 0.45)
    f = lambda x, y: np.stack([np.full_like(x, 1.2), 1.2 * Uw * y / H, 0.0 * x, 1.0 / 0.4 + 0.6 * (Uw * y / H) ** 2], -1)   # noqa: E731
    U0 = _cell_averages(mesh, f)
    heat = mu * (Uw / H) ** 2
    for fp in (["strict", "fast"] if fast_available() else ["strict"]):
        se = EmulatedSolver(mesh, "TENO", "HLLC", "SSPRK3", order=order, gas=_gas(mu), bcs=bcs, teno_fixed=True, fp_mode=fp)
        se.set_state(U0)
        rhs = se.calc_rhs()
        cof = mesh.arrays["cells_of_face"]
        inner = np.ones(mesh.n_cells, bool)
        inner[cof[cof[:, 1] < 0, 0]] = False                            # the wall cells take their flux from the boundary functor (first-order accurate there)
        tol = 1e-9                                                      # k-exactness of the reconstruction / cell size
        assert np.abs(rhs[inner, :3]).max() < tol, (fp, np.abs(rhs[inner, :3]).max())
        assert np.abs(rhs[inner, 3] - heat).max() < tol, (fp, np.abs(rhs[inner, 3] - heat).max())
