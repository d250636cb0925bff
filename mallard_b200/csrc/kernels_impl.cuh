// Device kernels of the explicit residual path.  This file is compiled twice, into namespace `strict` with
// -fmad=false (every product and sum rounds separately, exactly as the reference's x86-64 build does) and into
// namespace `fast` with FMA contraction enabled.  Summation orders follow the reference in both.
//
//   face_flux_kernel    one thread per face: quadrature loop over the Riemann flux of interior and boundary faces
//                       (replaces K2 FirstOrder, K4 interior flux, K5-K9 boundary fluxes of SURVEY §2.1)
//   gather_stage_kernel cell-centric, atomic-free residual gather fused with the RK stage update
//                       (replaces K1 zero, the atomic scatter of K4-K9, K10 divide-by-volume, K11 BLAS-1 stage
//                        combinations and K12 update_primitives)
//   teno_recon_kernel   TENO reconstruction (K3): one thread per (cell, conserved variable), warp = one 8-cell table tile
//   cfl_kernel          spectral radius + max reduction + dt (K13, K14)
//
// Layout: conserved states and residuals are AoS [cell][4] (one 32-byte sector per cell: every gather of a neighbour's
// state moves exactly the bytes it uses; per-cell streaming accesses are 2 x 128-bit); primitives are SoA [6][Npad]
// (u, v, p, T, h as the reference's primitives view, plus rho for the CFL kernel); TENO face values are cell-centred
// AoS Fc[cell][slot * Q + q][4].
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#ifndef MLB_HOST_EMULATION
#include <cooperative_groups.h>
#endif

#include "kernel_args.h"

#ifndef MLB_KNS
#error "define MLB_KNS before including kernels_impl.cuh"
#endif

namespace mlb {
namespace MLB_KNS {

// ---------------------------------------------------------------------------------------------------------------
// Physics — Euler::compute_primitives_from_conservatives_impl (physics/physics.h:852-867)
// ---------------------------------------------------------------------------------------------------------------
// FAST mode replaces repeated divisions by one reciprocal and multiplications (<= 2 ulp, inside the 1e-12 budget);
// STRICT mode performs the reference's divisions.
#ifdef MLB_STREAM_KERNELS
#define MLB_DIV_BY(inv, den, x) ((x) * (inv))
#else
#define MLB_DIV_BY(inv, den, x) ((void)(inv), (x) / (den))
#endif

__device__ __forceinline__ void cons_to_prim(const GasParams & g, const double * U, double * P) {
    const double rho = U[0];
    const double ir = 1.0 / rho;
    const double u0 = MLB_DIV_BY(ir, rho, U[1]), u1 = MLB_DIV_BY(ir, rho, U[2]);
    const double E = MLB_DIV_BY(ir, rho, U[3]);
    const double e = E - 0.5 * (u0 * u0 + u1 * u1);
    const double p = fmax(g.p_min, fmin(g.p_max, (g.gamma - 1.0) * rho * e));   // :842-845
    P[0] = u0; P[1] = u1; P[2] = p; P[3] = e / g.cv; P[4] = e + MLB_DIV_BY(ir, rho, p);
}

// ---------------------------------------------------------------------------------------------------------------
// Star-state estimators and Riemann fluxes (numerics/riemann_solver.h:206-519)
// ---------------------------------------------------------------------------------------------------------------
struct Wn { double rho, un, p, gam; };

__device__ __forceinline__ void star_pvrs(const Wn & l, const Wn & r, double & ps) {   // :206-240, averages branch
    const double al = sqrt(l.gam * l.p / l.rho), ar = sqrt(r.gam * r.p / r.rho);
    const double rho_avg = 0.5 * (l.rho + r.rho), a_avg = 0.5 * (al + ar);
    ps = 0.5 * (l.p + r.p) + 0.5 * (l.un - r.un) * rho_avg * a_avg;
}
__device__ __noinline__ void star_trrs(const Wn & l, const Wn & r, double & ps) {      // :243-271
    const double al = sqrt(l.gam * l.p / l.rho), ar = sqrt(r.gam * r.p / r.rho);
    const double zl = (l.gam - 1.0) / (2.0 * l.gam), zr = (r.gam - 1.0) / (2.0 * r.gam);
    const double Plr = pow((l.p / r.p), zl);
    const double us = (Plr * l.un / al + r.un / ar + 2.0 * (1.0 - Plr) / (l.gam - 1.0)) / (Plr / al + 1.0 / ar);
    ps = 0.5 * (l.p * pow((1.0 + (l.gam - 1.0) / (2.0 * al) * (l.un - us)), (1.0 / zl)) +
                r.p * pow((1.0 + (r.gam - 1.0) / (2.0 * ar) * (us - r.un)), (1.0 / zr)));
}
__device__ __forceinline__ void star_tsrs(const Wn & l, const Wn & r, double & ps) {   // :274-306, p0 = PVRS guess
    const double gl = (l.gam - 1.0) / (l.gam + 1.0), gr = (r.gam - 1.0) / (r.gam + 1.0);
    const double Al = 2.0 / (l.gam + 1.0) / l.rho, Bl = gl * l.p;
    const double Ar = 2.0 / (r.gam + 1.0) / r.rho, Br = gr * r.p;
    const double p0 = fmax(0.0, ps);
    const double ql = sqrt(Al / (p0 + Bl)), qr = sqrt(Ar / (p0 + Br));
    ps = (ql * l.p + qr * r.p - (r.un - l.un)) / (ql + qr);
}
__device__ __forceinline__ double star_pressure(const Wn & l, const Wn & r) {           // ANRS :309-329
    const double pmax = fmax(l.p, r.p), pmin = fmin(l.p, r.p);
    const double qmax = pmax / pmin;
    double ps;
    star_pvrs(l, r, ps);
    if (!((qmax < 2.0) && (pmin <= ps) && (ps <= pmax))) {
        if (ps < pmin) star_trrs(l, r, ps); else star_tsrs(l, r, ps);
    }
    return ps;
}

// One face state as RiemannSolver::calc_flux receives it (riemann_solver.h:85-90)
struct FaceState { double rho, u, v, p, h; };

template <int RS>
__device__ __forceinline__ void riemann_flux(double * F, double nx, double ny, const FaceState & L, const FaceState & R, double gam) {
    const double uln = L.u * nx + L.v * ny, urn = R.u * nx + R.v * ny;
    const double ulul = L.u * L.u + L.v * L.v, urur = R.u * R.u + R.v * R.v;
    const double al = sqrt(gam * L.p / L.rho), ar = sqrt(gam * R.p / R.rho);
    const double El = (L.h + 0.5 * ulul) * L.rho - L.p, Er = (R.h + 0.5 * urur) * R.rho - R.p;
    const double Ul[4] = {L.rho, L.rho * L.u, L.rho * L.v, El}, Ur[4] = {R.rho, R.rho * R.u, R.rho * R.v, Er};
    const double Fl[4] = {L.rho * uln, L.rho * L.u * uln + L.p * nx, L.rho * L.v * uln + L.p * ny, (El + L.p) * uln};
    const double Fr[4] = {R.rho * urn, R.rho * R.u * urn + R.p * nx, R.rho * R.v * urn + R.p * ny, (Er + R.p) * urn};
    if (RS == MLB_RIEMANN_RUSANOV) {   // :332-375
        const double smax = fmax(fabs(uln) + al, fabs(urn) + ar);
#pragma unroll
        for (int i = 0; i < 4; i++) F[i] = 0.5 * (Fl[i] + Fr[i] + smax * (Ul[i] - Ur[i]));
        return;
    }
    const Wn wl = {L.rho, uln, L.p, gam}, wr = {R.rho, urn, R.p, gam};
    const double ps = star_pressure(wl, wr);
    const double ql = (ps <= L.p) ? 1.0 : sqrt(1.0 + (gam + 1.0) / (2.0 * gam) * (ps / L.p - 1.0));
    const double qr = (ps <= R.p) ? 1.0 : sqrt(1.0 + (gam + 1.0) / (2.0 * gam) * (ps / R.p - 1.0));
    const double Sl = uln - al * ql, Sr = urn + ar * qr;
    if (RS == MLB_RIEMANN_HLL) {       // :378-439
        if (0.0 <= Sl) {
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = Fl[i];
        } else if (Sr <= 0.0) {
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = Fr[i];
        } else {
            const double den = Sr - Sl, inv = 1.0 / den;
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = MLB_DIV_BY(inv, den, Sr * Fl[i] - Sl * Fr[i] + Sl * Sr * (Ur[i] - Ul[i]));
        }
        return;
    }
    // HLLC "variant 2" :442-519
    const double Ss = (R.p - L.p + L.rho * uln * (Sl - uln) - R.rho * urn * (Sr - urn)) / (L.rho * (Sl - uln) - R.rho * (Sr - urn));
    if (0.0 <= Sl) {
#pragma unroll
        for (int i = 0; i < 4; i++) F[i] = Fl[i];
    } else if (Sr <= 0.0) {
#pragma unroll
        for (int i = 0; i < 4; i++) F[i] = Fr[i];
    } else {
        const double D[4] = {0.0, nx, ny, Ss};
        const double Plr = 0.5 * (L.p + R.p + L.rho * (Sl - uln) * (Ss - uln) + R.rho * (Sr - urn) * (Ss - urn));
        if (Ss >= 0.0) {
            const double den = Sl - Ss, inv = 1.0 / den;
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = MLB_DIV_BY(inv, den, Ss * (Sl * Ul[i] - Fl[i]) + Sl * Plr * D[i]);
        } else {
            const double den = Sr - Ss, inv = 1.0 / den;
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = MLB_DIV_BY(inv, den, Ss * (Sr * Ur[i] - Fr[i]) + Sr * Plr * D[i]);
        }
    }
}

#ifdef MLB_STREAM_KERNELS
// ---------------------------------------------------------------------------------------------------------------
// FAST mode: the same wave-speed estimates and fluxes (numerics/riemann_solver.h:206-519) with the algebra arranged for the
// FP64 pipe, which is what bounds the face kernel (a double-precision division or square root is a ~10-instruction
// dependent chain): one reciprocal per state instead of four divisions, a^2 = gamma p (1/rho), a q = sqrt(gamma (1/rho)
// (p + c (p* - p))) instead of sqrt(..) * sqrt(1 + c (p*/p - 1)), and only the side of the fan the face lies in is evaluated.  Differences to riemann_flux<RS>() are a
// few ulp (inside the 1e-12 per-step budget; tests/test_gpu_parity.py::test_riemann_*).  The rarely taken TRRS / TSRS
// estimators stay out of line.
// ---------------------------------------------------------------------------------------------------------------
struct FaceCons { double rho, u, v, p, E, ir; };   // E = rho * total energy per mass, ir = 1 / rho

__device__ __forceinline__ FaceCons face_cons(const GasParams & g, const double * U) {
    FaceCons s;
    s.rho = U[0]; s.ir = 1.0 / U[0];
    s.u = U[1] * s.ir; s.v = U[2] * s.ir;
    const double ke = 0.5 * (s.u * s.u + s.v * s.v);
    const double e = U[3] * s.ir - ke;
    s.p = fmax(g.p_min, fmin(g.p_max, (g.gamma - 1.0) * s.rho * e));   // physics.h:842-845
    // rho E as the reference rebuilds it from h (flux_functor.h:140-149 -> riemann_solver.h): with reference-faithful TENO weights
    //  a face state can be far outside the physical range, e - ke cancels, and U[3] itself would be a DIFFERENT
    // (more accurate) number than the one the reference's flux sees
    s.E = ((e + s.p * s.ir) + ke) * s.rho - s.p;
    return s;
}
__device__ __forceinline__ FaceCons face_cons(const FaceState & f) {   // ghost states arrive as (rho, u, v, p, h)
    FaceCons s;
    s.rho = f.rho; s.ir = 1.0 / f.rho; s.u = f.u; s.v = f.v; s.p = f.p;
    s.E = (f.h + 0.5 * (f.u * f.u + f.v * f.v)) * f.rho - f.p;
    return s;
}
__device__ __noinline__ double star_pressure_rare(double rl, double unl, double pl, double rr, double unr, double pr, double gam, double ps) {
    const Wn l = {rl, unl, pl, gam}, r = {rr, unr, pr, gam};
    if (ps < fmin(pl, pr)) star_trrs(l, r, ps); else star_tsrs(l, r, ps);
    return ps;
}

template <int RS>
__device__ __forceinline__ void riemann_flux_lean(double * F, double nx, double ny, const FaceCons & L, const FaceCons & R, double gam) {
    const double uln = L.u * nx + L.v * ny, urn = R.u * nx + R.v * ny;
    const double al = sqrt(gam * L.p * L.ir), ar = sqrt(gam * R.p * R.ir);
    const double ml = L.rho * uln, mr = R.rho * urn;
    if (RS == MLB_RIEMANN_RUSANOV) {   // :332-375
        const double smax = fmax(fabs(uln) + al, fabs(urn) + ar);
        F[0] = 0.5 * (ml + mr + smax * (L.rho - R.rho));
        F[1] = 0.5 * ((ml * L.u + L.p * nx) + (mr * R.u + R.p * nx) + smax * (L.rho * L.u - R.rho * R.u));
        F[2] = 0.5 * ((ml * L.v + L.p * ny) + (mr * R.v + R.p * ny) + smax * (L.rho * L.v - R.rho * R.v));
        F[3] = 0.5 * ((L.E + L.p) * uln + (R.E + R.p) * urn + smax * (L.E - R.E));
        return;
    }
    // ANRS :309-329 with the PVRS guess :206-240 inline
    const double pmax = fmax(L.p, R.p), pmin = fmin(L.p, R.p);
    double ps = 0.5 * (L.p + R.p) + 0.5 * (uln - urn) * (0.5 * (L.rho + R.rho)) * (0.5 * (al + ar));
    const bool q_small = pmin > 0.0 ? pmax < 2.0 * pmin : pmax / pmin < 2.0;
    if (!(q_small && (pmin <= ps) && (ps <= pmax))) ps = star_pressure_rare(L.rho, uln, L.p, R.rho, urn, R.p, gam, ps);
    const double c = (gam + 1.0) / (2.0 * gam);
    const double Sl = uln - ((ps <= L.p) ? al : sqrt(gam * L.ir * (L.p + c * (ps - L.p))));
    const double Sr = urn + ((ps <= R.p) ? ar : sqrt(gam * R.ir * (R.p + c * (ps - R.p))));
    if (RS == MLB_RIEMANN_HLL) {       // :378-439
        const double Fl[4] = {ml, ml * L.u + L.p * nx, ml * L.v + L.p * ny, (L.E + L.p) * uln};
        const double Fr[4] = {mr, mr * R.u + R.p * nx, mr * R.v + R.p * ny, (R.E + R.p) * urn};
        if (0.0 <= Sl) {
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = Fl[i];
        } else if (Sr <= 0.0) {
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = Fr[i];
        } else {
            const double Ul[4] = {L.rho, L.rho * L.u, L.rho * L.v, L.E}, Ur[4] = {R.rho, R.rho * R.u, R.rho * R.v, R.E};
            const double inv = 1.0 / (Sr - Sl);
#pragma unroll
            for (int i = 0; i < 4; i++) F[i] = (Sr * Fl[i] - Sl * Fr[i] + Sl * Sr * (Ur[i] - Ul[i])) * inv;
        }
        return;
    }
    // HLLC "variant 2" :442-519: F = F_K, or F*_K = (S* (S_K U_K - F_K) + S_K P_LR D) / (S_K - S*), K the side of the contact
    const double dl = L.rho * (Sl - uln), dr = R.rho * (Sr - urn);
    const double Ss = (R.p - L.p + dl * uln - dr * urn) / (dl - dr);
    const bool pure = (0.0 <= Sl) || (Sr <= 0.0);
    const bool left = (0.0 <= Sl) || (!(Sr <= 0.0) && (Ss >= 0.0));
    const double rK = left ? L.rho : R.rho, uK = left ? L.u : R.u, vK = left ? L.v : R.v, pK = left ? L.p : R.p, EK = left ? L.E : R.E;
    const double unK = left ? uln : urn, mK = left ? ml : mr, SK = left ? Sl : Sr;
    const double FK[4] = {mK, mK * uK + pK * nx, mK * vK + pK * ny, (EK + pK) * unK};
    if (pure) {
#pragma unroll
        for (int i = 0; i < 4; i++) F[i] = FK[i];
    } else {
        const double Plr = 0.5 * (L.p + R.p + dl * (Ss - uln) + dr * (Ss - urn));
        const double inv = 1.0 / (SK - Ss);
        const double UK[4] = {rK, rK * uK, rK * vK, EK}, D[4] = {0.0, nx, ny, Ss};
        const double sp = SK * Plr;
#pragma unroll
        for (int i = 0; i < 4; i++) F[i] = (Ss * (SK * UK[i] - FK[i]) + sp * D[i]) * inv;
    }
}
#endif  // MLB_STREAM_KERNELS

// Ghost state of a boundary face from the interior face state (boundary/*.cpp calc_lr_states_impl)
__device__ __forceinline__ void ghost_state(const BcParams & bc, const GasParams & g, double nx, double ny, const double * Ul,
                                            const double * Pl, FaceState & R) {
    switch (bc.type) {
        case MLB_BC_SYMMETRY: {        // boundary_symmetry.cpp:39-69
            const double un = Pl[0] * nx + Pl[1] * ny;
            R.rho = Ul[0]; R.u = Pl[0] - 2.0 * un * nx; R.v = Pl[1] - 2.0 * un * ny; R.p = Pl[2]; R.h = Pl[4];
            break;
        }
        case MLB_BC_UPT:               // boundary_upt.cpp:84-100
            R.rho = bc.data[0]; R.u = bc.data[1]; R.v = bc.data[2]; R.p = bc.data[3]; R.h = bc.data[5];
            break;
        case MLB_BC_P_OUT: {           // boundary_p_out.cpp:58-100
            const double umag = sqrt(Pl[0] * Pl[0] + Pl[1] * Pl[1]);
            const double sos = sqrt(g.gamma * Pl[2] / Ul[0]);
            const double p_out = (umag < sos) ? bc.data[0] : Pl[2];
            const double rho_bc = p_out / (g.R * Pl[3]);
            const double e_bc = g.cv * Pl[3];
            R.rho = rho_bc; R.u = Pl[0]; R.v = Pl[1]; R.p = p_out; R.h = e_bc + p_out / rho_bc;
            break;
        }
        default:                       // extrapolation: boundary_extrapolation.cpp:39-55
            R.rho = Ul[0]; R.u = Pl[0]; R.v = Pl[1]; R.p = Pl[2]; R.h = Pl[4];
            break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// RK stage combination (numerics/time_integrator.cpp:57-163; KokkosBlas axpy: y = a*x + y, axpby: y = a*x + b*y)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const double * p, size_t i, double * out) {
    const double4 t = reinterpret_cast<const double4 *>(p)[i];
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
__device__ __forceinline__ void st4(double * p, size_t i, const double * v) {
    reinterpret_cast<double4 *>(p)[i] = make_double4(v[0], v[1], v[2], v[3]);
}

__device__ __forceinline__ void rk_update(const RkArgs & rk, const double * Uin, uint32_t i, const double * k, double dt, double * Unew) {
    double base[4];
    ld4(rk.base, i, base);
    if (rk.mode == 0) {
#pragma unroll
        for (int v = 0; v < 4; v++) Unew[v] = (dt * rk.coef) * k[v] + base[v];
    } else if (rk.mode == 1) {
        double uin[4];
        ld4(Uin, i, uin);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double y = rk.c0 * base[v] + rk.c1 * uin[v];
            Unew[v] = (dt * rk.coef) * k[v] + y;
        }
    } else {
#pragma unroll
        for (int v = 0; v < 4; v++) Unew[v] = base[v];
        for (int j = 0; j < rk.n_prev; j++) {
            double kp[4];
            ld4(rk.kprev[j], i, kp);
#pragma unroll
            for (int v = 0; v < 4; v++) Unew[v] = (dt * rk.cprev[j]) * kp[v] + Unew[v];
        }
#pragma unroll
        for (int v = 0; v < 4; v++) Unew[v] = (dt * rk.coef) * k[v] + Unew[v];
    }
    st4(rk.out, i, Unew);
}

// ---------------------------------------------------------------------------------------------------------------
// Viscous terms (new: the reference is Euler only, physics/physics.h:23-29; SURVEY 8f N4, BASELINE configs[4]).  Compiled in
// only for mu > 0: with mu == 0 the instantiations below are not launched and the path is the reference's, bit for bit.
//   visc_grad_kernel   least-squares gradients of (u, v, T) of every owned and first-ring ghost cell from the cell averages of its
//                      face neighbours (boundary faces: the mirror state that puts the boundary's value on the face), weights
//                      1 / distance^2: exact for linear fields on any mesh (Green-Gauss with face averages is not, on triangles)
//   viscous_flux()     at a face: gradient = mean of the two cells' gradients, with its component along the centroid line
//                      replaced by the two-point difference (no odd-even decoupling); Newtonian stress with Stokes' hypothesis,
//                      Fourier heat flux; returns F_v . n
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cons_to_uvT(const GasParams & g, const double * U, double * w) {
    const double ir = 1.0 / U[0];
    w[0] = U[1] * ir; w[1] = U[2] * ir;
    w[2] = (U[3] * ir - 0.5 * (w[0] * w[0] + w[1] * w[1])) / g.cv;
}
// value of (u, v, T) the boundary condition puts ON the face, given the interior cell's
__device__ __forceinline__ void boundary_face_uvT(const BcParams & bc, double nx, double ny, const double * wi, double * wf) {
    switch (bc.type) {
        case MLB_BC_WALL_NOSLIP: wf[0] = bc.data[1]; wf[1] = bc.data[2]; wf[2] = bc.data[4] > 0.0 ? bc.data[4] : wi[2]; break;
        case MLB_BC_SYMMETRY: case MLB_BC_WALL_ADIABATIC: {
            const double un = wi[0] * nx + wi[1] * ny;
            wf[0] = wi[0] - un * nx; wf[1] = wi[1] - un * ny; wf[2] = wi[2]; break;
        }
        case MLB_BC_UPT: wf[0] = 0.5 * (wi[0] + bc.data[1]); wf[1] = 0.5 * (wi[1] + bc.data[2]); wf[2] = 0.5 * (wi[2] + bc.data[4]); break;
        default: wf[0] = wi[0]; wf[1] = wi[1]; wf[2] = wi[2]; break;
    }
}

__global__ void __launch_bounds__(256) visc_grad_kernel(const __grid_constant__ StageArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.g.N_recon) return;
    const uint32_t Np = a.g.Npad;
    double Ui[4], wi[3];
    ld4(a.Uin, i, Ui);
    cons_to_uvT(a.ph.gas, Ui, wi);
    // inverse-distance weighted least squares over the face neighbours: min sum_j w_j (phi_j - phi_i - g . d_j)^2, w_j = 1 / |d_j|^2
    double axx = 0.0, axy = 0.0, ayy = 0.0, bx[3] = {0.0, 0.0, 0.0}, by[3] = {0.0, 0.0, 0.0};
    const int nf = a.g.nfc[i];
    for (int j = 0; j < nf; j++) {
        const size_t at = (size_t)j * Np + i;
        const double dx = a.g.slot_d[((size_t)j * 2) * Np + i], dy = a.g.slot_d[((size_t)j * 2 + 1) * Np + i];
        const double d2 = dx * dx + dy * dy;
        if (!(d2 > 0.0)) continue;
        const int32_t nbr = a.g.slot_nbr[at];
        double wn[3];
        if (nbr >= 0) {
            double Un[4];
            ld4(a.Uin, (size_t)nbr, Un);
            cons_to_uvT(a.ph.gas, Un, wn);
        } else if (nbr != INT32_MIN) {       // mirror state: the value ON the face is the boundary's
            const double id = 1.0 / sqrt(d2);
            double wf[3];
            boundary_face_uvT(a.ph.bcs[-nbr - 1], dx * id, dy * id, wi, wf);
#pragma unroll
            for (int v = 0; v < 3; v++) wn[v] = 2.0 * wf[v] - wi[v];
        } else {
#pragma unroll
            for (int v = 0; v < 3; v++) wn[v] = wi[v];
        }
        const double w = 1.0 / d2;
        axx += w * dx * dx; axy += w * dx * dy; ayy += w * dy * dy;
#pragma unroll
        for (int v = 0; v < 3; v++) { const double dphi = w * (wn[v] - wi[v]); bx[v] += dphi * dx; by[v] += dphi * dy; }
    }
    const double det = axx * ayy - axy * axy;
    const double idet = det != 0.0 ? 1.0 / det : 0.0;
    double * G = a.G + 6 * (size_t)i;
#pragma unroll
    for (int v = 0; v < 3; v++) {
        G[2 * v] = (ayy * bx[v] - axy * by[v]) * idet;
        G[2 * v + 1] = (axx * by[v] - axy * bx[v]) * idet;
    }
}

// F_v . n at face f (unit normal n out of cell cl); cr < 0: boundary with condition bc
__device__ __forceinline__ void viscous_flux(const StageArgs & a, uint32_t f, uint32_t cl, int32_t cr, double nx, double ny, double * Fv) {
    const GasParams & g = a.ph.gas;
    double Ul[4], wl[3], wr[3], Gl[6], Gr[6];
    ld4(a.Uin, cl, Ul);
    cons_to_uvT(g, Ul, wl);
#pragma unroll
    for (int v = 0; v < 6; v++) Gl[v] = a.G[6 * (size_t)cl + v];
    if (cr >= 0) {
        double Ur[4];
        ld4(a.Uin, (size_t)cr, Ur);
        cons_to_uvT(g, Ur, wr);
#pragma unroll
        for (int v = 0; v < 6; v++) Gr[v] = a.G[6 * (size_t)cr + v];
    } else {
        const BcParams & bc = a.ph.bcs[-cr - 1];
        if (bc.type == MLB_BC_SYMMETRY || bc.type == MLB_BC_WALL_ADIABATIC) {     // slip surfaces: no shear, no heat flux
            Fv[0] = 0.0; Fv[1] = 0.0; Fv[2] = 0.0; Fv[3] = 0.0;
            return;
        }
        double wf[3];
        boundary_face_uvT(bc, nx, ny, wl, wf);
#pragma unroll
        for (int v = 0; v < 3; v++) wr[v] = 2.0 * wf[v] - wl[v];                   // mirror state: the face value is the boundary's
#pragma unroll
        for (int v = 0; v < 6; v++) Gr[v] = Gl[v];
    }
    const double dx = a.g.face_d[f], dy = a.g.face_d[(size_t)a.g.NFpad + f];
    const double id = 1.0 / sqrt(dx * dx + dy * dy);
    const double ex = dx * id, ey = dy * id;
    double gx[3], gy[3];
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const double mx = 0.5 * (Gl[2 * v] + Gr[2 * v]), my = 0.5 * (Gl[2 * v + 1] + Gr[2 * v + 1]);
        const double corr = (wr[v] - wl[v]) * id - (mx * ex + my * ey);
        gx[v] = mx + corr * ex; gy[v] = my + corr * ey;
    }
    // velocity AT the face mid-point (the work of the stress): both cells' linear reconstructions, averaged - exact for linear
    // fields on skewed meshes, where the mean of the two cell values is not
    const double rx = a.g.face_d[2 * (size_t)a.g.NFpad + f], ry = a.g.face_d[3 * (size_t)a.g.NFpad + f];
    const double uf = 0.5 * ((wl[0] + Gl[0] * rx + Gl[1] * ry) + (wr[0] + Gr[0] * (rx - dx) + Gr[1] * (ry - dy)));
    const double vf = 0.5 * ((wl[1] + Gl[2] * rx + Gl[3] * ry) + (wr[1] + Gr[2] * (rx - dx) + Gr[3] * (ry - dy)));
    const double div = gx[0] + gy[1];
    const double txx = g.mu * (2.0 * gx[0] - (2.0 / 3.0) * div), tyy = g.mu * (2.0 * gy[1] - (2.0 / 3.0) * div), txy = g.mu * (gy[0] + gx[1]);
    Fv[0] = 0.0;
    Fv[1] = txx * nx + txy * ny;
    Fv[2] = txy * nx + tyy * ny;
    Fv[3] = (uf * txx + vf * txy + g.kappa * gx[2]) * nx + (uf * txy + vf * tyy + g.kappa * gy[2]) * ny;
}

// ---------------------------------------------------------------------------------------------------------------
// Face fluxes.  One thread per face: quadrature loop over the Riemann flux (BaseFluxFunctor::call_impl,
// numerics/flux_functor.h:124-162; boundary ghost states boundary/*.cpp), result A * (1/2 sum_q w_q F_q) stored once per
// face.  The reference scatters -+ that product into both cells with atomic_add; here the cells gather it (next kernel).
// ---------------------------------------------------------------------------------------------------------------
#ifndef MLB_FLUX_MINB
#define MLB_FLUX_MINB 8   // 64 registers: occupancy beats the 216 bytes of spills (A/B on B200: 0.48 -> 0.37 ms per launch, profiles/r01h)
#endif
// QT > 0: the number of face quadrature points is known at compile time and ONE THREAD PER (face, quadrature point) solves
// one Riemann problem; the QT lanes of a face then combine w_q F_q in the reference's q order with warp shuffles (same
// rounding as the sequential loop) and lane 0 stores.  Twice the parallelism and half the dependent sqrt/div chain per
// thread of a per-face loop.  QT = 0: one thread per face loops over a run-time Q.
// (the body takes the global thread index as a parameter: small_step_kernel below walks the faces with a grid-stride loop)
template <int RS, bool TENO, int QT, bool VISC>
__device__ __forceinline__ void face_flux_body(const StageArgs & a, const uint32_t gid) {
    constexpr int TPF = QT > 0 ? QT : 1;                 // threads per face (1, 2 or 4: divides the warp)
    const bool valid = gid / TPF < a.g.NF;
    const uint32_t f = valid ? gid / TPF : a.g.NF - 1;   // surplus lanes recompute the last face and do not store
    const int lane_q = gid % TPF;
    const int Q = !TENO ? 1 : (QT > 0 ? QT : a.g.Q);
    const int NPT = a.g.n_slots * Q;                     // face-value points per cell
    const uint32_t cl = a.g.face_cl[f];
    const int32_t cr = a.g.face_cr[f];
    const double nx = a.g.face_nx[f], ny = a.g.face_ny[f], area = a.g.face_area[f];
    const uint32_t slots = TENO ? a.g.face_slots[f] : 0u;
    const int sl = slots & 15u, sr = slots >> 4;
    const bool no_flux = cr == INT32_MIN;                // boundary zone without a [[boundaries]] entry
    const BcParams * bc = (cr < 0 && !no_flux) ? &a.ph.bcs[-cr - 1] : nullptr;

    auto flux_at = [&](int q, double * ft) {
        double Ul[4], Pl[5];
        if (TENO) ld4(a.Fc, (size_t)cl * NPT + (sl * Q + q), Ul); else ld4(a.Uin, cl, Ul);
        if (cr >= 0) {
            double Ur[4];
            if (TENO) ld4(a.Fc, (size_t)cr * NPT + (sr * Q + q), Ur); else ld4(a.Uin, (size_t)cr, Ur);
#ifdef MLB_STREAM_KERNELS
            riemann_flux_lean<RS>(ft, nx, ny, face_cons(a.ph.gas, Ul), face_cons(a.ph.gas, Ur), a.ph.gas.gamma);
#else
            double Pr[5];
            cons_to_prim(a.ph.gas, Ul, Pl);
            cons_to_prim(a.ph.gas, Ur, Pr);
            const FaceState L = {Ul[0], Pl[0], Pl[1], Pl[2], Pl[4]};
            const FaceState R = {Ur[0], Pr[0], Pr[1], Pr[2], Pr[4]};
            riemann_flux<RS>(ft, nx, ny, L, R, a.ph.gas.gamma);
#endif
            return;
        }
        cons_to_prim(a.ph.gas, Ul, Pl);
        if (bc->type == MLB_BC_WALL_ADIABATIC) {          // boundary_wall_adiabatic.cpp:39-70
            ft[0] = 0.0; ft[1] = Pl[2] * nx; ft[2] = Pl[2] * ny; ft[3] = 0.0;
        } else {
            FaceState gh;
            ghost_state(*bc, a.ph.gas, nx, ny, Ul, Pl, gh);
#ifdef MLB_STREAM_KERNELS
            riemann_flux_lean<RS>(ft, nx, ny, face_cons(a.ph.gas, Ul), face_cons(gh), a.ph.gas.gamma);
#else
            const FaceState L = {Ul[0], Pl[0], Pl[1], Pl[2], Pl[4]};
            riemann_flux<RS>(ft, nx, ny, L, gh, a.ph.gas.gamma);
#endif
        }
    };

    double fsum[4] = {0.0, 0.0, 0.0, 0.0};
    if (QT == 0) {
        if (!no_flux)
            for (int q = 0; q < Q; q++) {
                double ft[4];
                flux_at(q, ft);
                const double wq = a.ph.qf_w[q];
#pragma unroll
                for (int v = 0; v < 4; v++) fsum[v] += wq * ft[v];     // flux_functor.h:151
            }
    } else {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        if (!no_flux) {
            double ft[4];
            flux_at(lane_q, ft);
            const double wq = a.ph.qf_w[lane_q];
#pragma unroll
            for (int v = 0; v < 4; v++) t[v] = wq * ft[v];
        }
#pragma unroll
        for (int q = 0; q < TPF; q++) {                                // fsum = ((0 + t_0) + t_1) + ..., flux_functor.h:151
#pragma unroll
            for (int v = 0; v < 4; v++) fsum[v] += __shfl_sync(0xffffffffu, t[v], q, TPF);
        }
        if (lane_q != 0) return;
    }
    if (!valid) return;
    if (VISC) {   // the residual takes A (F_inviscid - F_viscous) . n
        double Fv[4] = {0.0, 0.0, 0.0, 0.0};
        if (!no_flux) viscous_flux(a, f, cl, cr, nx, ny, Fv);
#pragma unroll
        for (int v = 0; v < 4; v++) fsum[v] = area * (fsum[v] * 0.5 - Fv[v]);
    } else {
#pragma unroll
        for (int v = 0; v < 4; v++) fsum[v] = area * (fsum[v] * 0.5);   // flux_functor.h:153,156-161: (-A)*F == -(A*F) exactly
    }
    reinterpret_cast<double4 *>(a.AF)[f] = make_double4(fsum[0], fsum[1], fsum[2], fsum[3]);
}
template <int RS, bool TENO, int QT, bool VISC>
__global__ void __launch_bounds__(128, MLB_FLUX_MINB) face_flux_kernel(const __grid_constant__ StageArgs a) {
    face_flux_body<RS, TENO, QT, VISC>(a, blockIdx.x * blockDim.x + threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------
// Residual gather + RK update.  One thread per owned cell, no atomics: each cell sums -+ the stored face products in
// the reference's (Serial back-end) accumulation order, divides by its volume (DivideVolumeFunctor,
// solver_rhs.cpp:18-42) and applies the stage combination; the last stage also refreshes the primitives.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_stage_body(const StageArgs & a, const uint32_t i) {
    if (i >= a.g.N_owned) return;
    const uint32_t Np = a.g.Npad;
    double k[4];
    if (a.k_override) {
        ld4(a.k_override, i, k);
    } else {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const uint32_t order = a.g.rhs_order[i];
        const int nf = a.g.nfc[i];
        const double4 * AF = reinterpret_cast<const double4 *>(a.AF);
#pragma unroll
        for (int jj = 0; jj < MAX_SLOTS; jj++) {
            if (jj < nf) {
                const int s = (order >> (2 * jj)) & 3;
                const int32_t nbr = a.g.slot_nbr[(size_t)s * Np + i];
                if (nbr != INT32_MIN) {
                    const uint32_t fcode = a.g.slot_face[(size_t)s * Np + i];
                    const double4 P = AF[fcode & 0x7FFFFFFFu];
                    if (fcode >> 31) { acc[0] += P.x; acc[1] += P.y; acc[2] += P.z; acc[3] += P.w; }
                    else             { acc[0] -= P.x; acc[1] -= P.y; acc[2] -= P.z; acc[3] -= P.w; }
                }
            }
        }
        const double vol = a.g.cell_vol[i];
#pragma unroll
        for (int v = 0; v < 4; v++) k[v] = acc[v] / vol;
    }
    if (a.rk.k_store) st4(a.rk.k_store, i, k);
    if (a.rk.mode == 3) return;
    const double dt = a.scal[SC_DT];
    if (dt < 0.0) {   // the reference throws in calc_dt BEFORE take_step (solver.cpp:587-589): a step with a negative dt leaves the
        if (a.rk.last_stage) {   // solution, the time and the step counter untouched; the host reports the error when it next reads dt
            double b[4];
            ld4(a.rk.base, i, b);
            st4(a.rk.out, i, b);
        }
        return;
    }
    double Unew[4];
    rk_update(a.rk, a.Uin, i, k, dt, Unew);
    if (a.rk.last_stage) {
        if (a.rk.prim_out) {   // Solver::update_primitives solver.cpp:533-578
            double P[5];
            cons_to_prim(a.ph.gas, Unew, P);
#pragma unroll
            for (int v = 0; v < 5; v++) a.rk.prim_out[(size_t)v * Np + i] = P[v];
            a.rk.prim_out[5 * (size_t)Np + i] = Unew[0];
        }
        if (i == 0) { a.scal[SC_T] = a.scal[SC_T] + dt; *a.step_counter += 1ull; }   // solver.cpp:529-530
    }
}
__global__ void __launch_bounds__(256) gather_stage_kernel(const __grid_constant__ StageArgs a) {
    gather_stage_body(a, blockIdx.x * blockDim.x + threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------
// TENO reconstruction — TENOFunctor::operator() (numerics/face_reconstruction.cpp:866-1039).
// Thread = (cell, conserved variable); the four variable-threads of a cell share every table address (broadcast) and
// the eight cells of a warp read eight adjacent entries of the tile-interleaved tables (fully used 32 B sectors,
// 128 B per double2 request).  Each thread performs the reference's sums in the reference's order.
// ---------------------------------------------------------------------------------------------------------------
template <int ORDER>
__device__ __forceinline__ void legendre_values(double x, double * P) {   // basis.h:81-86, same expression shapes
    P[0] = 1.0 * (1.0);
    if (ORDER >= 1) P[1] = 1.0 * (1.0 * x);
    if (ORDER >= 2) P[2] = 0.5 * (3.0 * x * x - 1.0);
    if (ORDER >= 3) P[3] = 0.5 * (5.0 * x * x * x - 3.0 * x);
    if (ORDER >= 4) P[4] = 0.125 * (35.0 * x * x * x * x - 30.0 * x * x + 3.0);
}

// 1-D basis values psi_0..psi_ORDER at x: Legendre as above, or the reference's default monomials (basis.h:66-70,
// Kokkos::pow(x, p): libm pow in STRICT mode, repeated multiplication in FAST mode)
template <int ORDER>
__device__ __forceinline__ void basis_values(int basis, double x, double * P) {
    if (basis == MLB_BASIS_LEGENDRE) { legendre_values<ORDER>(x, P); return; }
#ifdef MLB_STREAM_KERNELS
    P[0] = 1.0;
#pragma unroll
    for (int d = 1; d <= ORDER; d++) P[d] = P[d - 1] * x;
#else
#pragma unroll
    for (int d = 0; d <= ORDER; d++) P[d] = pow(x, (double)d);
#endif
}

constexpr int RECON_THREADS = 128;

// exponents of the k-th basis function in the reference's graded ordering (p,0),(p-1,1),...,(0,p)
// (TENO::calc_polynomial_indices, face_reconstruction.cpp:182-215)
__host__ __device__ constexpr int dof_ey(int k) { int p = 0; while (k > p) { k -= p + 1; p++; } return k; }
__host__ __device__ constexpr int dof_ex(int k) { int p = 0; while (k > p) { k -= p + 1; p++; } return p - k; }

// out += w * dof_k * (psi_k(x, y) + cbar_k) for k = 0..K-1 in order (:1021-1030), exponents resolved at compile time
template <int K, int... Ks>
__device__ __forceinline__ double poly_accumulate(double out, double w, const double * dof_sm_tid, const double * Px, const double * Py,
                                                  const double * cbar, std::integer_sequence<int, Ks...>) {
    ((out += w * dof_sm_tid[Ks * RECON_THREADS] * (Px[std::integral_constant<int, dof_ex(Ks)>::value] * Py[std::integral_constant<int, dof_ey(Ks)>::value] + cbar[Ks])), ...);
    return out;
}

template <int ORDER, int MP>
__global__ void __launch_bounds__(RECON_THREADS) teno_recon_kernel(const __grid_constant__ ReconArgs a) {
    constexpr int K = (ORDER + 1) * (ORDER + 2) / 2;
    constexpr int MAXS = 1 + MAX_SLOTS;
    MLB_DYNAMIC_SMEM(double, dof_sm);    // [S][K][RECON_THREADS]
    const int tid = threadIdx.x;
    const uint32_t cell = blockIdx.x * (RECON_THREADS / 4) + (tid >> 2);
    const int var = tid & 3;
    if (cell >= a.g.N_recon) return;
    const uint32_t Np = a.g.Npad;
    const size_t tile = cell / TILE;
    const int lane = cell % TILE;
    const int S = a.S;
    const double * Uv = a.Uin + var;                                     // AoS [cell][4]: Uv[4 * c] is this thread's variable
    const double u_self = Uv[4 * (size_t)cell];

    double w[MAXS];
    double area0 = 0.5;
#pragma unroll 1
    for (int s = 0; s < S; s++) {
        const size_t sbase = (tile * S + s) * MP;
        const uint32_t * ids = a.st_ids + sbase * TILE + lane;
        if (ids[0] == NO_FACE) { w[s] = 0.0; continue; }               // empty stencil :896-899
        const double * areas = a.st_area + sbase * TILE + lane;
        double b[MP];
#pragma unroll
        for (int m = 0; m < MP; m++) b[m] = areas[m * TILE] * (Uv[4 * (size_t)ids[m * TILE]] - u_self);   // :903-910
        if (s == 0) area0 = areas[0];
        const double2 * mat = reinterpret_cast<const double2 *>(a.st_mat) + ((tile * S + s) * K * (MP / 2)) * TILE + lane;
        double dof[K];
#pragma unroll
        for (int k = 0; k < K; k++) {                                   // a = A+ b, k-ascending sums :915-918
            double sum = 0.0;
#pragma unroll
            for (int m2 = 0; m2 < MP / 2; m2++) {
                const double2 c = mat[(k * (MP / 2) + m2) * TILE];
                sum += c.x * b[2 * m2];
                sum += c.y * b[2 * m2 + 1];
            }
            dof[k] = sum;
            dof_sm[(s * K + k) * RECON_THREADS + tid] = sum;
        }
        double si = 0.0;                                                // SI = a . (OI a) :922-936
        double oa[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < K; j++) t += a.OI[k * K + j] * dof[j];
            oa[k] = t;
        }
#pragma unroll
        for (int k = 0; k < K; k++) si += dof[k] * oa[k];
        const double x = si + 1.0e-12;                                  // 1/(SI+eps)^6 :940-944 (pow → 3 multiplications)
        const double x2 = x * x, x3 = x2 * x;
        w[s] = 1.0 / (x3 * x3);
    }

    // non-linear weights :948-981 (reference-faithful: the central weight stays raw in the ENO branch)
    {
        double sd = 0.0;
        for (int s = 1; s < S; s++) sd += w[s];
        if (w[0] / (sd + w[0]) > 1.0e-7) {
            w[0] = 1.0;
            for (int s = 1; s < S; s++) w[s] = 0.0;
        } else {
            for (int s = 1; s < S; s++) {
                if (w[s] / sd > 1.0e-5) w[s] = (1.0 / K);
                else if (a.fixed_weights) w[s] = 0.0;
            }
            sd = 0.0;
            for (int s = 1; s < S; s++) sd += w[s];
            for (int s = 1; s < S; s++) w[s] /= sd;
            if (a.fixed_weights) w[0] = 0.0;
        }
    }

    double cbar[K];                                                     // psi_bar_k / area_t[s][0] :1028-1029
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double pb = a.psi_bar_cell ? a.psi_bar_cell[(size_t)cell * K + k] : a.psi_bar[k];
        cbar[k] = a.fixed_weights ? -pb : pb / area0;   // fixed: mean-free basis
    }

    const int nf = a.g.nfc[cell];
    const int Q = a.g.Q;
    for (int j = 0; j < nf; j++) {                                      // :985-1034
        const double * fx = a.g.slot_fx + ((size_t)j * 4) * Np + cell;
        const double x0 = fx[0], y0 = fx[Np], x1 = fx[2 * (size_t)Np], y1 = fx[3 * (size_t)Np];
        for (int q = 0; q < Q; q++) {
            const double tq = (a.qf_x[q] + 1.0) * 0.5;
            const double xq = tq * (x1 - x0) + x0, yq = tq * (y1 - y0) + y0;
            double Px[ORDER + 1], Py[ORDER + 1];
            basis_values<ORDER>(a.basis, xq, Px);
            basis_values<ORDER>(a.basis, yq, Py);
            double out = u_self;
#pragma unroll 1
            for (int s = 0; s < S; s++) {
                if (w[s] == 0.0) continue;
                out = poly_accumulate<K>(out, w[s], dof_sm + (size_t)(s * K) * RECON_THREADS + tid, Px, Py, cbar, std::make_integer_sequence<int, K>{});
            }
            a.Fc[((size_t)cell * (a.g.n_slots * Q) + (j * Q + q)) * 4 + var] = out;
        }
    }
}

#ifndef MLB_HOST_EMULATION           // the streaming kernels are PTX (TMA, mbarriers): not part of the host emulation
#ifdef MLB_STREAM_KERNELS
#include "teno_stream.cuh"
#else
#include "teno_strict_stream.cuh"
#endif
#endif
#include "teno_generic.cuh"

// ---------------------------------------------------------------------------------------------------------------
// Spectral radius + max + dt — SpectralRadiusFunctor / Solver::calc_dt (solver/solver.cpp:580-742)
// ---------------------------------------------------------------------------------------------------------------
// spectral radius of cell i (stored as cfl_local before scaling); -1 for i >= N_owned
__device__ __forceinline__ double spectral_radius_body(const CflArgs & a, const uint32_t i) {
    const uint32_t Np = a.g.Npad;
    double sr = -1.0;
    if (i < a.g.N_owned) {
        double conv = 0.0, acou = 0.0, visc = 0.0;
        const int nf = a.g.nfc[i];
        const double rho_s = a.prim[5 * (size_t)Np + i], u_s = a.prim[i], v_s = a.prim[(size_t)Np + i], p_s = a.prim[2 * (size_t)Np + i];
        const double sos_s = sqrt(a.gas.gamma * p_s / rho_s);
        for (int j = 0; j < nf; j++) {
            const uint32_t fcode = a.g.slot_face[(size_t)j * Np + i];
            const uint32_t f = fcode & 0x7FFFFFFFu, side = fcode >> 31;
            const double nx = a.g.face_nx[f], ny = a.g.face_ny[f];
            const int32_t nbr = a.g.slot_nbr[(size_t)j * Np + i];
            double sx, sy, sos_l, sos_r, ul, vl, ur, vr;
            if (nbr < 0) {   // boundary "hack" :662-666 — l = r = this cell
                sx = a.g.bnd_s[i]; sy = sx;
                sos_l = sos_s; sos_r = sos_s; ul = u_s; vl = v_s; ur = u_s; vr = v_s;
            } else {
                const double rho_n = a.prim[5 * (size_t)Np + nbr], u_n = a.prim[nbr], v_n = a.prim[(size_t)Np + nbr], p_n = a.prim[2 * (size_t)Np + nbr];
                const double sos_n = sqrt(a.gas.gamma * p_n / rho_n);
                const double dxs = a.g.cell_xy[i], dys = a.g.cell_xy[(size_t)Np + i];
                const double dxn = a.g.cell_xy[nbr], dyn = a.g.cell_xy[(size_t)Np + nbr];
                if (side == 0) { sx = dxn - dxs; sy = dyn - dys; sos_l = sos_s; sos_r = sos_n; ul = u_s; vl = v_s; ur = u_n; vr = v_n; }
                else           { sx = dxs - dxn; sy = dys - dyn; sos_l = sos_n; sos_r = sos_s; ul = u_n; vl = v_n; ur = u_s; vr = v_s; }
            }
            const double dx_n = fabs(sx * nx + sy * ny);
            const double sos_f = 0.5 * (sos_l + sos_r);
            const double uf = 0.5 * (ul + ur), vf = 0.5 * (vl + vr);
            const double un = fabs(uf * nx + vf * ny);
            conv += un / dx_n;
            const double r = sos_f / dx_n;
            acou += r * r;
            if (a.gas.mu > 0.0) visc += 1.0 / (dx_n * dx_n);
        }
        const double geom = 3.0 / nf;
        conv *= 1.37 * geom;
        acou = 1.37 * sqrt(geom * acou);
        sr = conv + acou;
        // new (the reference declares spectral_radius_viscous / _heat and leaves them commented out, solver.cpp:638-651): diffusion
        // number of the stiffer of momentum and heat diffusion, 2 nu' sum_faces 1/dx_n^2 with the same geometric factor
        if (a.gas.mu > 0.0) sr += 2.0 * geom * visc * fmax(4.0 / 3.0, a.gas.gamma * a.gas.kappa / (a.gas.mu * a.gas.cp)) * a.gas.mu / rho_s;
        a.sr_out[i] = sr;
    }
    return sr;
}

#include "small_step.cuh"

#ifndef MLB_HOST_EMULATION           // from here on: shared memory, atomics, launch syntax
__global__ void __launch_bounds__(256) cfl_kernel(const __grid_constant__ CflArgs a) {
    const double sr = spectral_radius_body(a, blockIdx.x * blockDim.x + threadIdx.x);
    // block max (NaN never wins: Kokkos::Max joins with `<`)
    __shared__ double red[256];
    red[threadIdx.x] = (sr == sr) ? sr : -1.0;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicMax(a.max_bits, __double_as_longlong(red[0]));
        __threadfence();
        const unsigned int done = atomicAdd(a.blocks_done, 1u);
        if (done == gridDim.x - 1) {   // last block: publish max and dt, reset the accumulators
            const double mx = __longlong_as_double(atomicAdd((unsigned long long *)a.max_bits, 0ull));
            a.scal[SC_MAX_SR] = mx;
            if (a.cfl > 0.0) { a.scal[SC_DT] = a.cfl / mx; a.scal[SC_CFL] = a.cfl; }
            *a.max_bits = __double_as_longlong(-1.0);
            *a.blocks_done = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Stateless kernels for the plug-in level known-answer tests
// ---------------------------------------------------------------------------------------------------------------
template <int RS>
__global__ void riemann_kernel(uint64_t n, const double * nunit, const double * L, const double * R, double gamma, double * flux) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FaceState l = {L[5 * i], L[5 * i + 1], L[5 * i + 2], L[5 * i + 3], L[5 * i + 4]};
    const FaceState r = {R[5 * i], R[5 * i + 1], R[5 * i + 2], R[5 * i + 3], R[5 * i + 4]};
    double F[4];
#ifdef MLB_STREAM_KERNELS
    riemann_flux_lean<RS>(F, nunit[2 * i], nunit[2 * i + 1], face_cons(l), face_cons(r), gamma);
#else
    riemann_flux<RS>(F, nunit[2 * i], nunit[2 * i + 1], l, r, gamma);
#endif
    for (int v = 0; v < 4; v++) flux[4 * i + v] = F[v];
}

__global__ void prims_aos_kernel(const GasParams g, uint64_t n, const double * U, double * P) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u[4] = {U[4 * i], U[4 * i + 1], U[4 * i + 2], U[4 * i + 3]}, p[5];
    cons_to_prim(g, u, p);
    for (int v = 0; v < 5; v++) P[5 * i + v] = p[v];
}

// primitives (SoA [6][npad]: u, v, p, T, h, rho) of AoS states
__global__ void prims_soa_kernel(const GasParams g, uint32_t n, uint32_t npad, const double * U, double * P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u[4], p[5];
    ld4(U, i, u);
    cons_to_prim(g, u, p);
    for (int v = 0; v < 5; v++) P[(size_t)v * npad + i] = p[v];
    P[5 * (size_t)npad + i] = u[0];
}

// ---------------------------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------------------------
template <int RS, bool VISC>
static void launch_faces_rs(const StageArgs & a, cudaStream_t st) {
    if (a.g.NF == 0) return;
    auto grid = [&](unsigned tpf) { return (unsigned)(((uint64_t)a.g.NF * tpf + 127u) / 128u); };
    if (!a.teno) face_flux_kernel<RS, false, 1, VISC><<<grid(1), 128, 0, st>>>(a);
    else if (a.g.Q == 1) face_flux_kernel<RS, true, 1, VISC><<<grid(1), 128, 0, st>>>(a);
    else if (a.g.Q == 2) face_flux_kernel<RS, true, 2, VISC><<<grid(2), 128, 0, st>>>(a);
    else face_flux_kernel<RS, true, 0, VISC><<<grid(1), 128, 0, st>>>(a);
}
static void launch_faces(const StageArgs & a, cudaStream_t st) {
    if (a.G) {
        switch (a.ph.riemann) {
            case MLB_RIEMANN_RUSANOV: launch_faces_rs<MLB_RIEMANN_RUSANOV, true>(a, st); break;
            case MLB_RIEMANN_HLL: launch_faces_rs<MLB_RIEMANN_HLL, true>(a, st); break;
            default: launch_faces_rs<MLB_RIEMANN_HLLC, true>(a, st); break;
        }
        return;
    }
    switch (a.ph.riemann) {
        case MLB_RIEMANN_RUSANOV: launch_faces_rs<MLB_RIEMANN_RUSANOV, false>(a, st); break;
        case MLB_RIEMANN_HLL: launch_faces_rs<MLB_RIEMANN_HLL, false>(a, st); break;
        default: launch_faces_rs<MLB_RIEMANN_HLLC, false>(a, st); break;
    }
}
static void launch_gradients(const StageArgs & a, cudaStream_t st) {
    const unsigned grid = (a.g.N_recon + 255u) / 256u;
    if (grid && a.G) visc_grad_kernel<<<grid, 256, 0, st>>>(a);
}
static void launch_stage(const StageArgs & a, cudaStream_t st) {
    const unsigned grid = (a.g.N_owned + 255u) / 256u;
    if (grid) gather_stage_kernel<<<grid, 256, 0, st>>>(a);
}

template <int ORDER, int MP>
static void launch_recon_t(const ReconArgs & a, cudaStream_t st) {
    constexpr int K = (ORDER + 1) * (ORDER + 2) / 2;
    const size_t smem = (size_t)a.S * K * RECON_THREADS * sizeof(double);
    ensure_dynamic_smem(reinterpret_cast<const void *>(teno_recon_kernel<ORDER, MP>), (size_t)(1 + MAX_SLOTS) * K * RECON_THREADS * sizeof(double));
    const unsigned cells_per_block = RECON_THREADS / 4;
    const unsigned grid = (a.g.N_recon + cells_per_block - 1) / cells_per_block;
    if (grid == 0) return;
    teno_recon_kernel<ORDER, MP><<<grid, RECON_THREADS, smem, st>>>(a);
}
static bool recon_specialised(int order, int Mp, int S) {
    return S <= 1 + MAX_SLOTS && ((order == 1 && Mp == 6) || (order == 2 && Mp == 12) || (order == 3 && Mp == 20) || (order == 4 && Mp == 30));
}
static bool recon_supported(int order, int K, int Mp, int S, int basis) {
    if (basis != MLB_BASIS_LEGENDRE && basis != MLB_BASIS_MONOMIAL) return false;
    return recon_specialised(order, Mp, S) || generic::generic_supported(order, K, Mp, S);
}
static void launch_recon(const ReconArgs & a, cudaStream_t st) {
    const char * fg = getenv("MLB_TENO_GENERIC");               // test hook: run the generic kernel where a specialised one exists
    const bool force_generic = fg && fg[0] == '1';
#ifndef MLB_STREAM_KERNELS
    if (!force_generic && sstream::strict_stream_supported(a)) {       // bit-faithful AND streaming (teno_strict_stream.cuh); same results as below
        if (a.order == 1 && a.Mp == 6) return sstream::launch_strict_stream<1, 6>(a, st);
        if (a.order == 2 && a.Mp == 12) return sstream::launch_strict_stream<2, 12>(a, st);
        if (a.order == 3 && a.Mp == 20) return sstream::launch_strict_stream<3, 20>(a, st);
        if (a.order == 4 && a.Mp == 30) return sstream::launch_strict_stream<4, 30>(a, st);
    }
#endif
    if (force_generic || !recon_specialised(a.order, a.Mp, a.S)) return generic::launch_generic(a, st);
    if (a.order == 1 && a.Mp == 6) launch_recon_t<1, 6>(a, st);
    else if (a.order == 2 && a.Mp == 12) launch_recon_t<2, 12>(a, st);
    else if (a.order == 3 && a.Mp == 20) launch_recon_t<3, 20>(a, st);
    else if (a.order == 4 && a.Mp == 30) launch_recon_t<4, 30>(a, st);
}

static void launch_cfl(const CflArgs & a, cudaStream_t st) {
    const unsigned grid = (a.g.N_owned + 255u) / 256u;
    if (grid == 0) return;
    cfl_kernel<<<grid, 256, 0, st>>>(a);
}

static void launch_riemann(int riemann, uint64_t n, const double * nunit, const double * L, const double * R, double gamma,
                           double * flux, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (!grid) return;
    if (riemann == MLB_RIEMANN_RUSANOV) riemann_kernel<MLB_RIEMANN_RUSANOV><<<grid, 128, 0, st>>>(n, nunit, L, R, gamma, flux);
    else if (riemann == MLB_RIEMANN_HLL) riemann_kernel<MLB_RIEMANN_HLL><<<grid, 128, 0, st>>>(n, nunit, L, R, gamma, flux);
    else riemann_kernel<MLB_RIEMANN_HLLC><<<grid, 128, 0, st>>>(n, nunit, L, R, gamma, flux);
}
static void launch_prims(const GasParams & g, uint64_t n, const double * U, double * P, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (grid) prims_aos_kernel<<<grid, 128, 0, st>>>(g, n, U, P);
}
static void launch_prims_soa(const GasParams & g, uint32_t n, uint32_t npad, const double * U, double * P, cudaStream_t st) {
    const unsigned grid = (n + 127u) / 128u;
    if (grid) prims_soa_kernel<<<grid, 128, 0, st>>>(g, n, npad, U, P);
}

#define MLB_STR2(x) #x
#define MLB_STR(x) MLB_STR2(x)
#ifdef MLB_STREAM_KERNELS
static const KernelTable table = {MLB_STR(MLB_KNS), launch_gradients, launch_faces, launch_stage, launch_recon, launch_cfl, launch_riemann, launch_prims,
                                  launch_prims_soa, recon_supported, stream::launch_stream, stream::stream_supported, small::launch_small_step};
#else
static const KernelTable table = {MLB_STR(MLB_KNS), launch_gradients, launch_faces, launch_stage, launch_recon, launch_cfl, launch_riemann, launch_prims,
                                  launch_prims_soa, recon_supported, nullptr, nullptr, small::launch_small_step};
#endif

#endif  // MLB_HOST_EMULATION

}  // namespace MLB_KNS
}  // namespace mlb
