// Mesh preprocessor (host, OpenMP): turns the reference's mesh arrays into the device layout.
//   1. cell renumbering for locality (reverse Cuthill–McKee on the cell–face dual graph) and face ordering by owner;
//   2. ELL ("slot") flattening of cell→face→neighbour connectivity with the boundary-condition binding and the
//      reference's residual accumulation order encoded per cell;
//   3. TENO stencil search (NCB + Type-4 directional, numerics/face_reconstruction.cpp:217-475) and pseudo-inverse
//      reconstruction matrices (:477-741), oscillation-indicator matrix (:743-788), written straight into the
//      tile-interleaved tables the reconstruction kernel streams.
// The arithmetic follows the reference operation by operation (this file is compiled with -ffp-contract=off), so the
// tables are bit-identical to the reference's; only the loop structure differs (cells in parallel, Householder updates
// restricted to the rows/columns they can change).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <numeric>
#include <queue>
#include <unordered_map>

#include "mlb_internal.h"

namespace mlb {

// ----------------------------------------------------------------------------------------------------------------
// Quadrature tables — published constants, carried to the digits the reference uses (numerics/quadrature.cpp).
// ----------------------------------------------------------------------------------------------------------------
void gauss_legendre_rule(int order, dvec & x, dvec & w) {
    static const double X2 = 0.577350269189625764509149;
    static const double X3 = 0.774596669241483377035853;
    static const double X4a = 0.861136311594052575223946, X4b = 0.339981043584856264802666;
    static const double X5a = 0.906179845938663992797627, X5b = 0.538469310105683091036314;
    static const double X6a = 0.932469514203152027812302, X6b = 0.661209386466264513661400, X6c = 0.238619186083196908630502;
    static const double X7a = 0.949107912342758524526190, X7b = 0.741531185599394439863865, X7c = 0.405845151377397166906607;
    switch (order) {
        case 1: x = {0.0}; w = {2.0}; break;
        case 2: x = {-X2, X2}; w = {1.0, 1.0}; break;
        case 3: x = {-X3, 0.0, X3}; w = {0.55555555555555555555556, 0.88888888888888888888889, 0.55555555555555555555556}; break;
        case 4: x = {-X4a, -X4b, X4b, X4a};
                w = {0.34785484513745385737306, 0.65214515486254614262694, 0.65214515486254614262694, 0.34785484513745385737306}; break;
        case 5: x = {-X5a, -X5b, 0.0, X5b, X5a};
                w = {0.23692688505618908751426, 0.47862867049936646804129, 0.56888888888888888888889, 0.47862867049936646804129,
                     0.23692688505618908751426}; break;
        case 6: x = {-X6a, -X6b, -X6c, X6c, X6b, X6a};
                w = {0.17132449237917034504030, 0.36076157304813860756983, 0.46791393457269104738987, 0.46791393457269104738987,
                     0.36076157304813860756983, 0.17132449237917034504030}; break;
        case 7: x = {-X7a, -X7b, -X7c, 0.0, X7c, X7b, X7a};
                w = {0.12948496616886969327061, 0.27970539148927666790147, 0.38183005050511894495037, 0.41795918367346938775510,
                     0.38183005050511894495037, 0.27970539148927666790147, 0.12948496616886969327061}; break;
        default: throw std::runtime_error("Gauss-Legendre quadrature rule of order " + std::to_string(order) + " not implemented.");
    }
}

void dunavant_rule(int order, dvec & xy, dvec & w) {
    switch (order) {
        case 1: xy = {1.0 / 3.0, 1.0 / 3.0}; w = {1.0}; break;
        case 2: xy = {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0}; w = {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0}; break;
        case 3: {
            const double b = 0.52083333333333333333333333333333;
            xy = {1.0 / 3.0, 1.0 / 3.0, 0.6, 0.2, 0.2, 0.6, 0.2, 0.2}; w = {-0.5625, b, b, b}; break;
        }
        case 4: {
            const double a1 = 0.108103018168070, b1 = 0.445948490915965, a2 = 0.816847572980459, b2 = 0.091576213509771;
            const double w1 = 0.223381589678011, w2 = 0.109951743655322;
            xy = {a1, b1, b1, a1, b1, b1, a2, b2, b2, a2, b2, b2}; w = {w1, w1, w1, w2, w2, w2}; break;
        }
        case 5: {
            const double c = 0.333333333333333, a1 = 0.059715871789770, b1 = 0.470142064105115, a2 = 0.797426985353087, b2 = 0.101286507323456;
            const double w0 = 0.225000000000000, w1 = 0.132394152788506, w2 = 0.125939180544827;
            xy = {c, c, a1, b1, b1, a1, b1, b1, a2, b2, b2, a2, b2, b2}; w = {w0, w1, w1, w1, w2, w2, w2}; break;
        }
        default: throw std::runtime_error("Triangle Dunavant quadrature rule of order " + std::to_string(order) + " not implemented.");
    }
}

namespace {

// ----------------------------------------------------------------------------------------------------------------
// 1-D bases (numerics/basis.h:66-176).  Legendre polynomials as coefficient tables; every monomial is built by
// repeated multiplication from the coefficient and the terms are summed left to right, which is the rounding order of
// the reference's expanded expressions.
// ----------------------------------------------------------------------------------------------------------------
struct PolyTerm { double c; int e; };
struct Poly { double scale; int nt; PolyTerm t[5]; };
const Poly LEGENDRE[10] = {
    {1.0, 1, {{1.0, 0}}},
    {1.0, 1, {{1.0, 1}}},
    {0.5, 2, {{3.0, 2}, {-1.0, 0}}},
    {0.5, 2, {{5.0, 3}, {-3.0, 1}}},
    {0.125, 3, {{35.0, 4}, {-30.0, 2}, {3.0, 0}}},
    {0.125, 3, {{63.0, 5}, {-70.0, 3}, {15.0, 1}}},
    {0.0625, 4, {{231.0, 6}, {-315.0, 4}, {105.0, 2}, {-5.0, 0}}},
    {0.0625, 4, {{429.0, 7}, {-693.0, 5}, {315.0, 3}, {-35.0, 1}}},
    {0.0078125, 5, {{6435.0, 8}, {-12012.0, 6}, {6930.0, 4}, {-1260.0, 2}, {35.0, 0}}},
    {0.0078125, 5, {{12155.0, 9}, {-25740.0, 7}, {18018.0, 5}, {-4620.0, 3}, {315.0, 1}}},
};

double legendre(int deriv, int p, double x) {
    if (deriv > p) return 0.0;
    const Poly & P = LEGENDRE[p];
    double acc = 0.0;
    bool have = false;
    for (int j = 0; j < P.nt; j++) {
        const int e = P.t[j].e;
        if (e < deriv) continue;
        // reference quirk (numerics/basis.h:122): the hard-coded d/dx P9 stops at "- 4620.0*3.0*x*x" - the constant 315.0*1.0 that the
        // 315 x term would leave is missing.  Reproduced: the oscillation-indicator matrix of basis_order 9 is built from these
        // derivatives (:743-788) and must be the reference's bit for bit (found by pinning against a dump of the reference at order 9).
        if (p == 9 && deriv == 1 && e == 1) continue;
        double term = P.t[j].c;
        for (int k = 0; k < deriv; k++) term *= (double)(e - k);
        for (int k = 0; k < e - deriv; k++) term *= x;
        acc = have ? acc + term : term;
        have = true;
    }
    return P.scale * acc;
}

double monomial(int deriv, int p, double x) {   // incl. the reference's derivative quirk (basis.h:72-78)
    if (deriv == 0) return std::pow(x, (double)p);
    double r = std::pow(x, (double)(p - deriv));
    for (int k = 0; k < deriv; k++) r *= (p - k + 1);
    return r;
}

inline double basis_1d(int type, int deriv, int p, double x) {
    return type == MLB_BASIS_MONOMIAL ? monomial(deriv, p, x) : legendre(deriv, p, x);
}

inline void mat2_inverse(const double * A, double * Ai) {   // common_math.h:146-160
    const double det = A[0] * A[3] - A[1] * A[2];
    if (det == 0.0) throw std::runtime_error("Matrix is singular.");
    const double r = 1.0 / det;
    Ai[0] = A[3] * r; Ai[1] = -A[1] * r; Ai[2] = -A[2] * r; Ai[3] = A[0] * r;
}
inline void mat2_apply(const double * A, const double * x, double * y) {   // common_math.h:207-213 (aliasing-safe)
    const double y0 = A[0] * x[0] + A[1] * x[1], y1 = A[2] * x[0] + A[3] * x[1];
    y[0] = y0; y[1] = y1;
}
inline void edge_frame(const double * o, const double * a, const double * b, double * J) {   // triangle_J, common_math.h:534-542
    J[0] = a[0] - o[0]; J[1] = b[0] - o[0]; J[2] = a[1] - o[1]; J[3] = b[1] - o[1];
}
inline double tri_area2(const double * a, const double * b, const double * c) {
    return 0.5 * std::fabs(a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1]));
}

// Upper-triangular factor of a rows x cols matrix by Householder reflections (common_math.h:355-422).  The reference
// forms each reflector densely and multiplies the whole matrix; only rows >= j of columns >= j can change, and the
// skipped terms are exact zeros, so restricting the update leaves every stored bit unchanged.
void householder_R(double * R, int rows, int cols, double * v, double * colbuf) {
    for (int j = 0; j < cols; j++) {
        double nrm = 0.0;
        for (int i = j; i < rows; i++) nrm += R[i * cols + j] * R[i * cols + j];
        nrm = std::sqrt(nrm);
        if (nrm < 1.0e-15) continue;
        const double sgn = (R[j * cols + j] >= 0.0) ? 1.0 : -1.0;
        const double alpha = -sgn * nrm;
        const int len = rows - j;
        double nu = 0.0;
        for (int k = 0; k < len; k++) {
            v[k] = R[(j + k) * cols + j];
            if (k == 0) v[k] -= alpha;
            nu += v[k] * v[k];
        }
        nu = std::sqrt(nu);
        for (int k = 0; k < len; k++) v[k] /= nu;
        for (int c = j; c < cols; c++) {
            for (int i = 0; i < len; i++) {
                double s = 0.0;
                for (int k = 0; k < len; k++) {
                    const double q = ((i == k) ? 1.0 : 0.0) - 2.0 * v[i] * v[k];
                    s += q * R[(j + k) * cols + c];
                }
                colbuf[i] = s;
            }
            for (int i = 0; i < len; i++) R[(j + i) * cols + c] = colbuf[i];
        }
    }
}

// Solve R^T Y = B^T (forward, common_math.h:437-457 with tL=tB=true) then R X = Y (backward, :472-493); B is rows x cols
// row-major, R its cols x cols upper factor stored with row length cols; X (cols x rows) is the pseudo-inverse of B.
void pseudo_inverse_from_R(const double * R, const double * B, int rows, int cols, double * Y, double * X) {
    for (int i = 0; i < cols; i++)
        for (int j = 0; j < rows; j++) {
            double s = 0.0;
            for (int k = 0; k < i; k++) s += R[k * cols + i] * Y[k * rows + j];
            Y[i * rows + j] = (B[j * cols + i] - s) / R[i * cols + i];
        }
    for (int i = cols - 1; i >= 0; i--)
        for (int j = 0; j < rows; j++) {
            double s = 0.0;
            for (int k = i + 1; k < cols; k++) s += R[i * cols + k] * X[k * rows + j];
            X[i * rows + j] = (Y[i * rows + j] - s) / R[i * cols + i];
        }
}

// ----------------------------------------------------------------------------------------------------------------
// Stencil search
// ----------------------------------------------------------------------------------------------------------------
struct RingSearch {
    const HostMesh & m;
    uint32_t target;
    std::vector<uvec> rings;
    RingSearch(const HostMesh & mesh, uint32_t t, const uvec & seed) : m(mesh), target(t) { rings.push_back(seed); }

    bool seen(uint32_t c) const {
        for (auto & r : rings) if (std::find(r.begin(), r.end(), c) != r.end()) return true;
        return false;
    }
    // face_reconstruction.cpp:217-265
    const uvec & grow() {
        uvec cand;
        uint32_t nb[MAX_SLOTS + 1];
        for (uint32_t c : rings.back()) {
            int n = 0;
            nb[n++] = c;
            for (int j = 0; j < m.nfc(c); j++) {
                const uint32_t f = m.foc[m.ofc[c] + j];
                const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
                if (b == -1) continue;
                if (b == CUT_FACE) throw std::runtime_error("rank-local mesh: the stencil search reached the cut (the mesh needs more ghost layers)");
                nb[n++] = (a == (int32_t)c) ? (uint32_t)b : (uint32_t)a;
            }
            for (int a = 1; a < n; a++)                  // Mesh::h_neighbors_of_cell: sorted, unique (mesh.cpp:158-165)
                for (int b = a; b > 0 && nb[b - 1] > nb[b]; b--) std::swap(nb[b - 1], nb[b]);
            n = (int)(std::unique(nb, nb + n) - nb);
            for (int i = 0; i < n; i++)
                if (!seen(nb[i]) && std::find(cand.begin(), cand.end(), nb[i]) == cand.end()) cand.push_back(nb[i]);
        }
        std::vector<std::pair<uint32_t, double>> byd;
        byd.reserve(cand.size());
        for (uint32_t c : cand) {
            double d2 = 0.0;
            for (int i = 0; i < 2; i++) {
                const double d = m.cell_xy[2 * (size_t)c + i] - m.cell_xy[2 * (size_t)target + i];
                d2 += d * d;
            }
            byd.push_back({c, d2});
        }
        // std::sort with the reference's comparator: identical tie order on the same libstdc++
        std::sort(byd.begin(), byd.end(), [](auto & l, auto & r) { return l.second < r.second; });
        uvec ring;
        ring.reserve(byd.size());
        for (auto & p : byd) ring.push_back(p.first);
        rings.push_back(std::move(ring));
        return rings.back();
    }
};

void fill_outwards(const HostMesh & m, uint32_t cell, uvec & st, int M) {   // :275-284 and :386-401
    RingSearch rs(m, cell, st);
    while ((int)st.size() < M) {
        const uvec & ring = rs.grow();
        if (ring.empty()) throw std::runtime_error("TENO: mesh too small to fill a stencil of " + std::to_string(M) + " cells");
        for (uint32_t x : ring) { if ((int)st.size() == M) break; st.push_back(x); }
    }
}

// stencils[0] = centred (NCB), stencils[1 + j] = directional stencil of face slot j (empty for boundary faces)
void cell_stencils(const HostMesh & m, uint32_t c, int M, std::vector<uvec> & out) {
    const int ns = m.nfc(c);
    out.assign(1 + ns, uvec());
    out[0] = {c};
    fill_outwards(m, c, out[0], M);

    double Jinv[MAX_SLOTS][4];
    bool bface[MAX_SLOTS];
    for (int s = 0; s < ns; s++) {
        const uint32_t f = m.foc[m.ofc[c] + s];
        double J[4];
        edge_frame(&m.cell_xy[2 * (size_t)c], &m.node_xy[2 * (size_t)m.nof[m.onf[f]]], &m.node_xy[2 * (size_t)m.nof[m.onf[f] + 1]], J);
        mat2_inverse(J, Jinv[s]);
        bface[s] = m.cof[2 * (size_t)f + 1] == -1;
        out[1 + s] = {c};
    }
    RingSearch rs(m, c, uvec{c});
    bool grew[MAX_SLOTS];
    for (bool done = false; !done;) {   // :320-375
        const uvec & ring = rs.grow();
        for (int s = 0; s < ns; s++) {
            grew[s] = false;
            uvec & st = out[1 + s];
            for (uint32_t x : ring) {
                if ((int)st.size() == M) break;
                const double d[2] = {m.cell_xy[2 * (size_t)x] - m.cell_xy[2 * (size_t)c], m.cell_xy[2 * (size_t)x + 1] - m.cell_xy[2 * (size_t)c + 1]};
                double t[2];
                mat2_apply(Jinv[s], d, t);
                if (!(t[0] < 0.0) && !(t[1] < 0.0)) { st.push_back(x); grew[s] = true; }
            }
        }
        done = true;
        for (int s = 0; s < ns; s++)
            if (!bface[s] && (int)out[1 + s].size() < M && grew[s]) { done = false; break; }
    }
    for (int s = 0; s < ns; s++) {      // :377-401
        if (bface[s]) { out[1 + s].clear(); continue; }
        fill_outwards(m, c, out[1 + s], M);
    }
}

struct MatrixScratch {
    dvec A, B, R, Y, X, v, colbuf, qp, px, py;
};

// Rows of the reconstruction matrix before the mean is removed: A[i][k] = area_t[i] * (mean of psi_k over stencil cell i,
// in the target cell's reference coordinates), by cell quadrature (:525-596).
// Cells with more than three nodes (new: the reference throws, :485-487): the target's reference frame is spanned by the
// edges node 0 -> node 1 and node 0 -> last node (for a triangle exactly the reference's frame), and a stencil cell is
// integrated as the fan of triangles (v0, v_j, v_j+1) - the split Mesh::compute_cell_volumes uses (mesh/mesh.cpp:196-215) -
// A[i][k] = sum_tri area_t(tri) * mean_tri(psi_k), area_t[i] = sum_tri area_t(tri).  For triangles nothing changes, bit for bit.
void integrate_basis_rows(const HostMesh & m, const TenoTables & t, uint32_t cell, const uint32_t * st, int M,
                          double * area_t, MatrixScratch & w) {
    const int K = t.K, nq = t.nq_cell, p = t.order;
    const double * X = m.node_xy.data();
    const uint32_t * cn = &m.noc[m.onc[cell]];
    const int kc = m.nnc(cell);
    const double * o = &X[2 * (size_t)cn[0]];
    double J[4], Ji[4];
    edge_frame(o, &X[2 * (size_t)cn[1]], &X[2 * (size_t)cn[kc - 1]], J);
    mat2_inverse(J, Ji);
    w.A.resize((size_t)M * K);
    w.px.resize((size_t)(p + 1) * nq); w.py.resize((size_t)(p + 1) * nq);
    for (int i = 0; i < M; i++) {
        const uint32_t * nn = &m.noc[m.onc[st[i]]];
        const int kn = m.nnc(st[i]);
        area_t[i] = 0.0;
        for (int tri = 0; tri + 2 < kn; tri++) {
            const double * v0 = &X[2 * (size_t)nn[0]], * v1 = &X[2 * (size_t)nn[tri + 1]], * v2 = &X[2 * (size_t)nn[tri + 2]];
            double Jn[4];
            edge_frame(v0, v1, v2, Jn);
            double a[2] = {v0[0] - o[0], v0[1] - o[1]}, b[2] = {v1[0] - o[0], v1[1] - o[1]}, c[2] = {v2[0] - o[0], v2[1] - o[1]};
            mat2_apply(Ji, a, a); mat2_apply(Ji, b, b); mat2_apply(Ji, c, c);
            const double at = tri_area2(a, b, c);
            for (int q = 0; q < nq; q++) {   // quadrature point: sub-triangle reference -> physical -> target reference
                double x[2] = {t.qc_xy[2 * q], t.qc_xy[2 * q + 1]};
                mat2_apply(Jn, x, x);
                x[0] += v0[0]; x[1] += v0[1];
                x[0] -= o[0]; x[1] -= o[1];
                mat2_apply(Ji, x, x);
                for (int d = 0; d <= p; d++) {
                    w.px[(size_t)d * nq + q] = basis_1d(t.basis, 0, d, x[0]);
                    w.py[(size_t)d * nq + q] = basis_1d(t.basis, 0, d, x[1]);
                }
            }
            for (int k = 0; k < K; k++) {
                const int ex = t.pidx[2 * k], ey = t.pidx[2 * k + 1];
                double s = 0.0;
                for (int q = 0; q < nq; q++) s += t.qc_w[q] * (w.px[(size_t)ex * nq + q] * w.py[(size_t)ey * nq + q]);
                if (tri == 0) w.A[(size_t)i * K + k] = s * at; else w.A[(size_t)i * K + k] += s * at;
            }
            if (tri == 0) area_t[i] = at; else area_t[i] += at;
        }
    }
}

// One stencil's reconstruction matrix (K x M row-major pseudo-inverse) and transformed areas, :503-706.
void stencil_matrix(const HostMesh & m, const TenoTables & t, uint32_t cell, const uint32_t * st, int M,
                    const double * psi_bar, double * area_t, double * Ainv, MatrixScratch & w) {
    const int K = t.K;
    integrate_basis_rows(m, t, cell, st, M, area_t, w);
    for (int i = 0; i < M; i++)
        for (int k = 0; k < K; k++) w.A[(size_t)i * K + k] -= area_t[i] * psi_bar[k];

    bool col0_zero = true;
    for (int i = 0; i < M; i++) if (w.A[(size_t)i * K] > 1.0e-12) { col0_zero = false; break; }
    w.v.resize(M); w.colbuf.resize(M);
    if (col0_zero) {
        const int r = M - 1, c = K - 1;
        w.B.resize((size_t)r * c); w.Y.resize((size_t)c * r); w.X.resize((size_t)c * r);
        for (int i = 1; i < M; i++)
            for (int k = 1; k < K; k++) w.B[(size_t)(i - 1) * c + (k - 1)] = w.A[(size_t)i * K + k];
        w.R = w.B;
        householder_R(w.R.data(), r, c, w.v.data(), w.colbuf.data());
        pseudo_inverse_from_R(w.R.data(), w.B.data(), r, c, w.Y.data(), w.X.data());
        for (int k = 0; k < K; k++)
            for (int i = 0; i < M; i++)
                Ainv[(size_t)k * M + i] = (k == 0 || i == 0) ? 0.0 : w.X[(size_t)(k - 1) * r + (i - 1)];
    } else {
        w.R = w.A;
        w.Y.resize((size_t)K * M);
        householder_R(w.R.data(), M, K, w.v.data(), w.colbuf.data());
        pseudo_inverse_from_R(w.R.data(), w.A.data(), M, K, w.Y.data(), Ainv);
    }
}

void oscillation_matrix(TenoTables & t) {   // :743-788
    const int K = t.K, nq = t.nq_cell;
    t.OI.assign((size_t)K * K, 0.0);
    for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) {
            double acc = 0.0;
            for (int k = 1; k < K; k++)
                for (int q = 0; q < nq; q++) {
                    double di = 1.0, dj = 1.0;
                    for (int d = 0; d < 2; d++) di *= basis_1d(t.basis, t.pidx[2 * k + d], t.pidx[2 * i + d], t.qc_xy[2 * q + d]);
                    for (int d = 0; d < 2; d++) dj *= basis_1d(t.basis, t.pidx[2 * k + d], t.pidx[2 * j + d], t.qc_xy[2 * q + d]);
                    acc += t.qc_w[q] * di * dj;
                }
            t.OI[(size_t)i * K + j] = acc;
        }
}

}  // namespace

// ----------------------------------------------------------------------------------------------------------------
// Reverse Cuthill–McKee over the sub-graph induced by `cells` (reference ids). Deterministic: components are started
// from the lowest-degree, lowest-id unvisited cell; neighbours are queued by (degree, id).
// ----------------------------------------------------------------------------------------------------------------
void rcm_order(const HostMesh & m, const std::vector<uint32_t> & cells, uvec & order) {
    const size_t n = cells.size();
    std::vector<int32_t> local(m.nc, -1);
    for (size_t i = 0; i < n; i++) local[cells[i]] = (int32_t)i;
    std::vector<uint32_t> adj_off(n + 1, 0), adj;
    adj.reserve(n * 3);
    for (size_t i = 0; i < n; i++) {
        const uint32_t c = cells[i];
        for (int j = 0; j < m.nfc(c); j++) {
            const uint32_t f = m.foc[m.ofc[c] + j];
            const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
            if (b < 0) continue;
            const uint32_t o = (a == (int32_t)c) ? (uint32_t)b : (uint32_t)a;
            if (local[o] >= 0) adj.push_back((uint32_t)local[o]);
        }
        adj_off[i + 1] = (uint32_t)adj.size();
    }
    auto deg = [&](uint32_t i) { return adj_off[i + 1] - adj_off[i]; };
    std::vector<uint32_t> by_deg(n);
    std::iota(by_deg.begin(), by_deg.end(), 0u);
    std::stable_sort(by_deg.begin(), by_deg.end(), [&](uint32_t a, uint32_t b) { return deg(a) < deg(b); });
    std::vector<char> vis(n, 0);
    std::vector<uint32_t> cm;
    cm.reserve(n);
    size_t cursor = 0;
    std::vector<uint32_t> nb;
    while (cm.size() < n) {
        while (vis[by_deg[cursor]]) cursor++;
        uint32_t start = by_deg[cursor];
        // pseudo-peripheral start: two BFS sweeps from the minimum-degree seed
        for (int sweep = 0; sweep < 2; sweep++) {
            std::vector<uint32_t> frontier{start}, next;
            std::vector<uint32_t> touched{start};
            std::vector<char> & mark = vis;   // temporary marks, undone below
            mark[start] = 2;
            uint32_t last = start;
            while (!frontier.empty()) {
                next.clear();
                for (uint32_t u : frontier)
                    for (uint32_t k = adj_off[u]; k < adj_off[u + 1]; k++) {
                        const uint32_t v = adj[k];
                        if (!mark[v]) { mark[v] = 2; next.push_back(v); touched.push_back(v); }
                    }
                if (!next.empty()) {
                    last = next[0];
                    for (uint32_t v : next) if (deg(v) < deg(last) || (deg(v) == deg(last) && v < last)) last = v;
                }
                frontier.swap(next);
            }
            for (uint32_t v : touched) mark[v] = 0;
            start = last;
        }
        size_t head = cm.size();
        cm.push_back(start);
        vis[start] = 1;
        while (head < cm.size()) {
            const uint32_t u = cm[head++];
            nb.clear();
            for (uint32_t k = adj_off[u]; k < adj_off[u + 1]; k++) if (!vis[adj[k]]) { vis[adj[k]] = 1; nb.push_back(adj[k]); }
            std::sort(nb.begin(), nb.end(), [&](uint32_t a, uint32_t b) { return deg(a) != deg(b) ? deg(a) < deg(b) : a < b; });
            cm.insert(cm.end(), nb.begin(), nb.end());
        }
    }
    order.resize(n);
    for (size_t i = 0; i < n; i++) order[i] = cells[cm[n - 1 - i]];
}

// ----------------------------------------------------------------------------------------------------------------
// Main entry
// ----------------------------------------------------------------------------------------------------------------
void preprocess(const HostMesh & m, const mlb_numerics & num, const std::vector<std::string> & bc_zones,
                const PrepOptions & opt, Prep & P) {
    const auto t0 = std::chrono::steady_clock::now();
    const bool timing = getenv("MLB_PREP_TIMING") != nullptr;
    auto tick = [&](const char * what) {
        if (timing) fprintf(stderr, "[mlb]   %-28s at %.2f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    const bool teno = num.recon == MLB_RECON_TENO;
    int n_slots = 0;
    for (uint32_t c = 0; c < m.nc; c++) n_slots = std::max(n_slots, m.nfc(c));
    if (n_slots > MAX_SLOTS) throw std::runtime_error("cells with more than 4 faces are not supported");
    P.n_slots = n_slots;

    // ---- face quadrature (FirstOrder: GaussLegendre(1), face_reconstruction.cpp:41-44; TENO: :115-116)
    if (!teno) gauss_legendre_rule(1, P.qf_x, P.qf_w);
    else {
        if (num.basis_order < 1 || num.basis_order > 9) throw std::runtime_error("TENO basis_order must be in 1..9");
        gauss_legendre_rule(num.quadrature_order_face > 0 ? num.quadrature_order_face : (num.basis_order + 1) / 2, P.qf_x, P.qf_w);
    }
    P.Q = (int)P.qf_w.size();

    // ---- TENO meta (face_reconstruction.cpp:170-215)
    TenoTables & T = P.teno;
    if (teno) {
        T.basis = num.basis; T.order = num.basis_order;
        int nd = 1, den = 1;
        for (int i = 1; i <= 2; i++) { den *= i; nd *= T.order + i; }
        T.K = nd / den;
        const double factor = num.max_stencil_size_factor > 0 ? num.max_stencil_size_factor : 2.0;
        T.M = (int)(uint16_t)(factor * T.K);
        T.Mp = T.M + (T.M & 1);
        T.S = 1 + n_slots;
        T.pidx.clear();
        for (int p = 0; p <= T.order; p++) {
            int a = p, b = 0;
            T.pidx.push_back((uint8_t)a); T.pidx.push_back((uint8_t)b);
            while (b < p) { if (a > 0) { a--; b++; } T.pidx.push_back((uint8_t)a); T.pidx.push_back((uint8_t)b); }
        }
        dunavant_rule(num.quadrature_order_cell > 0 ? num.quadrature_order_cell : T.order + 1, T.qc_xy, T.qc_w);
        T.nq_cell = (int)T.qc_w.size();
        T.keep_ref = opt.keep_ref_tables;
        for (uint32_t c = 0; c < m.nc; c++)
            if (m.nnc(c) != 3) { T.mixed = true; if (m.nnc(c) != 4) throw std::runtime_error("TENO: cells must be triangles or quadrilaterals."); }
    }

    // ---- which cells this context holds: owned | ring-1 ghosts (reconstructed, not updated) | state-only ghosts
    std::vector<uint32_t> owned, g1, g2;
    std::vector<uint8_t> cls(m.nc, 0);   // 1 owned, 2 g1, 3 g2
    if (!opt.part) { owned.resize(m.nc); std::iota(owned.begin(), owned.end(), 0u); std::fill(cls.begin(), cls.end(), 1); }
    else {
        for (uint32_t c = 0; c < m.nc; c++) if (opt.part[c] == opt.rank) { owned.push_back(c); cls[c] = 1; }
        for (uint32_t c : owned)
            for (int j = 0; j < m.nfc(c); j++) {
                const uint32_t f = m.foc[m.ofc[c] + j];
                const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
                if (b < 0) continue;
                const uint32_t o = (a == (int32_t)c) ? (uint32_t)b : (uint32_t)a;
                if (!cls[o]) { cls[o] = 2; g1.push_back(o); }
            }
        std::sort(g1.begin(), g1.end());
    }
    std::vector<uint32_t> gv;            // viscous: face neighbours of the first ghost ring (their states feed the ring's Green-Gauss gradients)
    if (opt.viscous && opt.part) {
        for (uint32_t c : g1)
            for (int j = 0; j < m.nfc(c); j++) {
                const uint32_t f = m.foc[m.ofc[c] + j];
                const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
                if (b == CUT_FACE) throw std::runtime_error("rank-local mesh: a first-ring ghost cell has a cut face (viscous runs need one more ghost layer)");
                if (b < 0) continue;
                const uint32_t o = (a == (int32_t)c) ? (uint32_t)b : (uint32_t)a;
                if (!cls[o]) { cls[o] = 3; gv.push_back(o); }
            }
        std::sort(gv.begin(), gv.end());
    }
    uvec order;
    if (num.renumber == MLB_RENUMBER_RCM && owned.size() > 1) rcm_order(m, owned, order); else order = owned;
    const uint32_t n_owned = (uint32_t)order.size();
    order.insert(order.end(), g1.begin(), g1.end());
    const uint32_t n_recon = (uint32_t)order.size();
    uint32_t n_interior = opt.part ? 0u : n_owned;   // first order: no reconstruction kernel to overlap

    tick("renumbering done");
    // ---- TENO stencils for owned + ring-1 cells (parallel), written into the tile layout with REFERENCE ids first;
    //      state-only ghosts are whatever else those stencils touch
    const size_t n_tiles = (n_recon + TILE - 1) / TILE;
    if (teno) {
        T.st_ids.assign(n_tiles * T.S * T.Mp * TILE, NO_FACE);
        std::string err;
#pragma omp parallel
        {
            std::vector<uvec> cs;
#pragma omp for schedule(dynamic, 64)
            for (int64_t ii = 0; ii < (int64_t)n_recon; ii++) {
                const size_t tile = (size_t)ii / TILE, lane = (size_t)ii % TILE;
                try {
                    cell_stencils(m, order[ii], T.M, cs);
                    for (int s = 0; s < (int)cs.size(); s++) {
                        if (cs[s].empty()) continue;
                        if ((int)cs[s].size() < T.M) throw std::runtime_error("Stencil is not full.");
                        for (int k = 0; k < T.M; k++) T.st_ids[((tile * T.S + s) * T.Mp + k) * TILE + lane] = cs[s][k];
                    }
                } catch (const std::exception & e) {
#pragma omp critical
                    err = e.what();
                }
            }
        }
        if (!err.empty()) throw std::runtime_error(err);
        P.seconds_stencils = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (opt.part) {
            // Interior-first order of the owned cells: a cell whose stencils hold owned cells only can be reconstructed
            // while the ghost states are still in flight.  The partition is stable, so each class keeps its RCM order.
            std::vector<uint8_t> rim(n_owned, 0);
#pragma omp parallel for schedule(static)
            for (int64_t ii = 0; ii < (int64_t)n_owned; ii++) {
                const size_t tile = (size_t)ii / TILE, lane = (size_t)ii % TILE;
                for (int s = 0; s < T.S && !rim[ii]; s++)
                    for (int k = 0; k < T.M; k++) {
                        const uint32_t x = T.st_ids[((tile * T.S + s) * T.Mp + k) * TILE + lane];
                        if (x != NO_FACE && cls[x] != 1) { rim[ii] = 1; break; }
                    }
            }
            uvec newpos(n_recon);
            uint32_t a = 0;
            for (uint32_t ii = 0; ii < n_owned; ii++) if (!rim[ii]) newpos[ii] = a++;
            n_interior = a;
            for (uint32_t ii = 0; ii < n_owned; ii++) if (rim[ii]) newpos[ii] = a++;
            for (uint32_t ii = n_owned; ii < n_recon; ii++) newpos[ii] = ii;
            uvec ids2(T.st_ids.size(), NO_FACE), order2(order.size());
#pragma omp parallel for schedule(static)
            for (int64_t ii = 0; ii < (int64_t)n_recon; ii++) {
                const size_t tile = (size_t)ii / TILE, lane = (size_t)ii % TILE, jj = newpos[ii], tile2 = jj / TILE, lane2 = jj % TILE;
                order2[jj] = order[ii];
                for (int s = 0; s < T.S; s++)
                    for (int k = 0; k < T.Mp; k++)
                        ids2[((tile2 * T.S + s) * T.Mp + k) * TILE + lane2] = T.st_ids[((tile * T.S + s) * T.Mp + k) * TILE + lane];
            }
            T.st_ids.swap(ids2);
            order.swap(order2);
            for (uint32_t x : T.st_ids) if (x != NO_FACE && !cls[x]) { cls[x] = 3; g2.push_back(x); }
            g2.insert(g2.end(), gv.begin(), gv.end());
            gv.clear();
            std::sort(g2.begin(), g2.end());
            order.insert(order.end(), g2.begin(), g2.end());
        }
    }
    order.insert(order.end(), gv.begin(), gv.end());      // first order: the second ring is all the state-only ghosts there are
    const uint32_t N = (uint32_t)order.size();
    P.N = N; P.N_owned = n_owned; P.N_recon = n_recon; P.N_interior = n_interior;
    P.Npad = (N + 31u) & ~31u;
    P.perm_cells = order;
    P.iperm_cells.assign(m.nc, NO_FACE);
    for (uint32_t i = 0; i < N; i++) P.iperm_cells[order[i]] = i;

    tick("stencils done");
    // ---- faces of owned cells, ordered by (owner = lower library cell id, other)
    {
        std::vector<std::pair<uint64_t, uint32_t>> keyed;
        std::vector<char> taken(m.nf, 0);
        for (uint32_t i = 0; i < n_owned; i++) {
            const uint32_t c = order[i];
            for (int j = 0; j < m.nfc(c); j++) {
                const uint32_t f = m.foc[m.ofc[c] + j];
                if (taken[f]) continue;
                taken[f] = 1;
                const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
                const uint32_t la = P.iperm_cells[a], lb = b >= 0 ? P.iperm_cells[b] : 0xFFFFFFFEu;
                keyed.push_back({((uint64_t)std::min(la, lb) << 32) | std::max(la, lb), f});
            }
        }
        std::sort(keyed.begin(), keyed.end());
        P.NF = (uint32_t)keyed.size();
        P.NFpad = (P.NF + 31u) & ~31u;
        P.perm_faces.resize(P.NF);
        for (uint32_t i = 0; i < P.NF; i++) P.perm_faces[i] = keyed[i].second;
    }
    std::vector<uint32_t> iperm_faces(m.nf, NO_FACE);
    for (uint32_t i = 0; i < P.NF; i++) iperm_faces[P.perm_faces[i]] = i;

    tick("faces ordered");
    // ---- zone binding: order key of every face = (0, pos in interior zone) or (1 + bc index, pos in its zone)
    std::vector<int32_t> face_bc(m.nf, INT32_MIN);          // bc index for boundary faces with a [[boundaries]] entry
    std::vector<uint64_t> face_key(m.nf, UINT64_MAX);
    if (const HostZone * zi = m.zone("interior"))
        for (size_t k = 0; k < zi->faces.size(); k++) face_key[zi->faces[k]] = (uint64_t)k;
    if ((int)bc_zones.size() > MAX_BCS) throw std::runtime_error("too many boundaries");
    for (size_t b = 0; b < bc_zones.size(); b++) {
        const HostZone * z = m.zone(bc_zones[b]);
        if (!z) throw std::runtime_error("Boundary name " + bc_zones[b] + " not found in mesh.");
        for (size_t k = 0; k < z->faces.size(); k++) {
            const uint32_t f = z->faces[k];
            if (face_bc[f] == INT32_MIN) { face_bc[f] = (int32_t)b; face_key[f] = ((uint64_t)(b + 1) << 40) | (uint64_t)k; }
        }
    }

    tick("zones bound");
    // ---- geometry in library numbering
    const uint32_t Np = P.Npad;
    P.cell_vol.assign(Np, 1.0);
    P.cell_xy.assign(2 * (size_t)Np, 0.0);
    P.n_faces_of_cell.assign(Np, 0);
    for (uint32_t i = 0; i < N; i++) {
        const uint32_t c = order[i];
        P.cell_vol[i] = m.cell_vol[c];
        P.cell_xy[i] = m.cell_xy[2 * (size_t)c];
        P.cell_xy[(size_t)Np + i] = m.cell_xy[2 * (size_t)c + 1];
        P.n_faces_of_cell[i] = (uint8_t)m.nfc(c);
    }
    P.face_nx.assign(P.NFpad, 0.0); P.face_ny.assign(P.NFpad, 0.0); P.face_area.assign(P.NFpad, 0.0);
    for (uint32_t i = 0; i < P.NF; i++) {
        const uint32_t f = P.perm_faces[i];
        const double nx = m.face_n[2 * (size_t)f], ny = m.face_n[2 * (size_t)f + 1];
        const double inv = 1 / std::sqrt(nx * nx + ny * ny);   // unit<2>, common_math.h:101-106
        P.face_nx[i] = nx * inv;
        P.face_ny[i] = ny * inv;
        P.face_area[i] = m.face_area[f];
    }

    tick("geometry done");
    // ---- slots
    P.slot_face.assign((size_t)n_slots * Np, NO_FACE);
    P.slot_nbr.assign((size_t)n_slots * Np, INT32_MIN);
    P.slot_nslot.assign((size_t)n_slots * Np, 0);
    P.rhs_order.assign(Np, 0);
    if (teno) P.slot_fx.assign((size_t)n_slots * 4 * Np, 0.0);
    std::string slot_err;
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t)n_recon; ii++) {
        const uint32_t i = (uint32_t)ii, c = order[i];
        const int nfc = m.nfc(c);
        uint64_t keys[MAX_SLOTS];
        for (int j = 0; j < nfc; j++) {
            const uint32_t f = m.foc[m.ofc[c] + j];
            const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
            if (b == CUT_FACE) {   // rank-local mesh: fine on a first-ring ghost of a first-order context (its faces are never used)
                if (i < n_owned || teno) {
#pragma omp critical
                    slot_err = "rank-local mesh: an owned or reconstructed ghost cell has a cut face (the mesh needs more ghost layers)";
                }
                keys[j] = UINT64_MAX;
                continue;
            }
            const uint32_t side = (a == (int32_t)c) ? 0u : 1u;
            const size_t at = (size_t)j * Np + i;
            if (iperm_faces[f] != NO_FACE) P.slot_face[at] = iperm_faces[f] | (side << 31);
            keys[j] = face_key[f];
            if (b >= 0) {
                const uint32_t o = side == 0 ? (uint32_t)b : (uint32_t)a;
                if (face_key[f] != UINT64_MAX && P.iperm_cells[o] != NO_FACE) P.slot_nbr[at] = (int32_t)P.iperm_cells[o];
                for (int k = 0; k < m.nfc(o); k++) if (m.foc[m.ofc[o] + k] == f) P.slot_nslot[at] = (uint8_t)k;
            } else if (face_bc[f] != INT32_MIN) {
                P.slot_nbr[at] = -(face_bc[f] + 1);
            }
            if (teno) {   // face end points in the cell's reference frame (face_reconstruction.cpp:877-886,990-997)
                const double * X = m.node_xy.data();
                const uint32_t * cn = &m.noc[m.onc[c]];
                const double * o0 = &X[2 * (size_t)cn[0]];
                double J[4], Ji[4];
                edge_frame(o0, &X[2 * (size_t)cn[1]], &X[2 * (size_t)cn[m.nnc(c) - 1]], J);
                mat2_inverse(J, Ji);
                const uint32_t n0 = m.nof[m.onf[f]], n1 = m.nof[m.onf[f] + 1];
                double x0[2] = {X[2 * (size_t)n0] - o0[0], X[2 * (size_t)n0 + 1] - o0[1]};
                double x1[2] = {X[2 * (size_t)n1] - o0[0], X[2 * (size_t)n1 + 1] - o0[1]};
                mat2_apply(Ji, x0, x0); mat2_apply(Ji, x1, x1);
                double * fx = &P.slot_fx[((size_t)j * 4) * Np + i];
                fx[0] = x0[0]; fx[Np] = x0[1]; fx[2 * (size_t)Np] = x1[0]; fx[3 * (size_t)Np] = x1[1];
            }
        }
        // accumulation order of the reference's Serial backend: interior zone order, then boundaries in input order
        int idx[MAX_SLOTS] = {0, 1, 2, 3};
        std::stable_sort(idx, idx + nfc, [&](int a, int b) { return keys[a] < keys[b]; });
        uint8_t code = 0;
        for (int j = 0; j < nfc; j++) code |= (uint8_t)(idx[j] << (2 * j));
        P.rhs_order[i] = code;
    }

    if (!slot_err.empty()) throw std::runtime_error(slot_err);
    if (opt.viscous) {   // least-squares geometry of every reconstructed cell, centroid line of every held face
        P.slot_d.assign((size_t)n_slots * 2 * Np, 0.0);
        P.face_d.assign(4 * (size_t)P.NFpad, 0.0);
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < (int64_t)n_recon; ii++) {
            const uint32_t i = (uint32_t)ii, c = order[i];
            for (int j = 0; j < m.nfc(c); j++) {
                const uint32_t f = m.foc[m.ofc[c] + j];
                const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
                double dx, dy;
                if (b >= 0) {
                    const uint32_t o = a == (int32_t)c ? (uint32_t)b : (uint32_t)a;
                    dx = m.cell_xy[2 * (size_t)o] - m.cell_xy[2 * (size_t)c]; dy = m.cell_xy[2 * (size_t)o + 1] - m.cell_xy[2 * (size_t)c + 1];
                } else if (b == -1) {   // mirror image of the centroid in the boundary face
                    const double nx = m.face_n[2 * (size_t)f], ny = m.face_n[2 * (size_t)f + 1], inv = 1 / std::sqrt(nx * nx + ny * ny);
                    const double * n0 = &m.node_xy[2 * (size_t)m.nof[m.onf[f]]], * n1 = &m.node_xy[2 * (size_t)m.nof[m.onf[f] + 1]];
                    const double dist = (0.5 * (n0[0] + n1[0]) - m.cell_xy[2 * (size_t)c]) * nx * inv + (0.5 * (n0[1] + n1[1]) - m.cell_xy[2 * (size_t)c + 1]) * ny * inv;
                    dx = 2.0 * dist * nx * inv; dy = 2.0 * dist * ny * inv;
                } else continue;        // cut face of a rank-local mesh (first-ring ghosts of viscous runs were checked above)
                P.slot_d[((size_t)j * 2) * Np + i] = dx;
                P.slot_d[((size_t)j * 2 + 1) * Np + i] = dy;
            }
        }
        for (uint32_t i = 0; i < P.NF; i++) {
            const uint32_t f = P.perm_faces[i];
            const int32_t a = m.cof[2 * (size_t)f], b = m.cof[2 * (size_t)f + 1];
            const double * n0 = &m.node_xy[2 * (size_t)m.nof[m.onf[f]]], * n1 = &m.node_xy[2 * (size_t)m.nof[m.onf[f] + 1]];
            const double rx = 0.5 * (n0[0] + n1[0]) - m.cell_xy[2 * (size_t)a], ry = 0.5 * (n0[1] + n1[1]) - m.cell_xy[2 * (size_t)a + 1];
            double dx, dy;
            if (b >= 0) { dx = m.cell_xy[2 * (size_t)b] - m.cell_xy[2 * (size_t)a]; dy = m.cell_xy[2 * (size_t)b + 1] - m.cell_xy[2 * (size_t)a + 1]; }
            else {   // mirror image of the cell centroid in the face
                const double dist = rx * P.face_nx[i] + ry * P.face_ny[i];
                dx = 2.0 * dist * P.face_nx[i]; dy = 2.0 * dist * P.face_ny[i];
            }
            P.face_d[i] = dx; P.face_d[(size_t)P.NFpad + i] = dy;
            P.face_d[2 * (size_t)P.NFpad + i] = rx; P.face_d[3 * (size_t)P.NFpad + i] = ry;
        }
    }
    tick("slots done");
    // ---- face-centred view of the same connectivity (face flux kernel)
    P.face_cl.assign(P.NFpad, 0u);
    P.face_cr.assign(P.NFpad, INT32_MIN);
    P.face_slots.assign(P.NFpad, 0);
    for (uint32_t i = 0; i < n_recon; i++)
        for (int j = 0; j < (int)P.n_faces_of_cell[i]; j++) {
            const size_t at = (size_t)j * Np + i;
            const uint32_t fcode = P.slot_face[at];
            if (fcode == NO_FACE) continue;
            const uint32_t f = fcode & 0x7FFFFFFFu;
            if (fcode >> 31) P.face_slots[f] |= (uint8_t)(j << 4);
            else {
                P.face_cl[f] = i;
                P.face_slots[f] |= (uint8_t)j;
                if (P.slot_nbr[at] < 0) P.face_cr[f] = P.slot_nbr[at];
            }
            if (P.slot_nbr[at] >= 0) {   // interior face: side 1 cell, whichever side we are looking from
                if (fcode >> 31) { P.face_cr[f] = (int32_t)i; P.face_cl[f] = (uint32_t)P.slot_nbr[at]; }
                else P.face_cr[f] = P.slot_nbr[at];
            }
        }

    tick("face view done");
    // ---- TENO tables
    if (teno) {
        const int K = T.K, M = T.M, Mp = T.Mp, S = T.S;
        T.fast = opt.fast_tables;
        const bool strict_tables = !T.fast || T.keep_ref;        // bit-faithful layout (STRICT kernels, parity hooks)
        if (T.fast && (S != FAST_S || M != 2 * K)) throw std::runtime_error("streaming TENO tables need triangles and max_stencil_size_factor 2");
        if (strict_tables || !opt.device_tables) T.st_area.assign(n_tiles * S * Mp * TILE, 0.0);
        if (strict_tables) T.st_mat.assign(n_tiles * S * K * Mp * TILE, 0.0);
        const int KR = K - 1, MC = M - 1, NPAIR = MC / 2;
        const size_t n_ftiles = (n_recon + FAST_CT - 1) / FAST_CT;
        const size_t frow = (size_t)(2 * NPAIR + 1) * FAST_CT;   // doubles per stored row of one tile
        const bool host_fm = T.fast && !opt.device_tables;       // device_tables: teno_tables.cu builds fm_mat / fm_area0 in HBM
        if (host_fm) {
            T.fm_mat.assign(n_ftiles * S * KR * frow, 0.0);
            T.fm_area0.assign(n_ftiles * FAST_CT, 0.5);
        }
        if (T.fast && opt.device_tables) {   // node coordinates of every held cell, library numbering (the kernel's only geometry input)
            T.tri_xy.assign(6 * (size_t)P.Npad, 0.0);
#pragma omp parallel for schedule(static)
            for (int64_t ii = 0; ii < (int64_t)N; ii++) {
                const uint32_t * cn = &m.noc[m.onc[order[ii]]];
                for (int k2 = 0; k2 < 3; k2++) {
                    T.tri_xy[6 * (size_t)ii + 2 * k2] = m.node_xy[2 * (size_t)cn[k2]];
                    T.tri_xy[6 * (size_t)ii + 2 * k2 + 1] = m.node_xy[2 * (size_t)cn[k2] + 1];
                }
            }
        }
        T.psi_bar.assign(K, 0.0);
        {   // integral_psi_target: row 0 of REFERENCE cell 0's central stencil, i.e. the cell itself (:598-602)
            MatrixScratch w;
            const uint32_t self = 0;
            double a0;
            if (opt.psi_ref_tri) {   // rank-local mesh: the caller hands over the GLOBAL mesh's cell 0
                HostMesh one;
                one.nc = 1; one.nn = 3;
                one.node_xy.assign(opt.psi_ref_tri, opt.psi_ref_tri + 6);
                one.onc = {0u, 3u}; one.noc = {0u, 1u, 2u};
                integrate_basis_rows(one, T, 0, &self, 1, &a0, w);
            } else {
                uint32_t ref = 0;                          // the reference's cell 0 (a triangle there; the first triangle of a mixed mesh)
                while (ref < m.nc && m.nnc(ref) != 3) ref++;
                if (ref == m.nc) ref = 0;
                integrate_basis_rows(m, T, ref, &ref, 1, &a0, w);
            }
            for (int k = 0; k < K; k++) T.psi_bar[k] = w.A[k] / a0;
        }
        if (T.mixed) {   // quadrilaterals do not share one reference shape: their own mean of every basis function
            T.psi_bar_cell.assign((size_t)P.Npad * K, 0.0);
#pragma omp parallel
            {
                MatrixScratch w;
#pragma omp for schedule(static)
                for (int64_t ii = 0; ii < (int64_t)n_recon; ii++) {
                    const uint32_t c = order[ii];
                    double a0;
                    if (m.nnc(c) == 3) { for (int k = 0; k < K; k++) T.psi_bar_cell[(size_t)ii * K + k] = T.psi_bar[k]; continue; }
                    integrate_basis_rows(m, T, c, &c, 1, &a0, w);
                    for (int k = 0; k < K; k++) T.psi_bar_cell[(size_t)ii * K + k] = w.A[k] / a0;
                }
            }
        }
        std::string err;
        const auto t_mat = std::chrono::steady_clock::now();
#pragma omp parallel
        {
            MatrixScratch w;
            dvec at(M), Ai((size_t)K * M);
            uvec st(M);
#pragma omp for schedule(dynamic, 32)
            for (int64_t ii = 0; ii < (int64_t)n_recon; ii++) {
                const uint32_t i = (uint32_t)ii;
                const size_t tile = i / TILE, lane = i % TILE;
                if (!strict_tables && !host_fm) continue;           // every matrix is built on the device
                try {
                    for (int s = 0; s < S; s++) {
                        const size_t base = (tile * S + s) * Mp;
                        if (T.st_ids[base * TILE + lane] == NO_FACE) continue;
                        for (int k2 = 0; k2 < M; k2++) st[k2] = T.st_ids[(base + k2) * TILE + lane];
                        stencil_matrix(m, T, order[i], st.data(), M, T.mixed ? &T.psi_bar_cell[(size_t)i * K] : T.psi_bar.data(), at.data(), Ai.data(), w);
                        for (int k2 = 0; k2 < M; k2++) T.st_area[(base + k2) * TILE + lane] = at[k2];
                        if (strict_tables)
                            for (int k = 0; k < K; k++)
                                for (int k2 = 0; k2 < M; k2++)
                                    T.st_mat[((((tile * S + s) * K + k) * (Mp / 2) + k2 / 2) * TILE + lane) * 2 + (k2 & 1)] = Ai[(size_t)k * M + k2];
                        if (host_fm) {   // rows 1..K-1, columns 1..M-1, transformed areas folded into the columns
                            const size_t ft = i / FAST_CT, fl = i % FAST_CT;
                            for (int k2 = 0; k2 < M; k2++) if (Ai[k2] != 0.0) throw std::runtime_error("TENO: reconstruction matrix has a non-zero first row; compact tables unavailable (use fp_mode strict)");
                            for (int k = 0; k < K; k++) if (Ai[(size_t)k * M] != 0.0) throw std::runtime_error("TENO: reconstruction matrix has a non-zero first column; compact tables unavailable (use fp_mode strict)");
                            if (s == 0) T.fm_area0[i] = at[0];
                            for (int k = 1; k < K; k++) {
                                double * row = &T.fm_mat[((ft * S + s) * KR + (k - 1)) * frow];
                                for (int c = 0; c < MC; c++) {
                                    const double v = Ai[(size_t)k * M + (c + 1)] * at[c + 1];
                                    if (c < 2 * NPAIR) row[((size_t)(c / 2) * FAST_CT + fl) * 2 + (c & 1)] = v;
                                    else row[(size_t)2 * NPAIR * FAST_CT + fl] = v;
                                }
                            }
                        }
                    }
                } catch (const std::exception & e) {
#pragma omp critical
                    err = e.what();
                }
            }
        }
        if (!err.empty()) throw std::runtime_error(err);
        P.seconds_matrices = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_mat).count();
        oscillation_matrix(T);
        if (T.keep_ref) {   // re-emit in the reference's CSR layout and numbering (face_reconstruction.h:201-260)
            if (opt.part) throw std::runtime_error("reference-layout TENO tables are only kept for unpartitioned contexts");
            T.ref_off_groups.assign(1, 0); T.ref_off_stencils.assign(1, 0); T.ref_off_mats.assign(1, 0);
            T.ref_stencils.clear(); T.ref_mats.clear(); T.ref_areas.clear();
            for (uint32_t c = 0; c < m.nc; c++) {
                const uint32_t i = P.iperm_cells[c];
                const size_t tile = i / TILE, lane = i % TILE;
                for (int s = 0; s < 1 + m.nfc(c); s++) {
                    const size_t base = (tile * S + s) * Mp;
                    if (T.st_ids[base * TILE + lane] != NO_FACE) {
                        for (int k2 = 0; k2 < M; k2++) {
                            T.ref_stencils.push_back(T.st_ids[(base + k2) * TILE + lane]);
                            T.ref_areas.push_back(T.st_area[(base + k2) * TILE + lane]);
                        }
                        for (int k = 0; k < K; k++)
                            for (int k2 = 0; k2 < M; k2++)
                                T.ref_mats.push_back(T.st_mat[((((tile * S + s) * K + k) * (Mp / 2) + k2 / 2) * TILE + lane) * 2 + (k2 & 1)]);
                    }
                    T.ref_off_stencils.push_back((uint32_t)T.ref_stencils.size());
                    T.ref_off_mats.push_back((uint32_t)T.ref_mats.size());
                }
                T.ref_off_groups.push_back((uint32_t)T.ref_off_stencils.size() - 1);
            }
        }
        // finally: reference ids -> library ids; the odd-M padding column reads the cell itself with area 0
        for (size_t tile = 0; tile < n_tiles; tile++)
            for (int s = 0; s < S; s++)
                for (size_t lane = 0; lane < TILE; lane++) {
                    const size_t base = (tile * S + s) * Mp;
                    if (T.st_ids[base * TILE + lane] == NO_FACE) continue;
                    for (int k2 = 0; k2 < M; k2++) {
                        uint32_t & id = T.st_ids[(base + k2) * TILE + lane];
                        id = P.iperm_cells[id];
                    }
                    for (int k2 = M; k2 < Mp; k2++) T.st_ids[(base + k2) * TILE + lane] = (uint32_t)(tile * TILE + lane);
                }
        if (T.fast) {
            // an empty stencil (boundary face) and the padding cells of the last tile list the cell itself: the gather
            // stays in bounds without a predicate, and "first id == own id" marks the stencil as empty
            T.fm_ids.resize(n_ftiles * S * MC * FAST_CT);
#pragma omp parallel for schedule(static)
            for (int64_t ii = 0; ii < (int64_t)(n_ftiles * FAST_CT); ii++) {
                const size_t tile = (size_t)ii / TILE, lane = (size_t)ii % TILE, ft = (size_t)ii / FAST_CT, fl = (size_t)ii % FAST_CT;
                for (int s = 0; s < S; s++) {
                    const size_t base = (tile * S + s) * Mp;
                    const bool empty = ii >= (int64_t)n_recon || T.st_ids[base * TILE + lane] == NO_FACE;
                    for (int c = 0; c < MC; c++)
                        T.fm_ids[((ft * S + s) * MC + c) * FAST_CT + fl] = empty ? (uint32_t)ii : T.st_ids[(base + c + 1) * TILE + lane];
                }
            }
            T.OIs.assign((size_t)KR * KR, 0.0);
            for (int k = 1; k < K; k++)
                for (int j = k; j < K; j++)   // a^T OI a = sum_k OI_kk a_k^2 + sum_{k<j} (OI_kj + OI_jk) a_k a_j
                    T.OIs[(size_t)(k - 1) * KR + (j - 1)] = j == k ? T.OI[(size_t)k * K + k] : T.OI[(size_t)k * K + j] + T.OI[(size_t)j * K + k];
            if (!strict_tables) { uvec().swap(T.st_ids); dvec().swap(T.st_area); }
        }
    }
    P.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("MLB_PREP_TIMING"))
        fprintf(stderr, "[mlb] preprocess: %u cells, total %.2f s (stencil search %.2f s, matrices %.2f s)\n", P.N, P.seconds, P.seconds_stencils, P.seconds_matrices);
}

}  // namespace mlb
