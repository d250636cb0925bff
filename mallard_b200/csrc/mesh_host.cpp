// Host-side mesh producers: structured generators and geometry, in the REFERENCE's numbering so that the library
// can be handed either these arrays or the reference Mesh object's own (mesh/mesh.h:228-253).
//
// The generators are written face-centrically (every face's owner, neighbour and node order is given in closed form)
// instead of the reference's cell loop with overwrite-by-later-cell semantics (mesh/mesh.cpp:349-399,603-667); the
// resulting arrays are identical — tests/test_host_mesh.py checks them bit for bit against the oracle.
// Compiled with -ffp-contract=off: geometry must round exactly like the reference's x86-64 build.
#include <algorithm>
#include <cmath>

#include "mlb_internal.h"

namespace mlb {

namespace {

inline double tri_area2(const double * a, const double * b, const double * c) {   // common_math.h:507-513
    return 0.5 * std::fabs(a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1]));
}

struct Grid {
    uint32_t nx, ny;
    uint32_t node(uint32_t i, uint32_t j) const { return i * (ny + 1) + j; }
};

void grid_nodes(HostMesh & m, const Grid & g, double Lx, double Ly) {
    m.nn = (g.nx + 1) * (g.ny + 1);
    m.node_xy.assign(2 * (size_t)m.nn, 0.0);
    const double dx = Lx / g.nx, dy = Ly / g.ny;
    for (uint32_t i = 0; i <= g.nx; i++)
        for (uint32_t j = 0; j <= g.ny; j++) {
            m.node_xy[2 * (size_t)g.node(i, j)] = i * dx;
            m.node_xy[2 * (size_t)g.node(i, j) + 1] = j * dy;
        }
}

void set_face(HostMesh & m, uint32_t f, int32_t c0, int32_t c1, uint32_t n0, uint32_t n1) {
    m.cof[2 * (size_t)f] = c0; m.cof[2 * (size_t)f + 1] = c1;
    m.nof[2 * (size_t)f] = n0; m.nof[2 * (size_t)f + 1] = n1;
}

void uniform_offsets(uvec & off, uint32_t n, uint32_t stride) {
    off.resize((size_t)n + 1);
    for (uint32_t i = 0; i <= n; i++) off[i] = i * stride;
}

void make_zones(HostMesh & m, uvec & right, uvec & top, uvec & left, uvec & bottom, uint32_t n_real_faces) {
    uvec interior;
    for (uint32_t f = 0; f < n_real_faces; f++)
        if (m.cof[2 * (size_t)f + 1] >= 0) interior.push_back(f);   // ascending = the reference's sort+unique (mesh.cpp:461-464)
    m.zones.clear();
    m.zones.push_back({"interior", interior});
    m.zones.push_back({"right", right});
    m.zones.push_back({"top", top});
    m.zones.push_back({"left", left});
    m.zones.push_back({"bottom", bottom});
}

// Quads, mesh/mesh.cpp:305-557. Column ic owns faces [(2ny+1)ic, (2ny+1)(ic+1)): ny vertical "left" faces, then per
// cell the bottom face at +ny+jc (the top of the last cell closes the column); the last column's right faces follow.
void gen_quads(HostMesh & m, uint32_t nx, uint32_t ny, double Lx, double Ly) {
    Grid g{nx, ny};
    grid_nodes(m, g, Lx, Ly);
    m.nc = nx * ny;
    m.nf = 2 * m.nc + nx + ny;
    m.cof.assign(2 * (size_t)m.nf, 0);
    m.nof.assign(2 * (size_t)m.nf, 0);
    m.noc.resize(4 * (size_t)m.nc);
    m.foc.resize(4 * (size_t)m.nc);
    uniform_offsets(m.onc, m.nc, 4); uniform_offsets(m.ofc, m.nc, 4); uniform_offsets(m.onf, m.nf, 2);
    const uint32_t col = 2 * ny + 1;
    uvec zr, zt, zl, zb;
    for (uint32_t ic = 0; ic < nx; ic++)
        for (uint32_t jc = 0; jc < ny; jc++) {
            const uint32_t c = ic * ny + jc;
            const uint32_t tr = g.node(ic + 1, jc + 1), tl = g.node(ic, jc + 1), bl = g.node(ic, jc), br = g.node(ic + 1, jc);
            const uint32_t fR = col * (ic + 1) + jc, fT = col * ic + jc + ny + 1, fL = col * ic + jc, fB = col * ic + jc + ny;
            const uint32_t nodes[4] = {tr, tl, bl, br}, faces[4] = {fR, fT, fL, fB};
            for (int k = 0; k < 4; k++) { m.noc[4 * (size_t)c + k] = nodes[k]; m.foc[4 * (size_t)c + k] = faces[k]; }
            // A shared face ends up described by its higher-numbered cell: that cell's left / bottom face.
            if (ic == 0) { set_face(m, fL, c, -1, tl, bl); zl.push_back(fL); } else set_face(m, fL, c, c - ny, tl, bl);
            if (jc == 0) { set_face(m, fB, c, -1, bl, br); zb.push_back(fB); } else set_face(m, fB, c, c - 1, bl, br);
            if (ic == nx - 1) { set_face(m, fR, c, -1, br, tr); zr.push_back(fR); }
            if (jc == ny - 1) { set_face(m, fT, c, -1, tr, tl); zt.push_back(fT); }
        }
    make_zones(m, zr, zt, zl, zb, m.nf);
}

// Triangles, mesh/mesh.cpp:559-826: each quad is cut along bl–tr into a "cr" (lower-right) and "cl" (upper-left) cell.
// The reference allocates 3*nc + nx + ny faces although only (3ny+1)nx + ny exist; the tail stays zero-initialised
// ("phantom" faces).  We keep the allocation so face ids and array shapes match, and mark nothing there.
void gen_tris(HostMesh & m, uint32_t nx, uint32_t ny, double Lx, double Ly) {
    Grid g{nx, ny};
    grid_nodes(m, g, Lx, Ly);
    m.nc = 2 * nx * ny;
    m.nf = 3 * m.nc + nx + ny;
    m.cof.assign(2 * (size_t)m.nf, 0);
    m.nof.assign(2 * (size_t)m.nf, 0);
    m.noc.resize(3 * (size_t)m.nc);
    m.foc.resize(3 * (size_t)m.nc);
    uniform_offsets(m.onc, m.nc, 3); uniform_offsets(m.ofc, m.nc, 3); uniform_offsets(m.onf, m.nf, 2);
    const uint32_t col = 3 * ny + 1;
    uvec zr, zt, zl, zb;
    for (uint32_t ic = 0; ic < nx; ic++)
        for (uint32_t jc = 0; jc < ny; jc++) {
            const uint32_t q = ic * ny + jc, cr = 2 * q, cl = cr + 1;
            const uint32_t tr = g.node(ic + 1, jc + 1), tl = g.node(ic, jc + 1), bl = g.node(ic, jc), br = g.node(ic + 1, jc);
            const uint32_t fR = col * (ic + 1) + jc, fL = col * ic + jc;
            const uint32_t fB = col * ic + 2 * jc + ny, fD = fB + 1, fT = fB + 2;
            const uint32_t ncr[3] = {br, tr, bl}, ncl[3] = {tl, bl, tr}, fcr[3] = {fB, fR, fD}, fcl[3] = {fT, fL, fD};
            for (int k = 0; k < 3; k++) {
                m.noc[3 * (size_t)cr + k] = ncr[k]; m.noc[3 * (size_t)cl + k] = ncl[k];
                m.foc[3 * (size_t)cr + k] = fcr[k]; m.foc[3 * (size_t)cl + k] = fcl[k];
            }
            set_face(m, fD, cr, cl, bl, tr);
            if (ic == 0) { set_face(m, fL, cl, -1, tl, bl); zl.push_back(fL); } else set_face(m, fL, cl, cr - 2 * ny, tl, bl);
            if (jc == 0) { set_face(m, fB, cr, -1, bl, br); zb.push_back(fB); } else set_face(m, fB, cr, cr - 1, bl, br);
            if (ic == nx - 1) { set_face(m, fR, cr, -1, br, tr); zr.push_back(fR); }
            if (jc == ny - 1) { set_face(m, fT, cl, -1, tr, tl); zt.push_back(fT); }
        }
    make_zones(m, zr, zt, zl, zb, col * nx + ny);
}

}  // namespace

// Mesh::compute_face_areas / compute_cell_volumes / compute_cell_centroids / compute_face_normals
// (mesh/mesh.cpp:167-261), evaluated in the generators' order (:553-556).
void host_mesh_geometry(HostMesh & m) {
    const double * X = m.node_xy.data();
    m.face_area.assign(m.nf, 0.0);
    m.cell_vol.assign(m.nc, 0.0);
    m.cell_xy.assign(2 * (size_t)m.nc, 0.0);
    m.face_n.assign(2 * (size_t)m.nf, 0.0);
    for (uint32_t f = 0; f < m.nf; f++) {
        const double * a = &X[2 * (size_t)m.nof[m.onf[f]]], * b = &X[2 * (size_t)m.nof[m.onf[f] + 1]];
        const double ex = b[0] - a[0], ey = b[1] - a[1];
        m.face_area[f] = std::sqrt(ex * ex + ey * ey);
    }
    for (uint32_t c = 0; c < m.nc; c++) {
        const uint32_t * n = &m.noc[m.onc[c]];
        const int k = m.nnc(c);
        if (k == 3) m.cell_vol[c] = tri_area2(&X[2 * (size_t)n[0]], &X[2 * (size_t)n[1]], &X[2 * (size_t)n[2]]);
        else if (k == 4) {
            const double a1 = tri_area2(&X[2 * (size_t)n[0]], &X[2 * (size_t)n[1]], &X[2 * (size_t)n[2]]);
            const double a2 = tri_area2(&X[2 * (size_t)n[0]], &X[2 * (size_t)n[2]], &X[2 * (size_t)n[3]]);
            m.cell_vol[c] = a1 + a2;
        } else throw std::runtime_error("Unknown cell type.");
        double sx = 0.0, sy = 0.0;
        for (int i = 0; i < k; i++) { sx += X[2 * (size_t)n[i]]; sy += X[2 * (size_t)n[i] + 1]; }
        m.cell_xy[2 * (size_t)c] = sx / k;
        m.cell_xy[2 * (size_t)c + 1] = sy / k;
    }
    for (uint32_t f = 0; f < m.nf; f++) {
        const double * a = &X[2 * (size_t)m.nof[m.onf[f]]], * b = &X[2 * (size_t)m.nof[m.onf[f] + 1]];
        const double ex = b[0] - a[0], ey = b[1] - a[1];
        const double len = std::sqrt(ex * ex + ey * ey);
        double nx = ey / len * m.face_area[f], ny = -ex / len * m.face_area[f];
        const int32_t c0 = m.cof[2 * (size_t)f];
        const double rx = 0.5 * (a[0] + b[0]) - m.cell_xy[2 * (size_t)c0], ry = 0.5 * (a[1] + b[1]) - m.cell_xy[2 * (size_t)c0 + 1];
        if (rx * nx + ry * ny < 0) { nx *= -1; ny *= -1; }
        m.face_n[2 * (size_t)f] = nx;
        m.face_n[2 * (size_t)f + 1] = ny;
    }
}

void host_mesh_generate(HostMesh & m, int type, uint32_t nx, uint32_t ny, double Lx, double Ly) {
    if (nx == 0 || ny == 0) throw std::runtime_error("mesh: Nx and Ny must be positive");
    switch (type) {
        case MLB_MESH_CARTESIAN: gen_quads(m, nx, ny, Lx, Ly); break;
        case MLB_MESH_CARTESIAN_TRI: gen_tris(m, nx, ny, Lx, Ly); break;
        case MLB_MESH_WEDGE: {   // mesh/mesh.cpp:828-848: 8 degree ramp starting at x = 0.5
            gen_quads(m, nx, ny, Lx, Ly);
            const double theta = 8 * 3.141592653589793238462643383279502884 / 180.0, x_ramp = 0.5;
            for (uint32_t n = 0; n < m.nn; n++) {
                const double x = m.node_xy[2 * (size_t)n], y = m.node_xy[2 * (size_t)n + 1];
                if (x > x_ramp) {
                    const double floor_y = (x - x_ramp) * std::tan(theta);
                    m.node_xy[2 * (size_t)n + 1] = (y / Ly) * (Ly - floor_y) + floor_y;
                }
            }
            break;
        }
        default: throw std::runtime_error("Unknown mesh type.");
    }
    host_mesh_geometry(m);
}

void host_mesh_from_view(HostMesh & m, const mlb_mesh & v) {
    if (!v.node_coords || !v.offsets_nodes_of_cell || !v.nodes_of_cell || !v.offsets_faces_of_cell || !v.faces_of_cell ||
        !v.offsets_nodes_of_face || !v.nodes_of_face || !v.cells_of_face)
        throw std::runtime_error("mlb_mesh: connectivity pointers must not be NULL");
    m.nc = v.n_cells; m.nf = v.n_faces; m.nn = v.n_nodes;
    m.node_xy.assign(v.node_coords, v.node_coords + 2 * (size_t)m.nn);
    m.onc.assign(v.offsets_nodes_of_cell, v.offsets_nodes_of_cell + m.nc + 1);
    m.noc.assign(v.nodes_of_cell, v.nodes_of_cell + m.onc[m.nc]);
    m.ofc.assign(v.offsets_faces_of_cell, v.offsets_faces_of_cell + m.nc + 1);
    m.foc.assign(v.faces_of_cell, v.faces_of_cell + m.ofc[m.nc]);
    m.onf.assign(v.offsets_nodes_of_face, v.offsets_nodes_of_face + m.nf + 1);
    m.nof.assign(v.nodes_of_face, v.nodes_of_face + m.onf[m.nf]);
    m.cof.assign(v.cells_of_face, v.cells_of_face + 2 * (size_t)m.nf);
    for (uint32_t z = 0; z < v.n_zones; z++)
        m.zones.push_back({v.zones[z].name, uvec(v.zones[z].faces, v.zones[z].faces + v.zones[z].n_faces)});
    if (v.cell_coords && v.cell_volume && v.face_area && v.face_normals) {
        m.cell_xy.assign(v.cell_coords, v.cell_coords + 2 * (size_t)m.nc);
        m.cell_vol.assign(v.cell_volume, v.cell_volume + m.nc);
        m.face_area.assign(v.face_area, v.face_area + m.nf);
        m.face_n.assign(v.face_normals, v.face_normals + 2 * (size_t)m.nf);
    } else {
        host_mesh_geometry(m);
    }
}

GasParams make_gas(const mlb_physics & p) {
    GasParams g;
    g.gamma = p.gamma; g.p_min = p.p_min; g.p_max = p.p_max;
    g.R = p.p_ref / (p.T_ref * p.rho_ref);
    g.cp = g.R * p.gamma / (p.gamma - 1.0);
    g.cv = g.cp / p.gamma;
    g.mu = p.mu > 0.0 ? p.mu : 0.0;
    g.kappa = g.mu * g.cp / (p.Pr > 0.0 ? p.Pr : 0.72);
    return g;
}

}  // namespace mlb
