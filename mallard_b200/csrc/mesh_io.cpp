// Mesh file reader / writer behind the C ABI.  The reference's `[mesh] type = "file"` is a stub that throws
// (mesh/mesh.cpp:41-43), so there is no reference behaviour to be faithful to; what is produced are the arrays of the
// reference's Mesh (mesh/mesh.h:228-253) under these conventions:
//   * Gmsh MSH 2.2 ASCII, 2-D: triangles (element type 2) and quadrilaterals (type 3) become cells in file order, 2-node
//     lines (type 1) with a physical tag name the boundary zones ($PhysicalNames); point elements are ignored;
//   * faces are the unique cell edges, numbered by ascending (lower node, higher node); cells_of_face[f][0] is the
//     lower-numbered cell and the face's nodes run in that cell's orientation; face j of a cell is the edge (node j, node j+1);
//   * zone "interior" = the two-cell faces, ascending; one zone per tagged physical line group (name from $PhysicalNames, else
//     "zone_<tag>"), faces ascending; boundary faces no line element claims go to the zone "boundary";
//   * geometry by host_mesh_geometry(), i.e. exactly as Mesh::compute_* computes it for the generated meshes.
// Host only.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>

#include "mlb_internal.h"

namespace mlb {

namespace {

struct BoundaryEdge { uint32_t a, b; int tag; };

void build_from_cells(HostMesh & m, const dvec & node_xy, const uvec & onc, const uvec & noc, const std::vector<BoundaryEdge> & bedges,
                      const std::map<int, std::string> & names) {
    m = HostMesh();
    m.nn = (uint32_t)(node_xy.size() / 2);
    m.nc = (uint32_t)onc.size() - 1;
    m.node_xy = node_xy; m.onc = onc; m.noc = noc; m.ofc = onc;
    const uint64_t stride = (uint64_t)m.nn + 1;
    struct Edge { uint64_t key; uint32_t cell, a, b, pos; };
    std::vector<Edge> edges(noc.size());
    for (uint32_t c = 0; c < m.nc; c++) {
        const uint32_t n = onc[c + 1] - onc[c];
        for (uint32_t j = 0; j < n; j++) {
            const uint32_t a = noc[onc[c] + j], b = noc[onc[c] + (j + 1) % n];
            if (a >= m.nn || b >= m.nn) throw std::runtime_error("mesh file: node index out of range");
            edges[onc[c] + j] = {std::min<uint64_t>(a, b) * stride + std::max<uint64_t>(a, b), c, a, b, onc[c] + j};
        }
    }
    std::vector<Edge> sorted = edges;
    std::stable_sort(sorted.begin(), sorted.end(), [](const Edge & x, const Edge & y) { return x.key < y.key; });   // ties stay in ascending cell order
    m.foc.assign(noc.size(), 0);
    std::map<uint64_t, uint32_t> face_of_key;
    for (size_t i = 0; i < sorted.size();) {
        size_t j = i;
        while (j < sorted.size() && sorted[j].key == sorted[i].key) j++;
        if (j - i > 2) throw std::runtime_error("mesh file: an edge is shared by more than two cells");
        const uint32_t f = m.nf++;
        m.cof.push_back((int32_t)sorted[i].cell);
        m.cof.push_back(j - i == 2 ? (int32_t)sorted[i + 1].cell : -1);
        m.nof.push_back(sorted[i].a); m.nof.push_back(sorted[i].b);
        for (size_t k = i; k < j; k++) m.foc[sorted[k].pos] = f;
        face_of_key[sorted[i].key] = f;
        i = j;
    }
    m.onf.resize((size_t)m.nf + 1);
    for (uint32_t f = 0; f <= m.nf; f++) m.onf[f] = 2 * f;
    HostZone interior{"interior", {}};
    std::vector<char> claimed(m.nf, 0);
    for (uint32_t f = 0; f < m.nf; f++) if (m.cof[2 * (size_t)f + 1] >= 0) interior.faces.push_back(f);
    m.zones.push_back(interior);
    std::map<int, uvec> by_tag;
    for (const BoundaryEdge & e : bedges) {
        const auto it = face_of_key.find(std::min<uint64_t>(e.a, e.b) * stride + std::max<uint64_t>(e.a, e.b));
        if (it == face_of_key.end() || m.cof[2 * (size_t)it->second + 1] >= 0)
            throw std::runtime_error("mesh file: boundary line (" + std::to_string(e.a) + ", " + std::to_string(e.b) + ") is not a boundary edge of any cell");
        by_tag[e.tag].push_back(it->second);
        claimed[it->second] = 1;
    }
    for (auto & kv : by_tag) {
        std::sort(kv.second.begin(), kv.second.end());
        const auto nm = names.find(kv.first);
        m.zones.push_back({nm != names.end() ? nm->second : "zone_" + std::to_string(kv.first), kv.second});
    }
    HostZone rest{"boundary", {}};
    for (uint32_t f = 0; f < m.nf; f++) if (m.cof[2 * (size_t)f + 1] < 0 && !claimed[f]) rest.faces.push_back(f);
    if (!rest.faces.empty()) m.zones.push_back(rest);
    host_mesh_geometry(m);
}

std::string trimmed(const std::string & s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

}  // namespace

void host_mesh_read_gmsh(HostMesh & m, const char * path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error(std::string("mesh file: cannot open ") + path);
    std::map<std::string, std::vector<std::string>> sec;
    std::string line, cur;
    while (std::getline(in, line)) {
        line = trimmed(line);
        if (line.empty()) continue;
        if (line[0] == '$') {
            if (line.rfind("$End", 0) == 0) cur.clear(); else { cur = line.substr(1); sec[cur]; }
        } else if (!cur.empty()) sec[cur].push_back(line);
    }
    {
        const auto it = sec.find("MeshFormat");
        std::string ver, type;
        if (it != sec.end() && !it->second.empty()) { std::istringstream ss(it->second[0]); ss >> ver >> type; }
        if (ver.empty() || ver[0] != '2' || (!type.empty() && type != "0"))
            throw std::runtime_error("mesh file: only Gmsh MSH 2.x ASCII is supported (got '" + (it != sec.end() && !it->second.empty() ? it->second[0] : std::string("?")) + "')");
    }
    std::map<int, std::string> names;   // physical names of dimension 1
    if (sec.count("PhysicalNames"))
        for (size_t i = 1; i < sec["PhysicalNames"].size(); i++) {
            std::istringstream ss(sec["PhysicalNames"][i]);
            int dim, tag;
            ss >> dim >> tag;
            std::string name;
            std::getline(ss, name);
            name = trimmed(name);
            if (name.size() >= 2 && name.front() == '"' && name.back() == '"') name = name.substr(1, name.size() - 2);
            if (dim == 1) names[tag] = name;
        }
    if (!sec.count("Nodes") || sec["Nodes"].empty()) throw std::runtime_error("mesh file: no $Nodes section");
    const std::vector<std::string> & nl = sec["Nodes"];
    const size_t nn = std::stoul(nl[0]);
    if (nl.size() < nn + 1) throw std::runtime_error("mesh file: $Nodes is truncated");
    dvec xy(2 * nn);
    std::map<long, uint32_t> index;
    for (size_t k = 0; k < nn; k++) {
        std::istringstream ss(nl[k + 1]);
        long id; double x, y;
        if (!(ss >> id >> x >> y)) throw std::runtime_error("mesh file: bad node line '" + nl[k + 1] + "'");
        index[id] = (uint32_t)k; xy[2 * k] = x; xy[2 * k + 1] = y;
    }
    uvec onc{0}, noc;
    std::vector<BoundaryEdge> bedges;
    if (sec.count("Elements"))
        for (size_t i = 1; i < sec["Elements"].size(); i++) {
            std::istringstream ss(sec["Elements"][i]);
            long id; int typ, ntags;
            if (!(ss >> id >> typ >> ntags)) throw std::runtime_error("mesh file: bad element line");
            const int n_nodes = typ == 1 ? 2 : typ == 2 ? 3 : typ == 3 ? 4 : typ == 15 ? 1 : 0;
            if (!n_nodes) throw std::runtime_error("mesh file: unsupported element type " + std::to_string(typ));
            int tag = 0;
            for (int t = 0; t < ntags; t++) { int v; ss >> v; if (t == 0) tag = v; }
            uint32_t nodes[4];
            for (int k = 0; k < n_nodes; k++) {
                long nid;
                if (!(ss >> nid) || !index.count(nid)) throw std::runtime_error("mesh file: element refers to an unknown node");
                nodes[k] = index[nid];
            }
            if (typ == 2 || typ == 3) { for (int k = 0; k < n_nodes; k++) noc.push_back(nodes[k]); onc.push_back((uint32_t)noc.size()); }
            else if (typ == 1) bedges.push_back({nodes[0], nodes[1], tag});
        }
    if (onc.size() == 1) throw std::runtime_error("mesh file: no triangles or quadrilaterals");
    build_from_cells(m, xy, onc, noc, bedges, names);
}

void host_mesh_from_cells(HostMesh & m, uint32_t n_nodes, const double * node_xy, uint32_t n_cells, const uint32_t * onc, const uint32_t * noc,
                          uint32_t n_edges, const uint32_t * edge_nodes, const int32_t * edge_tags, uint32_t n_names, const int32_t * name_tags,
                          const char * const * names) {
    if (!node_xy || !onc || !noc || (n_edges && (!edge_nodes || !edge_tags)) || (n_names && (!name_tags || !names)))
        throw std::runtime_error("mesh from cells: NULL argument");
    std::vector<BoundaryEdge> be(n_edges);
    for (uint32_t i = 0; i < n_edges; i++) be[i] = {edge_nodes[2 * (size_t)i], edge_nodes[2 * (size_t)i + 1], edge_tags[i]};
    std::map<int, std::string> nm;
    for (uint32_t i = 0; i < n_names; i++) nm[name_tags[i]] = names[i];
    for (uint32_t c = 0; c < n_cells; c++) {
        const uint32_t k = onc[c + 1] - onc[c];
        if (k != 3 && k != 4) throw std::runtime_error("mesh from cells: cells must have three or four nodes");
    }
    build_from_cells(m, dvec(node_xy, node_xy + 2 * (size_t)n_nodes), uvec(onc, onc + n_cells + 1), uvec(noc, noc + onc[n_cells]), be, nm);
}

void host_mesh_write_gmsh(const mlb_mesh & v, const char * path) {
    FILE * fh = fopen(path, "w");
    if (!fh) throw std::runtime_error(std::string("mesh file: cannot write ") + path);
    std::vector<const mlb_zone *> bz;
    for (uint32_t z = 0; z < v.n_zones; z++) if (std::string(v.zones[z].name) != "interior") bz.push_back(&v.zones[z]);
    fprintf(fh, "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$PhysicalNames\n%zu\n", bz.size() + 1);
    for (size_t t = 0; t < bz.size(); t++) fprintf(fh, "1 %zu \"%s\"\n", t + 1, bz[t]->name);
    fprintf(fh, "2 %zu \"fluid\"\n$EndPhysicalNames\n$Nodes\n%u\n", bz.size() + 1, v.n_nodes);
    for (uint32_t k = 0; k < v.n_nodes; k++) fprintf(fh, "%u %.17g %.17g 0\n", k + 1, v.node_coords[2 * (size_t)k], v.node_coords[2 * (size_t)k + 1]);
    size_t n_el = v.n_cells;
    for (auto * z : bz) n_el += z->n_faces;
    fprintf(fh, "$EndNodes\n$Elements\n%zu\n", n_el);
    size_t e = 1;
    for (size_t t = 0; t < bz.size(); t++)
        for (uint32_t i = 0; i < bz[t]->n_faces; i++) {
            const uint32_t f = bz[t]->faces[i], o = v.offsets_nodes_of_face[f];
            fprintf(fh, "%zu 1 2 %zu %zu %u %u\n", e++, t + 1, t + 1, v.nodes_of_face[o] + 1, v.nodes_of_face[o + 1] + 1);
        }
    for (uint32_t c = 0; c < v.n_cells; c++) {
        const uint32_t o = v.offsets_nodes_of_cell[c], n = v.offsets_nodes_of_cell[c + 1] - o;
        if (n != 3 && n != 4) { fclose(fh); throw std::runtime_error("mesh file: only triangles and quadrilaterals can be written"); }
        fprintf(fh, "%zu %d 2 %zu %zu", e++, n == 3 ? 2 : 3, bz.size() + 1, bz.size() + 1);
        for (uint32_t k = 0; k < n; k++) fprintf(fh, " %u", v.nodes_of_cell[o + k] + 1);
        fprintf(fh, "\n");
    }
    fprintf(fh, "$EndElements\n");
    if (fclose(fh) != 0) throw std::runtime_error(std::string("mesh file: write failed: ") + path);
}

}  // namespace mlb
