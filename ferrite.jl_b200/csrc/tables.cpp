// Reference shapes, Lagrange interpolations (orders 1-2) and quadrature tables (host side).
//
// What the reference defines in src/Grid/grid.jl:196-229 (reference_vertices/edges/faces),
// src/interpolations.jl (Lagrange shape functions :600-1158, entity dof tables :584-590,
// :667-668,:766,:918,:1075-1083, facedof/edgedof index composition :349-365,:402-422) and
// src/Quadrature/* (tensor Gauss-Legendre quadrature.jl:85-106, Dunavant gaussquad_tri_table.jl:8-31,
// Keast gaussquad_tet_table.jl:2-68).  Shape-function gradients are analytic (the reference
// differentiates the same formulas with ForwardDiff, src/interpolations.jl:280-292).
#include <array>
#include <cmath>
#include <cstring>

#include "common.h"

static RefShapeInfo make_shape(int ct, int rdim, int nv, std::vector<std::array<int, 2>> edges,
                               std::vector<std::vector<int>> faces) {
    RefShapeInfo s;
    memset(&s, 0, sizeof(s));
    s.celltype = ct;
    s.rdim = rdim;
    s.nvertices = nv;
    s.nedges = (int)edges.size();
    s.nfaces = (int)faces.size();
    for (int e = 0; e < s.nedges; ++e) {
        s.edges[e][0] = edges[e][0] - 1;
        s.edges[e][1] = edges[e][1] - 1;
    }
    for (int f = 0; f < s.nfaces; ++f) {
        s.face_nverts[f] = (int)faces[f].size();
        for (int k = 0; k < 4; ++k) s.faces[f][k] = k < (int)faces[f].size() ? faces[f][k] - 1 : -1;
        // reference_face_edgenrs, src/Grid/grid.jl:261-275
        int nfv = s.face_nverts[f];
        for (int k = 0; k < nfv; ++k) {
            int v1 = s.faces[f][k], v2 = s.faces[f][(k + 1) % nfv];
            s.face_edges[f][k] = -1;
            for (int e = 0; e < s.nedges; ++e)
                if ((s.edges[e][0] == v1 && s.edges[e][1] == v2) || (s.edges[e][0] == v2 && s.edges[e][1] == v1)) {
                    s.face_edges[f][k] = e;
                    break;
                }
        }
    }
    return s;
}



const RefShapeInfo* fb2_refshape(int celltype) {
    static const RefShapeInfo line = make_shape(FB2_LINE, 1, 2, {{1, 2}}, {});
    static const RefShapeInfo tri = make_shape(FB2_TRIANGLE, 2, 3, {{1, 2}, {2, 3}, {3, 1}}, {{1, 2, 3}});
    static const RefShapeInfo quad =
        make_shape(FB2_QUADRILATERAL, 2, 4, {{1, 2}, {2, 3}, {3, 4}, {4, 1}}, {{1, 2, 3, 4}});
    static const RefShapeInfo tet = make_shape(FB2_TETRAHEDRON, 3, 4, {{1, 2}, {2, 3}, {3, 1}, {1, 4}, {2, 4}, {3, 4}},
                                               {{1, 3, 2}, {1, 2, 4}, {2, 3, 4}, {1, 4, 3}});
    static const RefShapeInfo hex = make_shape(
        FB2_HEXAHEDRON, 3, 8,
        {{1, 2}, {2, 3}, {3, 4}, {4, 1}, {5, 6}, {6, 7}, {7, 8}, {8, 5}, {1, 5}, {2, 6}, {3, 7}, {4, 8}},
        {{1, 4, 3, 2}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 4, 8, 7}, {1, 5, 8, 4}, {5, 6, 7, 8}});
    switch (celltype) {
        case FB2_LINE: return &line;
        case FB2_TRIANGLE: return &tri;
        case FB2_QUADRILATERAL: return &quad;
        case FB2_TETRAHEDRON: return &tet;
        case FB2_HEXAHEDRON: return &hex;
    }
    return nullptr;
}

// ---- Lagrange ---------------------------------------------------------------------------------
static const double RC_LINE2[3][3] = {{-1, 0, 0}, {1, 0, 0}, {0, 0, 0}};
static const double RC_QUAD2[9][3] = {{-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 0},
                                      {1, 0, 0},   {0, 1, 0},  {-1, 0, 0}, {0, 0, 0}};
static const double RC_HEX2[27][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1},  {-1, 1, 1}, {0, -1, -1},
    {1, 0, -1},   {0, 1, -1},  {-1, 0, -1}, {0, -1, 1},  {1, 0, 1},   {0, 1, 1},  {-1, 0, 1}, {-1, -1, 0}, {1, -1, 0},
    {1, 1, 0},    {-1, 1, 0},  {0, 0, -1}, {0, -1, 0},  {1, 0, 0},   {0, 1, 0},  {-1, 0, 0}, {0, 0, 1},  {0, 0, 0}};
static const double RC_TRI2[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 0}, {0.5, 0.5, 0}, {0, 0.5, 0}, {0.5, 0, 0}};
static const double RC_TET2[10][3] = {{0, 0, 0},   {1, 0, 0},     {0, 1, 0},   {0, 0, 1},   {0.5, 0, 0},
                                      {0.5, 0.5, 0}, {0, 0.5, 0}, {0, 0, 0.5}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};

bool fb2_lagrange(int celltype, int order, LagrangeInfo* out) {
    const RefShapeInfo* rs = fb2_refshape(celltype);
    if (!rs || (order != 1 && order != 2)) return false;
    LagrangeInfo ip;
    memset(&ip, 0, sizeof(ip));
    ip.celltype = celltype;
    ip.order = order;
    ip.rdim = rs->rdim;
    ip.nvertexdofs = 1;
    const double(*rc)[3] = nullptr;
    int full = 0;
    switch (celltype) {
        case FB2_LINE: rc = RC_LINE2; full = 3; break;
        case FB2_QUADRILATERAL: rc = RC_QUAD2; full = 9; break;
        case FB2_HEXAHEDRON: rc = RC_HEX2; full = 27; break;
        case FB2_TRIANGLE: rc = RC_TRI2; full = 6; break;
        case FB2_TETRAHEDRON: rc = RC_TET2; full = 10; break;
    }
    ip.nbase = order == 1 ? rs->nvertices : full;
    for (int i = 0; i < ip.nbase; ++i)
        for (int d = 0; d < 3; ++d) ip.refcoords[i][d] = rc[i][d];
    if (order == 2) {
        bool cube = celltype == FB2_LINE || celltype == FB2_QUADRILATERAL || celltype == FB2_HEXAHEDRON;
        ip.nedgedofs = 1;
        // the cell interior of a 2-D cell is its single face; of a 1-D cell its single edge
        ip.nfacedofs = (cube && rs->rdim >= 2) ? 1 : 0;
        ip.nvolumedofs = (celltype == FB2_HEXAHEDRON) ? 1 : 0;
    }
    ip.edge_first = rs->nvertices;
    ip.face_first = ip.edge_first + rs->nedges * ip.nedgedofs;
    ip.vol_first = ip.face_first + rs->nfaces * ip.nfacedofs;
    *out = ip;
    return true;
}

static inline void l1d(int order, double node, double x, double* v, double* d) {
    if (order == 1) {
        if (node < 0) { *v = (1 - x) / 2; *d = -0.5; }
        else { *v = (1 + x) / 2; *d = 0.5; }
        return;
    }
    if (node < 0) { *v = -x * (1 - x) / 2; *d = x - 0.5; }
    else if (node > 0) { *v = x * (1 + x) / 2; *d = x + 0.5; }
    else { *v = (1 + x) * (1 - x); *d = -2 * x; }
}

void fb2_lagrange_eval(const LagrangeInfo& ip, const double* xi, double* N, double* dN) {
    const int n = ip.nbase, rd = ip.rdim;
    const int ct = ip.celltype;
    if (ct == FB2_LINE || ct == FB2_QUADRILATERAL || ct == FB2_HEXAHEDRON) {
        for (int i = 0; i < n; ++i) {
            double v[3] = {1, 1, 1}, d[3] = {0, 0, 0};
            for (int k = 0; k < rd; ++k) l1d(ip.order, ip.refcoords[i][k], xi[k], &v[k], &d[k]);
            double val = 1;
            for (int k = 0; k < rd; ++k) val *= v[k];
            N[i] = val;
            for (int k = 0; k < rd; ++k) {
                double g = d[k];
                for (int e = 0; e < rd; ++e)
                    if (e != k) g *= v[e];
                dN[i * rd + k] = g;
            }
        }
        return;
    }
    if (ct == FB2_TRIANGLE) {
        double x = xi[0], y = xi[1], g = 1 - x - y;
        if (ip.order == 1) {
            double Nv[3] = {x, y, g};
            double dv[3][2] = {{1, 0}, {0, 1}, {-1, -1}};
            for (int i = 0; i < 3; ++i) { N[i] = Nv[i]; dN[i * 2] = dv[i][0]; dN[i * 2 + 1] = dv[i][1]; }
        } else {
            double Nv[6] = {x * (2 * x - 1), y * (2 * y - 1), g * (2 * g - 1), 4 * x * y, 4 * y * g, 4 * x * g};
            double dv[6][2] = {{4 * x - 1, 0}, {0, 4 * y - 1}, {-(4 * g - 1), -(4 * g - 1)},
                               {4 * y, 4 * x}, {-4 * y, 4 * g - 4 * y}, {4 * g - 4 * x, -4 * x}};
            for (int i = 0; i < 6; ++i) { N[i] = Nv[i]; dN[i * 2] = dv[i][0]; dN[i * 2 + 1] = dv[i][1]; }
        }
        return;
    }
    // tetrahedron
    double x = xi[0], y = xi[1], z = xi[2], g = 1 - x - y - z;
    if (ip.order == 1) {
        double Nv[4] = {g, x, y, z};
        double dv[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < 4; ++i) { N[i] = Nv[i]; for (int k = 0; k < 3; ++k) dN[i * 3 + k] = dv[i][k]; }
    } else {
        double a = -(4 * g - 1);
        double Nv[10] = {(2 * g - 1) * g, x * (2 * x - 1), y * (2 * y - 1), z * (2 * z - 1), 4 * x * g,
                         4 * x * y,       4 * y * g,       4 * z * g,       4 * x * z,       4 * y * z};
        double dv[10][3] = {{a, a, a}, {4 * x - 1, 0, 0}, {0, 4 * y - 1, 0}, {0, 0, 4 * z - 1},
                            {4 * g - 4 * x, -4 * x, -4 * x}, {4 * y, 4 * x, 0}, {-4 * y, 4 * g - 4 * y, -4 * y},
                            {-4 * z, -4 * z, 4 * g - 4 * z}, {4 * z, 0, 4 * x}, {0, 4 * z, 4 * y}};
        for (int i = 0; i < 10; ++i) { N[i] = Nv[i]; for (int k = 0; k < 3; ++k) dN[i * 3 + k] = dv[i][k]; }
    }
}

std::vector<std::vector<int>> fb2_boundarydof_indices(const LagrangeInfo& ip, int kind) {
    const RefShapeInfo* rs = fb2_refshape(ip.celltype);
    std::vector<std::vector<int>> out;
    if (kind == FB2_BC_FACET) kind = rs->rdim == 3 ? FB2_BC_FACE : (rs->rdim == 2 ? FB2_BC_EDGE : FB2_BC_VERTEX);
    if (kind == FB2_BC_VERTEX) {
        for (int v = 0; v < rs->nvertices; ++v) out.push_back({v});
    } else if (kind == FB2_BC_EDGE) {
        for (int e = 0; e < rs->nedges; ++e) {
            std::vector<int> d = {rs->edges[e][0], rs->edges[e][1]};
            for (int k = 0; k < ip.nedgedofs; ++k) d.push_back(ip.edge_first + e * ip.nedgedofs + k);
            out.push_back(d);
        }
    } else if (kind == FB2_BC_FACE) {
        for (int f = 0; f < rs->nfaces; ++f) {
            std::vector<int> d;
            for (int k = 0; k < rs->face_nverts[f]; ++k) d.push_back(rs->faces[f][k]);
            for (int k = 0; k < rs->face_nverts[f]; ++k)
                for (int j = 0; j < ip.nedgedofs; ++j) d.push_back(ip.edge_first + rs->face_edges[f][k] * ip.nedgedofs + j);
            for (int k = 0; k < ip.nfacedofs; ++k) d.push_back(ip.face_first + f * ip.nfacedofs + k);
            out.push_back(d);
        }
    }
    return out;
}

// ---- quadrature --------------------------------------------------------------------------------
static void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
    x.resize(n);
    w.resize(n);
    for (int i = 0; i < n; ++i) {
        // Newton iteration on P_n, ascending order
        double z = -std::cos(M_PI * (i + 0.75) / (n + 0.5));
        double pp = 0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1, p2 = 0;
            for (int j = 0; j < n; ++j) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1);
            }
            pp = n * (z * p1 - p2) / (z * z - 1);
            double z1 = z;
            z = z1 - p1 / pp;
            if (std::fabs(z - z1) < 1e-16) break;
        }
        // one more evaluation for the weight at the converged point
        double p1 = 1, p2 = 0;
        for (int j = 0; j < n; ++j) {
            double p3 = p2;
            p2 = p1;
            p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1);
        }
        pp = n * (z * p1 - p2) / (z * z - 1);
        x[i] = z;
        w[i] = 2.0 / ((1 - z * z) * pp * pp);
    }
    // enforce exact antisymmetry like a symmetric eigen-solve would
    for (int i = 0; i < n / 2; ++i) {
        double a = 0.5 * (x[n - 1 - i] - x[i]);
        x[i] = -a;
        x[n - 1 - i] = a;
        double ww = 0.5 * (w[i] + w[n - 1 - i]);
        w[i] = w[n - 1 - i] = ww;
    }
    if (n % 2) x[n / 2] = 0.0;
}

bool fb2_quadrature(int celltype, int order, std::vector<double>* w, std::vector<double>* pts) {
    w->clear();
    pts->clear();
    if (order < 1) return false;
    if (celltype == FB2_LINE || celltype == FB2_QUADRILATERAL || celltype == FB2_HEXAHEDRON) {
        int dim = celltype == FB2_LINE ? 1 : (celltype == FB2_QUADRILATERAL ? 2 : 3);
        if (order > 8) return false;
        std::vector<double> p1, w1;
        gauss_legendre(order, p1, w1);
        int nq = 1;
        for (int d = 0; d < dim; ++d) nq *= order;
        for (int q = 0; q < nq; ++q) {
            int idx[3], r = q;
            for (int d = 0; d < dim; ++d) { idx[d] = r % order; r /= order; }  // i_1 fastest
            double wt = 1.0;
            for (int d = 0; d < dim; ++d) { pts->push_back(p1[idx[d]]); wt *= w1[idx[d]]; }
            w->push_back(wt);
        }
        return true;
    }
    if (celltype == FB2_TRIANGLE) {
        // Dunavant constants as published by the reference (truncated to 14 digits)
        if (order == 1) { *pts = {0.33333333333333, 0.33333333333333}; *w = {1.00000000000000 / 2.0}; return true; }
        if (order == 2) {
            *pts = {0.16666666666667, 0.16666666666667, 0.16666666666667, 0.66666666666667, 0.66666666666667, 0.16666666666667};
            *w = {0.33333333333333 / 2.0, 0.33333333333333 / 2.0, 0.33333333333333 / 2.0};
            return true;
        }
        if (order == 3) {
            *pts = {0.33333333333333, 0.33333333333333, 0.20000000000000, 0.20000000000000,
                    0.20000000000000, 0.60000000000000, 0.60000000000000, 0.20000000000000};
            *w = {-0.56250000000000 / 2.0, 0.52083333333333 / 2.0, 0.52083333333333 / 2.0, 0.52083333333333 / 2.0};
            return true;
        }
        return false;
    }
    if (celltype == FB2_TETRAHEDRON) {
        auto add = [&](double a, double b, double c, double wt) {
            pts->push_back(a); pts->push_back(b); pts->push_back(c); w->push_back(wt);
        };
        if (order == 1) { add(0.25, 0.25, 0.25, 1.0 / 6.0); return true; }
        if (order == 2) {
            double a = (5.0 + 3.0 * std::sqrt(5.0)) / 20.0, b = (5.0 - std::sqrt(5.0)) / 20.0, wt = 1.0 / 24.0;
            add(a, b, b, wt); add(b, a, b, wt); add(b, b, a, wt); add(b, b, b, wt);
            return true;
        }
        if (order == 3) {
            double a1 = 1.0 / 4.0, a2 = 1.0 / 2.0, b2 = 1.0 / 6.0, w1 = -2.0 / 15.0, w2 = 3.0 / 40.0;
            add(a1, a1, a1, w1); add(a2, b2, b2, w2); add(b2, a2, b2, w2); add(b2, b2, a2, w2); add(b2, b2, b2, w2);
            return true;
        }
        if (order == 4) {
            double a1 = 1.0 / 4.0, w1 = -74.0 / 5625.0;
            double a2 = 5.0 / 70.0, b2 = 11.0 / 14.0, w2 = 343.0 / 45000.0;
            double a3 = (1.0 + std::sqrt(5.0 / 14.0)) / 4.0, b3 = (1.0 - std::sqrt(5.0 / 14.0)) / 4.0, w3 = 28.0 / 1125.0;
            add(a1, a1, a1, w1);
            add(b2, a2, a2, w2); add(a2, b2, a2, w2); add(a2, a2, b2, w2); add(a2, a2, a2, w2);
            add(a3, a3, b3, w3); add(a3, b3, a3, w3); add(a3, b3, b3, w3);
            add(b3, a3, a3, w3); add(b3, a3, b3, w3); add(b3, b3, a3, w3);
            return true;
        }
        return false;
    }
    return false;
}
