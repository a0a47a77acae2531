"""CellValues + element routines (oracle; test infrastructure only).

Restates, vectorised over cells with numpy:
  * `CellValues(qr, ip, ip_geo)` tables: src/FEValues/CellValues.jl:57-83,
    GeometryMapping.jl:33-102, FunctionValues.jl:41-98
  * `reinit!`: J = sum_j x_j (x) dM_j/dxi (GeometryMapping.jl:124-130), detJ > 0 check and
    dOmega = detJ*w (CellValues.jl:110-115,134-138), dNdx = dNdxi . inv(J)
    (FunctionValues.jl:163,186-192)
  * element routines = the reference's tutorial kernels:
      heat        docs/src/literate-tutorials/heat_equation.jl:143-164
      elasticity  docs/src/literate-howto/threaded_assembly.jl:105-119 with the isotropic
                  C of benchmark/helper.jl:281-287
      neo-hooke   docs/src/literate-tutorials/hyperelasticity.jl:162-176,241-276 (volume terms;
                  S and dS/dC in closed form instead of Tensors.jl AD of Psi)
Local dof of (scalar basis a, component c) = (a-1)*vdim + c (src/interpolations.jl:1819-1825).
"""
import numpy as np

from .interpolations import geometric_interpolation

__all__ = ["CellValues", "reinit", "element_heat", "element_mass", "element_elasticity",
           "element_neohooke", "element_elasticity_general", "isotropic_stiffness", "lame", "ELEMENTS", "DetJError"]


class DetJError(ArithmeticError):
    """det(J) is not positive (src/FEValues/common_values.jl:5 `throw_detJ_not_pos`)."""


class CellValues:
    def __init__(self, qr, ip, ip_geo=None):
        self.qr = qr
        self.ip = ip
        self.base = ip.base
        self.vdim = ip.vdim
        self.ip_geo = (ip_geo.base if ip_geo is not None else geometric_interpolation(ip.shape))
        nq = qr.nq
        self.N = np.zeros((nq, self.base.nbase))
        self.dNdxi = np.zeros((nq, self.base.nbase, self.base.rdim))
        self.M = np.zeros((nq, self.ip_geo.nbase))
        self.dMdxi = np.zeros((nq, self.ip_geo.nbase, self.ip_geo.rdim))
        for q in range(nq):
            self.N[q], self.dNdxi[q] = self.base.value_and_gradient(qr.points[q])
            self.M[q], self.dMdxi[q] = self.ip_geo.value_and_gradient(qr.points[q])
        self.w = qr.weights.copy()

    @property
    def nq(self):
        return len(self.w)

    @property
    def nbase(self):
        return self.base.nbase * self.vdim


def _det_inv(J):
    d = J.shape[-1]
    if d == 1:
        det = J[..., 0, 0]
        inv = 1.0 / J
    elif d == 2:
        a, b, c, e = J[..., 0, 0], J[..., 0, 1], J[..., 1, 0], J[..., 1, 1]
        det = a * e - b * c
        inv = np.empty_like(J)
        inv[..., 0, 0], inv[..., 0, 1], inv[..., 1, 0], inv[..., 1, 1] = e / det, -b / det, -c / det, a / det
    else:
        det = (J[..., 0, 0] * (J[..., 1, 1] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 1])
               - J[..., 0, 1] * (J[..., 1, 0] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 0])
               + J[..., 0, 2] * (J[..., 1, 0] * J[..., 2, 1] - J[..., 1, 1] * J[..., 2, 0]))
        inv = np.empty_like(J)
        inv[..., 0, 0] = (J[..., 1, 1] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 1])
        inv[..., 0, 1] = -(J[..., 0, 1] * J[..., 2, 2] - J[..., 0, 2] * J[..., 2, 1])
        inv[..., 0, 2] = (J[..., 0, 1] * J[..., 1, 2] - J[..., 0, 2] * J[..., 1, 1])
        inv[..., 1, 0] = -(J[..., 1, 0] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 0])
        inv[..., 1, 1] = (J[..., 0, 0] * J[..., 2, 2] - J[..., 0, 2] * J[..., 2, 0])
        inv[..., 1, 2] = -(J[..., 0, 0] * J[..., 1, 2] - J[..., 0, 2] * J[..., 1, 0])
        inv[..., 2, 0] = (J[..., 1, 0] * J[..., 2, 1] - J[..., 1, 1] * J[..., 2, 0])
        inv[..., 2, 1] = -(J[..., 0, 0] * J[..., 2, 1] - J[..., 0, 1] * J[..., 2, 0])
        inv[..., 2, 2] = (J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0])
        inv = inv / det[..., None, None]
    return det, inv


def reinit(cv, x):
    """x: (ncells, ngeo, sdim) -> dNdx (ncells, nq, n, sdim), dOmega (ncells, nq)."""
    assert x.shape[-1] == cv.base.rdim, "oracle: embedded elements out of scope"
    J = np.einsum("cja,qjb->cqab", x, cv.dMdxi)
    det, Jinv = _det_inv(J)
    if not np.all(det > 0):
        bad = np.argwhere(~(det > 0))[0]
        raise DetJError(f"det(J) is not positive: det(J) = {det[tuple(bad)]} in cell {bad[0] + 1}")
    dNdx = np.einsum("qia,cqab->cqib", cv.dNdxi, Jinv)
    dOm = det * cv.w[None, :]
    return dNdx, dOm


def lame(E, nu):
    """benchmark/helper.jl:281-287"""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    return lam, mu


def element_heat(cv, x, params=None, u=None):
    """Ke = int grad(Ni).grad(Nj), fe = int Ni * source.  params = {'source': 1.0, 'k': 1.0}"""
    p = dict(source=1.0, k=1.0)
    p.update(params or {})
    dNdx, dOm = reinit(cv, x)
    Ke = p["k"] * np.einsum("cqia,cqja,cq->cij", dNdx, dNdx, dOm)
    fe = p["source"] * np.einsum("qi,cq->ci", cv.N, dOm)
    return Ke, fe


def element_mass(cv, x, params=None, u=None):
    """Me = int rho Ni Nj (scalar or vector field), fe = 0."""
    p = dict(rho=1.0)
    p.update(params or {})
    _, dOm = reinit(cv, x)
    Ms = p["rho"] * np.einsum("qi,qj,cq->cij", cv.N, cv.N, dOm)
    v = cv.vdim
    nb = cv.base.nbase
    Ke = np.zeros((x.shape[0], nb * v, nb * v))
    for c in range(v):
        Ke[:, c::v, c::v] = Ms
    return Ke, np.zeros((x.shape[0], nb * v))


def element_elasticity(cv, x, params=None, u=None):
    """Ke[(a,c),(b,d)] = int lam g_a[c] g_b[d] + mu (g_a[d] g_b[c] + delta_cd g_a.g_b);
    fe[(a,c)] = int N_a b_c.   params = {'lambda','mu','b'}"""
    p = dict(b=None)
    p.update(params or {})
    lam, mu = p["lambda"], p["mu"]
    v = cv.vdim
    b = np.zeros(v) if p["b"] is None else np.asarray(p["b"], dtype=np.float64)
    dNdx, dOm = reinit(cv, x)
    assert dNdx.shape[-1] == v
    nb = cv.base.nbase
    t1 = np.einsum("cqae,cqbd,cq->caebd", dNdx, dNdx, dOm)   # t1[c,a,e,b,d] = int g_a[e] g_b[d]
    gg = np.einsum("cqae,cqbe,cq->cab", dNdx, dNdx, dOm)
    K5 = lam * t1 + mu * np.swapaxes(t1, 2, 4)       # lam g_a[e] g_b[d] + mu g_a[d] g_b[e]
    for e in range(v):
        K5[:, :, e, :, e] += mu * gg
    Ke = K5.reshape(x.shape[0], nb * v, nb * v)
    fe = np.einsum("qa,cq,e->cae", cv.N, dOm, b).reshape(x.shape[0], nb * v)
    return Ke, fe


def element_elasticity_general(cv, x, params=None, u=None):
    """Linear elasticity with an arbitrary 4th-order stiffness C (the routine of docs/src/literate-tutorials/
    linear_elasticity.jl:266-281 / benchmark/helper.jl:249-262: Ke[I,J] += (grad(dN_I) : C : grad^sym(N_J)) dOmega; with the minor
    symmetries of a SymmetricTensor{4} the symmetric part is implied): Ke[(a,c),(b,d)] = int g_a[j] C[c,j,d,n] g_b[n],
    fe[(a,c)] = int N_a b_c.   params = {'C': (dim,dim,dim,dim) array, 'b'}"""
    p = dict(b=None)
    p.update(params or {})
    v = cv.vdim
    C = np.asarray(p["C"], dtype=np.float64)
    assert C.shape == (v, v, v, v)
    b = np.zeros(v) if p["b"] is None else np.asarray(p["b"], dtype=np.float64)
    dNdx, dOm = reinit(cv, x)
    nb = cv.base.nbase
    Ke = np.einsum("cqaj,ejdn,cqbn,cq->caebd", dNdx, C, dNdx, dOm).reshape(x.shape[0], nb * v, nb * v)
    fe = np.einsum("qa,cq,e->cae", cv.N, dOm, b).reshape(x.shape[0], nb * v)
    return Ke, fe


def isotropic_stiffness(lam, mu, dim=3):
    """C_ijkl = lam d_ij d_kl + mu (d_ik d_jl + d_il d_jk)"""
    d = np.eye(dim)
    return lam * np.einsum("ij,kl->ijkl", d, d) + mu * (np.einsum("ik,jl->ijkl", d, d) + np.einsum("il,jk->ijkl", d, d))


def _inv3_sym(C):
    det, inv = _det_inv(C)
    return det, inv


def element_neohooke(cv, x, params=None, u=None):
    """Tangent ke and residual ge (volume terms) for Psi = mu/2 (Ic-3-2 ln J) + lam/2 (J-1)^2.
    u: (ncells, n*3) cell-local displacement dofs.  params = {'lambda','mu','b'}"""
    p = dict(b=None)
    p.update(params or {})
    lam, mu = p["lambda"], p["mu"]
    b = np.zeros(3) if p["b"] is None else np.asarray(p["b"], dtype=np.float64)
    dNdx, dOm = reinit(cv, x)
    nc, nq, nb, _ = dNdx.shape
    ue = u.reshape(nc, nb, 3)
    I3 = np.eye(3)
    gradu = np.einsum("cai,cqaj->cqij", ue, dNdx)
    F = I3 + gradu
    C = np.einsum("cqki,cqkj->cqij", F, F)
    detC, Ci = _inv3_sym(C)
    if not np.all(detC > 0):
        raise DetJError("det(C) is not positive")
    Jd = np.sqrt(detC)
    S = mu * (I3 - Ci) + (lam * Jd * (Jd - 1))[..., None, None] * Ci
    c1 = (mu - lam * Jd * (Jd - 1))[..., None, None, None, None]
    c2 = (lam * (2 * Jd - 1) * (Jd / 2))[..., None, None, None, None]
    dSdC = c1 * 0.5 * (np.einsum("cqik,cqlj->cqijkl", Ci, Ci) + np.einsum("cqil,cqkj->cqijkl", Ci, Ci)) \
        + c2 * np.einsum("cqij,cqkl->cqijkl", Ci, Ci)
    P = np.einsum("cqia,cqaj->cqij", F, S)
    dPdF = np.einsum("im,cqjn->cqijmn", I3, S) + 2.0 * np.einsum("cqia,cqajkn,cqmk->cqijmn", F, dSdC, F)
    ge = (np.einsum("cqaj,cqej,cq->cae", dNdx, P, dOm)
          - np.einsum("qa,e,cq->cae", cv.N, b, dOm)).reshape(nc, nb * 3)
    ke = np.einsum("cqaj,cqejdn,cqbn,cq->caebd", dNdx, dPdF, dNdx, dOm).reshape(nc, nb * 3, nb * 3)
    return ke, ge


ELEMENTS = {
    "heat": element_heat,
    "mass": element_mass,
    "elasticity": element_elasticity,
    "elasticity_general": element_elasticity_general,
    "neohooke": element_neohooke,
}
