"""Quadrature rules of the reference (host mirror of src/quadrature.jl).

`QuadratureRule(geometry, order)` returns the same points/weights the reference's
`QuadratureRule{T,EG}(order)` constructs (quadrature.jl:130-148 Edge1D, 173-195 Triangle2D, 268-325
Tetrahedron3D, 332-502 symmetric rules, 528-562 Stroud conical product).  In production
the Julia host passes `qf.xref` / `qf.w` to libgrmp_cuda unchanged; this mirror produces
them in-container.  Weights sum to 1 (the cell volume is applied later,
bilinearform.jl:320).
"""
from __future__ import annotations

import numpy as np


class QuadratureRule:
    def __init__(self, geometry: str, order: int):
        self.geometry = geometry
        self.order = order
        if geometry == "Edge1D":
            self.xref, self.w, self.name = _edge(order)
        elif geometry == "Triangle2D":
            self.xref, self.w, self.name = _triangle(order)
        elif geometry == "Tetrahedron3D":
            self.xref, self.w, self.name = _tetrahedron(order)
        else:
            raise ValueError(f"no quadrature rule for {geometry}")
        self.xref = np.ascontiguousarray(self.xref, dtype=np.float64)
        self.w = np.ascontiguousarray(self.w, dtype=np.float64)

    def __len__(self):
        return self.w.size


def _edge(order):
    """quadrature.jl:130-148, generic Gauss rule 506-525"""
    if order <= 1:
        return np.array([[0.5]]), np.array([1.0]), "midpoint rule"
    if order == 2:
        return np.array([[0.0], [0.5], [1.0]]), np.array([1.0 / 6, 2.0 / 3, 1.0 / 6]), "Simpson's rule"
    n = order // 2 + 1
    k = np.arange(1, n)
    gamma = k / np.sqrt(4.0 * k**2 - 1.0)
    r, V = np.linalg.eigh(np.diag(gamma, 1) + np.diag(gamma, -1))
    w = 2 * V[0, :] ** 2
    return (0.5 * r + 0.5)[:, None], 0.5 * w, f"generic Gauss rule of order {order}"


def _stroud(order):
    n = order // 2 + 1
    k = np.arange(1, n)
    gamma = k / np.sqrt(4.0 * k**2 - 1.0)
    r, V = np.linalg.eigh(np.diag(gamma, 1) + np.diag(gamma, -1))
    a = 2 * V[0, :] ** 2
    k1 = np.arange(1, n + 1)
    delta = -1.0 / (4.0 * k1**2 - 1.0)
    gamma = np.sqrt((k + 1.0) * k) / (2.0 * (k + 1.0) - 1.0)
    s, V = np.linalg.eigh(np.diag(delta) + np.diag(gamma, 1) + np.diag(gamma, -1))
    b = 2 * V[0, :] ** 2
    r = 0.5 * r + 0.5
    s = 0.5 * s + 0.5
    a = 0.5 * a
    b = 0.5 * b
    xref, w = [], []
    for js in range(n):
        for ir in range(n):
            t = r[ir] * (s[js] - 1)
            xref.append([s[js] * 1.0 - t * 0.0, s[js] * 0.0 - t * 1.0])
            w.append(a[ir] * b[js])
    return np.array(xref), np.array(w), f"generic Stroud rule of order {order}"


def _sym_tri8():
    wS3 = .1443156076777871682510911104890646
    aS21 = [.1705693077517602066222935014914645, .0505472283170309754584235505965989, .4592925882927231560288155144941693]
    wS21 = [.1032173705347182502817915502921290, .0324584976231980803109259283417806, .0950916342672846247938961043885843]
    a, b = .2631128296346381134217857862846436, .0083947774099576053372138345392944
    wS111 = .0272303141744349942648446900739089
    x, w = [[1.0 / 3, 1.0 / 3]], [wS3]
    for aj, wj in zip(aS21, wS21):
        x += [[aj, aj], [aj, 1 - 2 * aj], [1 - 2 * aj, aj]]
        w += [wj] * 3
    x += [[a, b], [b, a], [a, 1 - a - b], [b, 1 - a - b], [1 - a - b, a], [1 - a - b, b]]
    w += [wS111] * 6
    return np.array(x), np.array(w), "symmetric rule order 8"


def _sym_tet8():
    aS31 = [.0396754230703899012650713295393895, .3144878006980963137841605626971483, .1019866930627033000000000000000000, .1842036969491915122759464173489092]
    wS31 = [.0063971477799023213214514203351730, .0401904480209661724881611584798178, .0243079755047703211748691087719226, .0548588924136974404669241239903914]
    aS22, wS22 = .0634362877545398924051412387018983, .0357196122340991824649509689966176
    aS211 = [[.0216901620677280048026624826249302, .7199319220394659358894349533527348], [.2044800806367957142413355748727453, .5805771901288092241753981713906204]]
    wS211 = [.0071831906978525394094511052198038, .0163721819453191175409381397561191]
    x, w = [], []
    for a, wj in zip(aS31, wS31):
        c = 1 - 3 * a
        x += [[a, a, a], [a, a, c], [a, c, a], [c, a, a]]
        w += [wj] * 4
    a, h = aS22, 0.5 - aS22
    x += [[a, a, h], [a, h, a], [h, a, a], [h, a, h], [h, h, a], [a, h, h]]
    w += [wS22] * 6
    for (a, b), wj in zip(aS211, wS211):
        c = 1 - 2 * a - b
        x += [[a, a, b], [a, b, a], [b, a, a], [a, a, c], [a, c, a], [c, a, a],
              [a, b, c], [a, c, b], [c, a, b], [b, a, c], [b, c, a], [c, b, a]]
        w += [wj] * 12
    return np.array(x), np.array(w), "symmetric rule order 8"


def _triangle(order):
    if order <= 1:
        return np.array([[1.0 / 3, 1.0 / 3]]), np.array([1.0]), "midpoint rule"
    if order == 2:
        return (np.array([[0.5, 0.5], [0.0, 0.5], [0.5, 0.0]]), np.array([1.0 / 3] * 3), "face midpoints rule")
    if order == 8:
        return _sym_tri8()
    if order <= 11:
        return _stroud(order)
    raise NotImplementedError("triangle quadrature order > 11")


def _tetrahedron(order):
    if order <= 1:
        return np.array([[0.25, 0.25, 0.25]]), np.array([1.0]), "midpoint rule"
    if order == 2:
        a, b = 0.1381966011250105, 0.5854101966249685
        return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.array([0.25] * 4), "order 2 rule"
    if order <= 3:
        x = np.array([[1 / 4, 1 / 4, 1 / 4], [1 / 2, 1 / 6, 1 / 6], [1 / 6, 1 / 6, 1 / 6], [1 / 6, 1 / 6, 1 / 2], [1 / 6, 1 / 2, 1 / 6]])
        return x, np.array([-4 / 5, 9 / 20, 9 / 20, 9 / 20, 9 / 20]), "order 3 rule"
    if order <= 4:
        c, d, e, f, g = 0.2500000000000000, 0.7857142857142857, 0.0714285714285714, 0.1005964238332008, 0.3994035761667992
        x = np.array([[c, c, c], [d, e, e], [e, e, e], [e, e, d], [e, d, e], [f, g, g], [g, f, g], [g, g, f], [g, f, f], [f, g, f], [f, f, g]])
        w0, w1, w2 = -0.0789333333333333, 0.0457333333333333, 0.1493333333333333
        return x, np.array([w0] + [w1] * 4 + [w2] * 6), "order 4 rule"
    return _sym_tet8()
