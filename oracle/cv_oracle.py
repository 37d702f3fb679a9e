"""
CPU oracle (NumPy, float64) for the collective variables of cvpack -- TEST INFRASTRUCTURE ONLY.

Nothing under ``cvpack_b200/`` imports this module; only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the checker or
as the timed CPU baseline.

What it restates.  cvpack itself holds no arithmetic: each CV configures an OpenMM Force and OpenMM
evaluates it (``/root/reference/cvpack/utils.py:140``).  OpenMM (unpinned in the reference's
``pyproject.toml:23-27``; CI uses 7.7/8.0/8.1) and openmm-cpp-forces (unpinned, ``mdtools`` conda
channel) are absent from ``/root/reference`` and from this image, so every function below restates
the published semantics of the OpenMM force the corresponding cvpack constructor builds, and cites
the cvpack lines that fix the expression, the groups and the parameters.  Parity is pinned by the
reference's own offline known-answer values and independent recomputations (see
``tests/test_oracle.py``): ``distance.py:37-50``, ``angle.py:46-60``, ``torsion.py:53-67``,
``composite_rmsd.py:116-121``, the scipy ``Rotation.align_vectors`` check of
``tests/test_cvpack.py:275-305``, the explicit double loop of ``tests/test_cvpack.py:688-732``, the
numpy recomputations of ``tests/test_cvpack.py:172-229,232-272,735-770,807-860,890-928`` and central
finite differences for every gradient.  CompositeRMSD with more than one body has no numeric test in
the reference beyond the 0.0-invariance doctest: for that CV parity is **unpinned** (restated from
the docstring formula ``composite_rmsd.py:36-59``).

Every function takes one frame ``x`` of shape ``[N, 3]`` (nm, float64) and returns
``(value, gradient[N, 3])`` where ``gradient`` is +dCV/dx (OpenMM's forces are its negative).
``box`` is ``None`` or a 3x3 array of row vectors in OpenMM's reduced form.
"""

from __future__ import annotations

import typing as t

import numpy as np

ONE_4PI_EPS0 = 138.93545764438198  # attraction_strength.py:20


# ---- geometry ------------------------------------------------------------------------------------
def min_image(d: np.ndarray, box: t.Optional[np.ndarray]) -> np.ndarray:
    """
    Minimum-image displacement(s).  Rectangular: ``d - L floor(d/L + 0.5)`` per axis; triclinic:
    reduce along c, then b, then a (OpenMM ReferenceForce::getDeltaRPeriodic[Triclinic]).
    """
    if box is None:
        return d
    d = np.array(d, dtype=np.float64, copy=True)
    box = np.asarray(box, dtype=np.float64)
    for axis in (2, 1, 0):
        scale = np.floor(d[..., axis] / box[axis, axis] + 0.5)
        d -= scale[..., None] * box[axis]
    return d


def step_function(kind: str, x: np.ndarray, n: int = 6, m: int = 12) -> t.Tuple[np.ndarray, np.ndarray]:
    """
    S(x) and dS/dx for the whitelisted step functions:
    ``rational`` 1/(1+x^n) (number_of_contacts.py:132), ``heaviside`` step(1-x) (Lepton step(0)=1,
    number_of_contacts.py:113), ``rational_pair`` (1+x^n)/(1+x^n+x^2n) (helix_rmsd_content.py:135),
    ``plumed`` (1-x^n)/(1-x^m).
    """
    x = np.asarray(x, dtype=np.float64)
    if kind == "rational":
        s = 1.0 / (1.0 + x**n)
        return s, -n * x ** (n - 1) * s * s
    if kind == "heaviside":
        return np.where(x <= 1.0, 1.0, 0.0), np.zeros_like(x)
    if kind == "rational_pair":
        xn = x**n
        num = 1.0 + xn
        den = num + xn * xn
        dnum = n * x ** (n - 1)
        return num / den, (dnum * den - num * dnum * (1.0 + 2.0 * xn)) / den**2
    if kind == "plumed":
        num, den = 1.0 - x**n, 1.0 - x**m
        with np.errstate(divide="ignore", invalid="ignore"):
            s = num / den
            ds = (-n * x ** (n - 1) * den + num * m * x ** (m - 1)) / den**2
        near = np.abs(den) < 1e-6
        s = np.where(near, n / m * (1.0 + 0.5 * (n - m) * (x - 1.0)), s)
        ds = np.where(near, n / m * 0.5 * (n - m), ds)
        return s, ds
    raise ValueError(kind)


def switching(r: np.ndarray, rs: t.Optional[float], rc: float) -> t.Tuple[np.ndarray, np.ndarray]:
    """OpenMM's switching function 1-10t^3+15t^4-6t^5, t=(r-rs)/(rc-rs), and its r-derivative."""
    if rs is None:
        return np.ones_like(r), np.zeros_like(r)
    tt = np.clip((r - rs) / (rc - rs), 0.0, None)
    sw = 1.0 + tt**3 * (-10.0 + tt * (15.0 - 6.0 * tt))
    dsw = tt**2 * (-30.0 + tt * (60.0 - 30.0 * tt)) / (rc - rs)
    return sw, dsw


# ---- bonded terms (distance.py:57-59, angle.py:73-75, torsion.py:80-82) ---------------------------
def distance(x: np.ndarray, a1: int, a2: int, box: t.Optional[np.ndarray] = None):
    """|r2 - r1| (CustomBondForce "r")."""
    d = min_image(x[a2] - x[a1], box)
    r = np.linalg.norm(d)
    g = np.zeros_like(x)
    g[a2] += d / r
    g[a1] -= d / r
    return r, g


def angle(x: np.ndarray, a1: int, a2: int, a3: int, box: t.Optional[np.ndarray] = None):
    """acos of the angle at atom 2 (CustomAngleForce "theta")."""
    u = min_image(x[a1] - x[a2], box)
    v = min_image(x[a3] - x[a2], box)
    nu, nv = np.linalg.norm(u), np.linalg.norm(v)
    cosine = np.clip(np.dot(u, v) / (nu * nv), -1.0, 1.0)
    theta = np.arccos(cosine)
    sine = np.sqrt(max(1.0 - cosine * cosine, 1e-300))
    g = np.zeros_like(x)
    g1 = -(v / (nu * nv) - cosine * u / nu**2) / sine
    g3 = -(u / (nu * nv) - cosine * v / nv**2) / sine
    g[a1] += g1
    g[a3] += g3
    g[a2] -= g1 + g3
    return theta, g


def torsion(x: np.ndarray, a1: int, a2: int, a3: int, a4: int, box: t.Optional[np.ndarray] = None):
    """
    Dihedral with the sign convention of torsion.py:22-33 (pinned by tests/test_cvpack.py:161-167):
    atan2((r21 x r34).u23, r21.r34 - (r21.u23)(r34.u23)), r_ij = r_j - r_i.
    """
    f = min_image(x[a1] - x[a2], box)  # r21
    gv = min_image(x[a3] - x[a2], box)  # r23
    h = min_image(x[a4] - x[a3], box)  # r34
    ng = np.linalg.norm(gv)
    u = gv / ng
    theta = np.arctan2(np.dot(np.cross(f, h), u), np.dot(f, h) - np.dot(f, u) * np.dot(h, u))
    # gradient (Blondel & Karplus form) in terms of A = f x g, B = h x g
    a = np.cross(f, gv)
    b = np.cross(h, gv)
    a2n, b2n = np.dot(a, a), np.dot(b, b)
    fg, hg = np.dot(f, gv), np.dot(h, gv)
    d1 = ng / a2n * a
    d4 = -ng / b2n * b
    # projections of the outer-atom gradients along the central bond
    d2 = -d1 + (fg / (ng * ng)) * d1 + (hg / (ng * ng)) * d4
    d3 = -d4 - (fg / (ng * ng)) * d1 - (hg / (ng * ng)) * d4
    g = np.zeros_like(x)
    g[a1] += d1
    g[a2] += d2
    g[a3] += d3
    g[a4] += d4
    return theta, g


def torsion_similarity(
    x: np.ndarray,
    first: t.Sequence[t.Sequence[int]],
    second: t.Sequence[t.Sequence[int]],
    box: t.Optional[np.ndarray] = None,
):
    """
    sum_i 0.5*(1 + cos(min(delta_i, 2*pi - delta_i))), delta_i = dihedral(first_i) -
    dihedral(second_i)  (torsion_similarity.py:89-93).  Lepton differentiates min(a, b) by the
    branch taken: d/d delta is -sin(delta) on the first branch and +sin(2*pi - delta) = -sin(delta)
    on the second, so the gradient is -0.5 sin(delta_i) (grad phi1_i - grad phi2_i) either way.
    """
    total = 0.0
    grad = np.zeros_like(x)
    for one, two in zip(first, second):
        q1, g1 = torsion(x, *one, box=box)
        q2, g2 = torsion(x, *two, box=box)
        delta = q1 - q2
        total += 0.5 * (1.0 + np.cos(min(delta, 2 * np.pi - delta)))
        grad += -0.5 * np.sin(delta) * (g1 - g2)
    return total, grad


def path_in_cv_space(
    values: t.Sequence[float],
    gradients: t.Sequence[np.ndarray],
    milestones: np.ndarray,
    sigma: float,
    metric: str,
    periods: t.Optional[t.Sequence[float]] = None,
    scales: t.Optional[t.Sequence[float]] = None,
):
    """
    Progress / deviation along a path of milestones in the space of other CVs, from the inner CVs'
    values and gradients (path_in_cv_space.py:141-158 + base_path_cv.py:43-62):
    x_i = sum_j (delta_ij/scale_j)^2 with delta = m_ij - cv_j, or min(|delta|, P - |delta|) for a
    periodic variable (Lepton: min() differentiates the branch taken, the second on ties;
    abs'(0) = +1); w_i = exp(lambda (xmin - x_i)), lambda = 1/(2 sigma^2);
    s = sum_i i w_i/((n-1) sum w), z = xmin - ln(sum w)/lambda.
    """
    milestones = np.atleast_2d(np.asarray(milestones, dtype=float))
    n, nv = milestones.shape
    periods = [0.0] * nv if periods is None else list(periods)
    scales = [1.0] * nv if scales is None else list(scales)
    lam = 1.0 / (2.0 * sigma * sigma)
    x = np.zeros(n)
    dx = np.zeros((n, nv))  # d x_i / d cv_j
    for i in range(n):
        for j in range(nv):
            delta = milestones[i, j] - values[j]
            if periods[j] > 0:
                a = abs(delta)
                sign = 1.0 if delta >= 0 else -1.0
                if a < periods[j] - a:
                    m, dm = a, -sign
                else:
                    m, dm = periods[j] - a, sign
            else:
                m, dm = delta, -1.0
            x[i] += (m / scales[j]) ** 2
            dx[i, j] = 2.0 * m * dm / scales[j] ** 2
    xmin = x.min()
    w = np.exp(lam * (xmin - x))
    wsum = w.sum()
    index = np.arange(n)
    s = (index * w).sum() / ((n - 1) * wsum)
    if metric == "progress":
        value = s
        dvalue = -lam * w * (index / (n - 1) - s) / wsum
    else:
        value = xmin - np.log(wsum) / lam
        dvalue = w / wsum
    coef = dvalue @ dx
    grad = sum(c * g for c, g in zip(coef, gradients))
    return value, grad


def bonded_sum(
    x: np.ndarray,
    terms: t.Sequence[t.Sequence[int]],
    function: str = "identity",
    reference: t.Optional[t.Sequence[float]] = None,
    tolerance: float = 1.0,
    half_exponent: int = 3,
    coefficient: float = 1.0,
    box: t.Optional[np.ndarray] = None,
):
    """
    sum_t f(q_t) with q a distance / angle / torsion (2/3/4 atoms per term) and
    f = coefficient*q ("identity") or coefficient/(1+((q-ref_t)/tolerance)^(2m)) ("boxcar";
    helix_torsion_content.py:126-128, helix_angle_content.py:118-120, helix_hbond_content.py:104).
    The torsion boxcar does not wrap q - ref periodically (helix_torsion_content.py:128).
    """
    total = 0.0
    grad = np.zeros_like(x)
    fn = {2: distance, 3: angle, 4: torsion}
    for k, atoms in enumerate(terms):
        q, g = fn[len(atoms)](x, *atoms, box=box)
        if function == "identity":
            total += coefficient * q
            grad += coefficient * g
        else:
            xx = (q - reference[k]) / tolerance
            p = 2 * half_exponent
            s = 1.0 / (1.0 + xx**p)
            total += coefficient * s
            grad += coefficient * (-p * xx ** (p - 1) * s * s / tolerance) * g
    return total, grad


# ---- radius of gyration (radius_of_gyration.py:85-92, radius_of_gyration_sq.py:87-89) -------------
def radius_of_gyration(
    x: np.ndarray,
    group: t.Sequence[int],
    weights: t.Optional[np.ndarray] = None,
    box: t.Optional[np.ndarray] = None,
    squared: bool = False,
):
    """
    rg^2 = (1/n) sum |d_i|^2, d_i = minimum image of r_i - c, c = sum w_i r_i with w = 1/n or m/M
    (the sum stays unweighted; tests/test_cvpack.py:185-187,215-217).
    """
    group = np.asarray(group)
    n = len(group)
    w = np.full(n, 1.0 / n) if weights is None else np.asarray(weights, dtype=np.float64)
    c = w @ x[group]
    d = min_image(x[group] - c, box)
    rg2 = np.sum(d * d) / n
    g2 = (2.0 / n) * (d - w[:, None] * d.sum(axis=0))
    grad = np.zeros_like(x)
    if squared:
        grad[group] = g2
        return rg2, grad
    rg = np.sqrt(rg2)
    grad[group] = g2 / (2.0 * rg) if rg > 0 else 0.0
    return rg, grad


# ---- pair sums -------------------------------------------------------------------------------------
def interaction_pairs(
    group1: t.Sequence[int], group2: t.Sequence[int], exclusions: t.Iterable[t.Tuple[int, int]] = ()
) -> np.ndarray:
    """
    Unordered distinct pairs of a CustomNonbondedForce interaction group (number_of_contacts.py:
    47-56): i in group1, j in group2, i != j, each pair once, minus exclusions.  Integer-exact.
    """
    excluded = {(min(i, j), max(i, j)) for i, j in exclusions}
    pairs = set()
    set2 = set(int(j) for j in group2)
    for i in set(int(i) for i in group1):
        for j in set2:
            if i != j:
                pair = (min(i, j), max(i, j))
                if pair not in excluded:
                    pairs.add(pair)
    return np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)


def _pair_geometry(x, pairs, box, cutoff):
    d = min_image(x[pairs[:, 1]] - x[pairs[:, 0]], box)
    r = np.linalg.norm(d, axis=1)
    keep = r < cutoff if cutoff is not None else np.ones(len(r), dtype=bool)
    return pairs[keep], d[keep], r[keep]


def _scatter_pair_gradient(x, pairs, d, r, de):
    grad = np.zeros_like(x)
    # r == 0 (coincident atoms): the pair counts S(0) but its force direction d/r is undefined and
    # d == 0 -- it contributes nothing (OpenMM's Reference platform would divide 0 by 0 here)
    contrib = np.divide(de, r, out=np.zeros_like(r), where=r > 0)[:, None] * d
    np.add.at(grad, pairs[:, 1], contrib)
    np.add.at(grad, pairs[:, 0], -contrib)
    return grad


def number_of_contacts(
    x: np.ndarray,
    pairs: np.ndarray,
    r0: float = 0.3,
    cutoff_factor: float = 2.0,
    switch_factor: t.Optional[float] = 1.5,
    step: str = "rational",
    n: int = 6,
    m: int = 12,
    reference: float = 1.0,
    box: t.Optional[np.ndarray] = None,
):
    """(1/reference) sum_pairs S(r/r0) sw(r) over r < cutoff_factor*r0 (number_of_contacts.py:143-161)."""
    rc = cutoff_factor * r0
    rs = None if switch_factor is None else switch_factor * r0
    pairs, d, r = _pair_geometry(x, pairs, box, rc)
    s, ds = step_function(step, r / r0, n, m)
    sw, dsw = switching(r, rs, rc)
    value = np.sum(s * sw) / reference
    de = (ds / r0 * sw + s * dsw) / reference
    return value, _scatter_pair_gradient(x, pairs, d, r, de)


def attraction_strength(
    x: np.ndarray,
    pairs: np.ndarray,
    charge: np.ndarray,
    sigma: np.ndarray,
    epsilon: np.ndarray,
    cutoff: float,
    sign: t.Optional[np.ndarray] = None,
    switch_distance: t.Optional[float] = None,
    reference: float = 1.0,
    box: t.Optional[np.ndarray] = None,
):
    """
    -(1/ref) sum sgn_i sgn_j [4 eps (1/y^2-1/y) + ke min(0,q_i q_j)(1/x+(x^2-3)/2)] sw(r), with
    y=|(r/sigma)^6-2|+2, x=r/rc, eps=sqrt(eps_i eps_j), sigma=(sigma_i+sigma_j)/2
    (attraction_strength.py:189-199; formula-level check tests/test_cvpack.py:688-732).
    """
    pairs, d, r = _pair_geometry(x, pairs, box, cutoff)
    i, j = pairs[:, 0], pairs[:, 1]
    sgn = np.ones(len(r)) if sign is None else sign[i] * sign[j]
    sig = 0.5 * (sigma[i] + sigma[j])
    eps = np.sqrt(epsilon[i] * epsilon[j])
    qq = np.minimum(0.0, charge[i] * charge[j])
    u6 = (r / sig) ** 6
    a = u6 - 2.0
    y = np.abs(a) + 2.0
    lj = 4.0 * eps * (1.0 / y**2 - 1.0 / y)
    dy = np.where(a >= 0.0, 6.0, -6.0) * u6 / r
    dlj = 4.0 * eps * (-2.0 / y**3 + 1.0 / y**2) * dy
    xx = r / cutoff
    coul = ONE_4PI_EPS0 * qq * (1.0 / xx + 0.5 * (xx * xx - 3.0))
    dcoul = ONE_4PI_EPS0 * qq * (-1.0 / xx**2 + xx) / cutoff
    e = -sgn * (lj + coul)
    de = -sgn * (dlj + dcoul)
    sw, dsw = switching(r, switch_distance, cutoff)
    value = np.sum(e * sw) / reference
    return value, _scatter_pair_gradient(x, pairs, d, r, (de * sw + e * dsw) / reference)


def shortest_distance(
    x: np.ndarray,
    pairs: np.ndarray,
    beta: float = 100.0,
    cutoff: float = 1.0,
    box: t.Optional[np.ndarray] = None,
):
    """
    (rc e^-beta + sum r w)/(e^-beta + sum w), w = exp(-beta r/rc), r < rc (shortest_distance.py:
    122-127; the form pinned by tests/test_cvpack.py:919-927).
    """
    pairs, d, r = _pair_geometry(x, pairs, box, cutoff)
    w = np.exp(-beta * r / cutoff)
    guard = np.exp(-beta)
    denom = guard + w.sum()
    value = (cutoff * guard + np.sum(r * w)) / denom
    dw = -beta / cutoff * w
    de = ((w + r * dw) - value * dw) / denom
    return value, _scatter_pair_gradient(x, pairs, d, r, de)


def residue_coordination(
    x: np.ndarray,
    groups1: t.Sequence[t.Sequence[int]],
    groups2: t.Sequence[t.Sequence[int]],
    weights1: t.Sequence[np.ndarray],
    weights2: t.Sequence[np.ndarray],
    r0: float = 1.0,
    step: str = "rational",
    n: int = 6,
    m: int = 12,
    reference: float = 1.0,
    box: t.Optional[np.ndarray] = None,
):
    """
    (1/refval) sum_I sum_J S(|R_J - R_I|/r0), R = sum_a w_a r_a from raw coordinates, minimum image
    on the centroid-centroid vector only, no cutoff (residue_coordination.py:124-144).
    """
    c1 = np.array([np.asarray(w) @ x[np.asarray(g)] for g, w in zip(groups1, weights1)])
    c2 = np.array([np.asarray(w) @ x[np.asarray(g)] for g, w in zip(groups2, weights2)])
    d = min_image(c2[None, :, :] - c1[:, None, :], box)
    r = np.linalg.norm(d, axis=2)
    s, ds = step_function(step, r / r0, n, m)
    value = s.sum() / reference
    coef = np.divide(ds / r0 / reference, r, out=np.zeros_like(r), where=r > 0)[:, :, None] * d
    g1 = -coef.sum(axis=1)
    g2 = coef.sum(axis=0)
    grad = np.zeros_like(x)
    for grp, w, gc in zip(groups1, weights1, g1):
        grad[np.asarray(grp)] += np.asarray(w)[:, None] * gc
    for grp, w, gc in zip(groups2, weights2, g2):
        grad[np.asarray(grp)] += np.asarray(w)[:, None] * gc
    return value, grad


# ---- RMSD family -----------------------------------------------------------------------------------
def _horn_matrix(r: np.ndarray) -> np.ndarray:
    (sxx, sxy, sxz), (syx, syy, syz), (szx, szy, szz) = r
    return np.array(
        [
            [sxx + syy + szz, syz - szy, szx - sxz, sxy - syx],
            [syz - szy, sxx - syy - szz, sxy + syx, szx + sxz],
            [szx - sxz, sxy + syx, -sxx + syy - szz, syz + szy],
            [sxy - syx, szx + sxz, syz + szy, -sxx - syy + szz],
        ]
    )


def _rotation_from_quaternion(q: np.ndarray) -> np.ndarray:
    q0, q1, q2, q3 = q / np.linalg.norm(q)
    return np.array(
        [
            [q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
            [2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)],
            [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3],
        ]
    )


def composite_rmsd(
    x: np.ndarray, reference: np.ndarray, bodies: t.Sequence[t.Sequence[int]]
):
    """
    sqrt((1/n) min_U sum_bodies sum_i |U x^_i - y^_i|^2), every body centred on its own centroid in
    both structures, one shared rotation (composite_rmsd.py:36-59).  One body = openmm.RMSDForce
    (rmsd.py:24-38): R = sum x^ (x) y^, Horn's 4x4 matrix, msd = (sum|x^|^2 + sum|y^|^2 - 2 lambda)/n,
    value 0 with zero gradient when msd < 1e-20, d/dx_i = (x^_i - U^T y^_i)/(n rmsd).
    ``reference`` rows pair with the atoms of ``bodies`` in order (concatenated).
    """
    atoms = np.concatenate([np.asarray(b) for b in bodies])
    n = len(atoms)
    xh = np.empty((n, 3))
    yh = np.empty((n, 3))
    k = 0
    for body in bodies:
        m = len(body)
        xs = x[np.asarray(body)]
        ys = np.asarray(reference[k : k + m], dtype=np.float64)
        xh[k : k + m] = xs - xs.mean(axis=0)
        yh[k : k + m] = ys - ys.mean(axis=0)
        k += m
    w, v = np.linalg.eigh(_horn_matrix(xh.T @ yh))
    msd = (np.sum(xh * xh) + np.sum(yh * yh) - 2.0 * w[-1]) / n
    grad = np.zeros_like(x)
    if msd < 1e-20:
        return 0.0, grad
    value = np.sqrt(msd)
    u = _rotation_from_quaternion(v[:, -1])
    grad[atoms] = (xh - yh @ u) / (n * value)
    return value, grad


def rmsd(x: np.ndarray, reference: np.ndarray, group: t.Sequence[int]):
    """RMSD of one group after optimal superposition; ``reference`` rows pair with ``group``."""
    return composite_rmsd(x, reference, [group])


def rmsd_content(
    x: np.ndarray,
    blocks: t.Sequence[t.Sequence[int]],
    ideal: np.ndarray,
    r0: float = 0.08,
    step: str = "rational_pair",
    n: int = 4,
    m: int = 8,
    normalize: bool = False,
):
    """sum_blocks S(rmsd_b/r0) (/#blocks), every block against ``ideal`` (base_rmsd_content.py:44-79)."""
    total = 0.0
    grad = np.zeros_like(x)
    for block in blocks:
        d, g = rmsd(x, ideal, block)
        s, ds = step_function(step, np.array(d / r0), n, m)
        total += float(s)
        grad += float(ds) / r0 * g
    scale = 1.0 / len(blocks) if normalize else 1.0
    return total * scale, grad * scale


def path_in_rmsd_space(
    x: np.ndarray,
    milestones: t.Sequence[t.Tuple[t.Sequence[int], np.ndarray]],
    sigma: float,
    metric: str = "progress",
):
    """
    x_i = rmsd_i^2, lambda = 1/(2 sigma^2), w_i = exp(lambda (xmin - x_i));
    progress s = sum i w_i/((n-1) sum w), deviation z = xmin - ln(sum w)/lambda
    (base_path_cv.py:44-59, path_in_rmsd_space.py:137-141).  ``milestones`` = [(atoms, reference)].
    """
    n = len(milestones)
    lam = 1.0 / (2.0 * sigma**2)
    ds, gs = zip(*(rmsd(x, ref, atoms) for atoms, ref in milestones))
    xs = np.array(ds) ** 2
    xmin = xs.min()
    w = np.exp(lam * (xmin - xs))
    wsum = w.sum()
    s = np.dot(np.arange(n), w) / ((n - 1) * wsum)
    if metric == "progress":
        value = s
        dx = -lam * w * (np.arange(n) / (n - 1) - s) / wsum
    else:
        value = xmin - np.log(wsum) / lam
        dx = w / wsum
    grad = np.zeros_like(x)
    for k in range(n):
        grad += dx[k] * 2.0 * ds[k] * gs[k]
    return value, grad


# ---- effective mass (utils.py:172-184) ---------------------------------------------------------------
def effective_mass(gradient: np.ndarray, masses: np.ndarray) -> float:
    """1/sum_i |g_i|^2/m_i over atoms with non-zero gradient; inf if there is none."""
    squared = np.sum(np.square(gradient), axis=1)
    nonzeros = np.nonzero(squared)[0]
    if nonzeros.size == 0:
        return np.inf
    return 1.0 / np.sum(squared[nonzeros] / np.asarray(masses, dtype=np.float64)[nonzeros])


def finite_difference(fn: t.Callable[[np.ndarray], float], x: np.ndarray, atoms: t.Iterable[int], h: float = 1e-6) -> np.ndarray:
    """Central finite-difference gradient of ``fn`` at ``x`` for the given atoms."""
    grad = np.zeros_like(x)
    for a in atoms:
        for c in range(3):
            xp = x.copy()
            xm = x.copy()
            xp[a, c] += h
            xm[a, c] -= h
            grad[a, c] = (fn(xp) - fn(xm)) / (2 * h)
    return grad
