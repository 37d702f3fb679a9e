"""
A stand-in for the ``openmm`` package, sufficient to *run the reference's own constructors and
evaluation path* (``/root/reference/cvpack``) without OpenMM.

TEST INFRASTRUCTURE ONLY (part of ``oracle/``; never imported by ``cvpack_b200``).

What this buys: everything cvpack itself decides -- Lepton expression strings, per-term parameters,
atom / group / bond index lists, exclusions, interaction groups, cutoffs, switching distances,
normalisation constants, nesting of ``CustomCVForce`` objects, ``getValue`` / ``getEffectiveMass``
plumbing -- is produced by the **reference code**, not restated.  What is still restated here is
OpenMM's Reference-platform semantics of each ``Force`` class (OpenMM is an absent, unpinned
third-party dependency), written independently of ``oracle/cv_oracle.py``:

* ``CustomNonbondedForce``: interaction groups (per group: ``i in set1, j in set2``, skip ``i == j``,
  skip excluded pairs, skip the mirrored duplicate when both atoms are in both sets), skip
  ``r >= cutoff``, minimum image when ``CutoffPeriodic``, switching ``1 - 10t^3 + 15t^4 - 6t^5``
  applied as ``E*sw`` and ``E'*sw + E*sw'``;
* ``CustomBondForce`` / ``CustomAngleForce`` / ``CustomTorsionForce`` / ``CustomCompoundBondForce``:
  ``r``, ``theta`` (acos), IUPAC dihedral in (-pi, pi], each bond vector minimum-imaged separately;
* ``CustomCentroidBondForce``: weighted centres of raw coordinates (default weights = masses,
  normalised per group), PBC only on centre-centre vectors;
* ``RMSDForce``: reference centred on the selected particles, optimal rotation (Kabsch/SVD here;
  OpenMM uses Horn's quaternion -- same minimum), ``msd < 1e-20 -> 0`` with zero forces;
* ``CustomCVForce``: expression of the inner energies, gradient by the chain rule with Lepton's
  derivative rules (``oracle/refshim/lepton.py``);
* ``openmmcppforces.CompositeRMSDForce``: restated from the docstring formula
  (``composite_rmsd.py:36-59``); the plugin's source is absent -> **parity unpinned** for it.

All arithmetic is FP64.  Periodic displacement follows the Reference platform's triclinic
reduction: subtract ``round(d_z/c_z) c``, then ``round(d_y/b_y) b``, then ``round(d_x/a_x) a``.
"""

from __future__ import annotations

import copy
import itertools
import typing as t

import numpy as np

from cvpack_b200 import mmunit

from . import lepton

_THIS = itertools.count(1)
_OWNERS: t.Dict[int, t.Any] = {}

Vec3 = t.NamedTuple("Vec3", [("x", float), ("y", float), ("z", float)])


def _md(value: t.Any) -> t.Any:
    if mmunit.is_quantity(value):
        return value.value_in_unit_system(mmunit.md_unit_system)
    return value


def _positions_array(value: t.Any) -> np.ndarray:
    return np.array(_md(value), dtype=np.float64).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------
# geometry (FP64) with gradients
# ------------------------------------------------------------------------------------------------
def delta(a: np.ndarray, b: np.ndarray, box: t.Optional[np.ndarray]) -> np.ndarray:
    """``b - a`` (rows), minimum-imaged in the Reference platform's order when ``box`` is given."""
    d = np.asarray(b, dtype=np.float64) - np.asarray(a, dtype=np.float64)
    if box is not None:
        d = d - np.floor(d[..., 2:3] / box[2, 2] + 0.5) * box[2]
        d = d - np.floor(d[..., 1:2] / box[1, 1] + 0.5) * box[1]
        d = d - np.floor(d[..., 0:1] / box[0, 0] + 0.5) * box[0]
    return d


def distance_terms(x: np.ndarray, idx: np.ndarray, box: t.Optional[np.ndarray]):
    """r and d r/d x for rows of ``idx`` [(i, j)]: returns (r[T], grad[T, 2, 3])."""
    d = delta(x[idx[:, 0]], x[idx[:, 1]], box)
    r = np.sqrt(np.sum(d * d, axis=1))
    with np.errstate(all="ignore"):
        u = d / r[:, None]
    return r, np.stack([-u, u], axis=1)


def angle_terms(x: np.ndarray, idx: np.ndarray, box: t.Optional[np.ndarray]):
    """theta(p1, p2, p3) at p2 and its gradient [T, 3, 3]."""
    a = delta(x[idx[:, 1]], x[idx[:, 0]], box)
    b = delta(x[idx[:, 1]], x[idx[:, 2]], box)
    na, nb = np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1)
    cos = np.clip(np.sum(a * b, axis=1) / (na * nb), -1.0, 1.0)
    theta = np.arccos(cos)
    with np.errstate(all="ignore"):
        minus_inv_sin = -1.0 / np.sqrt(1.0 - cos * cos)
        ua, ub = a / na[:, None], b / nb[:, None]
        g1 = minus_inv_sin[:, None] * (ub - cos[:, None] * ua) / na[:, None]
        g3 = minus_inv_sin[:, None] * (ua - cos[:, None] * ub) / nb[:, None]
    return theta, np.stack([g1, -(g1 + g3), g3], axis=1)


def dihedral_terms(x: np.ndarray, idx: np.ndarray, box: t.Optional[np.ndarray]):
    """IUPAC dihedral of (p1, p2, p3, p4) in (-pi, pi] and its gradient [T, 4, 3]."""
    b1 = delta(x[idx[:, 0]], x[idx[:, 1]], box)
    b2 = delta(x[idx[:, 1]], x[idx[:, 2]], box)
    b3 = delta(x[idx[:, 2]], x[idx[:, 3]], box)
    n1, n2 = np.cross(b1, b2), np.cross(b2, b3)
    lb2 = np.linalg.norm(b2, axis=1)
    y = np.sum(np.cross(n1, n2) * b2, axis=1) / lb2
    xx = np.sum(n1 * n2, axis=1)
    phi = np.arctan2(y, xx)
    # Blondel & Karplus (1996) with F = -b1, G = -b2, H = b3, A = F x G = n1, B = H x G = n2
    n1sq, n2sq = np.sum(n1 * n1, axis=1), np.sum(n2 * n2, axis=1)
    g1 = -(lb2 / n1sq)[:, None] * n1
    g4 = (lb2 / n2sq)[:, None] * n2
    s12 = (np.sum(b1 * b2, axis=1) / (lb2 * lb2))[:, None]
    s32 = (np.sum(b3 * b2, axis=1) / (lb2 * lb2))[:, None]
    g2 = -g1 - s12 * g1 + s32 * g4
    g3 = -g4 + s12 * g1 - s32 * g4
    return phi, np.stack([g1, g2, g3, g4], axis=1)


# ------------------------------------------------------------------------------------------------
# forces
# ------------------------------------------------------------------------------------------------
class Force:
    """Base class: name, force group, identity (``this``) and the evaluation protocol."""

    def __init__(self) -> None:
        self._omm_init()

    def _omm_init(self) -> None:
        self.this = next(_THIS)
        _OWNERS[self.this] = self
        self._omm_name = type(self).__name__
        self._omm_group = 0
        self._omm_pbc = False

    def _owner(self) -> "Force":
        """The object that owns ``this`` (SWIG methods of ``openmm.Force`` dispatch on the C++ object,
        so a wrapper that only copies ``this`` -- serialization.py:104-112 -- still reaches it)."""
        return _OWNERS.get(self.__dict__.get("this"), self)

    def getName(self) -> str:
        return self._owner()._omm_name

    def setName(self, name: str) -> None:
        self._owner()._omm_name = name

    def getForceGroup(self) -> int:
        return self._owner()._omm_group

    def setForceGroup(self, group: int) -> None:
        if not 0 <= group < 32:
            raise ValueError("Force group must be between 0 and 31")
        self._owner()._omm_group = int(group)

    def usesPeriodicBoundaryConditions(self) -> bool:
        owner = self._owner()
        if owner is not self:
            return owner.usesPeriodicBoundaryConditions()
        return self._omm_pbc

    def setUsesPeriodicBoundaryConditions(self, periodic: bool) -> None:
        self._omm_pbc = bool(periodic)

    # -- evaluation protocol --------------------------------------------------------------------
    def _evaluate(self, x: np.ndarray, box: t.Optional[np.ndarray], masses: np.ndarray,
                  params: t.Mapping[str, float]) -> t.Tuple[float, np.ndarray]:
        """(energy, dE/dx[N, 3]) at one configuration."""
        raise NotImplementedError(type(self).__name__)

    def _plain_copy(self) -> "Force":
        """What ``XmlSerializer.deserialize(XmlSerializer.serialize(force))`` returns: an object
        of the OpenMM class (not of the cvpack subclass) with the same settings."""
        base = next(c for c in type(self).__mro__ if c.__module__ == __name__)
        new = base.__new__(base)
        for key, value in self.__dict__.items():
            if key.startswith("_omm_"):
                setattr(new, key, _copy_state(value))
        new.this = next(_THIS)
        _OWNERS[new.this] = new
        return new


def _copy_state(value: t.Any) -> t.Any:
    if isinstance(value, Force):
        return value._plain_copy()  # pylint: disable=protected-access
    if isinstance(value, list):
        return [_copy_state(v) for v in value]
    if isinstance(value, tuple):
        return tuple(_copy_state(v) for v in value)
    if isinstance(value, dict):
        return {k: _copy_state(v) for k, v in value.items()}
    return copy.deepcopy(value)


class _CustomForce(Force):
    """Shared by the Custom* classes: energy function, global and per-term parameters."""

    def __init__(self, energy: str) -> None:
        Force.__init__(self)
        self._omm_energy = energy
        self._omm_globals: t.List[t.Tuple[str, float]] = []
        self._omm_per_term: t.List[str] = []
        self._omm_param_derivs: t.List[str] = []

    def getEnergyFunction(self) -> str:
        return self._omm_energy

    def setEnergyFunction(self, energy: str) -> None:
        self._omm_energy = energy

    def addGlobalParameter(self, name: str, default: t.Any) -> int:
        self._omm_globals.append((name, float(_md(default))))
        return len(self._omm_globals) - 1

    def getNumGlobalParameters(self) -> int:
        return len(self._omm_globals)

    def getGlobalParameterName(self, index: int) -> str:
        return self._omm_globals[index][0]

    def getGlobalParameterDefaultValue(self, index: int) -> float:
        return self._omm_globals[index][1]

    def addEnergyParameterDerivative(self, name: str) -> None:
        self._omm_param_derivs.append(name)

    def _add_per_term(self, name: str) -> int:
        self._omm_per_term.append(name)
        return len(self._omm_per_term) - 1

    def _global_values(self, params: t.Mapping[str, float]) -> t.Dict[str, float]:
        return {name: params.get(name, default) for name, default in self._omm_globals}

    def _term_parameters(self, rows: t.Sequence[t.Sequence[t.Any]]) -> t.Dict[str, np.ndarray]:
        if not self._omm_per_term:
            return {}
        table = np.array([[float(_md(v)) for v in row] for row in rows], dtype=np.float64)
        table = table.reshape(len(rows), len(self._omm_per_term))
        return {name: table[:, k] for k, name in enumerate(self._omm_per_term)}


class CustomBondForce(_CustomForce):
    """E = sum over bonds of f(r; per-bond parameters)."""

    def __init__(self, energy: str) -> None:
        super().__init__(energy)
        self._omm_terms: t.List[t.Tuple[t.Tuple[int, ...], t.Tuple[float, ...]]] = []

    addPerBondParameter = _CustomForce._add_per_term
    _arity = 2
    _variable = "r"
    _geometry = staticmethod(distance_terms)

    def addBond(self, *args: t.Any) -> int:
        *atoms, parameters = args if len(args) == self._arity + 1 else (*args, [])
        self._omm_terms.append((tuple(int(a) for a in atoms), tuple(parameters or ())))
        return len(self._omm_terms) - 1

    def getNumBonds(self) -> int:
        return len(self._omm_terms)

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        if not self._omm_terms:
            return 0.0, grad
        idx = np.array([atoms for atoms, _ in self._omm_terms], dtype=np.int64)
        value, dvalue = self._geometry(x, idx, box if self._omm_pbc else None)
        variables: t.Dict[str, t.Any] = dict(self._global_values(params))
        variables.update(self._term_parameters([p for _, p in self._omm_terms]))
        variables[self._variable] = lepton.Dual.variable(self._variable, value)
        result = lepton.Expression(self._omm_energy).evaluate(variables)
        energy = np.broadcast_to(result.v, value.shape)
        slope = np.broadcast_to(result.d.get(self._variable, 0.0), value.shape)
        for k in range(idx.shape[1]):
            np.add.at(grad, idx[:, k], slope[:, None] * dvalue[:, k])
        return float(np.sum(energy)), grad


class CustomAngleForce(CustomBondForce):
    """E = sum over angles of f(theta)."""

    addPerAngleParameter = _CustomForce._add_per_term
    addAngle = CustomBondForce.addBond
    _arity = 3
    _variable = "theta"
    _geometry = staticmethod(angle_terms)

    def getNumAngles(self) -> int:
        return len(self._omm_terms)


class CustomTorsionForce(CustomBondForce):
    """E = sum over torsions of f(theta), theta in (-pi, pi]."""

    addPerTorsionParameter = _CustomForce._add_per_term
    addTorsion = CustomBondForce.addBond
    _arity = 4
    _variable = "theta"
    _geometry = staticmethod(dihedral_terms)

    def getNumTorsions(self) -> int:
        return len(self._omm_terms)


_GEOMETRY_CALLS = {"distance": (2, distance_terms), "angle": (3, angle_terms), "dihedral": (4, dihedral_terms)}


def _evaluate_site_expression(expression: str, prefix: str, sites: np.ndarray, terms: np.ndarray,
                              variables: t.Dict[str, t.Any], box: t.Optional[np.ndarray]):
    """
    Evaluate an expression over ``terms`` [T, k] rows of site indices, where the expression refers to
    sites as ``<prefix>1 .. <prefix>k`` inside distance / angle / dihedral calls and as ``x1, y1, z1``.
    Returns (energy[T], dE/dsite[T, k, 3]).
    """
    expr = lepton.Expression(expression)
    count, arity = terms.shape
    dsite = np.zeros((count, arity, 3))
    leaves: t.Dict[str, t.Tuple[np.ndarray, np.ndarray]] = {}

    def leaf_for(node: lepton.Node) -> t.Optional[lepton.Dual]:
        if node.value not in _GEOMETRY_CALLS:
            return None
        nargs, function = _GEOMETRY_CALLS[node.value]
        cols = []
        for arg in node.args:
            if arg.kind != "var" or not arg.value.startswith(prefix):
                raise ValueError(f"bad argument to {node.value}: {arg}")
            cols.append(int(arg.value[len(prefix):]) - 1)
        if len(cols) != nargs:
            raise ValueError(f"{node.value} takes {nargs} arguments")
        key = f"{node.value}({','.join(map(str, cols))})"
        if key not in leaves:
            value, dvalue = function(sites, terms[:, cols], box)
            leaves[key] = (np.array(cols), dvalue)
            leaf_values[key] = lepton.Dual.variable(key, value)
        return leaf_values[key]

    leaf_values: t.Dict[str, lepton.Dual] = {}
    coords = {}
    for k in range(arity):
        for axis, letter in enumerate("xyz"):
            name = f"{letter}{k + 1}"
            if name in expr.free_variables() and name not in variables:
                coords[name] = (k, axis)
                variables[name] = lepton.Dual.variable(name, sites[terms[:, k], axis])
    result = expr.evaluate(variables, leaf_for)
    energy = np.broadcast_to(np.asarray(result.v, dtype=np.float64), (count,)).copy()
    for key, (cols, dvalue) in leaves.items():
        slope = np.broadcast_to(result.d.get(key, 0.0), (count,))
        for pos, col in enumerate(cols):
            dsite[:, col] += slope[:, None] * dvalue[:, pos]
    for name, (k, axis) in coords.items():
        dsite[:, k, axis] += np.broadcast_to(result.d.get(name, 0.0), (count,))
    return energy, dsite


class CustomCompoundBondForce(_CustomForce):
    """E = sum over bonds of f(distance(p1,p2), angle(...), dihedral(...), x1, ...)."""

    def __init__(self, numParticles: int, energy: str) -> None:
        super().__init__(energy)
        self._omm_arity = int(numParticles)
        self._omm_terms: t.List[t.Tuple[t.Tuple[int, ...], t.Tuple[float, ...]]] = []

    addPerBondParameter = _CustomForce._add_per_term

    def addBond(self, particles: t.Sequence[int], parameters: t.Sequence[t.Any] = ()) -> int:
        if len(particles) != self._omm_arity:
            raise ValueError("wrong number of particles for a bond")
        self._omm_terms.append((tuple(int(p) for p in particles), tuple(parameters or ())))
        return len(self._omm_terms) - 1

    def getNumBonds(self) -> int:
        return len(self._omm_terms)

    def getNumParticlesPerBond(self) -> int:
        return self._omm_arity

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        if not self._omm_terms:
            return 0.0, grad
        terms = np.array([atoms for atoms, _ in self._omm_terms], dtype=np.int64)
        variables: t.Dict[str, t.Any] = dict(self._global_values(params))
        variables.update(self._term_parameters([p for _, p in self._omm_terms]))
        energy, dsite = _evaluate_site_expression(
            self._omm_energy, "p", x, terms, variables, box if self._omm_pbc else None
        )
        for k in range(terms.shape[1]):
            np.add.at(grad, terms[:, k], dsite[:, k])
        return float(np.sum(energy)), grad


class CustomCentroidBondForce(_CustomForce):
    """Like CustomCompoundBondForce, but the sites are weighted centres of atom groups."""

    def __init__(self, numGroups: int, energy: str) -> None:
        super().__init__(energy)
        self._omm_arity = int(numGroups)
        self._omm_groups: t.List[t.Tuple[t.Tuple[int, ...], t.Optional[t.Tuple[float, ...]]]] = []
        self._omm_terms: t.List[t.Tuple[t.Tuple[int, ...], t.Tuple[float, ...]]] = []

    addPerBondParameter = _CustomForce._add_per_term

    def addGroup(self, particles: t.Sequence[int], weights: t.Optional[t.Sequence[float]] = None) -> int:
        weights = None if weights is None or len(weights) == 0 else tuple(float(w) for w in weights)
        if weights is not None and len(weights) != len(particles):
            raise ValueError("wrong number of weights for a group")
        self._omm_groups.append((tuple(int(p) for p in particles), weights))
        return len(self._omm_groups) - 1

    def addBond(self, groups: t.Sequence[int], parameters: t.Sequence[t.Any] = ()) -> int:
        if len(groups) != self._omm_arity:
            raise ValueError("wrong number of groups for a bond")
        self._omm_terms.append((tuple(int(g) for g in groups), tuple(parameters or ())))
        return len(self._omm_terms) - 1

    def getNumGroups(self) -> int:
        return len(self._omm_groups)

    def getNumBonds(self) -> int:
        return len(self._omm_terms)

    def getNumGroupsPerBond(self) -> int:
        return self._omm_arity

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        if not self._omm_terms:
            return 0.0, grad
        for groups, _ in self._omm_terms:
            if max(groups) >= len(self._omm_groups):
                raise IndexError("CustomCentroidBondForce: bond refers to a group that does not exist")
        normalised = []
        centres = np.zeros((len(self._omm_groups), 3))
        for g, (atoms, weights) in enumerate(self._omm_groups):
            atoms_a = np.array(atoms, dtype=np.int64)
            w = masses[atoms_a] if weights is None else np.array(weights, dtype=np.float64)
            w = w / np.sum(w)
            normalised.append((atoms_a, w))
            centres[g] = w @ x[atoms_a]
        terms = np.array([groups for groups, _ in self._omm_terms], dtype=np.int64)
        variables: t.Dict[str, t.Any] = dict(self._global_values(params))
        variables.update(self._term_parameters([p for _, p in self._omm_terms]))
        energy, dsite = _evaluate_site_expression(
            self._omm_energy, "g", centres, terms, variables, box if self._omm_pbc else None
        )
        dcentre = np.zeros_like(centres)
        for k in range(terms.shape[1]):
            np.add.at(dcentre, terms[:, k], dsite[:, k])
        for g, (atoms_a, w) in enumerate(normalised):
            np.add.at(grad, atoms_a, w[:, None] * dcentre[g])
        return float(np.sum(energy)), grad


class CustomNonbondedForce(_CustomForce):
    """Pair sum over interaction groups with cutoff, switching and exclusions."""

    NoCutoff, CutoffNonPeriodic, CutoffPeriodic = 0, 1, 2

    def __init__(self, energy: str) -> None:
        super().__init__(energy)
        self._omm_method = self.NoCutoff
        self._omm_cutoff = 1.0
        self._omm_use_switch = False
        self._omm_switch = -1.0
        self._omm_lrc = False
        self._omm_particles: t.List[t.Tuple[t.Any, ...]] = []
        self._omm_exclusions: t.List[t.Tuple[int, int]] = []
        self._omm_igroups: t.List[t.Tuple[t.Tuple[int, ...], t.Tuple[int, ...]]] = []

    addPerParticleParameter = _CustomForce._add_per_term

    def setNonbondedMethod(self, method: int) -> None:
        self._omm_method = int(method)

    def getNonbondedMethod(self) -> int:
        return self._omm_method

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._omm_method == self.CutoffPeriodic

    def setCutoffDistance(self, distance: t.Any) -> None:
        self._omm_cutoff = float(_md(distance))

    def getCutoffDistance(self) -> mmunit.Quantity:
        return self._omm_cutoff * mmunit.nanometers

    def setUseSwitchingFunction(self, use: bool) -> None:
        self._omm_use_switch = bool(use)

    def getUseSwitchingFunction(self) -> bool:
        return self._omm_use_switch

    def setSwitchingDistance(self, distance: t.Any) -> None:
        self._omm_switch = float(_md(distance))

    def getSwitchingDistance(self) -> mmunit.Quantity:
        return self._omm_switch * mmunit.nanometers

    def setUseLongRangeCorrection(self, use: bool) -> None:
        self._omm_lrc = bool(use)

    def addParticle(self, parameters: t.Sequence[t.Any] = ()) -> int:
        self._omm_particles.append(tuple(parameters))
        return len(self._omm_particles) - 1

    def getNumParticles(self) -> int:
        return len(self._omm_particles)

    def addExclusion(self, particle1: int, particle2: int) -> int:
        self._omm_exclusions.append((int(particle1), int(particle2)))
        return len(self._omm_exclusions) - 1

    def getNumExclusions(self) -> int:
        return len(self._omm_exclusions)

    def addInteractionGroup(self, set1: t.Iterable[int], set2: t.Iterable[int]) -> int:
        self._omm_igroups.append((tuple(int(i) for i in set1), tuple(int(j) for j in set2)))
        return len(self._omm_igroups) - 1

    def getNumInteractionGroups(self) -> int:
        return len(self._omm_igroups)

    def _pairs(self, num_atoms: int) -> np.ndarray:
        excluded = {(min(i, j), max(i, j)) for i, j in self._omm_exclusions}
        pairs: t.List[t.Tuple[int, int]] = []
        if not self._omm_igroups:
            for i in range(num_atoms):
                for j in range(i + 1, num_atoms):
                    if (i, j) not in excluded:
                        pairs.append((i, j))
        for set1, set2 in self._omm_igroups:
            s1, s2 = set(set1), set(set2)
            for i in sorted(s1):
                for j in sorted(s2):
                    if i == j or (min(i, j), max(i, j)) in excluded:
                        continue
                    if i > j and j in s1 and i in s2:
                        continue  # both atoms are in both sets: the mirrored pair is the one kept
                    pairs.append((i, j))
        return np.array(pairs, dtype=np.int64).reshape(-1, 2)

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        if len(self._omm_particles) != x.shape[0]:
            raise ValueError("CustomNonbondedForce must have one particle per system particle")
        pairs = self._pairs(x.shape[0])
        if pairs.shape[0] == 0:
            return 0.0, grad
        periodic = self._omm_method == self.CutoffPeriodic
        if periodic and box is not None and self._omm_cutoff > 0.5 * min(box[0, 0], box[1, 1], box[2, 2]):
            raise ValueError("The cutoff distance cannot be greater than half the periodic box size.")
        r, dr = distance_terms(x, pairs, box if periodic else None)
        if self._omm_method != self.NoCutoff:
            keep = r < self._omm_cutoff
            pairs, r, dr = pairs[keep], r[keep], dr[keep]
        variables: t.Dict[str, t.Any] = dict(self._global_values(params))
        per_particle = self._term_parameters(self._omm_particles)
        for name, column in per_particle.items():
            variables[f"{name}1"] = column[pairs[:, 0]]
            variables[f"{name}2"] = column[pairs[:, 1]]
        variables["r"] = lepton.Dual.variable("r", r)
        result = lepton.Expression(self._omm_energy).evaluate(variables)
        energy = np.broadcast_to(np.asarray(result.v, dtype=np.float64), r.shape).copy()
        slope = np.broadcast_to(np.asarray(result.d.get("r", 0.0), dtype=np.float64), r.shape).copy()
        if self._omm_use_switch and self._omm_method != self.NoCutoff:
            width = self._omm_cutoff - self._omm_switch
            tt = np.where(r > self._omm_switch, (r - self._omm_switch) / width, 0.0)
            sw = 1.0 + tt**3 * (-10.0 + tt * (15.0 - tt * 6.0))
            dsw = tt**2 * (-30.0 + tt * (60.0 - tt * 30.0)) / width
            slope = slope * sw + energy * dsw
            energy = energy * sw
        np.add.at(grad, pairs[:, 0], slope[:, None] * dr[:, 0])
        np.add.at(grad, pairs[:, 1], slope[:, None] * dr[:, 1])
        return float(np.sum(energy)), grad


class NonbondedForce(Force):
    """Container of charges / LJ parameters / exceptions (only read by cvpack, never evaluated)."""

    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    def __init__(self) -> None:
        super().__init__()
        self._omm_particles: t.List[t.Tuple[float, float, float]] = []
        self._omm_exceptions: t.List[t.Tuple[int, int, float, float, float]] = []
        self._omm_method = self.NoCutoff
        self._omm_cutoff = 1.0
        self._omm_use_switch = False
        self._omm_switch = -1.0

    def addParticle(self, charge, sigma, epsilon) -> int:
        self._omm_particles.append((float(_md(charge)), float(_md(sigma)), float(_md(epsilon))))
        return len(self._omm_particles) - 1

    def getNumParticles(self) -> int:
        return len(self._omm_particles)

    def getParticleParameters(self, index: int):
        q, s, e = self._omm_particles[index]
        return [q * mmunit.elementary_charge, s * mmunit.nanometer, e * mmunit.kilojoule_per_mole]

    def addException(self, p1, p2, chargeProd, sigma, epsilon) -> int:
        self._omm_exceptions.append(
            (int(p1), int(p2), float(_md(chargeProd)), float(_md(sigma)), float(_md(epsilon)))
        )
        return len(self._omm_exceptions) - 1

    def getNumExceptions(self) -> int:
        return len(self._omm_exceptions)

    def getExceptionParameters(self, index: int):
        i, j, q, s, e = self._omm_exceptions[index]
        return [i, j, q * mmunit.elementary_charge**2, s * mmunit.nanometer, e * mmunit.kilojoule_per_mole]

    def setNonbondedMethod(self, method: int) -> None:
        self._omm_method = int(method)

    def getNonbondedMethod(self) -> int:
        return self._omm_method

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._omm_method in (self.CutoffPeriodic, self.Ewald, self.PME, self.LJPME)

    def setCutoffDistance(self, distance) -> None:
        self._omm_cutoff = float(_md(distance))

    def getCutoffDistance(self):
        return self._omm_cutoff * mmunit.nanometer

    def setUseSwitchingFunction(self, use: bool) -> None:
        self._omm_use_switch = bool(use)

    def getUseSwitchingFunction(self) -> bool:
        return self._omm_use_switch

    def setSwitchingDistance(self, distance) -> None:
        self._omm_switch = float(_md(distance))

    def getSwitchingDistance(self):
        return self._omm_switch * mmunit.nanometer

    def _evaluate(self, x, box, masses, params):
        raise NotImplementedError("the stand-in NonbondedForce is a parameter container only")


class HarmonicBondForce(Force):
    """E = sum k/2 (r - r0)^2 (used by ``RMSD.getNullBondForce`` with k = 0)."""

    def __init__(self) -> None:
        super().__init__()
        self._omm_bonds: t.List[t.Tuple[int, int, float, float]] = []

    def addBond(self, p1: int, p2: int, length: t.Any, k: t.Any) -> int:
        self._omm_bonds.append((int(p1), int(p2), float(_md(length)), float(_md(k))))
        return len(self._omm_bonds) - 1

    def getNumBonds(self) -> int:
        return len(self._omm_bonds)

    def getBondParameters(self, index: int):
        return list(self._omm_bonds[index])

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        if not self._omm_bonds:
            return 0.0, grad
        idx = np.array([(i, j) for i, j, _, _ in self._omm_bonds], dtype=np.int64)
        r0 = np.array([b[2] for b in self._omm_bonds])
        k = np.array([b[3] for b in self._omm_bonds])
        r, dr = distance_terms(x, idx, box if self._omm_pbc else None)
        slope = k * (r - r0)
        np.add.at(grad, idx[:, 0], slope[:, None] * dr[:, 0])
        np.add.at(grad, idx[:, 1], slope[:, None] * dr[:, 1])
        return float(np.sum(0.5 * k * (r - r0) ** 2)), grad


def _kabsch(xh: np.ndarray, yh: np.ndarray) -> np.ndarray:
    """Proper rotation R minimising sum |xh_i - R yh_i|^2 (rows are atoms)."""
    u, _, vt = np.linalg.svd(xh.T @ yh)
    d = np.sign(np.linalg.det(u @ vt))
    return u @ np.diag([1.0, 1.0, d]) @ vt


class RMSDForce(Force):
    """RMSD to reference positions after optimal superposition of the selected particles."""

    def __init__(self, referencePositions: t.Any, particles: t.Sequence[int] = ()) -> None:
        super().__init__()
        self._omm_reference = _positions_array(referencePositions)
        self._omm_particles = [int(p) for p in particles]

    def getReferencePositions(self):
        return [Vec3(*row) for row in self._omm_reference]

    def setReferencePositions(self, positions) -> None:
        self._omm_reference = _positions_array(positions)

    def getParticles(self):
        return list(self._omm_particles)

    def setParticles(self, particles) -> None:
        self._omm_particles = [int(p) for p in particles]

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        if self._omm_reference.shape[0] != x.shape[0]:
            raise ValueError("RMSDForce: wrong number of reference positions")
        atoms = np.array(self._omm_particles or range(x.shape[0]), dtype=np.int64)
        n = len(atoms)
        yh = self._omm_reference[atoms] - self._omm_reference[atoms].mean(axis=0)
        xh = x[atoms] - x[atoms].mean(axis=0)
        rot = _kabsch(xh, yh)
        # msd through the eigenvalue form used by OpenMM: (sum|x|^2 + sum|y|^2 - 2 lambda_max)/n,
        # with lambda_max = sum_i x_i . R y_i at the optimum
        lam = np.sum(xh * (yh @ rot.T))
        msd = (np.sum(xh * xh) + np.sum(yh * yh) - 2.0 * lam) / n
        if msd < 1e-20:
            return 0.0, grad
        rmsd = float(np.sqrt(msd))
        np.add.at(grad, atoms, (xh - yh @ rot.T) / (n * rmsd))
        return rmsd, grad


class CompositeRMSDForce(Force):
    """``openmmcppforces.CompositeRMSDForce``, restated from ``composite_rmsd.py:36-59``:
    every group is centred on its own centroid in both structures; one rotation for all."""

    def __init__(self, referencePositions: t.Any) -> None:
        super().__init__()
        self._omm_reference = _positions_array(referencePositions)
        self._omm_groups: t.List[t.List[int]] = []

    def addGroup(self, particles: t.Sequence[int]) -> int:
        self._omm_groups.append([int(p) for p in particles])
        return len(self._omm_groups) - 1

    def getNumGroups(self) -> int:
        return len(self._omm_groups)

    def _evaluate(self, x, box, masses, params):
        grad = np.zeros_like(x)
        groups = self._omm_groups or [list(range(x.shape[0]))]
        atoms = np.array(sum(groups, []), dtype=np.int64)
        xh = np.concatenate([x[g] - x[g].mean(axis=0) for g in groups])
        yh = np.concatenate([self._omm_reference[g] - self._omm_reference[g].mean(axis=0) for g in groups])
        n = len(atoms)
        rot = _kabsch(xh, yh)
        lam = np.sum(xh * (yh @ rot.T))
        msd = (np.sum(xh * xh) + np.sum(yh * yh) - 2.0 * lam) / n
        if msd < 1e-20:
            return 0.0, grad
        rmsd = float(np.sqrt(msd))
        np.add.at(grad, atoms, (xh - yh @ rot.T) / (n * rmsd))
        return rmsd, grad


class CustomCVForce(_CustomForce):
    """E = f(cv_1, ..., cv_k; globals) where each cv is the energy of an inner Force."""

    def __init__(self, energy: str) -> None:
        super().__init__(energy)
        self._omm_variables: t.List[t.Tuple[str, Force]] = []

    def addCollectiveVariable(self, name: str, variable: Force) -> int:
        self._omm_variables.append((name, variable))
        return len(self._omm_variables) - 1

    def getNumCollectiveVariables(self) -> int:
        return len(self._omm_variables)

    def getCollectiveVariableName(self, index: int) -> str:
        return self._omm_variables[index][0]

    def getCollectiveVariable(self, index: int) -> Force:
        return self._omm_variables[index][1]

    def usesPeriodicBoundaryConditions(self) -> bool:
        return any(f.usesPeriodicBoundaryConditions() for _, f in self._omm_variables)

    def _inner(self, x, box, masses, params):
        return [(name, *force._evaluate(x, box, masses, params))  # pylint: disable=protected-access
                for name, force in self._omm_variables]

    def _evaluate(self, x, box, masses, params):
        inner = self._inner(x, box, masses, params)
        variables: t.Dict[str, t.Any] = dict(self._global_values(params))
        for name, value, _ in inner:
            variables[name] = lepton.Dual.variable(name, value)
        result = lepton.Expression(self._omm_energy).evaluate(variables)
        grad = np.zeros_like(x)
        for name, _, inner_grad in inner:
            grad += float(result.d.get(name, 0.0)) * inner_grad
        return float(result.v), grad

    def _parameter_derivatives(self, x, box, masses, params) -> t.Dict[str, float]:
        inner = self._inner(x, box, masses, params)
        variables: t.Dict[str, t.Any] = {name: value for name, value, _ in inner}
        for name, value in self._global_values(params).items():
            variables[name] = lepton.Dual.variable(name, value) if name in self._omm_param_derivs else value
        result = lepton.Expression(self._omm_energy).evaluate(variables)
        return {name: float(result.d.get(name, 0.0)) for name in self._omm_param_derivs}

    def getCollectiveVariableValues(self, context: "Context") -> t.Tuple[float, ...]:
        x, box, masses, params = context._configuration()  # pylint: disable=protected-access
        return tuple(value for _, value, _ in self._inner(x, box, masses, params))

    def getInnerContext(self, context: "Context") -> "Context":
        system = System()
        for mass in context.getSystem()._masses:  # pylint: disable=protected-access
            system.addParticle(mass)
        for group, (_, force) in enumerate(self._omm_variables):
            force.setForceGroup(group)
            system.addForce(force)
        inner = Context(system, VerletIntegrator(1.0))
        inner._positions = context._positions  # pylint: disable=protected-access
        inner._box = context._box  # pylint: disable=protected-access
        inner._parameters = dict(context._parameters)  # pylint: disable=protected-access
        return inner


# ------------------------------------------------------------------------------------------------
# system, integrators, context, state
# ------------------------------------------------------------------------------------------------
class System:
    """Masses, default box and forces."""

    def __init__(self) -> None:
        self._masses: t.List[float] = []
        self._forces: t.List[Force] = []
        self._box: t.Optional[np.ndarray] = None

    def addParticle(self, mass: t.Any) -> int:
        self._masses.append(float(_md(mass)))
        return len(self._masses) - 1

    def getNumParticles(self) -> int:
        return len(self._masses)

    def getParticleMass(self, index: int):
        return self._masses[index] * mmunit.dalton

    def addForce(self, force: Force) -> int:
        self._forces.append(force)
        return len(self._forces) - 1

    def getForces(self) -> t.List[Force]:
        return list(self._forces)

    def getNumForces(self) -> int:
        return len(self._forces)

    def getForce(self, index: int) -> Force:
        return self._forces[index]

    def setDefaultPeriodicBoxVectors(self, a, b, c) -> None:
        self._box = np.array([_md(a), _md(b), _md(c)], dtype=np.float64).reshape(3, 3)

    def getDefaultPeriodicBoxVectors(self):
        box = self._box if self._box is not None else np.diag([2.0, 2.0, 2.0])
        return [mmunit.Quantity(np.array(row), mmunit.nanometers) for row in box]

    def usesPeriodicBoundaryConditions(self) -> bool:
        return any(f.usesPeriodicBoundaryConditions() for f in self._forces)


class VerletIntegrator:
    """Placeholder: nothing is integrated."""

    def __init__(self, stepSize: t.Any = 0.001) -> None:
        self._step = _md(stepSize)


CustomIntegrator = VerletIntegrator


class Platform:
    """Placeholder platform registry."""

    def __init__(self, name: str) -> None:
        self._name = name

    def getName(self) -> str:
        return self._name

    @staticmethod
    def getPlatformByName(name: str) -> "Platform":
        return Platform(name)


class State:
    """Result of ``Context.getState``."""

    def __init__(self, energy, forces, positions, box, derivatives) -> None:
        self._energy, self._forces, self._positions = energy, forces, positions
        self._box, self._derivatives = box, derivatives

    def getPotentialEnergy(self):
        return self._energy * mmunit.kilojoules_per_mole

    def getForces(self, asNumpy: bool = False):
        unit = mmunit.kilojoules_per_mole / mmunit.nanometers
        value = self._forces if asNumpy else [Vec3(*row) for row in self._forces]
        return mmunit.Quantity(value, unit)

    def getPositions(self, asNumpy: bool = False):
        value = self._positions if asNumpy else [Vec3(*row) for row in self._positions]
        return mmunit.Quantity(value, mmunit.nanometers)

    def getPeriodicBoxVectors(self, asNumpy: bool = False):
        if asNumpy:
            return mmunit.Quantity(self._box, mmunit.nanometers)
        return [mmunit.Quantity(np.array(row), mmunit.nanometers) for row in self._box]

    def getEnergyParameterDerivatives(self):
        return dict(self._derivatives)


class Context:
    """One configuration of a System."""

    def __init__(self, system: System, integrator: t.Any, platform: t.Any = None) -> None:
        self._system = system
        self._integrator = integrator
        self._positions: t.Optional[np.ndarray] = None
        box = system._box  # pylint: disable=protected-access
        self._box = np.diag([2.0, 2.0, 2.0]) if box is None else np.array(box)
        self._parameters: t.Dict[str, float] = {}

    def getSystem(self) -> System:
        return self._system

    def setPositions(self, positions: t.Any) -> None:
        self._positions = _positions_array(positions)
        if self._positions.shape[0] != self._system.getNumParticles():
            raise ValueError("Called setPositions() on a Context with the wrong number of positions")

    def setPeriodicBoxVectors(self, a, b, c) -> None:
        self._box = np.array([_md(a), _md(b), _md(c)], dtype=np.float64).reshape(3, 3)

    def setParameter(self, name: str, value: t.Any) -> None:
        self._parameters[name] = float(_md(value))

    def getParameter(self, name: str) -> float:
        if name in self._parameters:
            return self._parameters[name]
        for force in self._system.getForces():
            for pname, default in getattr(force, "_omm_globals", []):
                if pname == name:
                    return default
        raise KeyError(name)

    def reinitialize(self, preserveState: bool = False) -> None:
        if not preserveState:
            self._positions = None

    def _configuration(self):
        if self._positions is None:
            raise RuntimeError("Particle positions have not been set")
        masses = np.array(self._system._masses, dtype=np.float64)  # pylint: disable=protected-access
        return self._positions, self._box, masses, self._parameters

    def getState(self, getPositions=False, getEnergy=False, getForces=False,
                 getParameterDerivatives=False, groups=-1, **_ignored) -> State:
        x, box, masses, params = self._configuration()
        if isinstance(groups, (set, list, tuple)):
            groups = sum(1 << g for g in groups)
        energy, grad, derivatives = 0.0, np.zeros_like(x), {}
        if getEnergy or getForces or getParameterDerivatives:
            for force in self._system.getForces():
                if not (groups >> force.getForceGroup()) & 1:
                    continue
                e, g = force._evaluate(x, box, masses, params)  # pylint: disable=protected-access
                energy += e
                grad += g
                if getParameterDerivatives and isinstance(force, CustomCVForce):
                    for name, value in force._parameter_derivatives(x, box, masses, params).items():  # pylint: disable=protected-access
                        derivatives[name] = derivatives.get(name, 0.0) + value
        return State(energy, -grad, x.copy(), box.copy(), derivatives)


class XmlSerializer:
    """In-process stand-in: ``serialize`` parks a plain copy and returns a token."""

    _store: t.Dict[str, Force] = {}

    @classmethod
    def serialize(cls, force: Force) -> str:
        inner = getattr(force, "force", force)  # SerializableForce wrapper
        token = f'<?xml version="1.0" ?>\n<Force refshim="{len(cls._store)}" type="{type(inner).__name__}"/>\n'
        cls._store[token] = inner._plain_copy()  # pylint: disable=protected-access
        return token

    @classmethod
    def deserialize(cls, text: str) -> Force:
        return cls._store[text]._plain_copy()  # pylint: disable=protected-access


# the SWIG layer the reference reaches into (collective_variable.py:14,68,120,126; utils.py:171)
class _Swig:
    @staticmethod
    def Force_setName(force: Force, name: str) -> None:
        Force.setName(force, name)

    @staticmethod
    def Force_getName(force: Force) -> str:
        return Force.getName(force)

    @staticmethod
    def Force_setForceGroup(force: Force, group: int) -> None:
        Force.setForceGroup(force, group)

    @staticmethod
    def System_getParticleMass(system: System, index: int) -> float:
        return system._masses[int(index)]  # pylint: disable=protected-access
