"""
Stand-ins for the few OpenMM host objects the collective-variable constructors consume.

cvpack's constructors read a ``NonbondedForce`` only through ``getNumParticles``,
``usesPeriodicBoundaryConditions``, ``getNumExceptions``, ``getExceptionParameters``,
``getParticleParameters``, ``getCutoffDistance``, ``getUseSwitchingFunction`` and
``getSwitchingDistance`` (``/root/reference/cvpack/number_of_contacts.py:138-151``,
``attraction_strength.py:188-224``), and a ``System`` through ``getNumParticles``,
``getParticleMass``, ``getForces`` and ``addForce`` (``collective_variable.py:116-120,168``,
``utils.py:52-56,172-183``).  These classes provide exactly that surface (duck-type compatible with
the OpenMM originals, which may be passed instead) plus the XML form used when a force is embedded
in a serialized CV (``serialization/serialization.py:115-121``).
"""

from __future__ import annotations

import typing as t
import xml.etree.ElementTree as ET
from collections import namedtuple

import numpy as np

from . import mmunit

Vec3 = namedtuple("Vec3", ["x", "y", "z"])


def _md(value: t.Any) -> t.Any:
    if mmunit.is_quantity(value):
        return value.value_in_unit_system(mmunit.md_unit_system)
    if hasattr(value, "value_in_unit_system") and hasattr(value, "unit"):
        return _foreign_quantity(value)
    return value


def _foreign_quantity(value: t.Any) -> t.Any:
    """Convert a quantity of another unit package (e.g. openmm.unit) through its unit name."""
    unit = mmunit.parse_unit(str(value.unit).replace("elementary charge", "elementary_charge"))
    raw = value.value_in_unit(value.unit)
    return mmunit.Quantity(raw, unit).value_in_unit_system(mmunit.md_unit_system)


class Force:
    """Base of everything that can be added to a :class:`System`."""

    def __init__(self) -> None:
        self._name = type(self).__name__
        self._force_group = 0

    def getName(self) -> str:
        """Name of this force."""
        return self._name

    def setName(self, name: str) -> None:
        """Set the name of this force."""
        self._name = name

    def getForceGroup(self) -> int:
        """Force group (0..31) of this force."""
        return self._force_group

    def setForceGroup(self, group: int) -> None:
        """Set the force group (0..31) of this force."""
        if not 0 <= group < 32:
            raise ValueError("Force group must be between 0 and 31")
        self._force_group = int(group)

    def usesPeriodicBoundaryConditions(self) -> bool:
        """Whether this force uses periodic boundary conditions."""
        return False


class NonbondedForce(Force):
    """Per-particle charge/sigma/epsilon, exceptions, cutoff and switching settings."""

    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    def __init__(self) -> None:
        super().__init__()
        self._particles: t.List[t.Tuple[float, float, float]] = []
        self._exceptions: t.List[t.Tuple[int, int, float, float, float]] = []
        self._method = self.NoCutoff
        self._cutoff = 1.0
        self._use_switch = False
        self._switch_distance = -1.0

    # particles
    def addParticle(self, charge: t.Any, sigma: t.Any, epsilon: t.Any) -> int:
        """Append a particle; returns its index."""
        self._particles.append((float(_md(charge)), float(_md(sigma)), float(_md(epsilon))))
        return len(self._particles) - 1

    def getNumParticles(self) -> int:
        """Number of particles."""
        return len(self._particles)

    def getParticleParameters(self, index: int) -> t.List[mmunit.Quantity]:
        """Charge, sigma and epsilon of a particle."""
        charge, sigma, epsilon = self._particles[index]
        return [
            charge * mmunit.elementary_charge,
            sigma * mmunit.nanometer,
            epsilon * mmunit.kilojoule_per_mole,
        ]

    # exceptions
    def addException(
        self, particle1: int, particle2: int, chargeProd: t.Any, sigma: t.Any, epsilon: t.Any
    ) -> int:
        """Append an exception (a pair whose interaction is replaced)."""
        self._exceptions.append(
            (
                int(particle1),
                int(particle2),
                float(_md(chargeProd)),
                float(_md(sigma)),
                float(_md(epsilon)),
            )
        )
        return len(self._exceptions) - 1

    def getNumExceptions(self) -> int:
        """Number of exceptions."""
        return len(self._exceptions)

    def getExceptionParameters(self, index: int) -> t.List[t.Any]:
        """Particle indices, charge product, sigma and epsilon of an exception."""
        i, j, charge_prod, sigma, epsilon = self._exceptions[index]
        return [
            i,
            j,
            charge_prod * mmunit.elementary_charge**2,
            sigma * mmunit.nanometer,
            epsilon * mmunit.kilojoule_per_mole,
        ]

    # cutoff, method, switching
    def getNonbondedMethod(self) -> int:
        """Nonbonded method."""
        return self._method

    def setNonbondedMethod(self, method: int) -> None:
        """Set the nonbonded method."""
        self._method = int(method)

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._method in (self.CutoffPeriodic, self.Ewald, self.PME, self.LJPME)

    def getCutoffDistance(self) -> mmunit.Quantity:
        """Cutoff distance."""
        return self._cutoff * mmunit.nanometer

    def setCutoffDistance(self, distance: t.Any) -> None:
        """Set the cutoff distance."""
        self._cutoff = float(_md(distance))

    def getUseSwitchingFunction(self) -> bool:
        """Whether a switching function is applied."""
        return self._use_switch

    def setUseSwitchingFunction(self, use: bool) -> None:
        """Enable/disable the switching function."""
        self._use_switch = bool(use)

    def getSwitchingDistance(self) -> mmunit.Quantity:
        """Distance at which switching starts."""
        return self._switch_distance * mmunit.nanometer

    def setSwitchingDistance(self, distance: t.Any) -> None:
        """Set the distance at which switching starts."""
        self._switch_distance = float(_md(distance))


class XmlSerializer:
    """Reads and writes the OpenMM XML form of a :class:`NonbondedForce`."""

    @staticmethod
    def serialize(force: t.Any) -> str:
        """XML text of a nonbonded force (attribute names as written by OpenMM)."""
        if hasattr(force, "force"):
            force = force.force
        root = ET.Element(
            "Force",
            {
                "cutoff": repr(_md(force.getCutoffDistance())),
                "forceGroup": str(force.getForceGroup()),
                "method": str(int(force.getNonbondedMethod())),
                "name": force.getName(),
                "switchingDistance": repr(_md(force.getSwitchingDistance())),
                "type": "NonbondedForce",
                "useSwitchingFunction": str(int(force.getUseSwitchingFunction())),
                "version": "4",
            },
        )
        particles = ET.SubElement(root, "Particles")
        for index in range(force.getNumParticles()):
            charge, sigma, epsilon = map(_md, force.getParticleParameters(index))
            ET.SubElement(
                particles, "Particle", {"eps": repr(epsilon), "q": repr(charge), "sig": repr(sigma)}
            )
        exceptions = ET.SubElement(root, "Exceptions")
        for index in range(force.getNumExceptions()):
            i, j, charge_prod, sigma, epsilon = force.getExceptionParameters(index)
            ET.SubElement(
                exceptions,
                "Exception",
                {
                    "eps": repr(_md(epsilon)),
                    "p1": str(i),
                    "p2": str(j),
                    "q": repr(_md(charge_prod)),
                    "sig": repr(_md(sigma)),
                },
            )
        ET.indent(root, space="\t")
        return '<?xml version="1.0" ?>\n' + ET.tostring(root, encoding="unicode") + "\n"

    @staticmethod
    def deserialize(xml_code: str) -> NonbondedForce:
        """Rebuild a nonbonded force from XML text (ours or OpenMM's)."""
        root = ET.fromstring(xml_code)
        if root.get("type") != "NonbondedForce":
            raise ValueError(
                f"Only NonbondedForce can be embedded in a collective variable; got {root.get('type')}"
            )
        force = NonbondedForce()
        force.setName(root.get("name", "NonbondedForce"))
        force.setForceGroup(int(root.get("forceGroup", "0")))
        force.setNonbondedMethod(int(root.get("method", "0")))
        force.setCutoffDistance(float(root.get("cutoff", "1")))
        force.setUseSwitchingFunction(bool(int(root.get("useSwitchingFunction", "0"))))
        force.setSwitchingDistance(float(root.get("switchingDistance", "-1")))
        for particle in root.iterfind("Particles/Particle"):
            force.addParticle(
                float(particle.get("q")), float(particle.get("sig")), float(particle.get("eps"))
            )
        for exc in root.iterfind("Exceptions/Exception"):
            force.addException(
                int(exc.get("p1")),
                int(exc.get("p2")),
                float(exc.get("q")),
                float(exc.get("sig")),
                float(exc.get("eps")),
            )
        return force


class System:
    """Particle masses, default box and the list of forces (collective variables)."""

    def __init__(self) -> None:
        self._masses: t.List[float] = []
        self._forces: t.List[t.Any] = []
        self._box: t.Optional[np.ndarray] = None

    def addParticle(self, mass: t.Any) -> int:
        """Append a particle of the given mass (Da)."""
        self._masses.append(float(_md(mass)))
        return len(self._masses) - 1

    def getNumParticles(self) -> int:
        """Number of particles."""
        return len(self._masses)

    def getParticleMass(self, index: int) -> mmunit.Quantity:
        """Mass of a particle."""
        return self._masses[index] * mmunit.dalton

    def setParticleMass(self, index: int, mass: t.Any) -> None:
        """Set the mass of a particle."""
        self._masses[index] = float(_md(mass))

    def getMasses(self) -> np.ndarray:
        """All masses as a float64 array (Da)."""
        return np.asarray(self._masses, dtype=np.float64)

    def addForce(self, force: t.Any) -> int:
        """Append a force; returns its index."""
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self) -> int:
        """Number of forces."""
        return len(self._forces)

    def getForce(self, index: int) -> t.Any:
        """The force at a given index."""
        return self._forces[index]

    def getForces(self) -> t.List[t.Any]:
        """All forces."""
        return list(self._forces)

    def removeForce(self, index: int) -> None:
        """Remove the force at a given index."""
        del self._forces[index]

    def setDefaultPeriodicBoxVectors(self, a: t.Any, b: t.Any, c: t.Any) -> None:
        """Set the default periodic box (three vectors, nm)."""
        self._box = np.asarray([_md(a), _md(b), _md(c)], dtype=np.float64).reshape(3, 3)

    def getDefaultPeriodicBoxVectors(self) -> t.Optional[np.ndarray]:
        """Default periodic box as a 3x3 array of row vectors (nm), or None."""
        return self._box

    def usesPeriodicBoundaryConditions(self) -> bool:
        """Whether any force uses periodic boundary conditions."""
        return any(force.usesPeriodicBoundaryConditions() for force in self._forces)
