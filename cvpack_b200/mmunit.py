"""
A small, dependency-free physical-unit algebra for the MD unit system.

The reference leans on ``openmm.unit`` for this (``/root/reference/cvpack/units/units.py:15``);
OpenMM is not a dependency of this package, so this module supplies the subset of that behaviour the
collective-variable constructors need: named units (``nanometers``, ``daltons``, ``radians`` …),
products / quotients / powers of units, quantities that carry a unit, conversion between compatible
units, and conversion to the MD unit system (nm, ps, Da, kJ/mol, K, e, rad).

Formatting follows what the reference's doctests show (``collective_variable.py:299-304``,
``meta_collective_variable.py:101-111``): symbols of factors with positive exponents separated by
spaces, then ``/`` and the negative-exponent factors (parenthesised unless it is one bare symbol),
base units before scaled units -- e.g. ``nm**2 Da/(rad**2)`` and ``kJ/(mol rad**2)``.

Design: every named unit is an :class:`_Atom` with a dimension vector and a *scale* (the size of
one such unit measured in MD base units).  A :class:`Unit` is an immutable mapping atom->exponent.
"""

from __future__ import annotations

import ast
import math
import typing as t

import numpy as np

# order used when printing base units (mirrors the deterministic dimension order of openmm.unit)
_DIMENSIONS = (
    "mass",
    "length",
    "time",
    "temperature",
    "amount",
    "charge",
    "luminous intensity",
    "angle",
)
_NDIM = len(_DIMENSIONS)
AVOGADRO = 6.02214076e23


class _Atom:
    """A named unit that cannot be decomposed further when printing."""

    __slots__ = ("name", "symbol", "dims", "scale", "is_base", "order")

    def __init__(
        self, name: str, symbol: str, dims: t.Dict[str, float], scale: float, is_base: bool
    ) -> None:
        self.name = name
        self.symbol = symbol
        vec = [0.0] * _NDIM
        for dim, power in dims.items():
            vec[_DIMENSIONS.index(dim)] = power
        self.dims = tuple(vec)
        self.scale = float(scale)
        self.is_base = is_base
        first = next((i for i, p in enumerate(vec) if p), _NDIM)
        # base units print first, ordered by dimension; scaled units after, ordered by name
        self.order = (0, first, name) if is_base else (1, 0, name)

    def __repr__(self) -> str:
        return f"_Atom({self.name})"


def _is_number(value: t.Any) -> bool:
    return isinstance(value, (int, float, np.integer, np.floating)) and not isinstance(value, bool)


class Unit:
    """A product of named units raised to powers."""

    __slots__ = ("_factors", "_hash")
    __array_priority__ = 100

    def __init__(self, factors: t.Union["Unit", t.Mapping[_Atom, float]]) -> None:
        if isinstance(factors, Unit):
            factors = factors._factors  # pylint: disable=protected-access
        clean = {atom: power for atom, power in factors.items() if power != 0}
        self._factors = dict(sorted(clean.items(), key=lambda item: item[0].order))
        self._hash = None

    # ---- structure ----------------------------------------------------------------------------
    def iter_base_or_scaled_units(self) -> t.Iterator[t.Tuple[_Atom, float]]:
        """Iterate over (named unit, exponent) in printing order."""
        return iter(self._factors.items())

    def dimension_vector(self) -> t.Tuple[float, ...]:
        """Exponents of (mass, length, time, temperature, amount, charge, luminous, angle)."""
        vec = [0.0] * _NDIM
        for atom, power in self._factors.items():
            for k in range(_NDIM):
                vec[k] += atom.dims[k] * power
        return tuple(round(v, 12) for v in vec)

    def scale(self) -> float:
        """Size of one of this unit in MD base units."""
        result = 1.0
        for atom, power in self._factors.items():
            result *= atom.scale**power
        return result

    def is_compatible(self, other: "Unit") -> bool:
        """Whether the two units measure the same physical dimension."""
        return self.dimension_vector() == other.dimension_vector()

    def is_dimensionless(self) -> bool:
        """Whether all dimension exponents vanish (e.g. ``dimensionless`` or ``nm/angstrom``)."""
        return not any(self.dimension_vector())

    def conversion_factor_to(self, other: "Unit") -> float:
        """The number f such that ``x self == f*x other``."""
        if not self.is_compatible(other):
            raise TypeError(f'Unit "{self}" is not compatible with Unit "{other}".')
        if self._factors == other._factors:  # pylint: disable=protected-access
            return 1.0
        return self.scale() / other.scale()

    def in_unit_system(self, system: "UnitSystem") -> "Unit":
        """The unit of the given system with this unit's dimension."""
        return system.express(self)

    # ---- algebra ------------------------------------------------------------------------------
    def __mul__(self, other: t.Any) -> t.Any:
        if isinstance(other, Unit):
            factors = dict(self._factors)
            for atom, power in other._factors.items():  # pylint: disable=protected-access
                factors[atom] = factors.get(atom, 0) + power
            return type(self)._from_factors(factors)
        if isinstance(other, Quantity):
            return NotImplemented
        return Quantity(other, self)

    def __rmul__(self, other: t.Any) -> t.Any:
        if isinstance(other, (Unit, Quantity)):
            return NotImplemented
        return Quantity(other, self)

    def __truediv__(self, other: t.Any) -> t.Any:
        if isinstance(other, Unit):
            return self * other**-1
        if isinstance(other, Quantity):
            return NotImplemented
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other: t.Any) -> t.Any:
        if isinstance(other, (Unit, Quantity)):
            return NotImplemented
        return Quantity(other, self**-1)

    def __pow__(self, exponent: float) -> "Unit":
        if not _is_number(exponent):
            raise TypeError("Unit exponents must be numbers")
        return type(self)._from_factors(
            {atom: _tidy(power * exponent) for atom, power in self._factors.items()}
        )

    def sqrt(self) -> "Unit":
        """Square root of this unit."""
        return self**0.5

    @classmethod
    def _from_factors(cls, factors: t.Mapping[_Atom, float]) -> "Unit":
        unit = Unit.__new__(cls)
        Unit.__init__(unit, factors)
        return unit

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Unit):
            return False
        if self._factors == other._factors:
            return True
        return self.is_compatible(other) and math.isclose(
            self.scale(), other.scale(), rel_tol=1e-14
        )

    def __ne__(self, other: object) -> bool:
        return not self == other

    def __hash__(self) -> int:
        if self._hash is None:
            self._hash = hash(self.dimension_vector())
        return self._hash

    # ---- formatting ---------------------------------------------------------------------------
    def _format(self, use_symbol: bool) -> str:
        joiner = " " if use_symbol else "*"
        positive, negative = [], []
        for atom, power in self._factors.items():
            label = atom.symbol if use_symbol else atom.name
            magnitude = abs(power)
            text = label if magnitude == 1 else f"{label}**{_power_str(magnitude)}"
            (positive if power > 0 else negative).append(text)
        if not positive and not negative:
            return "dimensionless"
        head = joiner.join(positive) if positive else ""
        if not negative:
            return head
        tail = joiner.join(negative)
        if len(negative) > 1 or "**" in tail:
            tail = f"({tail})"
        return f"{head}/{tail}"

    def get_name(self) -> str:
        """Long form, e.g. ``nanometer**2*dalton/(radian**2)``."""
        return self._format(False)

    def get_symbol(self) -> str:
        """Short form, e.g. ``nm**2 Da/(rad**2)``."""
        return self._format(True)

    def __str__(self) -> str:
        return self.get_name()

    def __repr__(self) -> str:
        return f"Unit({self.get_name()!r})"


def _tidy(power: float) -> t.Union[int, float]:
    rounded = round(power)
    return int(rounded) if abs(power - rounded) < 1e-12 else power


def _power_str(power: float) -> str:
    return str(_tidy(power))


class Quantity:
    """A value (scalar, sequence, array or tensor) tagged with a :class:`Unit`."""

    __array_priority__ = 99

    def __init__(self, value: t.Any = None, unit: t.Optional[Unit] = None) -> None:
        if unit is None:
            if isinstance(value, Quantity):
                value, unit = value._value, value.unit
            elif (
                isinstance(value, (list, tuple))
                and len(value) > 0
                and all(isinstance(item, Quantity) for item in value)
            ):
                unit = value[0].unit
                value = type(value)(item.value_in_unit(unit) for item in value)
            else:
                unit = dimensionless
        elif isinstance(value, Quantity):
            value = value.value_in_unit(unit) if value.unit.is_compatible(unit) else value
            if isinstance(value, Quantity):
                unit = value.unit * unit
                value = value._value
        self._value = value
        self.unit = unit

    # ---- conversion ---------------------------------------------------------------------------
    def value_in_unit(self, unit: Unit) -> t.Any:
        """The bare value after conversion to ``unit``."""
        factor = self.unit.conversion_factor_to(unit)
        return _scaled(self._value, factor)

    def in_units_of(self, unit: Unit) -> "Quantity":
        """The same quantity expressed in ``unit``."""
        return type(self)(self.value_in_unit(unit), unit)

    def value_in_unit_system(self, system: "UnitSystem") -> t.Any:
        """The bare value after conversion to the given unit system."""
        return self.value_in_unit(system.express(self.unit))

    def in_unit_system(self, system: "UnitSystem") -> "Quantity":
        """The same quantity expressed in the given unit system."""
        return self.in_units_of(system.express(self.unit))

    # ---- algebra ------------------------------------------------------------------------------
    @staticmethod
    def _wrap(value: t.Any, unit: Unit) -> t.Any:
        if unit.is_dimensionless():
            factor = unit.scale()
            return value if factor == 1.0 else _scaled(value, factor)
        return Quantity(value, unit)

    def __mul__(self, other: t.Any) -> t.Any:
        if isinstance(other, Unit):
            return self._wrap(self._value, self.unit * other)
        if isinstance(other, Quantity):
            return self._wrap(_mul(self._value, other._value), self.unit * other.unit)
        return Quantity(_mul(self._value, other), self.unit)

    def __rmul__(self, other: t.Any) -> t.Any:
        if isinstance(other, Unit):
            return self._wrap(self._value, other * self.unit)
        return Quantity(_mul(other, self._value), self.unit)

    def __truediv__(self, other: t.Any) -> t.Any:
        if isinstance(other, Unit):
            return self._wrap(self._value, self.unit / other)
        if isinstance(other, Quantity):
            return self._wrap(_div(self._value, other._value), self.unit / other.unit)
        return Quantity(_div(self._value, other), self.unit)

    def __rtruediv__(self, other: t.Any) -> t.Any:
        if isinstance(other, Unit):
            return self._wrap(_div(1.0, self._value), other / self.unit)
        return Quantity(_div(other, self._value), self.unit**-1)

    def __pow__(self, exponent: float) -> "Quantity":
        return Quantity(self._value**exponent, self.unit**exponent)

    def sqrt(self) -> "Quantity":
        """Square root of the quantity."""
        return Quantity(_scaled(self._value, 1.0) ** 0.5, self.unit.sqrt())

    def __add__(self, other: "Quantity") -> "Quantity":
        if not isinstance(other, Quantity):
            raise TypeError("Cannot add a Quantity and a bare number")
        return Quantity(self._value + other.value_in_unit(self.unit), self.unit)

    def __sub__(self, other: "Quantity") -> "Quantity":
        if not isinstance(other, Quantity):
            raise TypeError("Cannot subtract a bare number from a Quantity")
        return Quantity(self._value - other.value_in_unit(self.unit), self.unit)

    def __neg__(self) -> "Quantity":
        return Quantity(-self._value, self.unit)

    def __abs__(self) -> "Quantity":
        return Quantity(abs(self._value), self.unit)

    def _other_value(self, other: t.Any) -> t.Any:
        if isinstance(other, Quantity):
            return other.value_in_unit(self.unit)
        if self.unit.is_dimensionless():
            return other
        raise TypeError("Cannot compare a Quantity with a bare number")

    def __eq__(self, other: object) -> t.Any:
        if not isinstance(other, Quantity) or not self.unit.is_compatible(other.unit):
            return False
        result = self._value == other.value_in_unit(self.unit)
        if isinstance(result, np.ndarray):
            return bool(result.all())
        return result

    def __ne__(self, other: object) -> t.Any:
        return not self == other

    __hash__ = None  # type: ignore[assignment]

    def __lt__(self, other: t.Any) -> t.Any:
        return self._value < self._other_value(other)

    def __le__(self, other: t.Any) -> t.Any:
        return self._value <= self._other_value(other)

    def __gt__(self, other: t.Any) -> t.Any:
        return self._value > self._other_value(other)

    def __ge__(self, other: t.Any) -> t.Any:
        return self._value >= self._other_value(other)

    def __float__(self) -> float:
        if not self.unit.is_dimensionless():
            raise TypeError("Only dimensionless quantities convert to float")
        return float(self._value) * self.unit.scale()

    # ---- container protocol -------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._value)

    def __getitem__(self, key: t.Any) -> "Quantity":
        return Quantity(self._value[key], self.unit)

    def __iter__(self) -> t.Iterator["Quantity"]:
        for item in self._value:
            yield Quantity(item, self.unit)

    # ---- formatting ---------------------------------------------------------------------------
    def __str__(self) -> str:
        return f"{self._value} {self.unit.get_symbol()}"

    def __repr__(self) -> str:
        return f"Quantity(value={self._value!r}, unit={self.unit})"


def _scaled(value: t.Any, factor: float) -> t.Any:
    if factor == 1.0:
        return value
    if isinstance(value, (list, tuple)):
        return type(value)(_scaled(item, factor) for item in value)
    return value * factor


def _mul(a: t.Any, b: t.Any) -> t.Any:
    if isinstance(a, (list, tuple)):
        a = np.asarray(a)
    if isinstance(b, (list, tuple)):
        b = np.asarray(b)
    return a * b


def _div(a: t.Any, b: t.Any) -> t.Any:
    if isinstance(a, (list, tuple)):
        a = np.asarray(a)
    if isinstance(b, (list, tuple)):
        b = np.asarray(b)
    return a / b


def is_quantity(obj: t.Any) -> bool:
    """Whether ``obj`` carries a unit."""
    return isinstance(obj, Quantity)


def is_unit(obj: t.Any) -> bool:
    """Whether ``obj`` is a unit."""
    return isinstance(obj, Unit)


class UnitSystem:
    """A choice of one unit per base dimension plus preferred derived units."""

    def __init__(self, base: t.Sequence[Unit], derived: t.Sequence[Unit] = ()) -> None:
        self._base: t.Dict[int, Unit] = {}
        for unit in base:
            vec = unit.dimension_vector()
            (index,) = [i for i, p in enumerate(vec) if p]
            self._base[index] = unit
        self._derived = {unit.dimension_vector(): unit for unit in derived}

    def express(self, unit: Unit) -> Unit:
        """The unit of this system with the same dimension as ``unit``."""
        vec = unit.dimension_vector()
        if vec in self._derived:
            return self._derived[vec]
        result = Unit({})
        for index, power in enumerate(vec):
            if power:
                result = result * self._base[index] ** _tidy(power)
        return result


def _base(name: str, symbol: str, dimension: str, scale: float = 1.0) -> Unit:
    return Unit({_Atom(name, symbol, {dimension: 1}, scale, True): 1})


def _scaled_unit(name: str, symbol: str, scale: float, master: Unit) -> Unit:
    dims = dict(zip(_DIMENSIONS, master.dimension_vector()))
    return Unit({_Atom(name, symbol, dims, scale * master.scale(), False): 1})


dimensionless = Unit({})

# --- length (MD base: nanometer) ---
nanometer = nanometers = _base("nanometer", "nm", "length")
angstrom = angstroms = _base("angstrom", "A", "length", 0.1)
meter = meters = _base("meter", "m", "length", 1e9)
centimeter = centimeters = _base("centimeter", "cm", "length", 1e7)
picometer = picometers = _base("picometer", "pm", "length", 1e-3)
bohr = bohrs = _base("bohr", "r_0", "length", 0.052917721067)
# --- time (MD base: picosecond) ---
picosecond = picoseconds = _base("picosecond", "ps", "time")
femtosecond = femtoseconds = _base("femtosecond", "fs", "time", 1e-3)
nanosecond = nanoseconds = _base("nanosecond", "ns", "time", 1e3)
second = seconds = _base("second", "s", "time", 1e12)
# --- mass: gram is the base unit; dalton is gram/mole so that kJ/mol = Da nm^2/ps^2 ---
gram = grams = _base("gram", "g", "mass")
kilogram = kilograms = _base("kilogram", "kg", "mass", 1e3)
# --- temperature, amount, charge, angle ---
kelvin = kelvins = _base("kelvin", "K", "temperature")
mole = moles = _base("mole", "mol", "amount")
elementary_charge = elementary_charges = _base("elementary charge", "e", "charge")
coulomb = coulombs = _base("coulomb", "C", "charge", 1.0 / 1.602176634e-19)
radian = radians = _base("radian", "rad", "angle")
degree = degrees = _base("degree", "deg", "angle", math.pi / 180.0)
candela = candelas = _base("candela", "cd", "luminous intensity")
# --- scaled units ---
dalton = daltons = amu = amus = atom_mass_units = _scaled_unit(
    "dalton", "Da", 1.0, gram / mole
)
_md_energy = gram * nanometer**2 / picosecond**2
kilojoule = kilojoules = _scaled_unit("kilojoule", "kJ", 1.0, _md_energy)
joule = joules = _scaled_unit("joule", "J", 1e-3, _md_energy)
kilocalorie = kilocalories = _scaled_unit("kilocalorie", "kcal", 4.184, _md_energy)
calorie = calories = _scaled_unit("calorie", "cal", 4.184e-3, _md_energy)
kilojoule_per_mole = kilojoules_per_mole = kilojoule / mole
kilocalorie_per_mole = kilocalories_per_mole = kilocalorie / mole
item = items = _scaled_unit("item", "", 1.0 / AVOGADRO, mole)

md_unit_system = UnitSystem(
    [nanometer, picosecond, kelvin, mole, elementary_charge, radian, candela, gram],
    derived=[kilojoule_per_mole, dalton],
)

_NAMESPACE = {
    name: value for name, value in list(globals().items()) if isinstance(value, Unit)
}
_NAMESPACE["elementary charge"] = elementary_charge


def parse_unit(text: str) -> Unit:
    """
    Build a unit from an expression in unit names, e.g. ``"nanometer**2*dalton/(radian**2)"``.

    Plays the role of the ``ast.NodeTransformer`` + ``eval`` in the reference
    (``/root/reference/cvpack/units/units.py:38-44,55-68``) without evaluating arbitrary code:
    only names of known units, numbers, ``*``, ``/`` and ``**`` are accepted.
    """
    text = text.strip().replace("elementary charge", "elementary_charge")
    tree = ast.parse(text, mode="eval")

    def walk(node: ast.AST) -> t.Any:
        if isinstance(node, ast.Expression):
            return walk(node.body)
        if isinstance(node, ast.Name):
            try:
                return _NAMESPACE[node.id]
            except KeyError as error:
                raise ValueError(f"Unknown unit name: {node.id}") from error
        if isinstance(node, ast.Constant) and _is_number(node.value):
            return node.value
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            return -walk(node.operand)
        if isinstance(node, ast.BinOp):
            left, right = walk(node.left), walk(node.right)
            if isinstance(node.op, ast.Mult):
                return left * right
            if isinstance(node.op, ast.Div):
                return left / right
            if isinstance(node.op, ast.Pow):
                return left**right
        raise ValueError(f"Unsupported syntax in unit expression: {text!r}")

    result = walk(tree)
    if isinstance(result, Quantity):
        if result._value == 1:  # pylint: disable=protected-access
            return result.unit
        raise ValueError(f"Unit expression has a numeric prefactor: {text!r}")
    if not isinstance(result, Unit):
        raise ValueError(f"Not a unit expression: {text!r}")
    return result
