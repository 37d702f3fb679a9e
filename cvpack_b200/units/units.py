"""
Serializable units and quantities in the MD unit system.

Mirrors ``/root/reference/cvpack/units/units.py:27-152`` -- ``Unit``/``Quantity`` with YAML tags
``!cvpack.Unit`` (state ``{data: <unit name expression>}``) and ``!cvpack.Quantity`` (state
``{value, unit}``), ``in_md_units`` and ``value_in_md_units`` -- on top of :mod:`cvpack_b200.mmunit`
instead of ``openmm.unit``.
"""

from __future__ import annotations

import typing as t
from numbers import Real

import numpy as np

from .. import mmunit
from ..mm import Vec3, _foreign_quantity
from ..serialization import Serializable

ScalarQuantity = t.Union[mmunit.Quantity, Real]
VectorQuantity = t.Union[mmunit.Quantity, np.ndarray, Vec3]
MatrixQuantity = t.Union[mmunit.Quantity, np.ndarray, t.Sequence[Vec3], t.Sequence[np.ndarray]]


class Unit(mmunit.Unit, Serializable):
    """A unit of measurement that can be written to and read from YAML."""

    def __init__(self, data: t.Union[str, mmunit.Unit, dict]) -> None:
        if isinstance(data, str):
            data = mmunit.parse_unit(data)
        if isinstance(data, mmunit.Unit):
            data = dict(data.iter_base_or_scaled_units())
        super().__init__(data)

    def __repr__(self) -> str:
        return self.get_symbol()

    def __getstate__(self) -> t.Dict[str, str]:
        return {"data": str(self)}

    def __setstate__(self, keywords: t.Dict[str, str]) -> None:
        self.__init__(keywords["data"])


Unit.registerTag("!cvpack.Unit")


class Quantity(mmunit.Quantity, Serializable):
    """A quantity that can be written to and read from YAML."""

    def __init__(self, *args: t.Any) -> None:
        if len(args) == 1 and mmunit.is_quantity(args[0]):
            super().__init__(args[0].value_in_unit(args[0].unit), Unit(args[0].unit))
        else:
            super().__init__(*args)

    def __repr__(self) -> str:
        return str(self)

    def __getstate__(self) -> t.Dict[str, t.Any]:
        value = self.value
        if isinstance(value, np.ndarray):
            value = value.tolist()
        return {"value": value, "unit": str(self.unit)}

    def __setstate__(self, keywords: t.Dict[str, t.Any]) -> None:
        self.__init__(keywords["value"], Unit(keywords["unit"]))

    @property
    def value(self) -> t.Any:
        """The value of the quantity."""
        return self._value

    def in_md_units(self) -> t.Any:
        """The value of the quantity in MD units."""
        return self.value_in_unit_system(mmunit.md_unit_system)


Quantity.registerTag("!cvpack.Quantity")


def in_md_units(quantity: t.Union[ScalarQuantity, VectorQuantity, MatrixQuantity]) -> Quantity:
    """
    Return a quantity in the MD unit system (mass in Da, distance in nm, time in ps,
    temperature in K, energy in kJ/mol, angle in rad).
    """
    if mmunit.is_quantity(quantity):
        unit_in_md_system = quantity.unit.in_unit_system(mmunit.md_unit_system)
        if quantity.unit.conversion_factor_to(unit_in_md_system) == 1:
            return Quantity(quantity)
        return Quantity(quantity.in_units_of(unit_in_md_system))
    return Quantity(quantity, Unit("dimensionless"))


def value_in_md_units(quantity: t.Union[ScalarQuantity, VectorQuantity, MatrixQuantity]) -> t.Any:
    """Return the bare value of a quantity in the MD unit system (numbers pass through)."""
    if mmunit.is_quantity(quantity):
        return quantity.value_in_unit_system(mmunit.md_unit_system)
    if hasattr(quantity, "value_in_unit_system") and hasattr(quantity, "unit"):
        return _foreign_quantity(quantity)
    return quantity
