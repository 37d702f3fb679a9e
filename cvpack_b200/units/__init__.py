"""Units subpackage (mirrors ``cvpack.units``)."""

from .units import (  # noqa: F401
    MatrixQuantity,
    Quantity,
    ScalarQuantity,
    Unit,
    VectorQuantity,
    in_md_units,
    value_in_md_units,
)
