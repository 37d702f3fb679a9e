"""Serialization subpackage (mirrors ``cvpack.serialization``)."""

from .serialization import (  # noqa: F401
    Serializable,
    SerializableAtom,
    SerializableForce,
    SerializableResidue,
    deserialize,
    serialize,
)
