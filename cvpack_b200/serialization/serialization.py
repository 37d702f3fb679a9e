"""
YAML (de)serialization of collective variables and of the host objects embedded in them.

Same wire format as the reference (``/root/reference/cvpack/serialization/serialization.py:17-186``):
PyYAML ``SafeDumper``/``SafeLoader`` tags ``!cvpack.<Class>``; the state of an object is what its
``__getstate__`` returns (for a collective variable: its constructor keyword arguments,
``collective_variable.py:32-36``); atoms and residues are stored by value; an embedded
``NonbondedForce`` is stored as OpenMM XML text under ``xml_code``.
"""

from __future__ import annotations

import typing as t

import yaml

from .. import app as mmapp
from .. import mm


class Serializable(yaml.YAMLObject):
    """Mixin that lets PyYAML's safe dumper/loader handle a class."""

    @classmethod
    def registerTag(cls, tag: str) -> None:
        """Attach a YAML tag to this class for safe dumping and loading."""
        cls.yaml_tag = tag
        yaml.SafeDumper.add_representer(cls, cls.to_yaml)
        yaml.SafeLoader.add_constructor(tag, cls.from_yaml)


class SerializableAtom(Serializable):
    """An atom stored by value (name, index, element symbol, residue index, id)."""

    def __init__(self, atom: t.Any) -> None:  # pylint: disable=super-init-not-called
        self.name = atom.name
        self.index = atom.index
        if hasattr(atom, "_element_symbol"):
            self._element_symbol = atom._element_symbol
            self._residue_index = atom._residue_index
        else:
            self._element_symbol = None if atom.element is None else atom.element.symbol
            self._residue_index = atom.residue.index
        self.id = atom.id

    def __getstate__(self) -> t.Dict[str, t.Any]:
        return self.__dict__

    def __setstate__(self, keywords: t.Dict[str, t.Any]) -> None:
        self.__dict__.update(keywords)

    @property
    def element(self) -> t.Optional[mmapp.Element]:
        """The chemical element of this atom."""
        if self._element_symbol is None:
            return None
        return mmapp.Element.getBySymbol(self._element_symbol)


SerializableAtom.registerTag("!cvpack.Atom")


class SerializableResidue(Serializable):
    """A residue stored by value together with its atoms."""

    def __init__(self, residue: t.Any) -> None:  # pylint: disable=super-init-not-called
        self.name = residue.name
        self.index = residue.index
        if hasattr(residue, "_chain_index"):
            self._chain_index = residue._chain_index
        else:
            self._chain_index = residue.chain.index
        self.id = residue.id
        self._atoms = [SerializableAtom(atom) for atom in residue.atoms()]

    def __getstate__(self) -> t.Dict[str, t.Any]:
        return self.__dict__

    def __setstate__(self, keywords: t.Dict[str, t.Any]) -> None:
        self.__dict__.update(keywords)

    def __len__(self) -> int:
        return len(self._atoms)

    def atoms(self) -> t.Iterator[SerializableAtom]:
        """Iterate over the atoms of this residue."""
        return iter(self._atoms)


SerializableResidue.registerTag("!cvpack.Residue")


class SerializableForce(Serializable):
    """A nonbonded force stored as OpenMM XML; attribute access falls through to the force."""

    def __init__(self, force: t.Any) -> None:  # pylint: disable=super-init-not-called
        self.force = force.force if isinstance(force, SerializableForce) else force

    def __getattr__(self, name: str) -> t.Any:
        if name == "force":
            raise AttributeError(name)
        return getattr(self.force, name)

    def __getstate__(self) -> t.Dict[str, str]:
        return {"xml_code": mm.XmlSerializer.serialize(self.force)}

    def __setstate__(self, keywords: t.Dict[str, str]) -> None:
        self.__init__(mm.XmlSerializer.deserialize(keywords["xml_code"]))


SerializableForce.registerTag("!cvpack.Force")


def serialize(obj: t.Any, iostream: t.IO) -> None:
    """
    Write an object of this package to a text stream.

    Example (the golden text of ``serialization.py:138-154`` in the reference)
    -------
    >>> import io, cvpack_b200 as cvpack
    >>> stream = io.StringIO()
    >>> cvpack.serialization.serialize(cvpack.RadiusOfGyration([0, 1, 2]), stream)
    >>> print(stream.getvalue())
    !cvpack.RadiusOfGyration
    group:
    - 0
    - 1
    - 2
    name: radius_of_gyration
    pbc: false
    weighByMass: false
    <BLANKLINE>
    """
    iostream.write(yaml.safe_dump(obj))


def deserialize(iostream: t.IO) -> t.Any:
    """Read an object of this package back from a text stream."""
    return yaml.safe_load(iostream.read())
