"""
Minimal molecular-topology containers (chains, residues, atoms, elements).

The reference receives ``openmm.app.Topology`` residues and atoms and only ever touches
``residue.atoms()``, ``residue.name``, ``residue.id``, ``residue.index``, ``residue.chain.index``,
``atom.name``, ``atom.index``, ``atom.id``, ``atom.element`` and ``atom.residue.index``
(``/root/reference/cvpack/serialization/serialization.py:37-99``,
``helix_torsion_content.py:115-121``, ``residue_coordination.py:136-140``,
``base_rmsd_content.py:89-102``).  OpenMM is not a dependency here, so this module provides
duck-type compatible stand-ins; real ``openmm.app`` objects work equally well wherever these do.
"""

from __future__ import annotations

import typing as t

from . import mmunit


class Element:
    """A chemical element; instances are singletons so ``a.element is hydrogen`` works."""

    _by_symbol: t.Dict[str, "Element"] = {}

    def __init__(self, number: int, name: str, symbol: str, mass: float) -> None:
        self.atomic_number = number
        self.name = name
        self.symbol = symbol
        self._mass = mass
        Element._by_symbol[symbol.upper()] = self

    @property
    def mass(self) -> mmunit.Quantity:
        """Atomic mass."""
        return self._mass * mmunit.daltons

    @staticmethod
    def getBySymbol(symbol: str) -> "Element":
        """Look an element up by its chemical symbol."""
        try:
            return Element._by_symbol[symbol.upper()]
        except KeyError as error:
            raise KeyError(f"Unknown element symbol: {symbol}") from error

    def __repr__(self) -> str:
        return f"<Element {self.name}>"


hydrogen = Element(1, "hydrogen", "H", 1.007947)
helium = Element(2, "helium", "He", 4.003)
lithium = Element(3, "lithium", "Li", 6.9412)
carbon = Element(6, "carbon", "C", 12.01078)
nitrogen = Element(7, "nitrogen", "N", 14.00672)
oxygen = Element(8, "oxygen", "O", 15.99943)
fluorine = Element(9, "fluorine", "F", 18.99840325)
sodium = Element(11, "sodium", "Na", 22.989769282)
magnesium = Element(12, "magnesium", "Mg", 24.30506)
phosphorus = Element(15, "phosphorus", "P", 30.9737622)
sulfur = Element(16, "sulfur", "S", 32.0655)
chlorine = Element(17, "chlorine", "Cl", 35.4532)
potassium = Element(19, "potassium", "K", 39.09831)
calcium = Element(20, "calcium", "Ca", 40.0784)
iron = Element(26, "iron", "Fe", 55.8452)
zinc = Element(30, "zinc", "Zn", 65.4094)
bromine = Element(35, "bromine", "Br", 79.9041)
iodine = Element(53, "iodine", "I", 126.904473)


class Atom:
    """An atom of a :class:`Topology`."""

    def __init__(
        self, name: str, element: t.Optional[Element], index: int, residue: "Residue", id: str
    ) -> None:  # pylint: disable=redefined-builtin
        self.name = name
        self.element = element
        self.index = index
        self.residue = residue
        self.id = id

    def __repr__(self) -> str:
        return f"<Atom {self.index} ({self.name}) of residue {self.residue.index}>"


class Residue:
    """A residue of a :class:`Topology`."""

    def __init__(self, name: str, index: int, chain: "Chain", id: str) -> None:  # pylint: disable=redefined-builtin
        self.name = name
        self.index = index
        self.chain = chain
        self.id = id
        self._atoms: t.List[Atom] = []

    def atoms(self) -> t.Iterator[Atom]:
        """Iterate over the atoms of this residue."""
        return iter(self._atoms)

    def __len__(self) -> int:
        return len(self._atoms)

    def __repr__(self) -> str:
        return f"<Residue {self.index} ({self.name}) of chain {self.chain.index}>"


class Chain:
    """A chain of a :class:`Topology`."""

    def __init__(self, index: int, topology: "Topology", id: str) -> None:  # pylint: disable=redefined-builtin
        self.index = index
        self.topology = topology
        self.id = id
        self._residues: t.List[Residue] = []

    def residues(self) -> t.Iterator[Residue]:
        """Iterate over the residues of this chain."""
        return iter(self._residues)


class Topology:
    """Chains → residues → atoms, with running indices like ``openmm.app.Topology``."""

    def __init__(self) -> None:
        self._chains: t.List[Chain] = []
        self._num_residues = 0
        self._num_atoms = 0

    def addChain(self, id: t.Optional[str] = None) -> Chain:  # pylint: disable=redefined-builtin
        """Append a chain."""
        index = len(self._chains)
        chain = Chain(index, self, str(index + 1) if id is None else id)
        self._chains.append(chain)
        return chain

    def addResidue(self, name: str, chain: Chain, id: t.Optional[str] = None) -> Residue:  # pylint: disable=redefined-builtin
        """Append a residue to a chain."""
        index = self._num_residues
        residue = Residue(name, index, chain, str(index + 1) if id is None else id)
        self._num_residues += 1
        chain._residues.append(residue)  # pylint: disable=protected-access
        return residue

    def addAtom(
        self,
        name: str,
        element: t.Optional[Element],
        residue: Residue,
        id: t.Optional[str] = None,  # pylint: disable=redefined-builtin
    ) -> Atom:
        """Append an atom to a residue."""
        index = self._num_atoms
        atom = Atom(name, element, index, residue, str(index + 1) if id is None else id)
        self._num_atoms += 1
        residue._atoms.append(atom)  # pylint: disable=protected-access
        return atom

    def chains(self) -> t.Iterator[Chain]:
        """Iterate over chains."""
        return iter(self._chains)

    def residues(self) -> t.Iterator[Residue]:
        """Iterate over residues."""
        for chain in self._chains:
            yield from chain.residues()

    def atoms(self) -> t.Iterator[Atom]:
        """Iterate over atoms."""
        for residue in self.residues():
            yield from residue.atoms()

    def getNumAtoms(self) -> int:
        """Number of atoms."""
        return self._num_atoms

    def getNumResidues(self) -> int:
        """Number of residues."""
        return self._num_residues
