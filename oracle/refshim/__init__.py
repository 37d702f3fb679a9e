"""
``oracle.refshim``: run the *reference's own Python code* (``/root/reference/cvpack``) without OpenMM.

TEST INFRASTRUCTURE ONLY.  ``install()`` registers stand-in ``openmm``, ``openmm.unit``,
``openmm.app`` (+ ``.topology``, ``.element``), ``openmm._openmm`` and ``openmmcppforces`` modules in
``sys.modules`` (see ``fake_openmm.py`` for what is restated and what is not), puts the reference
checkout on ``sys.path`` and returns the imported reference ``cvpack`` package.  It is used by
``tests/golden/make_reference_golden.py`` to generate the committed golden vectors and by
``tests/test_refshim.py``; it needs ``/root/reference`` and therefore never runs on the GPU box.
"""

from __future__ import annotations

import importlib
import os
import sys
import types
import typing as t

REFERENCE_ROOT = os.environ.get("CVPACK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    """Whether the reference checkout is present (it is not on the GPU box)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cvpack", "__init__.py"))


def install() -> types.ModuleType:
    """Register the stand-in modules and import the reference package; returns ``cvpack``."""
    if "openmm" in sys.modules and not getattr(sys.modules["openmm"], "__refshim__", False):
        raise RuntimeError("a real openmm is already imported; the shim must not shadow it")
    from cvpack_b200 import app as app_module  # pylint: disable=import-outside-toplevel
    from cvpack_b200 import mmunit  # pylint: disable=import-outside-toplevel

    from . import fake_openmm  # pylint: disable=import-outside-toplevel

    if "openmm" not in sys.modules:
        openmm = types.ModuleType("openmm")
        openmm.__refshim__ = True
        openmm.__path__ = []  # a package, so that ``from openmm import unit`` resolves via sys.modules
        for name, value in vars(fake_openmm).items():
            if isinstance(value, type) and value.__module__ == fake_openmm.__name__ and not name.startswith("_"):
                setattr(openmm, name, value)
        openmm.Vec3 = fake_openmm.Vec3
        openmm.CustomIntegrator = fake_openmm.CustomIntegrator
        openmm.unit = mmunit
        openmm.app = app_module
        openmm._openmm = fake_openmm._Swig  # pylint: disable=protected-access
        app_module.topology = app_module
        app_module.element = app_module
        plugin = types.ModuleType("openmmcppforces")
        plugin.CompositeRMSDForce = fake_openmm.CompositeRMSDForce
        sys.modules.update(
            {
                "openmm": openmm,
                "openmm.unit": mmunit,
                "openmm.app": app_module,
                "openmm.app.topology": app_module,
                "openmm.app.element": app_module,
                "openmm._openmm": fake_openmm._Swig,  # pylint: disable=protected-access
                "openmmcppforces": plugin,
            }
        )
    if not available():
        raise FileNotFoundError(f"reference checkout not found under {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    return importlib.import_module("cvpack")


def namespaces() -> t.Tuple[t.Any, t.Any]:
    """(reference namespace, product namespace) with the same attribute names, for building the
    same collective variable through both packages: ``.cvpack``, ``.System``, ``.NonbondedForce``,
    ``.app``, ``.unit``."""
    reference = install()
    import cvpack_b200  # pylint: disable=import-outside-toplevel
    from cvpack_b200 import app as app_module  # pylint: disable=import-outside-toplevel
    from cvpack_b200 import mmunit  # pylint: disable=import-outside-toplevel

    openmm = sys.modules["openmm"]
    ref_ns = types.SimpleNamespace(
        cvpack=reference, System=openmm.System, NonbondedForce=openmm.NonbondedForce,
        app=app_module, unit=mmunit, is_reference=True,
    )
    own_ns = types.SimpleNamespace(
        cvpack=cvpack_b200, System=cvpack_b200.System, NonbondedForce=cvpack_b200.NonbondedForce,
        app=app_module, unit=mmunit, is_reference=False,
    )
    return ref_ns, own_ns
