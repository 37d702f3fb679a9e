"""MetaCollectiveVariable: the reference's extra methods
(meta_collective_variable.py:143-261) on the batched GPU path."""

import io

import numpy as np
import pytest

import cvpack_b200 as cvpack
from cvpack_b200 import mmunit as unit
from cvpack_b200 import serialization
from tests import helpers


def build(rng, n=30):
    system = helpers.make_system(n, rng)
    phi = cvpack.Torsion(3, 17, 5, 9, name="phi")
    dist = cvpack.Distance(2, 11, name="dist")
    rg = cvpack.RadiusOfGyration(range(8, 20), name="rg")
    cv = cvpack.MetaCollectiveVariable(
        "0.5*kappa*min(delta,6.283185307179586-delta)^2 + dist/(rg+shift); delta=abs(phi-phi0)",
        [phi, dist, rg], unit.kilojoules_per_mole,
        kappa=1000 * unit.kilojoules_per_mole / unit.radian**2, phi0=120 * unit.degrees,
        shift=1.0 * unit.angstroms, name="drive",
    )
    return system, cv


def test_construction_and_serialization():
    system, cv = build(np.random.default_rng(71))
    assert cv.getName() == "drive" and str(cv.getUnit()) == "kilojoule/mole"
    assert cv.getMassUnit().get_symbol() == "nm**2 mol**2 Da/(kJ**2)"
    defaults = cv.getParameterDefaultValues()
    assert defaults["phi0"].unit == unit.radians and abs(defaults["phi0"]._value - 2 * np.pi / 3) < 1e-15
    assert defaults["shift"]._value == pytest.approx(0.1) and defaults["shift"].unit == unit.nanometers
    assert [v.getName() for v in cv.getInnerVariables()] == ["phi", "dist", "rg"]
    stream = io.StringIO()
    serialization.serialize(cv, stream)
    assert stream.getvalue().startswith("!cvpack.MetaCollectiveVariable\n")
    stream.seek(0)
    clone = serialization.deserialize(stream)
    assert isinstance(clone, cvpack.MetaCollectiveVariable)
    again = io.StringIO()
    serialization.serialize(clone, again)
    assert again.getvalue() == stream.getvalue()
    with pytest.raises(ValueError, match="unknown variable"):
        cvpack.MetaCollectiveVariable("phi+psi", [cvpack.Torsion(0, 1, 2, 3, name="phi")], unit.radians)
    with pytest.raises(ValueError, match="distinct names"):
        cvpack.MetaCollectiveVariable("d", [cvpack.Distance(0, 1, name="d"), cvpack.Distance(1, 2, name="d")],
                                      unit.nanometers)
    del system


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["single", "double"])
def test_meta_methods(precision):
    import torch

    rng = np.random.default_rng(72)
    system, cv = build(rng)
    x = rng.uniform(0, 1.5, size=(9, 30, 3))
    tol = 1e-5 if precision == "single" else 1e-10
    ctx = helpers.check_against_oracle(cv, system, x, None, precision, tol, tol)
    stored = ctx.getPositions().double().cpu().numpy()
    inner = cv.getInnerValues(ctx)
    masses = cv.getInnerEffectiveMasses(ctx)
    for variable in cv.getInnerVariables():
        name = variable.getName()
        assert inner[name].unit == variable.getUnit() and masses[name].unit == variable.getMassUnit()
        for b in range(9):
            value, grad = helpers.oracle_evaluate(variable, system, stored[b], None)
            assert float(inner[name]._value[b]) == pytest.approx(value, rel=tol, abs=tol)
            expected = helpers.oracle.effective_mass(grad, system.getMasses())
            assert float(masses[name]._value[b]) == pytest.approx(expected, rel=4 * tol)
    # d value / d parameter against central differences of rebuilt variables
    derivs = cv.getParameterDerivatives(ctx)
    value0 = cv.getValue(ctx)._value.double().cpu().numpy()
    defaults = cv.getParameterDefaultValues()
    for key, h in (("kappa", 1e-2), ("phi0", 1e-6), ("shift", 1e-6)):
        vals = []
        for sign in (+1, -1):
            ctx.setParameter(key, defaults[key]._value + sign * h)
            vals.append(cv.getValue(ctx)._value.double().cpu().numpy())
        ctx.setParameter(key, defaults[key]._value)
        fd = (vals[0] - vals[1]) / (2 * h)
        got = derivs[key]._value.double().cpu().numpy()
        if precision == "double":
            assert np.allclose(got, fd, rtol=1e-6, atol=1e-6 * np.abs(fd).max())
        assert derivs[key].unit == cv.getUnit() / defaults[key].unit
    assert np.array_equal(cv.getValue(ctx)._value.double().cpu().numpy(), value0)
    assert cv.getParameterValues(ctx)["kappa"]._value == 1000.0
    ctx.setParameter("kappa", 500.0)
    assert cv.getParameterValues(ctx)["kappa"]._value == 500.0
    half = cv.getValue(ctx)._value
    assert torch.all(half <= torch.as_tensor(value0, device=half.device, dtype=half.dtype) + 1e-6)
    # host residency goes through the same kernels
    host = cvpack.BatchContext(system, x, precision=precision, residency="host")
    v, g = cv.getValueAndGradient(host)
    assert not v._value.is_cuda and not g.is_cuda
    ctx.setParameter("kappa", 1000.0)
    v0, g0 = cv.getValueAndGradient(ctx)
    assert torch.equal(v._value, v0._value.cpu()) and torch.equal(g, g0.cpu())
