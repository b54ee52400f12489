"""State containers (reference: src/fluids2d/states.py:7-79): namedtuples of
float64 numpy arrays of shape (n2, n1).  The arrays are page-locked when a CUDA
library is available so that host<->device copies run at PCIe speed."""
from collections import namedtuple

import numpy as np

Specs = namedtuple("specs", ("variables", "prognostic"))

model_specs = {
    "euler": Specs(("u", "U", "omega", "ke", "p", "div", "flx"), ("u",)),
    "eulerpsi": Specs(("omega", "U", "psi", "vomega", "flx"), ("omega",)),
    "boussinesq": Specs(("b", "u", "U", "omega", "ke", "p", "div", "flx"), ("b", "u")),
    "hydrostatic": Specs(("b", "uh", "U", "omega", "ke", "p", "div", "flx"), ("b", "uh")),
    "rsw": Specs(("u", "h", "U", "omega", "ke", "p", "flx", "pv"), ("u", "h")),
    "qgrsw": Specs(("u", "h", "U", "omega", "ke", "p", "flx", "pv", "psi"), ("u", "h")),
    "qg": Specs(("pv", "U", "h", "flx", "work", "psi"), ("pv",)),
    "advection": Specs(("q", "U", "flx"), ("q",)),
    "vectoradv": Specs(("v", "U", "omega", "q"), ("v",)),
}

vectors = ["u", "U", "flx", "v"]


def get_specs(param):
    specs = model_specs[param.model]
    if (param.tracer is None) or (param.tracer == "None"):
        return specs
    p = specs.prognostic + (param.tracer,)
    return Specs(p + specs.variables[len(specs.prognostic):], p)


def _named(name, fields):
    class Namedtuple(namedtuple(name, fields)):
        def __repr__(self):
            return f"{name} with {fields}"
    return Namedtuple


Vector = _named("vector", ("x", "y"))

_pinned = True


def zeros(shape):
    """float64 zeros, page-locked if libf2d is loadable (falls back to pageable
    memory on a box without a CUDA driver: that only affects copy speed)."""
    global _pinned
    if _pinned:
        try:
            from ._cabi import pinned_empty
            a = pinned_empty(shape)
            a[...] = 0.0
            return a
        except Exception:
            _pinned = False
    return np.zeros(shape)


def allocate_var(name, shape):
    if name in vectors:
        return Vector(x=zeros(shape), y=zeros(shape))
    return zeros(shape)


def allocate_state(name, variables, shape):
    T = _named("state", variables)
    return T(**{v: allocate_var(v, shape) for v in variables})


def State(param, shape):
    specs = get_specs(param)
    assert specs.variables[:len(specs.prognostic)] == specs.prognostic
    return allocate_state(param.model, specs.variables, shape)


def Prognostic(param, shape):
    return allocate_state(param.model, get_specs(param).prognostic, shape)


def leaves(s, names=None):
    """[(leaf_name, array)] e.g. ("u.x", arr) over the fields `names` of s."""
    out = []
    for n in (names if names is not None else s._fields):
        v = getattr(s, n)
        if hasattr(v, "_fields"):
            out += [(f"{n}.x", v.x), (f"{n}.y", v.y)]
        else:
            out.append((n, v))
    return out
