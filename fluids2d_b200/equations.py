"""Equation-set helpers that user scripts import (reference:
src/fluids2d/equations.py:4-6, 229-238).  The model right-hand sides themselves
are CUDA kernels (csrc/step.cu); this module only keeps the host-side hooks."""


def fill(mesh, *variables):
    for v in variables:
        mesh.fill(v)


def addforcingterm(param, mesh, rhs, forcing):
    """wrap a (device) rhs with a host forcing callback ``forcing(param, mesh, s, ds)``"""
    print("[INFO] add a forcing term")

    def newrhs(s, ds):
        rhs(s, ds)
        forcing(param, mesh, s, ds)

    newrhs.__doc__ = "\n".join([rhs.__doc__ or "", "with forcing term"])
    return newrhs


class DeviceForcing:
    """A forcing of the form ``ds.<field> += amplitude(t) * pattern`` (what
    forced_convection.py:9-24 and warm_bubble.py:8 write) that stays on the device.

        heat = DeviceForcing({"b": Q_pattern}, amplitude=lambda t: np.tanh(t / 50))
        model.add_forcing(heat)

    ``patterns`` maps a prognostic field (``"b"``, ``"h"``, the tracer's name, or
    ``"u"`` with an ``(x, y)`` pair / ``"u.x"``) to a full haloed array.  The
    integrator keeps its fused, resident step and only sends the amplitude once
    per time step (as ``mesh.time.t`` is constant within a step in the reference).
    The object is also a valid HOST callback ``forcing(param, mesh, s, ds)``, so
    the same forcing can be handed to the reference (set ``mesh.time`` as
    forced_convection.py:6 does)."""

    def __init__(self, patterns, amplitude=1.0):
        self.patterns = {}
        for name, a in patterns.items():
            if isinstance(a, (tuple, list)) or hasattr(a, "_fields"):
                self.patterns[name + ".x"], self.patterns[name + ".y"] = a[0], a[1]
            else:
                self.patterns[name] = a
        self.amplitude = amplitude

    def amp(self, t):
        return float(self.amplitude(t)) if callable(self.amplitude) else float(self.amplitude)

    # host-callback form
    def __call__(self, param, mesh, s, ds):
        time = getattr(mesh, "time", None)
        a = self.amp(time.t if time is not None else 0.0)
        for leaf, F in self.patterns.items():
            tgt = ds
            for part in leaf.split("."):
                tgt = getattr(tgt, part)
            tgt += a * F

    # device form
    def install(self, engine, t):
        for leaf, F in self.patterns.items():
            engine.set_forcing(leaf, F, self.amp(t))

    def update(self, engine, t):
        if callable(self.amplitude):
            a = self.amp(t)
            for leaf in self.patterns:
                engine.set_forcing(leaf, None, a)
