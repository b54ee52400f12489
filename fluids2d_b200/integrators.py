"""Time integrators (reference: src/fluids2d/integrators.py).

``get_integrator(param, mesh, state)`` returns an object with the reference's
surface -- ``.step(state, time)``, ``.rhs(s, ds)``, ``.diag(s)``, ``.scratch``
-- whose work runs in libf2d.so.  ``state`` stays a namedtuple of numpy arrays
owned by the caller; how often it is synchronised with the device copy:

  * ``integrator.step(state, time)``  (the reference's per-step call): uploads
    the state, runs one fused step on the device, downloads the state.
  * ``Model.run()`` / ``Model.step(n)`` keep the state resident and only
    download at observation points (see model.py).
  * after ``model.add_forcing(f)`` the host callback must see ``s`` and ``ds``
    every stage (equations.py:229-238), so the step runs stage by stage with a
    download / callback / upload around each tendency.
"""
from collections import namedtuple

import numpy as np

from .states import Prognostic, leaves

_specs = namedtuple("specs", ("caller", "nstages"))


def get_integrator(param, mesh, state):
    if param.integrator in RKintegrators:
        return RKIntegrator(param, mesh, state)
    if param.integrator == "LFRA":
        return LFRAintegrator(param, mesh, state)
    raise NotImplementedError(f"{param.integrator} is not implemented on the device path")


def rk_coefficients(name, dt):
    """incremental-form coefficients, written as in integrators.py:82-124 so the
    host computes the very same doubles the reference passes to addto()"""
    if name == "ef":
        return [(dt,)]
    if name == "rk3":
        return [(dt,), (-3 * dt / 4, dt / 4), (-dt / 12, -dt / 12, 2 * dt / 3)]
    if name == "enrk3":
        return [(dt / 3,), (-dt / 3 - 5 * dt / 48, 15 * dt / 16),
                (5 * dt / 48 + dt / 10, -7 * dt / 16, 2 * dt / 5)]
    raise NotImplementedError(name)


class RKIntegrator:
    """Runge-Kutta integrator (rk3: SSP RK3, ef: Euler forward, enrk3: energy-preserving RK3)"""

    def __init__(self, param, mesh, state):
        specs = RKintegrators[param.integrator]
        self.param, self.mesh = param, mesh
        self.engine = mesh.engine
        self.name = param.integrator
        self.scratch = [Prognostic(param, mesh.shape) for _ in range(specs.nstages)]
        self._prognostic = type(self.scratch[0])._fields
        self._device_rhs = self._rhs_on_device
        self.rhs = self._device_rhs
        self.diag = self._diag_on_device
        self._fields = None
        self.device_forcings = []      # equations.DeviceForcing objects (Model.add_forcing)
        self.time = None

    # ---- host <-> device ----------------------------------------------------
    def _names(self, state):
        if self._fields is None:
            e = self.engine
            names = []
            for n, _ in leaves(state):
                try:
                    e.field_ptr(n)
                    names.append(n)
                except Exception:
                    pass           # e.g. euler's flx: never touched by the step
            self._fields = names
        return self._fields

    def upload(self, state, names=None):
        e = self.engine
        want = set(names) if names is not None else None
        for n, a in leaves(state):
            if n in self._names(state) and (want is None or n in want):
                e.upload_async(n, a)
        e.sync()

    def download(self, state, names=None):
        e = self.engine
        want = set(names) if names is not None else None
        for n, a in leaves(state):
            if n in self._names(state) and (want is None or n in want):
                e.download_async(n, a)
        e.sync()

    def _scratch_index(self, ds):
        for k, s in enumerate(self.scratch):
            if s is ds:
                return k
        raise ValueError("ds must be one of integrator.scratch")

    # ---- the reference's callables ---------------------------------------------
    def _rhs_on_device(self, s, ds):
        """ds = rhs(s)   (equations.py:11-15, 29-35, 121-127, 141-148)"""
        k = self._scratch_index(ds)
        self.upload(s)
        self.engine.rhs(k)
        for n, a in leaves(ds):
            self.engine.download_async(f"ds{k}.{n}", a)
        if self.param.model == "qgrsw":
            self.download(s, ["pv", "psi"])
        self.engine.sync()

    def _diag_on_device(self, s):
        """diag(s)   (equations.py:17-22, 37-43, 129-134, 150-155)"""
        self.upload(s)
        self.engine.diag()
        self.download(s)

    def _step_inputs(self, state):
        """fields a step READS before it writes them: everything except the pure
        outputs / work arrays (U = sharp(u) where the model carries u, div, flx,
        vomega, work), which need not cross PCIe on the way in"""
        has_u = "u" in state._fields or "uh" in state._fields
        skip = {"div", "flx.x", "flx.y", "vomega", "work"} | ({"U.x", "U.y"} if has_u else set())
        # solutions of the step's own elliptic solves: the reference's direct solve
        # overwrites them whatever they hold (elliptic.py:80-87); here they only seed
        # the iteration, and after the first step the device copy of the previous
        # step (and the first-guess history behind it) does that better
        if not getattr(self, "_seeded", False):
            pass
        elif self.param.model in ("euler", "boussinesq"):
            skip |= {"p"}
        elif self.param.model == "qgrsw":
            skip |= {"pv", "psi"}
        elif self.param.model in ("eulerpsi", "qg"):
            skip |= {"psi"}
        return [n for n in self._names(state) if n not in skip]

    def _step_outputs(self, state):
        """fields a step leaves different on the host: all of them by default, as
        the reference's in-place step does.  ``integrator.skip_outputs`` (a set of
        leaf names, empty by default) lets a script that never reads e.g. ``div``
        or ``U`` between steps keep them off the PCIe bus."""
        skip = getattr(self, "skip_outputs", None) or ()
        return [n for n in self._names(state) if n not in skip]

    def _update_forcings(self, t):
        for f in self.device_forcings:
            f.update(self.engine, t)

    def step(self, state, time):
        if self.rhs is self._device_rhs:
            self.upload(state, self._step_inputs(state))
            self._seeded = True
            self._update_forcings(time.t)
            self.engine.step(time.dt, 1)
            self.download(state, self._step_outputs(state))
        else:
            self._step_with_host_rhs(state, time.dt)
        time.pushforward()

    def step_resident(self, dt, nsteps=1):
        """device-only step(s): the caller guarantees the device state is current"""
        if self.rhs is not self._device_rhs:
            raise RuntimeError("a host forcing is installed: use step()")
        if self.device_forcings and self.time is not None:
            # the amplitude follows the clock: one call per step, t as the reference's
            # callback would read it (constant within a step)
            for k in range(nsteps):
                self._update_forcings(self.time.t + k * dt)
                self.engine.step(dt, 1)
            return
        self.engine.step(dt, nsteps)

    def _step_with_host_rhs(self, state, dt):
        """stage-by-stage step for a user-wrapped rhs (model.add_forcing)"""
        e = self.engine
        for k, coefs in enumerate(rk_coefficients(self.name, dt)):
            ds = self.scratch[k]
            self.rhs(state, ds)                      # device tendency + host forcing on ds
            for n, a in leaves(ds):
                e.upload_async(f"ds{k}.{n}", a)
            self.upload(state)
            e.addto(coefs)
            e.diag()
            self.download(state)


class LFRAintegrator(RKIntegrator):
    """Leap-Frog integrator combined with a Robert-Asselin filter; the first
    iteration is an Euler forward step (integrators.py:20-53).
    ``scratch = [sb, sa, ds]`` as in the reference; scripts that edit it
    (tracer_advection.fliptime) see and set the host copies."""

    def __init__(self, param, mesh, state):
        self.param, self.mesh = param, mesh
        self.engine = mesh.engine
        self.name = "LFRA"
        self.scratch = [Prognostic(param, mesh.shape) for _ in range(3)]
        self.RAgamma = param.RAgamma
        self._prognostic = type(self.scratch[0])._fields
        self._device_rhs = self._rhs_on_device
        self.rhs = self._device_rhs
        self.diag = self._diag_on_device
        self._fields = None
        self.device_forcings = []      # equations.DeviceForcing objects (Model.add_forcing)
        self.time = None

    def _scratch_io(self, upload):
        e = self.engine
        for k in (0, 1):      # sb, sa carry the leap-frog history
            for n, a in leaves(self.scratch[k]):
                (e.upload_async if upload else e.download_async)(f"ds{k}.{n}", a)
        e.sync()

    def step(self, state, time):
        if self.rhs is not self._device_rhs:
            raise NotImplementedError("host forcing with the LFRA integrator is not on the device path")
        self.upload(state, self._step_inputs(state))
        self._scratch_io(True)
        for f in self.device_forcings:
            f.update(self.engine, time.t)
        self.engine.step_lfra(time.dt, time.ite == 0, self.RAgamma)
        self.download(state)
        self._scratch_io(False)
        time.pushforward()

    def step_resident(self, dt, nsteps=1, first=False):
        for k in range(nsteps):
            if self.time is not None:
                for f in self.device_forcings:
                    f.update(self.engine, self.time.t + k * dt)
            self.engine.step_lfra(dt, first and k == 0, self.RAgamma)


RKintegrators = {"rk3": _specs("rk3", 3), "ef": _specs("ef", 1), "enrk3": _specs("enrk3", 3)}


def copyto(x, y):
    """copy x into y (integrators.py:140-151); works on nested namedtuples"""
    if hasattr(y, "_fields"):
        for k in range(min(len(y), len(x))):
            copyto(x[k], y[k])
    else:
        assert isinstance(y, np.ndarray)
        y[:] = x[:]


def addto(y, *args):
    """y += c0*x0 + c1*x1 + ...  on host arrays / nested namedtuples
    (integrators.py:154-196); a host utility for scripts."""
    assert len(args) % 2 == 0
    coefs, xs = args[::2], args[1::2]
    if hasattr(y, "_fields"):
        for k in range(len(xs[0])):
            addto(y[k], *[v for c, x in zip(coefs, xs) for v in (c, x[k])])
    else:
        assert isinstance(y, np.ndarray)
        y[:] += sum((c * x for c, x in zip(coefs, xs)))
