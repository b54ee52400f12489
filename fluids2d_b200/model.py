"""Model driver (reference: src/fluids2d/model.py:12-123): same attributes and
loop, with the state resident on the device between observation points."""
import signal
from time import time as _wall

from . import weno as _weno
from .equations import DeviceForcing, addforcingterm
from .integrators import get_integrator
from .io import IO
from .meshes import Mesh
from .param import DEVICE_MODELS
from .states import State
from .timeline import Time


class Model:
    def __init__(self, param):
        param.check()
        if param.model not in DEVICE_MODELS:
            raise NotImplementedError(
                f"model '{param.model}' is not on the B200 hot path ({', '.join(DEVICE_MODELS)})")
        self.param = param
        self.mesh = Mesh(param)
        _weno.bind(self.mesh.engine)
        self.state = State(param, self.mesh.shape)
        self.set_integrator()
        self.time = Time(param)
        self.integrator.time = self.time      # time-dependent device forcings read the clock
        self.io = IO(param, self.mesh, self.state, self.time)
        self._resident = False        # True while run() keeps the state on the device
        self.callbacks = []
        self.diags = []
        self.stop = False

    def set_integrator(self):
        self.integrator = get_integrator(self.param, self.mesh, self.state)

    # ------------------------------------------------------------------ run ---
    def _observation_due(self):
        """does anything at this observation point (model.py:48-51) read the HOST
        state?  Device-side observers (diagnostics.Bulk, io.IO) do not."""
        t = self.time
        host_diags = any(not getattr(d, "on_device", False) for d in self.diags)
        host_io = t.save_to_file and not getattr(self.io, "on_device", False)
        return host_diags or t.update_anim or host_io or t.finished

    def run(self):
        if self.param.animation:
            self.execute_callbacks()
        self.stop = False

        def handler(sig, frame):
            print("\n hit ctrl-C, stopping", end="")
            self.stop = True

        try:
            signal.signal(signal.SIGINT, handler)
        except ValueError:
            pass                      # not in the main thread
        tic = _wall()
        integ = self.integrator
        resident = integ.rhs is integ._device_rhs
        if resident:
            integ.upload(self.state)
            if self.param.integrator == "LFRA":
                integ._scratch_io(True)
        self._resident = resident
        if hasattr(self.io, "resident"):
            self.io.resident = resident
        self.save_to_file()
        while (not self.time.finished) and (not self.stop):
            if resident:
                self.set_dt(on_device=True)
                if self.param.integrator == "LFRA":
                    integ.step_resident(self.time.dt, 1, first=self.time.ite == 0)
                else:
                    integ.step_resident(self.time.dt, 1)
                self.time.pushforward()
                if self._observation_due():
                    integ.download(self.state)
            else:
                self.set_dt()
                self.step(1)
            self.progress()
            self.animation()
            self.compute_diags()
            self.save_to_file()
        if resident:
            integ.download(self.state)
            if self.param.integrator == "LFRA":
                integ._scratch_io(False)
        self._resident = False
        if hasattr(self.io, "resident"):
            self.io.resident = False
        if hasattr(self.io, "flush"):
            self.io.flush()
        self.progress()
        self.print_perf(_wall() - tic)
        self.finalize()

    def finalize(self):
        for d in self.diags:
            if hasattr(d, "finalize"):
                d.finalize()

    def step(self, nsteps=1):
        integ = self.integrator
        if nsteps > 1 and integ.rhs is integ._device_rhs and self.param.dt > 0 and self.param.integrator != "LFRA":
            integ.upload(self.state)
            integ.step_resident(self.time.dt, nsteps)
            integ.download(self.state)
            for _ in range(nsteps):
                self.time.pushforward()
            return
        for _ in range(nsteps):
            integ.step(self.state, self.time)

    def set_dt(self, on_device=False):
        """model.py:71-87"""
        p = self.param
        if p.dt > 0:
            self.time.dt = p.dt
            return
        if p.model == "rsw":
            c = (p.g * p.H) ** 0.5
            maxU = c / self.mesh.dx + c / self.mesh.dy
        elif on_device or getattr(p, "nranks", 1) > 1:
            # device reduction (all-reduced over the slabs when there are several)
            maxU = self.mesh.engine.max_abs_U() + 1e-99
        else:
            import numpy as np
            U = self.state.U
            maxU = np.max(np.abs(U.x)) + np.max(np.abs(U.y)) + 1e-99
        self.time.dt = min(p.cfl / maxU, p.dtmax)

    def print_perf(self, elapsed):
        print()
        nite = max(self.time.ite - self.time.ite0, 1)
        perf = elapsed / (self.mesh.nx * self.mesh.ny * nite)
        print(f"Elapsed: {elapsed:.2f} s  perf: {perf:.2e} s/dof")

    def progress(self):
        if (self.time.ite % self.param.nprint == 0) or self.time.finished:
            print(" ".join([f"\rite={self.time.ite}", self.time.tostring(),
                            f"dt={self.time.dt:.2g}"]), end="")

    def animation(self):
        if self.time.update_anim:
            self.execute_callbacks()
            fig = getattr(self, "figure", None)
            if fig is not None:
                fig.update(self.state, self.time)

    def save_to_file(self):
        if self.time.save_to_file:
            self.io.write(self.state, self.time)

    def execute_callbacks(self):
        for func in self.callbacks:
            func(self.param, self.mesh, self.state, self.time)

    def compute_diags(self):
        for d in self.diags:
            d()

    def add_forcing(self, forcing):
        """model.py:121-123.  A host callable wraps ``integrator.rhs`` (the step then
        runs stage by stage through host buffers); an ``equations.DeviceForcing``
        is installed on the device and keeps the fused resident step."""
        integ = self.integrator
        if isinstance(forcing, DeviceForcing) and integ.rhs is integ._device_rhs:
            print("[INFO] add a forcing term (device)")
            forcing.install(self.mesh.engine, self.time.t)
            integ.device_forcings.append(forcing)
            return
        integ.rhs = addforcingterm(self.param, self.mesh, integ.rhs, forcing)
