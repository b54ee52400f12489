"""Mesh (reference: src/fluids2d/meshes.py:7-143).

Same attributes as the reference -- shape, nx, ny, dx, dy, area, xshift, yshift,
msk, mskx, msky, mskv, slipcoef, oc/ov/ok, poisson_centers, poisson_vertices,
qg_helmholtz, hb, qgcoef -- but every derived array is computed on the device
by ``f2d_set_mask`` and mirrored to numpy, and the three solvers are multigrid
hierarchies in HBM instead of SuperLU factors.  Edit ``msk`` in place and call
``finalize()`` exactly as with the reference.
"""
from collections import namedtuple

import numpy as np

from . import slabs
from ._cabi import Engine
from .elliptic import Poisson2D

Stencil = namedtuple("stencil", ("x", "y"))


class Mesh:
    def __init__(self, param):
        self.param = param
        nranks = getattr(param, "nranks", 1) or 1
        self.slab = slabs.Slab(param.ny, param.halowidth, getattr(param, "rank", 0) or 0, nranks)
        # one GPU: the reference's (ny+2nh, nx+2nh); several: this rank's slab
        self.shape = (self.slab.n2, param.nx + 2 * param.halowidth)
        self.nx, self.ny = param.nx, param.ny
        self.dx = param.Lx / self.nx
        self.dy = param.Ly / self.ny
        self.area = self.dx * self.dy
        self.xshift = 1
        self.yshift = self.shape[-1]
        kind = {"pcg": 0, "mg": 1}[getattr(param, "solver", "pcg")]
        self.engine = Engine(param, device=getattr(param, "device", 0), slab=self.slab,
                             comm=slabs.communicator() if nranks > 1 else None,
                             solver_rtol=getattr(param, "solver_rtol", 0.0),
                             solver_maxit=getattr(param, "solver_maxit", 0), solver_kind=kind,
                             nu1=getattr(param, "solver_nu", 0), nu2=getattr(param, "solver_nu", 0),
                             guess_order=getattr(param, "solver_guess", None))
        self._hb = 0
        self.set_default_mask()
        self.finalize()

    # mesh.hb is assigned by user scripts (rsw_with_topo.py:96-99); keep the device copy in step
    @property
    def hb(self):
        return self._hb

    @hb.setter
    def hb(self, value):
        self._hb = value
        self.engine.set_topography(value)

    def _allocate(self):
        return np.zeros(self.shape, dtype="i1")

    def set_default_mask(self):
        nh = self.param.halowidth
        self.msk = self._allocate()
        xs = slice(None) if self.param.xperiodic else slice(nh, -nh)
        ys = slice(None) if (self.param.yperiodic or getattr(self.param, "ywrap", False)) else slice(nh, -nh)
        if self.slab.nranks > 1:
            # rows of the global default mask that fall in this rank's window
            g = np.zeros((self.ny + 2 * nh, self.shape[1]), dtype="i1")
            g[ys, xs] = 1
            self.msk[:] = g[self.slab.window()]
        else:
            self.msk[ys, xs] = 1

    def finalize(self):
        """Re-derive masks, slip coefficient, stencil orders and solvers from
        ``self.msk`` (meshes.py:34-52); everything is rebuilt on the device."""
        e = self.engine
        e.set_mask(self.msk)
        self.mskx, self.msky, self.mskv = (e.mesh_array(k) for k in ("mskx", "msky", "mskv"))
        self.slipcoef = e.mesh_array("slip")      # 0/1, as in noslip.py:4-36
        self.oc = Stencil(e.mesh_array("oc.x"), e.mesh_array("oc.y"))
        self.ov = Stencil(e.mesh_array("ov.x"), e.mesh_array("ov.y"))
        self.ok = Stencil(e.mesh_array("ok.x"), e.mesh_array("ok.y"))
        self.poisson_centers = Poisson2D(self, "c")
        self.poisson_vertices = Poisson2D(self, "v")
        if self.param.model in ("qg", "qgrsw", "rsw"):
            p = self.param
            self.qg_helmholtz = Poisson2D(self, "v", maindiag=self.area * p.f0 ** 2 / (p.g * p.H))
            self.hb = 0
            self.qgcoef = p.f0 / +p.H

    def x(self, which):
        idx = np.arange(self.nx + 2 * self.param.halowidth) - self.param.halowidth
        return (idx + (0.5 if which in ("c", "y") else 0.0)) * self.dx

    def y(self, which):
        idx = np.arange(self.slab.n2) + self.slab.row0 - self.param.halowidth
        return (idx + (0.5 if which in ("c", "x") else 0.0)) * self.dy

    def xy(self, which="c"):
        return np.meshgrid(self.x(which), self.y(which))

    def fill(self, variable):
        """x-periodic halo copy on HOST arrays (meshes.py:106-111, 135-143): a
        convenience for initial conditions; the time step does it on the device."""
        if hasattr(variable, "_fields"):
            for v in variable:
                self.fill(v)
        else:
            fill_halo_array(self.param, variable)


def get_shape(param):
    return (param.ny + 2 * param.halowidth, param.nx + 2 * param.halowidth)


def fill_halo_array(param, array):
    n = param.halowidth
    if param.xperiodic:
        array[..., :n] = array[..., -2 * n:-n]
        array[..., -n:] = array[..., n:2 * n]
    if getattr(param, "ywrap", False):        # NEW: a true periodic y direction (after x: corners are images too)
        array[..., :n, :] = array[..., -2 * n:-n, :]
        array[..., -n:, :] = array[..., n:2 * n, :]
