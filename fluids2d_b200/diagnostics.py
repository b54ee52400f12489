"""Bulk diagnostics (reference: src/fluids2d/diagnostics.py:8-84): domain
averages of kinetic energy, enstrophy, vorticity and angular momentum every
third iteration, buffered and appended to ``bulk.nc``.

The four averages need six whole-array sums; while ``Model.run()`` keeps the
state on the device they are one fused reduction there (``f2d_bulk_sums``, 48
bytes back per call) instead of a download of five fields.  From host arrays
(outside ``run()``) the reference's numpy expressions are used."""
from collections import namedtuple

import numpy as np

from . import _nc
from .io import get_atts_from_param


def create_file(ncfile, param, variables):
    with _nc.Dataset(ncfile, "w", format="NETCDF4") as nc:
        nc.setncatts(get_atts_from_param(param))
        nc.createDimension("t", None)
        for name in variables:
            v = nc.createVariable(name, "f", ("t",))
            v.standard_name = name


def get_number_of_records(ncfile):
    return _nc.number_of_records(ncfile)


class Bulk:
    on_device = True          # Model.run(): no state download on its account

    def __init__(self, model, ncfile="bulk.nc"):
        self.model = model
        msk = model.mesh.msk

        def domsum(x):
            return np.sum(x, axis=None)

        self.domavg = lambda x: domsum(x) / domsum(msk)
        variables = ("time", "ke", "ens", "vort", "angular")
        self.Bulk = namedtuple("bulk", variables)
        self.ndiags = 1_000
        self.data = self.Bulk(*[np.zeros((self.ndiags,)) for _ in variables])
        xv, _ = model.mesh.xy("y")
        _, yu = model.mesh.xy("x")
        self.xvyu = (xv, yu)
        self.kt = 0
        self.ncfile = ncfile
        # slab mode: the device sums are all-reduced, every rank holds the same
        # averages -- only rank 0 owns bulk.nc
        slab = getattr(model.mesh, "slab", None)
        self.nranks = slab.nranks if slab is not None else 1
        self.writer = slab is None or slab.rank == 0
        self.k0 = get_number_of_records(self.ncfile) if self.writer else 0
        if self.writer and self.k0 == 0:
            self.create_newfile()

    def create_newfile(self):
        create_file(self.ncfile, self.model.param, self.data._fields)
        self.k0 = 0

    def averages(self):
        """(ke, ens, vort, angular) of the current state (diagnostics.py:52-56)"""
        model = self.model
        mesh = model.mesh
        dx, dy, area = mesh.dx, mesh.dy, mesh.area
        resident = getattr(model, "_resident", False)
        if self.nranks > 1 and not resident:
            # a slab's host arrays carry ghost rows and a local mask sum: the global
            # averages come from the owned rows of every rank (f2d_bulk_sums all-reduces)
            model.integrator.upload(model.state, ["ke", "omega", "U.x", "U.y"])
            resident = True
        if resident:
            ske, som2, som, suyx, suxy, smsk = mesh.engine.bulk_sums()
            return (ske / smsk, 0.5 * (som2 / smsk) / area ** 2, (som / smsk) / area,
                    (suyx / smsk) * dx - (suxy / smsk) * dy)
        s = model.state
        xv, yu = self.xvyu
        return (self.domavg(s.ke), 0.5 * self.domavg(s.omega ** 2) / area ** 2, self.domavg(s.omega) / area,
                self.domavg(s.U.y * xv) * dx - self.domavg(s.U.x * yu) * dy)

    def __call__(self):
        if self.model.time.ite % 3 > 0:
            return
        kt = self.kt
        time, ke, ens, vort, angular = self.data
        ke[kt], ens[kt], vort[kt], angular[kt] = self.averages()
        time[kt] = self.model.time.t
        self.kt += 1
        if self.kt == self.ndiags:
            self.write()

    def finalize(self):
        self.write()

    def write(self):
        n = self.kt
        idx = slice(self.k0, self.k0 + n)
        if self.writer:
            with _nc.Dataset(self.ncfile, "r+") as nc:
                for name, x in zip(self.data._fields, self.data):
                    nc.variables[name][idx] = x[:n]
        self.k0 += n
        self.kt = 0

    def read(self):
        with _nc.Dataset(self.ncfile, "r") as nc:
            return self.Bulk(*[np.array(nc.variables[name][:]) for name in self.Bulk._fields])
