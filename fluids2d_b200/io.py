"""History output (reference: src/fluids2d/io.py:5-93): same file layout --
dimensions x, y, t; xc, yc; t, ite, dt; one float32 (t, y, x) variable per entry
of ``param.var_to_store`` (vectors as <name>x / <name>y).

What differs is how the data reaches the host.  While ``Model.run()`` keeps the
state on the device, ``write`` does not download float64 fields and stall the
step loop: each stored field is converted to float32 ON the device
(``f2d_download_f32``), copied into pinned host memory on a separate copy
stream, and a writer thread waits for the copy and appends the record while the
GPU goes on stepping.  Outside ``run()`` it writes the host arrays it is given,
like the reference."""
import threading

import numpy as np

from . import _nc
from .states import vectors


def get_atts_from_param(param):
    """io.py:40-50"""
    atts = {}
    for k, v in param.__dict__.items():
        if k.startswith("_"):
            continue
        if isinstance(v, bool):
            v = int(v)
        atts[k] = "None" if v is None else v
    return atts


def _leaf_names(varname):
    return [(varname + "x", varname + ".x"), (varname + "y", varname + ".y")] if varname in vectors \
        else [(varname, varname)]


def get_data_from_state(state, varname):
    v = getattr(state, varname)
    return (v.x, v.y) if varname in vectors else v


def history_path(param, mesh):
    """param.outputfile; one file per slab (<root>_<rank>.nc) when the mesh is
    decomposed over several GPUs"""
    slab = getattr(mesh, "slab", None)
    if slab is not None and slab.nranks > 1:
        import os
        root, ext = os.path.splitext(param.outputfile)
        return f"{root}_{slab.rank:02d}{ext}"
    return param.outputfile


def create_file(param, mesh, state, time):
    """io.py:53-93"""
    with _nc.Dataset(history_path(param, mesh), "w", format="NETCDF4") as nc:
        nc.setncatts(get_atts_from_param(param))
        for dim, size in (("t", None), ("y", mesh.shape[0]), ("x", mesh.shape[1])):   # NetCDF-3: record dimension first
            nc.createDimension(dim, size)
        for varname in ("xc", "yc"):
            nc.createVariable(varname, "f", ("y", "x"))
        for varname in ("t", "ite", "dt"):
            nc.createVariable(varname, "i4" if varname == "ite" else "f", ("t",))
        for varname in param.var_to_store:
            for ncname, _ in _leaf_names(varname):
                v = nc.createVariable(ncname, "f", ("t", "y", "x"))
                v.standard_name = ncname
        x, y = mesh.xy()
        nc.variables["xc"][:, :] = x
        nc.variables["yc"][:, :] = y


class IO:
    on_device = True          # Model.run(): no full-state download for a history record

    def __init__(self, param, mesh, state, time):
        self.param = param
        self.mesh = mesh
        self.kt = 0
        self.resident = False          # set by Model.run() while the device state is current
        self._thread = None
        self._bufs = None
        self._flip = 0
        if param.nhis > 0:
            create_file(param, mesh, state, time)

    # ---- records ------------------------------------------------------------
    def _append(self, kt, stamp, arrays):
        with _nc.Dataset(history_path(self.param, self.mesh), "r+") as nc:
            nc.variables["t"][kt] = stamp[0]
            nc.variables["ite"][kt] = stamp[1]
            nc.variables["dt"][kt] = stamp[2]
            for ncname, a in arrays:
                nc.variables[ncname][kt, :, :] = a

    def write(self, state, time):
        stamp = (time.t, time.ite, time.dt)
        kt, self.kt = self.kt, self.kt + 1
        if not self.resident:
            self.flush()
            arrays = []
            for varname in self.param.var_to_store:
                v = getattr(state, varname)
                for (ncname, _), a in zip(_leaf_names(varname), (v if varname in vectors else (v,))):
                    arrays.append((ncname, a))
            self._append(kt, stamp, arrays)
            return
        # device path: float32 conversion + copy stream + writer thread, double-buffered
        e = self.mesh.engine
        names = [ln for v in self.param.var_to_store for ln in _leaf_names(v)]
        if self._bufs is None:
            from ._cabi import pinned_empty
            self._bufs = [{nc: pinned_empty(self.mesh.shape, np.float32) for nc, _ in names} for _ in range(2)]
        self.flush()                   # the record two writes back has long left its buffers
        bufs = self._bufs[self._flip]
        self._flip ^= 1
        for ncname, leaf in names:
            e.download_f32_async(leaf, bufs[ncname])

        def work():
            e.io_sync()
            self._append(kt, stamp, [(n, bufs[n]) for n, _ in names])

        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()

    def flush(self):
        """wait for the record in flight (Model.run() calls it when the loop ends)"""
        if self._thread is not None:
            self._thread.join()
            self._thread = None
