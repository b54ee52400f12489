"""NetCDF access for history / bulk files: netCDF4 when it is installed (what
the reference uses, io.py:1), otherwise scipy's NetCDF-3 writer -- same
dimensions, variable names and float32 storage either way."""
import numpy as np

try:
    from netCDF4 import Dataset as _Dataset4
except Exception:                      # pragma: no cover - depends on the image
    _Dataset4 = None


class _Scipy:
    """the handful of netCDF4.Dataset calls io.py / diagnostics.py make"""

    def __init__(self, path, mode):
        from scipy.io import netcdf_file
        self._f = netcdf_file(path, {"r+": "a"}.get(mode, mode), version=2, mmap=False)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._f.close()

    def setncatts(self, atts):
        for k, v in atts.items():
            if isinstance(v, (list, tuple)):
                v = ",".join(str(x) for x in v)
            if isinstance(v, (bool, np.bool_)):
                v = int(v)
            if isinstance(v, float):
                v = np.float64(v)
            try:
                setattr(self._f, k, v)
            except Exception:
                setattr(self._f, k, str(v))

    def createDimension(self, name, size):
        self._f.createDimension(name, size)

    def createVariable(self, name, dtype, dims):
        return self._f.createVariable(name, {"i4": "i"}.get(dtype, dtype), dims)

    @property
    def variables(self):
        return self._f.variables

    @property
    def dimensions(self):
        return self._f.dimensions

    def nrecords(self, dim="t"):
        recs = [v.shape[0] for v in self._f.variables.values() if v.isrec]
        return max(recs) if recs else 0


def Dataset(path, mode="r", format="NETCDF4"):
    if _Dataset4 is not None:
        return _Dataset4(path, mode, format=format) if mode == "w" else _Dataset4(path, mode)
    return _Scipy(path, mode)


def number_of_records(path, dim="t"):
    import os
    if not os.path.isfile(path):
        return 0
    with Dataset(path, "r") as nc:
        if _Dataset4 is not None:
            return len(nc.dimensions[dim])
        return nc.nrecords(dim)
