"""ctypes binding of libf2d.so (include/f2d.h).  No torch, no numpy magic:
plain pointers and sizes.  Importing this module does not need a GPU; creating
an :class:`Engine` does, and fails loudly if the library or the device is
missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

MODELS = {"euler": 0, "boussinesq": 1, "rsw": 2, "qgrsw": 3, "eulerpsi": 4, "qg": 5, "advection": 6,
          "vectoradv": 7}
METHODS = {"weno": 0, "upwind": 1, "centered": 2, "cweno": 3, "classic": 4}
INTEGRATORS = {"rk3": 0, "ef": 1, "enrk3": 2, "LFRA": 3}
SOLVERS = {"c": 0, "v": 1, "h": 2}
NOSLIP = {"left": 1, "right": 2, "bottom": 4, "top": 8}
NOSLIP_ALL = 16
MESH_ARRAYS = ("msk", "mskx", "msky", "mskv", "slip", "oc.x", "oc.y", "ov.x", "ov.y", "ok.x", "ok.y")


class F2DError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libf2d status {status}: {msg}")
        self.status = status


class NotConverged(F2DError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nh", C.c_int32),
        ("Lx", C.c_double), ("Ly", C.c_double),
        ("xperiodic", C.c_int32), ("yperiodic", C.c_int32), ("noslip", C.c_int32),
        ("f0", C.c_double), ("g", C.c_double), ("H", C.c_double),
        ("integrator", C.c_int32), ("compflux", C.c_int32), ("vortexforce", C.c_int32),
        ("innerproduct", C.c_int32), ("maxorder", C.c_int32), ("device", C.c_int32),
        ("solver_rtol", C.c_double), ("solver_maxit", C.c_int32), ("solver_kind", C.c_int32),
        ("nu1", C.c_int32), ("nu2", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


# name -> (restype, argtypes); every symbol include/f2d.h declares
_P, _I, _I64, _D, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
SIGNATURES = {
    "f2d_version": (_I, []),
    "f2d_last_error": (C.c_char_p, []),
    "f2d_device_count": (_I, [C.POINTER(_I)]),
    "f2d_create": (_I, [C.POINTER(Config), C.POINTER(_P)]),
    "f2d_destroy": (_I, [_P]),
    "f2d_set_stream": (_I, [_P, _P]),
    "f2d_sync": (_I, [_P]),
    "f2d_set_mask": (_I, [_P, _P]),
    "f2d_get_mesh_array": (_I, [_P, C.c_char_p, _P]),
    "f2d_set_topography": (_I, [_P, _P]),
    "f2d_upload": (_I, [_P, C.c_char_p, _P]),
    "f2d_download": (_I, [_P, C.c_char_p, _P]),
    "f2d_field_ptr": (_I, [_P, C.c_char_p, C.POINTER(_P)]),
    "f2d_download_f32": (_I, [_P, C.c_char_p, _P]),
    "f2d_io_sync": (_I, [_P]),
    "f2d_set_forcing": (_I, [_P, C.c_char_p, _P, _D]),
    "f2d_bulk_sums": (_I, [_P, _I, C.POINTER(_D)]),
    "f2d_step": (_I, [_P, _D, _I]),
    "f2d_step_lfra": (_I, [_P, _D, _I, _D]),
    "f2d_rhs": (_I, [_P, _I]),
    "f2d_addto": (_I, [_P, _I, C.POINTER(_D)]),
    "f2d_diag": (_I, [_P]),
    "f2d_max_abs_U": (_I, [_P, C.POINTER(_D)]),
    "f2d_solve": (_I, [_P, _I, _P, _D, _P, C.POINTER(_I), C.POINTER(_D)]),
    "f2d_apply_laplacian": (_I, [_P, _I, _P, _P]),
    "f2d_solver_stats": (_I, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_D)]),
    "f2d_solver_info": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_D)]),
    "f2d_compflux": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I]),
    "f2d_vortexforce": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _I]),
    "f2d_innerproduct": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I]),
    "f2d_fill": (_I, [_P, _P]),
    "f2d_malloc": (_I, [_P, _SZ, C.POINTER(_P)]),
    "f2d_free": (_I, [_P, _P]),
    "f2d_memcpy_h2d": (_I, [_P, _P, _P, _SZ]),
    "f2d_memcpy_d2h": (_I, [_P, _P, _P, _SZ]),
    "f2d_host_alloc": (_I, [_SZ, C.POINTER(_P)]),
    "f2d_host_free": (_I, [_P]),
    "f2d_timer_start": (_I, [_P]),
    "f2d_timer_stop": (_I, [_P, C.POINTER(C.c_float)]),
    "f2d_bench_kernel": (_I, [_P, C.c_char_p, _I, C.POINTER(C.c_float), C.POINTER(_D)]),
    "f2d_dist_unique_id": (_I, [C.c_char_p]),
    "f2d_dist_init": (_I, [_P, _I, _I, C.c_char_p]),
    "f2d_dist_exchange": (_I, [_P, C.c_char_p]),
    "f2d_exchange_count": (_I, [_P, C.POINTER(_I64)]),
    "f2d_launch_count": (_I, [_P, C.POINTER(_I64)]),
}

_libs = {}


def load(exact=False):
    """dlopen libf2d.so (or libf2d_exact.so); raises if it is not built."""
    key = bool(exact)
    if key not in _libs:
        path = os.environ.get("F2D_LIB_PATH") or _build.lib_path(exact)     # F2D_LIB_PATH: a development build to A/B
        if not os.path.exists(path):
            raise F2DError(-1, f"{path} is not built: run `python -m fluids2d_b200.build` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError = missing symbol
            fn.restype, fn.argtypes = res, args
        _libs[key] = lib
    return _libs[key]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _noslip_flags(noslip):
    if noslip is None or noslip is False:
        return 0
    if noslip is True:
        return NOSLIP_ALL
    f = 0
    for k, v in NOSLIP.items():
        if k in noslip:
            f |= v
    return f


def config_from_param(param, device=0, solver_rtol=0.0, solver_maxit=0, solver_kind=0, nu1=0, nu2=0,
                      guess_order=None, slab=None):
    """param.py:13-59 attributes -> f2d_config."""
    if param.model not in MODELS:
        raise NotImplementedError(
            f"model '{param.model}' is not on the device path ({', '.join(MODELS)})")
    if param.integrator not in INTEGRATORS:
        raise NotImplementedError(f"integrator '{param.integrator}' is not on the device path")
    cfg = Config()
    cfg.reserved[5] = int(getattr(param, "tracer", None) not in (None, "None"))
    cfg.model = MODELS[param.model]
    cfg.nx, cfg.ny, cfg.nh = param.nx, param.ny, param.halowidth
    if slab is not None and slab.nranks > 1:
        # the context sees the local slab; dy stays Ly / (global ny)
        cfg.ny = slab.ny_ctx
        cfg.reserved[1], cfg.reserved[2], cfg.reserved[3] = slab.gs, slab.gn, slab.ny
    cfg.Lx, cfg.Ly = param.Lx, param.Ly
    cfg.xperiodic = int(bool(param.xperiodic))
    cfg.yperiodic = 2 if getattr(param, "ywrap", False) else int(bool(param.yperiodic))     # 2: true wrap (f2d.h)
    cfg.noslip = _noslip_flags(param.noslip)
    if slab is not None and slab.nranks > 1 and not (cfg.noslip & NOSLIP_ALL):
        if slab.rank > 0:
            cfg.noslip &= ~NOSLIP["bottom"]
        if slab.rank < slab.nranks - 1:
            cfg.noslip &= ~NOSLIP["top"]
    cfg.f0, cfg.g, cfg.H = param.f0, param.g, param.H
    cfg.reserved[4] = int(getattr(param, "beta", 0.0) != 0)
    cfg.integrator = INTEGRATORS[param.integrator]
    cfg.compflux = METHODS[param.compflux]
    cfg.vortexforce = METHODS[param.vortexforce]
    cfg.innerproduct = METHODS[param.innerproduct]
    cfg.maxorder = param.maxorder
    cfg.device = device
    cfg.solver_rtol, cfg.solver_maxit, cfg.solver_kind = solver_rtol, solver_maxit, solver_kind
    cfg.nu1, cfg.nu2 = nu1, nu2
    if guess_order is not None:
        cfg.reserved[0] = int(guess_order) + 1
    return cfg


class Engine:
    """One f2d_ctx: device-resident mesh + state + solvers for one Model."""

    def __init__(self, param, device=0, exact=False, slab=None, comm=None, **solver_kw):
        self.lib = load(exact)
        self.slab = slab if (slab is not None and slab.nranks > 1) else None
        self.cfg = config_from_param(param, device=device, slab=self.slab, **solver_kw)
        self.shape = (self.cfg.ny + 2 * param.halowidth, param.nx + 2 * param.halowidth)
        self.size = self.shape[0] * self.shape[1]
        self._h = C.c_void_p()
        # the device calls the extra scalar "tracer" whatever param.tracer names it
        t = getattr(param, "tracer", None)
        self._tracer = t if t not in (None, "None") else None
        self._chk(self.lib.f2d_create(C.byref(self.cfg), C.byref(self._h)))
        self._dev_allocs = []
        if self.slab is not None:
            if comm is None:
                raise F2DError(-2, "a slab engine needs the NCCL identity (slabs.set_communicator)")
            rank, world, uid = comm
            _preload_nccl()
            assert (rank, world) == (self.slab.rank, self.slab.nranks)
            self._chk(self.lib.f2d_dist_init(self._h, rank, world, uid))

    def _n(self, name):
        """host leaf name -> device field name (bytes)"""
        if self._tracer is not None:
            if name == self._tracer:
                name = "tracer"
            elif name.startswith("ds") and name.endswith("." + self._tracer) and name[2:3].isdigit():
                name = name[:name.index(".") + 1] + "tracer"
        return name.encode()

    # -- errors -------------------------------------------------------------
    def _chk(self, status):
        if status != 0:
            msg = self.lib.f2d_last_error().decode()
            raise (NotConverged if status == -4 else F2DError)(status, msg)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            for p in self._dev_allocs:
                self.lib.f2d_free(self._h, p)
            self.lib.f2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mesh ---------------------------------------------------------------
    def set_mask(self, msk=None):
        if msk is None:
            self._chk(self.lib.f2d_set_mask(self._h, None))
        else:
            m = np.ascontiguousarray(msk, dtype=np.int8)
            assert m.shape == self.shape
            self._chk(self.lib.f2d_set_mask(self._h, _ptr(m)))

    def mesh_array(self, name):
        out = np.empty(self.shape, dtype=np.int8)
        self._chk(self.lib.f2d_get_mesh_array(self._h, name.encode(), _ptr(out)))
        return out

    def set_topography(self, hb):
        if hb is None or (np.isscalar(hb) and not hb):
            self._chk(self.lib.f2d_set_topography(self._h, None))
        else:
            if np.isscalar(hb):            # a flat bottom at a non-zero level (qg_inversion uses hb * qgcoef)
                hb = np.full(self.shape, float(hb))
            a = np.ascontiguousarray(hb, dtype=np.float64)
            assert a.shape == self.shape
            self._chk(self.lib.f2d_set_topography(self._h, _ptr(a)))

    # -- state --------------------------------------------------------------
    def upload(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.shape, (name, a.shape, self.shape)
        self._chk(self.lib.f2d_upload(self._h, self._n(name), _ptr(a)))
        if self.slab is not None and not name.startswith("ds"):
            self._chk(self.lib.f2d_dist_exchange(self._h, self._n(name)))
        self.sync()     # `a` may be a temporary

    def upload_async(self, name, a):
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == self.shape
        self._chk(self.lib.f2d_upload(self._h, self._n(name), _ptr(a)))
        if self.slab is not None and not name.startswith("ds"):
            # host ghost rows may be stale: take them from their owners
            self._chk(self.lib.f2d_dist_exchange(self._h, self._n(name)))

    def download(self, name, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == self.shape
        self._chk(self.lib.f2d_download(self._h, self._n(name), _ptr(out)))
        self.sync()
        return out

    def download_async(self, name, out):
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == self.shape
        self._chk(self.lib.f2d_download(self._h, self._n(name), _ptr(out)))

    def field_ptr(self, name):
        p = C.c_void_p()
        self._chk(self.lib.f2d_field_ptr(self._h, self._n(name), C.byref(p)))
        return p.value

    def download_f32_async(self, name, out):
        """float32 copy of a device field into `out` (pinned float32) on the copy
        stream; io_sync() before reading it"""
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.size
        self._chk(self.lib.f2d_download_f32(self._h, self._n(name), _ptr(out)))

    def set_forcing(self, leaf, pattern, amplitude=1.0):
        """ds.<leaf> += amplitude * pattern in every stage; pattern None changes
        only the amplitude"""
        if pattern is not None:
            pattern = np.ascontiguousarray(pattern, dtype=np.float64)
            assert pattern.size == self.size
        self._chk(self.lib.f2d_set_forcing(self._h, self._n(leaf), None if pattern is None else _ptr(pattern),
                                           float(amplitude)))

    def io_sync(self):
        self._chk(self.lib.f2d_io_sync(self._h))

    def bulk_sums(self):
        """[sum ke, sum omega^2, sum omega, sum U.y*xv, sum U.x*yu, sum msk] of the
        device state (diagnostics.py:39-62)"""
        out = (_D * 6)()
        row0 = self.slab.row0 if self.slab is not None else 0
        self._chk(self.lib.f2d_bulk_sums(self._h, int(row0), out))
        return np.array(out[:])

    def sync(self):
        self._chk(self.lib.f2d_sync(self._h))

    def set_stream(self, stream_handle):
        self._chk(self.lib.f2d_set_stream(self._h, C.c_void_p(stream_handle or 0)))

    # -- stepping -----------------------------------------------------------
    def step(self, dt, nsteps=1):
        self._chk(self.lib.f2d_step(self._h, float(dt), int(nsteps)))

    def step_lfra(self, dt, first, gamma):
        self._chk(self.lib.f2d_step_lfra(self._h, float(dt), int(bool(first)), float(gamma)))

    def rhs(self, k):
        self._chk(self.lib.f2d_rhs(self._h, int(k)))

    def addto(self, coefs):
        arr = (C.c_double * len(coefs))(*[float(x) for x in coefs])
        self._chk(self.lib.f2d_addto(self._h, len(coefs), arr))

    def diag(self):
        self._chk(self.lib.f2d_diag(self._h))

    def max_abs_U(self):
        v = C.c_double()
        self._chk(self.lib.f2d_max_abs_U(self._h, C.byref(v)))
        return v.value

    # -- elliptic -----------------------------------------------------------
    def solve_dev(self, which, d_b, bscale, d_x):
        it, rr = C.c_int(), C.c_double()
        self._chk(self.lib.f2d_solve(self._h, SOLVERS[which], C.c_void_p(d_b), float(bscale),
                                     C.c_void_p(d_x), C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def solve(self, which, b, x, bscale=1.0):
        """Poisson2D.solve(b, x) on host arrays: x is updated in place."""
        db, dx = self.to_device(b), self.to_device(x)
        try:
            res = self.solve_dev(which, db, bscale, dx)
        finally:
            self.from_device(dx, x)
            self.free(db)
            self.free(dx)
        return res

    def apply_laplacian(self, which, x):
        dx, dy = self.to_device(x), self.to_device(np.zeros(self.shape))
        self._chk(self.lib.f2d_apply_laplacian(self._h, SOLVERS[which], C.c_void_p(dx), C.c_void_p(dy)))
        y = np.empty(self.shape)
        self.from_device(dy, y)
        self.free(dx)
        self.free(dy)
        return y

    def solver_stats(self):
        a, b, r = C.c_int64(), C.c_int64(), C.c_double()
        self._chk(self.lib.f2d_solver_stats(self._h, C.byref(a), C.byref(b), C.byref(r)))
        return dict(nsolves=a.value, niters=b.value, max_relres=r.value)

    def solver_info(self, which="c"):
        """connected components / multigrid levels of a solver, worst right-hand-side incompatibility so far"""
        nc, nl, inc = C.c_int(), C.c_int(), C.c_double()
        self._chk(self.lib.f2d_solver_info(self._h, SOLVERS[which], C.byref(nc), C.byref(nl), C.byref(inc)))
        return dict(components=nc.value, levels=nl.value, rhs_incompat=inc.value)

    # -- raw device memory (per-kernel tests, solves on host arrays) ----------
    def malloc(self, nbytes):
        p = C.c_void_p()
        self._chk(self.lib.f2d_malloc(self._h, nbytes, C.byref(p)))
        self._dev_allocs.append(p.value)
        return p.value

    def free(self, p):
        self._dev_allocs.remove(p)
        self._chk(self.lib.f2d_free(self._h, C.c_void_p(p)))

    def to_device(self, a):
        a = np.ascontiguousarray(a)
        p = self.malloc(a.nbytes)
        self._chk(self.lib.f2d_memcpy_h2d(self._h, C.c_void_p(p), _ptr(a), a.nbytes))
        self.sync()
        return p

    def from_device(self, p, out):
        assert out.flags.c_contiguous
        self._chk(self.lib.f2d_memcpy_d2h(self._h, _ptr(out), C.c_void_p(p), out.nbytes))
        return out

    # -- the three weno.py kernels on host arrays -----------------------------
    def _kernel(self, fn, first, others, o, tail):
        d_first = self.to_device(first)
        d_others = [self.to_device(a) for a in others]
        d_o = self.to_device(np.ascontiguousarray(o, dtype=np.int8))
        try:
            self._chk(fn(self._h, C.c_void_p(d_first), *[C.c_void_p(p) for p in d_others],
                         C.c_void_p(d_o), first.size, *tail))
            self.from_device(d_first, first)
        finally:
            for p in [d_first, d_o] + d_others:
                self.free(p)

    def compflux(self, flx, U, q, o, s, method):
        self._kernel(self.lib.f2d_compflux, flx, [U, q], o, (int(s), METHODS[method]))

    def vortexforce(self, du, V, omega, o, s, s2, sign, method):
        self._kernel(self.lib.f2d_vortexforce, du, [V, omega], o,
                     (int(s), int(s2), int(sign), METHODS[method]))

    def innerproduct(self, ke, U, u, o, s, method):
        self._kernel(self.lib.f2d_innerproduct, ke, [U, u], o, (int(s), METHODS[method]))

    def fill(self, a):
        d = self.to_device(a)
        self._chk(self.lib.f2d_fill(self._h, C.c_void_p(d)))
        self.from_device(d, a)
        self.free(d)

    # -- timing ---------------------------------------------------------------
    def timer_start(self):
        self._chk(self.lib.f2d_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._chk(self.lib.f2d_timer_stop(self._h, C.byref(ms)))
        return ms.value

    _MG_KERNELS = ("mg.down0", "mg.up0", "mg.down1", "mg.up1", "mg.tail", "cg.dir_apply", "cg.update")
    BENCH_KERNELS = {
        0: ("advection", "rk_update", "divergence", "project_diag") + _MG_KERNELS,                 # euler
        1: ("advection", "flux_div", "rk_update", "divergence", "project_diag") + _MG_KERNELS,     # boussinesq
        2: ("advection", "flux_div", "rk_update", "diag"),                                         # rsw: no solve in the step
        3: ("advection", "flux_div", "qg_pv", "qg_back", "rk_update", "diag") + _MG_KERNELS,       # qgrsw
    }

    def bench_kernel_names(self):
        """kernels f2d_bench_kernel can time alone for this model"""
        return list(self.BENCH_KERNELS.get(self.cfg.model, ()))

    def launches_per_step(self, name, iters_per_solve):
        """how often one RK3 step launches a benchmarked kernel (3 stages, one solve each)"""
        if name.startswith(("mg.", "cg.")):
            return 3.0 * iters_per_solve
        if name == "rk_update":
            # the projecting models fuse the velocity update into the tendency kernel
            # rsw and the Boussinesq buoyancy: fused as well (step.cu: fused_stage_rsw, launch_transport)
            return {0: 0, 1: 0, 2: 0, 3: 9}[self.cfg.model]
        return 3.0

    def bench_kernel(self, name, reps=20):
        ms, nbytes = C.c_float(), C.c_double()
        self._chk(self.lib.f2d_bench_kernel(self._h, name.encode(), int(reps), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def launch_count(self):
        n = C.c_int64()
        self._chk(self.lib.f2d_launch_count(self._h, C.byref(n)))
        return n.value

    def exchange_count(self):
        n = C.c_int64()
        self._chk(self.lib.f2d_exchange_count(self._h, C.byref(n)))
        return n.value


def _preload_nccl():
    """make the newest NCCL in the environment (torch's bundled copy) the one
    this process uses, before libf2d dlopens "libnccl.so.2" by soname"""
    import importlib.util
    import glob
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            for path in glob.glob(os.path.join(root, "lib", "libnccl.so*")):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return path
    except Exception:
        pass
    return None


def nccl_unique_id():
    """128-byte NCCL unique id (call on one rank, hand it to the others)"""
    _preload_nccl()
    lib = load()
    buf = C.create_string_buffer(128)
    st = lib.f2d_dist_unique_id(buf)
    if st != 0:
        raise F2DError(st, lib.f2d_last_error().decode())
    return buf.raw


class _PinnedOwner:
    """frees a cudaMallocHost block when the last numpy view of it is gone"""

    def __init__(self, lib, p):
        self.lib, self.p = lib, p

    def __del__(self):
        try:
            self.lib.f2d_host_free(self.p)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64, exact=False):
    """numpy array backed by page-locked host memory (cudaMallocHost).  The block
    belongs to the ctypes buffer the array is a view of: it is released
    (f2d_host_free) when the array and every view derived from it are gone."""
    lib = load(exact)
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    st = lib.f2d_host_alloc(max(n, 1), C.byref(p))
    if st != 0:
        raise F2DError(st, lib.f2d_last_error().decode())
    buf = (C.c_char * n).from_address(p.value)
    buf._owner = _PinnedOwner(lib, p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
