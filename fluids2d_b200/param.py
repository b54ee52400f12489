"""Param: the attribute bag of the reference (src/fluids2d/param.py:10-62), same
names, defaults and checks, plus a few device-side knobs (marked NEW)."""

_models = ["euler", "eulerpsi", "advection", "vectoradv",
           "boussinesq", "hydrostatic", "rsw", "qgrsw", "qg"]
_methods = ["weno", "upwind", "centered", "cweno"]
_methods_extended = _methods + ["classic"]
_integrators = ["rk3", "ef", "enrk3", "LFRA"]

# models whose time step runs on the device in this build
DEVICE_MODELS = ["euler", "boussinesq", "rsw", "qgrsw", "eulerpsi", "qg", "advection", "vectoradv"]


# name -> default, grouped as in the reference's constructor (param.py:13-59)
_REFERENCE_DEFAULTS = dict(
    # which equations, on which grid
    model="euler", nx=40, ny=40, Lx=1.0, Ly=1.0,
    xperiodic=False, yperiodic=False, halowidth=3, noslip=None,
    # physical constants of the rotating / stratified models
    f0=10.0, beta=0.0, g=1, H=1,
    # length of the run and step control
    tend=1.0, dt=0.0, maxite=100, dtmax=9e99, integrator="rk3", cfl=0.9, RAgamma=0.1,
    # observation points: print, plot, history file
    nprint=1, nplot=5, animation=False, generate_mp4=False, plotvar=None, clims=None,
    cmap="RdBu_r", outputfile="history.nc", var_to_store=[], nhis=0,
    # numerics of the three advective operators
    compflux="weno", vortexforce="weno", innerproduct="weno", maxorder=6,
    tracer=None, nthreads=1,
)

# NEW: device / elliptic-solver controls (the reference has a direct solve)
_DEVICE_DEFAULTS = dict(
    device=0,
    ywrap=False,            # a TRUE periodic y direction: halo rows are images, the Laplacian wraps.  (The
                            # reference's yperiodic only puts 1 in the mask of the halo rows -- meshes.py:74,
                            # :135-143, elliptic.py:142 -- and is reproduced as is.)  With xperiodic: doubly periodic.
    rank=0, nranks=1,       # y-slab decomposition: this process's slab out of nranks (slabs.py)
    solver="pcg",           # "pcg": multigrid-preconditioned CG; "mg": plain V-cycles
    solver_rtol=1e-12,      # ||b - A x|| <= rtol ||b||
    solver_maxit=100,
    solver_nu=2,            # red-black sweeps before and after each coarse correction
    solver_guess=4,         # first guess from the same RK stage of earlier steps: 0 off,
                            # 1 previous, 2 linear, 3 quadratic, 4 cubic extrapolation (max 6)
)


class Param:
    _quiet = False      # set Param._quiet = True to silence the help banner

    def __init__(self):
        for group in (_REFERENCE_DEFAULTS, _DEVICE_DEFAULTS):
            for name, value in group.items():
                setattr(self, name, list(value) if isinstance(value, list) else value)
        self.__parameters__ = _public_names(self)
        self.help()

    def add_parameter(self, name):
        """register a user parameter (param.py:64-66)"""
        setattr(self, name, None)
        self.__parameters__ = _public_names(self)

    def check_parameters_are_known(self):
        extra = set(_public_names(self)) - set(self.__parameters__)
        assert not extra, (f"parameter {extra} is unknown\n"
                           f"parameters are {self.__parameters__}")
        return True

    def check(self):
        assert self.model in _models
        assert self.compflux in _methods
        assert self.vortexforce in _methods
        assert self.innerproduct in _methods_extended
        assert self.integrator in _integrators
        assert self.solver in ("pcg", "mg")
        self.check_parameters_are_known()

    def help(self):
        if Param._quiet:
            return
        b = lambda s: "\033[1m\033[94m" + s + "\033[0m"
        print("\n".join([
            "Valid values for string parameters",
            f"  - {b('model')}: " + ", ".join(_models),
            f"  - {b('integrator')}: " + ", ".join(_integrators),
            f"  - {b('compflux')} (U*q): " + ", ".join(_methods),
            f"  - {b('vortexforce')} (omega x U): " + ", ".join(_methods),
            f"  - {b('innerproduct')} (U.u): " + ", ".join(_methods_extended),
            f"  - on the B200 path: " + ", ".join(DEVICE_MODELS) + " with rk3 / ef / enrk3",
            ""]))


def _public_names(obj):
    return [d for d in obj.__dir__() if "__" not in d]
