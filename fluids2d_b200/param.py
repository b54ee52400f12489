"""Param: the attribute bag of the reference (src/fluids2d/param.py:10-62), same
names, defaults and checks, plus a few device-side knobs (marked NEW)."""

_models = ["euler", "eulerpsi", "advection", "vectoradv",
           "boussinesq", "hydrostatic", "rsw", "qgrsw", "qg"]
_methods = ["weno", "upwind", "centered", "cweno"]
_methods_extended = _methods + ["classic"]
_integrators = ["rk3", "ef", "enrk3", "LFRA"]

# models whose time step runs on the device in this build
DEVICE_MODELS = ["euler", "boussinesq", "rsw", "qgrsw", "eulerpsi", "qg", "advection", "vectoradv"]


class Param:
    _quiet = False      # set Param._quiet = True to silence the help banner

    def __init__(self):
        self.model = "euler"
        self.nx = 40
        self.ny = 40
        self.Lx = 1.0
        self.Ly = 1.0
        self.xperiodic = False
        self.yperiodic = False
        self.halowidth = 3
        self.noslip = None
        self.f0 = 10.0
        self.beta = 0.0
        self.g = 1
        self.H = 1
        self.tend = 1.0
        self.dt = 0.0
        self.maxite = 100
        self.dtmax = 9e99
        self.nprint = 1
        self.nplot = 5
        self.animation = False
        self.generate_mp4 = False
        self.plotvar = None
        self.clims = None
        self.cmap = "RdBu_r"
        self.outputfile = "history.nc"
        self.var_to_store = []
        self.nhis = 0
        self.integrator = "rk3"
        self.cfl = 0.9
        self.RAgamma = 0.1
        self.compflux = "weno"
        self.vortexforce = "weno"
        self.innerproduct = "weno"
        self.maxorder = 6
        self.tracer = None
        self.nthreads = 1
        # NEW: device / elliptic-solver controls (the reference has a direct solve)
        self.device = 0
        self.rank = 0                # y-slab decomposition: this process's slab ...
        self.nranks = 1              # ... out of nranks (one GPU each); see slabs.py
        self.solver = "pcg"          # "pcg": multigrid-preconditioned CG; "mg": plain V-cycles
        self.solver_rtol = 1e-12     # ||b - A x|| <= rtol ||b||
        self.solver_maxit = 100
        self.solver_nu = 2           # red-black sweeps before and after each coarse correction
        self.solver_guess = 4        # first guess from the same RK stage of earlier steps: 0 off,
                                     # 1 previous, 2 linear, 3 quadratic, 4 cubic extrapolation (max 6)
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
