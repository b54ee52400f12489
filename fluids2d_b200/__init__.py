"""fluids2d_b200: the Fluids2d time-step hot path (WENO advection + RK stages +
masked elliptic solve) on NVIDIA B200, behind the reference's Python API.

    import fluids2d_b200 as f2d
    param = f2d.Param(); ...; model = f2d.Model(param); model.run()

``install_as_fluids2d()`` registers this package under the name ``fluids2d`` so
that unmodified experiment scripts (``import fluids2d as f2d``) pick it up.
"""
from .param import Param
from .model import Model
from . import tools

__version__ = "0.1.0"


def install_as_fluids2d():
    import importlib
    import sys
    sys.modules["fluids2d"] = sys.modules[__name__]
    for sub in ("param", "model", "tools", "meshes", "states", "integrators", "equations",
                "operators", "weno", "elliptic", "timeline", "io", "diagnostics"):
        sys.modules[f"fluids2d.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
