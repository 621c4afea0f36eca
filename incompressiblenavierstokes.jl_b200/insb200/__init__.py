"""insb200 -- B200-native backend for the per-time-step hot path of
IncompressibleNavierStokes.jl, host side.  See ``api`` for the mirrored interface."""
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, Insb200Error, lib  # noqa: F401
from .grid import (  # noqa: F401
    DirichletBC,
    PeriodicBC,
    PressureBC,
    Setup,
    SymmetricBC,
    cosine_grid,
    stretched_grid,
    tanh_grid,
)


def __getattr__(name):
    # the API layer needs torch; import it lazily so that `import insb200` (symbol checks,
    # grid set-up) works without touching CUDA
    import importlib

    if name == "api":
        return importlib.import_module(".api", __name__)
    api = importlib.import_module(".api", __name__)
    try:
        return getattr(api, name)
    except AttributeError:
        raise AttributeError(f"module 'insb200' has no attribute {name!r}") from None
