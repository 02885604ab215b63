"""mpopt_b200 -- B200-native collocation-transcription hot path behind mpopt's ``mp`` surface.

    from mpopt_b200 import mp, ca
    ocp = mp.OCP(n_states=2, n_controls=1)
    ...
    mpo, post = mp.solve(ocp, n_segments=20, poly_orders=3, scheme="LGR")
"""
from . import ca  # noqa: F401
from .ocp import OCP  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need the CUDA library
    if name == "mp":
        import importlib

        return importlib.import_module(".mp", __name__)
    raise AttributeError(name)
