"""``from mpopt_b200 import mp`` -- drop-in for the reference's ``from mpopt import mp`` on the hot path."""
from . import ca  # noqa: F401
from .collocation import Collocation, CollocationRoots  # noqa: F401
from .adaptive import mpopt_adaptive, mpopt_h_adaptive  # noqa: F401
from .mpopt import mpopt, post_process, solve  # noqa: F401
from .ocp import OCP  # noqa: F401
