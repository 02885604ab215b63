"""mpopt_b200 -- B200-native collocation-transcription hot path behind mpopt's ``mp`` surface."""
from . import ca  # noqa: F401
from .ocp import OCP  # noqa: F401
