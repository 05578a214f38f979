"""Same public names as reference keras/models/ranking/__init__.py:4-6."""
from .fm import FM  # noqa: F401
from .fm import FactorizationMachine  # noqa: F401
from .deepfm import DeepFM  # noqa: F401
