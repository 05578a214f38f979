"""Same public names as reference estimator/models/feature_interaction/__init__.py:4-6."""
from .fm import fm  # noqa: F401
from .fm import FM  # noqa: F401
from .dnn import dnn  # noqa: F401
