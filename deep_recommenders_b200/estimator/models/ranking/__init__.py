from .deepfm import DeepFM  # noqa: F401
