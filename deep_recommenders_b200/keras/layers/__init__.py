"""`deep_recommenders.keras.layers`: BASELINE.json's north_star names this package; the
reference does not have it.  It is provided here as a re-export
of the hot-path layers so either import path works."""
from .base import Layer, Model, Dense, Sequential, register_keras_serializable  # noqa: F401


def __getattr__(name):
    # lazy re-exports (avoid import cycles with keras.models.*)
    if name in ("FM", "FactorizationMachine", "DeepFM"):
        from ..models import ranking
        return getattr(ranking, name)
    if name in ("Cross", "DCN"):
        from ..models.ranking import dcn
        return getattr(dcn, name)
    if name == "DNN":
        from ..models.ranking.deepfm import DNN
        return DNN
    if name in ("Retrieval", "HardNegativeMining", "RemoveAccidentalNegative",
                "SamplingProbabilityCorrection", "TwoTower"):
        from ..models.retrieval import sbcnm
        return getattr(sbcnm, name)
    raise AttributeError(name)
