"""`import tensorflow.compat.v1 as tf` (estimator/models/feature_interaction/fm.py:6-7)."""
from tensorflow import *  # noqa: F401,F403
from tensorflow import (__version__, feature_column, layers, nn, keras, math, variable_scope,  # noqa: F401
                        disable_eager_execution, square, reduce_sum, subtract, stack, concat, float32)
