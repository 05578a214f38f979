"""Feature-column *specs* in the argument positions the reference passes tf.feature_column
objects (examples/train_deepfm_on_movielens_keras.py:10-35).  TensorFlow is not available
here, so a column is a small immutable description; the lookup arithmetic lives in the CUDA
gather kernel (deep_recommenders_b200/csrc/embed_fm.cu).

Id pipeline semantics restated from TensorFlow's documented behaviour:
  * hash_bucket:     id = FarmHash Fingerprint64(str(value)) mod hash_bucket_size
                     (deep_recommenders_b200/hashing.py).  Integer inputs that are already
                     ids in [0, N) can bypass hashing with `categorical_column_with_identity`.
  * vocabulary_list: id = index in the list, out-of-vocabulary -> -1 -> zero embedding and
                     all-zero indicator row (default num_oov_buckets=0, default_value=-1).
  * identity:        id = value, out of range -> treated as OOV (-1).
  * embedding_column: dimension D, combiner "mean", initializer truncated-normal(0, 1/sqrt(D)).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple


@dataclass
class RaggedFeature:
    """A multi-valued feature for one batch (what tf.io.VarLenFeature / a SparseTensor carries, e.g. "Genres",
    datasets/movielens.py): flat raw `values` (list / numpy / torch) and `row_splits` [B+1]; example b owns
    values[row_splits[b]:row_splits[b+1]].  Embedding columns mean-combine it, indicator columns count it."""
    values: object
    row_splits: object

    @property
    def batch_size(self) -> int:
        return len(self.row_splits) - 1


@dataclass
class PackedStrings:
    """n byte strings as ONE uint8 buffer + int64 offsets [n+1] (numpy, host): what the native TFRecord parser emits
    for a string feature and what the C-ABI hash / vocabulary entries consume -- no per-value Python objects."""
    data: object
    offsets: object

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def tolist(self):
        d, o = bytes(memoryview(self.data)), self.offsets
        return [d[int(o[i]):int(o[i + 1])] for i in range(len(self))]


@dataclass(frozen=True)
class CategoricalColumn:
    key: str
    kind: str                      # "hash" | "vocab" | "identity"
    num_buckets: int
    vocabulary_list: Tuple = field(default=())
    dtype: str = "string"

    @property
    def name(self) -> str:
        return self.key


def categorical_column_with_hash_bucket(key: str, hash_bucket_size: int, dtype: str = "string") -> CategoricalColumn:
    if hash_bucket_size is None or hash_bucket_size < 1:
        raise ValueError(f"hash_bucket_size must be at least 1. hash_bucket_size: {hash_bucket_size}, key: {key}")
    return CategoricalColumn(key, "hash", int(hash_bucket_size), (), dtype)


def categorical_column_with_vocabulary_list(key: str, vocabulary_list: Sequence, dtype=None) -> CategoricalColumn:
    if vocabulary_list is None or len(vocabulary_list) < 1:
        raise ValueError(f"vocabulary_list {vocabulary_list} must be non-empty, column_name: {key}")
    if len(set(vocabulary_list)) != len(vocabulary_list):
        raise ValueError(f"Duplicate keys in vocabulary_list: {vocabulary_list}, column_name: {key}")
    return CategoricalColumn(key, "vocab", len(vocabulary_list), tuple(vocabulary_list),
                             dtype or ("string" if isinstance(vocabulary_list[0], str) else "int64"))


def categorical_column_with_identity(key: str, num_buckets: int) -> CategoricalColumn:
    if num_buckets < 1:
        raise ValueError(f"num_buckets {num_buckets} < 1, column_name {key}")
    return CategoricalColumn(key, "identity", int(num_buckets), (), "int64")


@dataclass(frozen=True)
class IndicatorColumn:
    categorical_column: CategoricalColumn

    @property
    def name(self) -> str:
        return f"{self.categorical_column.key}_indicator"


@dataclass(frozen=True)
class EmbeddingColumn:
    categorical_column: CategoricalColumn
    dimension: int
    combiner: str = "mean"
    initializer_stddev: Optional[float] = None   # None -> 1/sqrt(dimension) (TF default)

    @property
    def name(self) -> str:
        return f"{self.categorical_column.key}_embedding"


def indicator_column(categorical_column: CategoricalColumn) -> IndicatorColumn:
    return IndicatorColumn(categorical_column)


def embedding_column(categorical_column: CategoricalColumn, dimension: int, combiner: str = "mean",
                     initializer_stddev: Optional[float] = None) -> EmbeddingColumn:
    if dimension is None or dimension < 1:
        raise ValueError(f"Invalid dimension {dimension}.")
    if combiner not in ("mean", "sqrtn", "sum"):
        raise ValueError(f"Invalid combiner {combiner}.")
    return EmbeddingColumn(categorical_column, int(dimension), combiner, initializer_stddev)
