"""Input formats of the hot path: TFRecord files of tf.train.Example protos and the
MovieLens dataset classes of the reference (datasets/movielens.py), on the native parser of the C-ABI library."""
from .tfrecord import TFRecordFile, TFRecordWriter, serialize_example  # noqa: F401
from .movielens import MovieLens, MovielensRanking, serialize_tfrecords  # noqa: F401
