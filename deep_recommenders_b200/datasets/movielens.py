"""MovieLens-1M dataset classes at the reference's import path ``deep_recommenders.datasets.movielens``.

Mirrors reference datasets/movielens.py: `serialize_tfrecords` (:65-96, same record schema :54-62), `MovieLens`
(:99-131: vocabularies, counts, `dataset(epochs, batch_size)`), `MovielensRanking` (:134-185: step arithmetic and
`input_fn` feature renaming + `rating > 3` labels).  Batches are plain dicts: fixed string features as PackedStrings,
int64 features as numpy arrays, "Genres" as a RaggedFeature -- exactly what the models' id pipeline consumes
(hash bucket / vocabulary lookup in the C-ABI library).  There is no network here: `download=True` raises.
"""
from __future__ import annotations

import os
import random
from typing import Dict, Iterator, Tuple

import numpy as np

from ..feature_column import PackedStrings, RaggedFeature
from .tfrecord import TFRecordFile, TFRecordWriter, serialize_example

_INT_COLS = ["Age", "Occupation", "Rating", "Timestamp"]
_STR_COLS = ["UserID", "MovieID", "Gender", "Zip-code", "Title"]


def _load_data(filename, columns):
    data = {}
    with open(filename, "r", encoding="unicode_escape") as f:
        for line in f:
            ls = line.strip("\n").split("::")
            data[ls[0]] = dict(zip(columns[1:], ls[1:]))
    return data


def _serialize_example(feature) -> bytes:
    out = {c: [int(feature[c])] for c in _INT_COLS}
    out.update({c: [feature[c]] for c in _STR_COLS})
    out["Genres"] = list(feature["Genres"])
    return serialize_example(out)


def serialize_tfrecords(tfrecords_fn, datadir="ml-1m", download=False, seed=None):
    """users.dat / movies.dat / ratings.dat ("::"-separated) -> one TFRecord file, ratings shuffled."""
    if download:
        raise RuntimeError("no network in this environment: place the extracted ml-1m directory at `datadir`")
    users = _load_data(os.path.join(datadir, "users.dat"), ["UserID", "Gender", "Age", "Occupation", "Zip-code"])
    movies = _load_data(os.path.join(datadir, "movies.dat"), ["MovieID", "Title", "Genres"])
    with open(os.path.join(datadir, "ratings.dat"), "r", encoding="unicode_escape") as f:
        lines = f.readlines()
    random.Random(seed).shuffle(lines)
    with TFRecordWriter(tfrecords_fn) as w:
        for line in lines:
            ls = line.strip().split("::")
            rating = dict(zip(["UserID", "MovieID", "Rating", "Timestamp"], ls))
            rating.update(users.get(ls[0]))
            rating.update(movies.get(ls[1]))
            for c in _STR_COLS:
                rating[c] = rating[c].encode("utf-8")
            rating["Genres"] = [x.encode("utf-8") for x in rating["Genres"].split("|")]
            w.write(_serialize_example(rating))


class MovieLens(object):

    def __init__(self, filename="movielens.tfrecords"):
        self._filename = filename if os.path.isabs(filename) else os.path.join(os.path.dirname(__file__), filename)
        self._columns = ["UserID", "MovieID", "Rating", "Timestamp", "Gender", "Age", "Occupation", "Zip-code",
                         "Title", "Genres"]
        self.num_ratings = 1000209
        self.num_users = 6040
        self.num_movies = 3952
        self.gender_vocab = ["F", "M"]
        self.age_vocab = [1, 18, 25, 35, 45, 50, 56]
        self.occupation_vocab = list(range(21))
        self.genres_vocab = ["Action", "Adventure", "Animation", "Children's", "Comedy", "Crime", "Documentary", "Drama",
                             "Fantasy", "Film-Noir", "Horror", "Musical", "Mystery", "Romance", "Sci-Fi", "Thriller",
                             "War", "Western"]
        self._file = None

    _SPEC = {**{c: ("int64", True) for c in _INT_COLS}, **{c: ("string", True) for c in _STR_COLS},
             "Genres": ("string", False)}

    def _open(self) -> TFRecordFile:
        if self._file is None:
            self._file = TFRecordFile(self._filename)
        return self._file

    def dataset(self, epochs=1, batch_size=256) -> Iterator[Tuple[Dict[str, object], np.ndarray]]:
        """TFRecordDataset(...).repeat(epochs).batch(batch_size).map(parse): batches run across epoch boundaries
        exactly as repeat-then-batch does; the last batch may be short."""
        f = self._open()
        n = len(f)
        total = n * epochs
        pos = 0
        while pos < total:
            take = min(batch_size, total - pos)
            lo = pos % n
            if lo + take <= n:
                example = f.parse(lo, lo + take, self._SPEC)
            else:       # the batch wraps into the next epoch: parse the two pieces and concatenate
                example = _concat(f.parse(lo, n, self._SPEC), f.parse(0, lo + take - n, self._SPEC))
            ratings = example.pop("Rating")
            yield example, ratings
            pos += take


def _concat(a: Dict[str, object], b: Dict[str, object]) -> Dict[str, object]:
    def cat(x, y):
        if isinstance(x, PackedStrings):
            return PackedStrings(np.concatenate([x.data, y.data]),
                                 np.concatenate([x.offsets, y.offsets[1:] + x.offsets[-1]]))
        if isinstance(x, RaggedFeature):
            return RaggedFeature(cat(x.values, y.values),
                                 np.concatenate([x.row_splits, y.row_splits[1:] + x.row_splits[-1]]))
        return np.concatenate([x, y])

    return {k: cat(a[k], b[k]) for k in a}


class MovielensRanking(MovieLens):

    def __init__(self, epochs: int = 10, batch_size: int = 1024, buffer_size: int = 1024, train_size: float = 0.8,
                 *args, **kwargs):
        super(MovielensRanking, self).__init__(*args, **kwargs)
        self._epochs = epochs
        self._batch_size = batch_size
        self._buffer_size = buffer_size
        self._train_size = train_size

    @property
    def train_steps(self):
        return int(self.num_ratings * self._epochs * self._train_size // self._batch_size)

    @property
    def train_steps_per_epoch(self):
        return int(self.num_ratings * self._train_size // self._batch_size)

    @property
    def test_steps(self):
        return self.num_ratings // self._batch_size - self.train_steps_per_epoch

    @property
    def training_input_fn(self):
        return _take(self.input_fn(), 0, self.train_steps)

    @property
    def testing_input_fn(self):
        return _take(self.input_fn(), self.train_steps, self.test_steps)

    def input_fn(self):
        for x, y in self.dataset(self._epochs, self._batch_size):
            yield ({"user_id": x["UserID"], "user_gender": x["Gender"], "user_age": x["Age"],
                    "user_occupation": x["Occupation"], "movie_id": x["MovieID"], "movie_genres": x["Genres"]},
                   (y > 3).astype(np.float32).reshape(-1, 1))


def _take(it, skip: int, count: int):
    for i, item in enumerate(it):
        if i >= skip + count:
            return
        if i >= skip:
            yield item
