#!/usr/bin/env python
"""BASELINE config C1 end to end, the reference's examples/train_fm_on_movielens_estimator.py (:10-54) with its real
input path: `movielens.tfrecords` (tf.train.Example records, datasets/movielens.py:54-62) -> MovielensRanking.input_fn
(:165-185: feature renaming, label = rating > 3) -> the six feature columns (hash bucket for the ids, vocabulary
lists, the multi-valued "Genres" slot) -> FM -> sigmoid cross-entropy -> Adam(0.01), batch 1024.

The TFRecord framing / Example parse and the FarmHash / vocabulary id pipeline run in libdeeprec_b200.so's host entry
points; lookup + FM forward / backward are the CUDA kernels.  There is no network here, so when no ml-1m directory is
given a small SYNTHETIC one with the real file formats ("::"-separated users.dat / movies.dat / ratings.dat) is written.

    python examples/train_fm_on_movielens_tfrecords.py [steps] [path/to/ml-1m]
"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import feature_column as fc  # noqa: E402
from deep_recommenders.datasets.movielens import MovielensRanking, serialize_tfrecords  # noqa: E402

GENRES = ["Action", "Adventure", "Animation", "Children's", "Comedy", "Crime", "Documentary", "Drama", "Fantasy",
          "Film-Noir", "Horror", "Musical", "Mystery", "Romance", "Sci-Fi", "Thriller", "War", "Western"]


def write_synthetic_ml1m(d, n_users=200, n_movies=120, n_ratings=5000, seed=0):
    """users.dat / movies.dat / ratings.dat in the MovieLens-1M text format with random content."""
    rng = np.random.default_rng(seed)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "users.dat"), "w") as f:
        for u in range(1, n_users + 1):
            f.write(f"{u}::{'FM'[int(rng.integers(0, 2))]}::{[1, 18, 25, 35, 45, 50, 56][int(rng.integers(0, 7))]}::"
                    f"{int(rng.integers(0, 21))}::{10000 + u}\n")
    with open(os.path.join(d, "movies.dat"), "w") as f:
        for m in range(1, n_movies + 1):
            g = "|".join(sorted(set(rng.choice(GENRES, size=int(rng.integers(1, 4))).tolist())))
            f.write(f"{m}::Movie {m} (19{m % 100:02d})::{g}\n")
    with open(os.path.join(d, "ratings.dat"), "w") as f:
        for _ in range(n_ratings):
            f.write(f"{int(rng.integers(1, n_users + 1))}::{int(rng.integers(1, n_movies + 1))}::"
                    f"{int(rng.integers(1, 6))}::{978300000 + int(rng.integers(0, 10 ** 6))}\n")
    return d


def build_columns(ml, fix_genre_vocab=False):
    """The reference's columns (examples/train_fm_on_movielens_estimator.py:10-35).  It builds `movie_genres` over the
    GENDER vocabulary (:22-23), so every genre is out-of-vocabulary; fix_genre_vocab=True uses the genre vocabulary."""
    user_id = fc.categorical_column_with_hash_bucket("user_id", ml.num_users)
    user_gender = fc.categorical_column_with_vocabulary_list("user_gender", ml.gender_vocab)
    user_age = fc.categorical_column_with_vocabulary_list("user_age", ml.age_vocab)
    user_occupation = fc.categorical_column_with_vocabulary_list("user_occupation", ml.occupation_vocab)
    movie_id = fc.categorical_column_with_hash_bucket("movie_id", ml.num_movies)
    movie_genres = fc.categorical_column_with_vocabulary_list(
        "movie_genres", ml.genres_vocab if fix_genre_vocab else ml.gender_vocab)
    base = [user_id, user_gender, user_age, user_occupation, movie_id, movie_genres]
    return [fc.indicator_column(c) for c in base], [fc.embedding_column(c, dimension=16) for c in base]


def main(steps=50, datadir=None):
    from deep_recommenders.estimator.models.feature_interaction import FM
    work = tempfile.mkdtemp(prefix="dr_ml1m_")
    datadir = datadir or write_synthetic_ml1m(os.path.join(work, "ml-1m"))
    records = os.path.join(work, "movielens.tfrecords")
    serialize_tfrecords(records, datadir=datadir, seed=0)
    ml = MovielensRanking(epochs=100, batch_size=1024, filename=records)
    indicator_columns, embedding_columns = build_columns(ml)
    model = FM(indicator_columns, embedding_columns, seed=42, device="cuda")
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    for step, (features, labels) in enumerate(ml.input_fn()):
        if step >= steps:
            break
        logits = model(features)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.from_numpy(labels).cuda())
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 10 == 0 or step == steps - 1:
            print(f"step {step:4d}  loss {float(loss):.4f}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50, sys.argv[2] if len(sys.argv) > 2 else None)
