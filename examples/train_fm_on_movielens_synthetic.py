#!/usr/bin/env python
"""BASELINE config C1 ("plumbing"): the reference's examples/train_fm_on_movielens_estimator.py
(:10-54: six MovieLens columns, embedding dim 16, FM, sigmoid cross-entropy, Adam(0.01), batch 1024)
on SYNTHETIC MovieLens-1M-shaped data -- the dataset download and the TFRecord reader are outside the
hot path, and there is no network here.  Column definitions are the reference's, including its quirk
that `movie_genres` is built with the *gender* vocabulary (:22-23), so every genre is
out-of-vocabulary and contributes a zero vector; the multi-valued genre list is therefore fed as one
OOV token per example (mean over all-zero rows == zero row).

    python examples/train_fm_on_movielens_synthetic.py [steps]
Needs a CUDA device: the FM forward/backward runs in libdeeprec_b200.so (no CPU fallback).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import feature_column as fc  # noqa: E402
from deep_recommenders.estimator.models.feature_interaction import FM  # noqa: E402

NUM_USERS, NUM_MOVIES = 6040, 3952
GENDER_VOCAB = ["F", "M"]
AGE_VOCAB = [1, 18, 25, 35, 45, 50, 56]
OCCUPATION_VOCAB = list(range(21))


def build_columns():
    user_id = fc.categorical_column_with_hash_bucket("user_id", NUM_USERS)
    user_gender = fc.categorical_column_with_vocabulary_list("user_gender", GENDER_VOCAB)
    user_age = fc.categorical_column_with_vocabulary_list("user_age", AGE_VOCAB)
    user_occupation = fc.categorical_column_with_vocabulary_list("user_occupation", OCCUPATION_VOCAB)
    movie_id = fc.categorical_column_with_hash_bucket("movie_id", NUM_MOVIES)
    movie_genres = fc.categorical_column_with_vocabulary_list("movie_genres", GENDER_VOCAB)   # sic (reference :22-23)
    base = [user_id, user_gender, user_age, user_occupation, movie_id, movie_genres]
    return [fc.indicator_column(c) for c in base], [fc.embedding_column(c, dimension=16) for c in base]


def synthetic_batch(rng, batch_size=1024):
    feats = {
        "user_id": np.asarray([str(u) for u in rng.integers(1, NUM_USERS + 1, batch_size)]),
        "user_gender": rng.choice(GENDER_VOCAB, batch_size),
        "user_age": rng.choice(AGE_VOCAB, batch_size),
        "user_occupation": rng.choice(OCCUPATION_VOCAB, batch_size),
        "movie_id": np.asarray([str(m) for m in rng.integers(1, NUM_MOVIES + 1, batch_size)]),
        "movie_genres": rng.choice(["Action", "Comedy", "Drama"], batch_size),     # never in GENDER_VOCAB
    }
    rating = rng.integers(1, 6, batch_size)
    labels = (rating > 3).astype(np.float32)[:, None]                              # movielens.py:181-183
    return feats, labels


def main(steps=50):
    rng = np.random.default_rng(42)
    indicator_columns, embedding_columns = build_columns()
    model = FM(indicator_columns, embedding_columns, seed=42, device="cuda")
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    for step in range(steps):
        feats, labels = synthetic_batch(rng)
        logits = model(feats)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.from_numpy(labels).cuda())
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 10 == 0 or step == steps - 1:
            print(f"step {step:4d}  loss {float(loss):.4f}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
