"""CPU checks of the SURVEY 8(f) "next" rows: the FarmHash host twin of the C-ABI library against the independent
Python restatement and the published answers, and the oracle's retrieval helpers against the reference's own
known-answer tests (tests/keras/test_factorized_top_k.py:17-34, :86-130)."""
import numpy as np
import pytest

from oracle import farmhash_py as F
from oracle import reference_np as R


def test_farmhash_oracle_known_answers():
    assert F.fingerprint64(b"") == 0x9AE16A3B2F90404F
    assert F.fingerprint64(b"abc") == 2640714258260161385            # pyfarmhash docs
    assert F.fingerprint64(b"hello") == 13009744463427800296
    assert F.hash_bucket_py(["Hello", "TensorFlow", "2.x"], 3) == [0, 2, 2]   # TF API docs, to_hash_bucket_fast


def test_farmhash_host_twin_matches_python_restatement_all_length_branches():
    """Every branch of Hash64 (0, 1-3, 4-7, 8-16, 17-32, 33-64, > 64 incl. multiples of 64): the C twin compiled
    from csrc/farmhash.cuh and the separately written Python restatement must agree bit for bit."""
    from deep_recommenders_b200.hashing import fingerprint64, hash_bucket
    rng = np.random.default_rng(7)
    lengths = list(range(0, 70)) + [127, 128, 129, 191, 192, 193, 255, 256, 257, 1000]
    strings = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in lengths for _ in range(3)]
    for s in strings:
        assert fingerprint64(s) == F.fingerprint64(s), len(s)
    for nb in (1, 3, 100, 1_000_003, (1 << 40) + 7):
        assert hash_bucket(strings, nb).tolist() == F.hash_bucket_py(strings, nb)


def test_hash_bucket_of_integers_goes_through_the_decimal_string():
    from deep_recommenders_b200.hashing import hash_bucket
    vals = np.array([0, 1, -1, 7, 10, 99, 6040, 3952, -12345, 2 ** 31, -2 ** 31, 2 ** 63 - 1, -2 ** 63], dtype=np.int64)
    got = hash_bucket(vals, 1000).tolist()
    assert got == F.hash_bucket_py([str(int(v)) for v in vals], 1000)
    assert got == hash_bucket([str(int(v)) for v in vals], 1000).tolist()


def test_hash_host_entries_validate_arguments(lib):
    out = np.zeros(2, dtype=np.int64)
    offs = np.array([0, 3, 1], dtype=np.int64)               # not monotone
    data = np.frombuffer(b"abc", dtype=np.uint8)
    assert lib.dr_hash_bucket_bytes_host(data.ctypes.data, offs.ctypes.data, 2, 10, out.ctypes.data) == -1
    assert b"monotone" in lib.dr_last_error()
    assert lib.dr_hash_bucket_i64_host(out.ctypes.data, 2, 0, out.ctypes.data) == -1       # num_buckets < 1
    assert lib.dr_hash_bucket_i64_host(None, 0, 5, None) == 0                              # empty is fine


def test_vocabulary_ids_oov_is_minus_one():
    from deep_recommenders_b200 import feature_column as fc
    from deep_recommenders_b200.hashing import vocabulary_ids
    col = fc.categorical_column_with_vocabulary_list("g", ["F", "M"])
    assert vocabulary_ids(col, ["M", "F", "x", b"M"]).tolist() == [1, 0, -1, 1]
    col = fc.categorical_column_with_vocabulary_list("age", [1, 18, 25, 35, 45, 50, 56])
    assert vocabulary_ids(col, [18, 56, 2, 1]).tolist() == [1, 6, -1, 0]


# ---- oracle retrieval helpers against the reference's own tests ------------------------------------------
def test_take_long_axis_kat():                                # tests/keras/test_factorized_top_k.py:17-23
    out = R.take_long_axis(np.float32([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]]), [[0, 1], [2, 1]])
    np.testing.assert_allclose(out, [[0.1, 0.2], [0.6, 0.5]])


def test_exclude_kat():                                       # tests/keras/test_factorized_top_k.py:25-34
    x, y = R.exclude(np.float32([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]]), [[0, 1, 2], [3, 4, 5]], [[1, 2], [3, 5]], 1)
    np.testing.assert_allclose(x, [[0.1], [0.5]])
    assert y.tolist() == [[0], [4]]


@pytest.mark.parametrize("index", ["streaming", "brute_force"])
def test_factorized_topk_metrics_property(index):             # tests/keras/test_factorized_top_k.py:86-130
    rng = np.random.RandomState(42)
    num_candidates, num_queries, dim = 100, 10, 4
    candidates = rng.normal(size=(num_candidates, dim)).astype(np.float32)
    queries = rng.normal(size=(num_queries, dim)).astype(np.float32)
    true_candidates = rng.normal(size=(num_queries, dim)).astype(np.float32)
    positive = (queries * true_candidates).sum(axis=1, keepdims=True)
    all_scores = np.concatenate([positive, queries @ candidates.T], axis=1)
    ks = [1, 5, 10, 50]
    if index == "streaming":
        batches = [candidates[i:i + 32] for i in range(0, num_candidates, 32)]
        top, ident = R.streaming_topk(queries, batches, None, max(ks))
    else:
        top, ident = R.brute_force_topk(queries, candidates, None, max(ks))
    # both indexes retrieve the same candidates
    bf, bf_ident = R.brute_force_topk(queries, candidates, None, max(ks))
    np.testing.assert_array_equal(ident, bf_ident)
    np.testing.assert_allclose(top, bf, rtol=1e-6)
    for k, value in zip(ks, R.factorized_topk_metric(queries, true_candidates, top, ks)):
        in_top_k = ((all_scores > all_scores[:, :1]).sum(axis=1) < k)
        assert value == pytest.approx(in_top_k.mean())


def test_topk_tie_order_is_lower_index_first():
    v, i = R.top_k(np.float32([[1, 3, 3, 2, 3]]), 4)
    assert i.tolist() == [[1, 2, 4, 3]] and v.tolist() == [[3, 3, 3, 2]]
    with pytest.raises(ValueError):
        R.top_k(np.zeros((2, 3), np.float32), 4)


def test_embedding_bag_oracle_semantics():
    table = np.arange(12, dtype=np.float32).reshape(4, 3)
    ids, splits = [0, 2, -1, 3, 3, 9], [0, 2, 3, 3, 6]
    mean = R.embedding_bag(table, ids, splits, "mean")
    np.testing.assert_allclose(mean, [(table[0] + table[2]) / 2, [0, 0, 0], [0, 0, 0], table[3]])
    np.testing.assert_allclose(R.embedding_bag(table, ids, splits, "sum")[3], 2 * table[3])
    np.testing.assert_allclose(R.embedding_bag(table, ids, splits, "sqrtn")[3], 2 * table[3] / np.sqrt(2), rtol=1e-6)
    g = np.ones((4, 3))
    gt = R.embedding_bag_grad(4, ids, splits, g, "mean")
    np.testing.assert_allclose(gt, [[.5] * 3, [0] * 3, [.5] * 3, [1.] * 3])


def test_adam_oracle_matches_torch_adam():
    """The ApplyAdam restatement against torch.optim.Adam (same recurrence up to where eps enters: TF adds eps to
    sqrt(v) un-corrected and folds both bias corrections into lr_t; with eps = 0 the two coincide).  The oracle
    rounds lr_t and the betas to float32 as TensorFlow's kernels receive them, so torch gets the same float32 betas
    and the comparison is at float32 resolution."""
    import torch
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(50)
    p = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    b1, b2 = float(np.float32(0.9)), float(np.float32(0.999))
    opt = torch.optim.Adam([p], lr=0.0078125, betas=(b1, b2), eps=0.0)
    pn, m, v = p0.copy(), np.zeros(50), np.zeros(50)
    for t in range(1, 6):
        g = rng.standard_normal(50)
        p.grad = torch.tensor(g)
        opt.step()
        pn, m, v = R.adam_dense(pn, g, m, v, R.adam_lr_t(0.0078125, t), eps=0.0, dtype=np.float64)
    np.testing.assert_allclose(pn, p.detach().numpy(), rtol=5e-7, atol=1e-9)
    # float32 hyper-parameters: 1 - float32(0.999) differs from 0.001 by 1.3e-5 relative (why the oracle rounds them)
    assert abs((1.0 - b2) / 0.001 - 1.0) > 1e-5


# ---- golden vectors produced by executing the reference's factorized_top_k.py under the TF stand-in -------------
def _retrieval_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retrieval_golden.npz"))


def test_oracle_reproduces_reference_retrieval_golden():
    g = _retrieval_golden()
    np.testing.assert_array_equal(R.take_long_axis(g["tla_arr"], g["tla_idx"]), g["tla_out"])
    np.testing.assert_allclose(g["tla_kat"], [[0.1, 0.2], [0.6, 0.5]], rtol=1e-6)
    assert g["exclude_kat_ids"].tolist() == [[0], [4]]
    for k in (5, 40, 60):
        xs, xi = R.exclude(g["exclude_scores"], g["exclude_ident"], g["exclude_excl"], k)
        np.testing.assert_array_equal(xi, g[f"exclude_k{k}_ids"])
        np.testing.assert_array_equal(xs, g[f"exclude_k{k}_scores"])
    q, c, names = g["idx_queries"], g["idx_candidates"], g["idx_names"]
    batches = [c[i:i + 32] for i in range(0, 100, 32)]
    nbatches = [names[i:i + 32] for i in range(0, 100, 32)]
    for tag, nb, nm in (("noid", None, None), ("id", nbatches, names)):
        s, i = R.streaming_topk(q, batches, nb, 10)
        np.testing.assert_array_equal(i, g[f"streaming_{tag}_ids"])
        np.testing.assert_allclose(s, g[f"streaming_{tag}_scores"], rtol=1e-6)
        s, i = R.brute_force_topk(q, c, nm, 10)
        np.testing.assert_array_equal(i, g[f"brute_{tag}_ids"])
        np.testing.assert_allclose(s, g[f"brute_{tag}_scores"], rtol=1e-6)
        np.testing.assert_array_equal(R.brute_force_topk(q, c, nm, 3)[1], g[f"brute_{tag}_k3_ids"])
    s7, i7 = R.brute_force_topk(q, c, names, 7)                         # query_with_exclusions: k + 2 then _exclude
    xs, xi = R.exclude(s7, i7, g["qwe_ban"], 7)
    np.testing.assert_array_equal(xi, g["qwe_ids"])
    s, i = R.streaming_topk(q, [c[:32], c[32:40]], None, 50)             # k clipped to what the stream holds
    np.testing.assert_array_equal(i, g["streaming_small_ids"])
    assert "candidate batch too small" in str(g["streaming_small_error"])
    top, _ = R.brute_force_topk(q, c, None, 50)
    want = R.factorized_topk_metric(q, g["idx_true"], top, g["metric_ks"].tolist())
    for tag in ("streaming", "brute", "dataset"):
        np.testing.assert_allclose(g[f"metric_{tag}"], want, rtol=1e-6)


# ---- host side of the id pipeline: feature dict -> (ids [B, S], bags) without a GPU ----------------------------------
def test_ids_and_bags_from_mixed_feature_dict_on_host():
    import torch
    from deep_recommenders_b200 import feature_column as fc
    from deep_recommenders_b200.feature_column import PackedStrings, RaggedFeature
    from deep_recommenders_b200.hashing import ids_and_bags, pack_strings
    cols = [fc.categorical_column_with_hash_bucket("user_id", 6040),
            fc.categorical_column_with_vocabulary_list("user_gender", ["F", "M"]),
            fc.categorical_column_with_vocabulary_list("user_age", [1, 18, 25, 35, 45, 50, 56]),
            fc.categorical_column_with_identity("slot", 10),
            fc.categorical_column_with_vocabulary_list("movie_genres", ["Action", "Comedy", "Drama"])]
    cats = {c.key: c for c in cols}
    keys = [c.key for c in cols]
    uid = [b"17", b"4711", b"1", b"6040"]
    data, offs = pack_strings(uid)
    gdata, goffs = pack_strings([b"Drama", b"Western", b"Action", b"Comedy", b"Drama"])
    feats = {
        "user_id": PackedStrings(data, offs),                          # what the TFRecord parser emits
        "user_gender": np.asarray([["M"], ["F"], ["?"], ["M"]]),       # [B, 1] strings, one OOV
        "user_age": np.asarray([18, 56, 2, 1]),
        "slot": torch.tensor([0, 9, 10, -3]),                          # identity column: out of range -> -1
        "movie_genres": RaggedFeature(PackedStrings(gdata, goffs), np.asarray([0, 2, 2, 3, 5])),
    }
    ids, bags = ids_and_bags(keys, cats, feats, torch.device("cpu"))
    assert ids.shape == (4, 5) and ids.dtype == torch.int64
    assert ids[:, 0].tolist() == F.hash_bucket_py(uid, 6040)
    assert ids[:, 1].tolist() == [1, 0, -1, 1]
    assert ids[:, 2].tolist() == [1, 6, -1, 0]
    assert ids[:, 3].tolist() == [0, 9, -1, -1]
    assert list(bags) == [4]
    bag_ids, splits = bags[4]
    assert bag_ids.tolist() == [2, -1, 0, 1, 2] and splits.tolist() == [0, 2, 2, 3, 5]
    with pytest.raises(KeyError):
        ids_and_bags(keys, cats, {k: v for k, v in feats.items() if k != "slot"}, torch.device("cpu"))
    with pytest.raises(ValueError, match="one value per example"):
        ids_and_bags(["user_age"], cats, {"user_age": np.zeros((4, 3), dtype=np.int64)}, torch.device("cpu"))
