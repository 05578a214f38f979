"""SURVEY 8(f) #4 on the CPU: native TFRecord framing + tf.train.Example parsing (csrc/tfrecord.cu) against the
published formats -- CRC-32C test vectors of RFC 3720 B.4, the protobuf library's own encoding of the Example
schema -- and the reference's MovieLens dataset classes (datasets/movielens.py) on a synthetic ml-1m directory."""
import os
import struct

import numpy as np
import pytest

from deep_recommenders_b200.datasets import TFRecordFile, TFRecordWriter, serialize_example
from deep_recommenders_b200.feature_column import PackedStrings, RaggedFeature


def test_crc32c_rfc3720_vectors(lib):
    crc = lambda b: lib.dr_crc32c_host(b, len(b))
    assert crc(b"123456789") == 0xE3069283
    assert crc(bytes(32)) == 0x8A9136AA
    assert crc(bytes([0xFF] * 32)) == 0x62A8AB43
    assert crc(bytes(range(32))) == 0x46DD794E
    assert crc(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert crc(b"") == 0
    c = crc(b"hello tfrecord")
    assert lib.dr_masked_crc32c_host(b"hello tfrecord", 14) == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def _example_classes():
    """tensorflow/core/example/{feature,example}.proto rebuilt with the protobuf library (TensorFlow is not installed)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="dr_example_test.proto", package="drtest", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname, packed in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
            if packed is not None:
                f.options.packed = packed
        return m

    msg("BytesList", [("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None, None)])
    msg("FloatList", [("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None, True)])
    msg("Int64List", [("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None, True)])
    feat = msg("Feature", [("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".drtest.BytesList", None),
                           ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".drtest.FloatList", None),
                           ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".drtest.Int64List", None)])
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", [("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".drtest.Features.FeatureEntry", None)])
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".drtest.Feature")
    msg("Example", [("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".drtest.Features", None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("drtest.Example"))


def _write(path, records):
    with TFRecordWriter(path) as w:
        for r in records:
            w.write(r)


SPEC = {"Age": ("int64", True), "UserID": ("string", True), "Genres": ("string", False), "w": ("float32", False)}


def test_parser_reads_what_the_protobuf_library_writes(tmp_path):
    Example = _example_classes()
    rng = np.random.default_rng(0)
    want, records = [], []
    for i in range(50):
        ex = Example()
        age = int(rng.integers(-2 ** 62, 2 ** 62)) if i % 3 else int(rng.integers(-5, 5))
        uid = bytes(rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8))
        genres = [bytes(rng.integers(65, 91, size=int(rng.integers(0, 12)), dtype=np.uint8)) for _ in range(int(rng.integers(0, 5)))]
        w = rng.standard_normal(int(rng.integers(0, 4))).astype(np.float32).tolist()
        ex.features.feature["Age"].int64_list.value.append(age)
        ex.features.feature["UserID"].bytes_list.value.append(uid)
        ex.features.feature["Genres"].bytes_list.value.extend(genres)
        ex.features.feature["w"].float_list.value.extend(w)
        ex.features.feature["unrelated"].int64_list.value.extend([1, 2, 3])
        records.append(ex.SerializeToString())
        want.append((age, uid, genres, w))
    path = str(tmp_path / "pb.tfrecords")
    _write(path, records)
    f = TFRecordFile(path)
    assert len(f) == 50 and f.record(7) == records[7]
    out = f.parse(0, 50, SPEC)
    assert out["Age"].tolist() == [w[0] for w in want]
    assert out["UserID"].tolist() == [w[1] for w in want]
    g = out["Genres"]
    assert isinstance(g, RaggedFeature) and isinstance(g.values, PackedStrings)
    flat = g.values.tolist()
    assert [flat[g.row_splits[i]:g.row_splits[i + 1]] for i in range(50)] == [w[2] for w in want]
    wf = out["w"]
    assert np.array_equal(wf.values, np.concatenate([np.asarray(w[3], np.float32) for w in want]))
    sub = f.parse(10, 20, {"Age": ("int64", True)})
    assert sub["Age"].tolist() == [w[0] for w in want[10:20]]
    f.close()


def test_our_serializer_is_readable_by_the_protobuf_library_and_handles_unpacked_scalars(tmp_path):
    Example = _example_classes()
    feats = {"Age": [25], "Occupation": [7], "UserID": [b"4711"], "Genres": [b"Drama", b"Sci-Fi"], "w": [0.5, -2.0],
             "neg": [-1, -2 ** 63, 2 ** 63 - 1]}
    blob = serialize_example(feats)
    ex = Example.FromString(blob)
    assert list(ex.features.feature["Age"].int64_list.value) == [25]
    assert list(ex.features.feature["Genres"].bytes_list.value) == [b"Drama", b"Sci-Fi"]
    assert list(ex.features.feature["w"].float_list.value) == [0.5, -2.0]
    assert list(ex.features.feature["neg"].int64_list.value) == [-1, -2 ** 63, 2 ** 63 - 1]
    assert blob == Example.FromString(blob).SerializeToString(deterministic=True)
    # unpacked encodings of repeated scalars (legal protobuf; old writers emit them)
    def ld(field, payload):
        return bytes([(field << 3) | 2, len(payload)]) + payload
    int_list = bytes([0x08, 5, 0x08, 6])                               # value: 5, value: 6 (wire type 0)
    flt_list = bytes([0x0D]) + struct.pack("<f", 1.5)                 # value: 1.5 (wire type 5)
    entry = lambda k, feat: ld(1, ld(1, k) + ld(2, feat))
    rec = ld(1, entry(b"Age", ld(3, int_list)) + entry(b"w", ld(2, flt_list)))
    path = str(tmp_path / "unpacked.tfrecords")
    _write(path, [rec])
    out = TFRecordFile(path).parse(0, 1, {"Age": ("int64", False), "w": ("float32", False)})
    assert out["Age"].values.tolist() == [5, 6] and out["w"].values.tolist() == [1.5]


def test_framing_errors_are_detected(tmp_path, lib):
    path = str(tmp_path / "a.tfrecords")
    _write(path, [serialize_example({"Age": [i]}) for i in range(4)])
    raw = bytearray(open(path, "rb").read())
    assert len(TFRecordFile(path)) == 4
    bad = bytearray(raw)
    bad[20] ^= 0x40                                                   # payload byte of record 0
    open(path, "wb").write(bad)
    with pytest.raises(ValueError, match="crc mismatch"):
        TFRecordFile(path)
    assert len(TFRecordFile(path, verify_crc=False)) == 4               # framing itself is intact
    open(path, "wb").write(raw[:-3])
    with pytest.raises(ValueError, match="truncated"):
        TFRecordFile(path)
    open(path, "wb").write(b"")
    assert len(TFRecordFile(path)) == 0
    # a required single-valued feature that is missing / a kind mismatch
    _write(path, [serialize_example({"Age": [1]}), serialize_example({"Other": [2]})])
    f = TFRecordFile(path)
    with pytest.raises(ValueError, match="required"):
        f.parse(0, 2, {"Age": ("int64", True)})
    with pytest.raises(ValueError, match="not of the requested kind"):
        f.parse(0, 1, {"Age": ("string", False)})


def _fake_ml1m(d):
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location(
        "c1_tfrecords_example", pathlib.Path(__file__).resolve().parent.parent / "examples" / "train_fm_on_movielens_tfrecords.py")
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    ex.write_synthetic_ml1m(d, n_users=30, n_movies=20, n_ratings=257, seed=0)
    return ex


def test_movielens_classes_on_a_synthetic_ml1m(tmp_path):
    from deep_recommenders.datasets.movielens import MovieLens, MovielensRanking, serialize_tfrecords
    from deep_recommenders_b200 import feature_column as fc
    from deep_recommenders_b200.hashing import hash_bucket, vocabulary_ids
    from oracle import farmhash_py as F
    d = str(tmp_path / "ml-1m")
    _fake_ml1m(d)
    rec = str(tmp_path / "movielens.tfrecords")
    serialize_tfrecords(rec, datadir=d, seed=1)
    with pytest.raises(RuntimeError, match="network"):
        serialize_tfrecords(rec, datadir=d, download=True)
    ml = MovieLens(rec)
    assert ml.num_users == 6040 and ml.gender_vocab == ["F", "M"] and len(ml.genres_vocab) == 18
    batches = list(ml.dataset(epochs=2, batch_size=100))                # 514 records -> 5 full + one of 14
    assert [len(y) for _, y in batches] == [100] * 5 + [14]
    x0, y0 = batches[0]
    assert set(x0) == {"UserID", "MovieID", "Timestamp", "Gender", "Age", "Occupation", "Zip-code", "Title", "Genres"}
    # repeat-then-batch: batch 2 wraps from the end of epoch 1 into epoch 2 (records 200..256, then 0..42)
    x2, y2 = batches[2]
    ref = ml._open().parse(0, 257, MovieLens._SPEC)
    assert x2["UserID"].tolist() == ref["UserID"].tolist()[200:] + ref["UserID"].tolist()[:43]
    assert y2.tolist() == ref["Rating"].tolist()[200:] + ref["Rating"].tolist()[:43]
    g2, gr = x2["Genres"], ref["Genres"]
    assert g2.row_splits[-1] == len(g2.values) and len(g2.row_splits) == 101
    flat2, flatr = g2.values.tolist(), gr.values.tolist()
    assert flat2[:int(g2.row_splits[57])] == flatr[int(gr.row_splits[200]):]
    # MovielensRanking: feature renaming, labels = rating > 3, step arithmetic (movielens.py:148-185)
    rk = MovielensRanking(epochs=1, batch_size=64, filename=rec)
    assert rk.train_steps_per_epoch == int(1000209 * 0.8 // 64) and rk.test_steps == 1000209 // 64 - rk.train_steps_per_epoch
    feats, labels = next(iter(rk.input_fn()))
    assert set(feats) == {"user_id", "user_gender", "user_age", "user_occupation", "movie_id", "movie_genres"}
    assert labels.shape == (64, 1) and labels.dtype == np.float32
    assert np.array_equal(labels.reshape(-1), (np.asarray(ref["Rating"][:64]) > 3).astype(np.float32))
    # packed strings feed the native id pipeline directly: same ids as the Python restatement of FarmHash
    ids = hash_bucket(feats["user_id"], ml.num_users)
    assert ids.tolist() == F.hash_bucket_py(feats["user_id"].tolist(), ml.num_users)
    gender = fc.categorical_column_with_vocabulary_list("user_gender", ml.gender_vocab)
    gids = vocabulary_ids(gender, feats["user_gender"])
    assert gids.tolist() == [ml.gender_vocab.index(s.decode()) for s in feats["user_gender"].tolist()]
    oov = vocabulary_ids(gender, feats["movie_genres"].values)          # the examples' always-OOV genre slot
    assert set(oov.tolist()) == {-1}


def test_batch_parser_equals_single_feature_parser_and_later_map_entry_wins(tmp_path):
    rng = np.random.default_rng(3)
    recs = []
    for i in range(64):
        recs.append(serialize_example({"Age": [int(rng.integers(-9, 9))], "UserID": [bytes(rng.integers(65, 91, size=int(rng.integers(0, 9)), dtype=np.uint8))],
                                       "Genres": [b"g%d" % j for j in range(int(rng.integers(0, 4)))],
                                       "w": [float(x) for x in rng.standard_normal(int(rng.integers(0, 3)))]}))
    # a record whose map repeats the key "Age": protobuf map semantics = the later entry wins
    def ld(field, payload):
        return bytes([(field << 3) | 2, len(payload)]) + payload
    entry = lambda k, feat: ld(1, ld(1, k) + ld(2, feat))
    recs.append(ld(1, entry(b"Age", ld(3, ld(1, bytes([7])))) + entry(b"UserID", ld(1, ld(1, b"u"))) + entry(b"Age", ld(3, ld(1, bytes([9]))))))
    path = str(tmp_path / "b.tfrecords")
    _write(path, recs)
    f = TFRecordFile(path)
    n = len(recs)
    off, ln = np.ascontiguousarray(f.offsets), np.ascontiguousarray(f.lengths)
    items = [("Age", 0), ("UserID", 1), ("Genres", 1), ("w", 2), ("absent", 0)]
    batch = f._columns(off, ln, items)
    for (name, kind), (sp_b, v_b) in zip(items, batch):
        sp_s, v_s = f._column(off, ln, name, kind)
        assert np.array_equal(sp_b, sp_s), name
        if isinstance(v_b, PackedStrings):
            assert v_b.tolist() == v_s.tolist(), name
        else:
            assert np.array_equal(v_b, v_s), name
    out = f.parse(n - 1, n, {"Age": ("int64", True), "UserID": ("string", True)})
    assert out["Age"].tolist() == [9] and out["UserID"].tolist() == [b"u"]


def test_threaded_batch_parse_is_identical_to_serial(tmp_path, lib):
    rng = np.random.default_rng(9)
    n = 20000
    recs = [serialize_example({"Age": [int(rng.integers(-99, 99))],
                               "UserID": [b"%d" % int(rng.integers(0, 10 ** 6))],
                               "Genres": [b"g%d" % j for j in range(int(rng.integers(0, 4)))],
                               "w": [float(i)] * int(rng.integers(0, 3))}) for i in range(n)]
    path = str(tmp_path / "big.tfrecords")
    _write(path, recs)
    f = TFRecordFile(path)
    spec = {"Age": ("int64", True), "UserID": ("string", True), "Genres": ("string", False), "w": ("float32", False)}
    outs = []
    try:
        for threads in (1, 2, 7):
            assert lib.dr_set_host_threads(threads) == 0
            outs.append(f.parse(0, n, spec))
    finally:
        lib.dr_set_host_threads(1)
    a = outs[0]
    for b in outs[1:]:
        assert np.array_equal(a["Age"], b["Age"])
        assert np.array_equal(a["UserID"].data, b["UserID"].data) and np.array_equal(a["UserID"].offsets, b["UserID"].offsets)
        assert np.array_equal(a["Genres"].row_splits, b["Genres"].row_splits)
        assert np.array_equal(a["Genres"].values.offsets, b["Genres"].values.offsets)
        assert np.array_equal(a["Genres"].values.data, b["Genres"].values.data)
        assert np.array_equal(a["w"].values, b["w"].values) and np.array_equal(a["w"].row_splits, b["w"].row_splits)
    assert a["UserID"].tolist()[:3] == [r for r in (TFRecordFile(path).parse(0, 3, {"UserID": ("string", True)})["UserID"].tolist())]
    # an error in a later range is reported with its record number, whichever thread finds it
    bad = list(recs)
    bad[15000] = b"\x0a\x05abc"                                     # truncated length-delimited field
    _write(path, bad)
    g = TFRecordFile(path)
    lib.dr_set_host_threads(4)
    try:
        with pytest.raises(ValueError, match="record 15000"):
            g.parse(0, n, {"Age": ("int64", False)})
    finally:
        lib.dr_set_host_threads(1)
    assert lib.dr_set_host_threads(-1) == -1
