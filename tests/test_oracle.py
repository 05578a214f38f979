"""Pins the oracle (oracle/reference_np.py) -- CPU only, runs with -m "not gpu".

1. against golden vectors produced by executing the reference's own source files under the
   numpy TensorFlow stand-in (tests/golden/make_golden.py -> hotpath_golden_{f32,f64}.npz);
2. against the reference's own known-answer / property tests;
3. every analytic backward against torch-CPU autograd of the same forward (float64).
"""
import os

import numpy as np
import pytest
import torch

from oracle import reference_np as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", params=["f32", "f64"])
def gold(request):
    g = np.load(os.path.join(GOLD, f"hotpath_golden_{request.param}.npz"))
    return request.param, (np.float32 if request.param == "f32" else np.float64), g


def tol(name):
    return dict(rtol=1e-5, atol=1e-5) if name == "f32" else dict(rtol=1e-12, atol=1e-12)


def test_golden_fm_layer(gold):
    name, dt, g = gold
    for tag in ("fm_a", "fm_b"):
        emb, sparse = g[f"{tag}_emb"], g[f"{tag}_sparse"]
        inter = R.fm_second_order(emb, dt)
        assert np.allclose(inter, g[f"{tag}_y_zero"], **tol(name))            # zero-init linear term
        lin = R.dense(sparse, g[f"{tag}_kernel"], g[f"{tag}_bias"], None, dt)
        assert np.allclose(lin, g[f"{tag}_y_linear_only"], **tol(name))
        assert np.allclose(lin + inter, g[f"{tag}_y"], **tol(name))


def test_golden_estimator_fm(gold):
    name, dt, g = gold
    for tag in ("estfm_a", "estfm_b"):
        y = R.fm_second_order(g[f"{tag}_x"], dt)
        assert y.shape == g[f"{tag}_y"].shape
        assert np.allclose(y, g[f"{tag}_y"], **tol(name))
    with pytest.raises(ValueError) as e:
        R.fm_second_order(np.zeros((4, 6)))
    assert str(e.value) == str(g["estfm_rank_error"])


def test_golden_cross(gold):
    name, dt, g = gold
    y, _ = R.cross(np.asarray([[0.1, 0.2, 0.3]], dt), np.asarray([[0.4, 0.5, 0.6]], dt), w=np.ones((3, 3), dt),
                   b=np.zeros(3, dt), dtype=dt)
    assert np.allclose(y, g["cross_kat_y"], **tol(name))
    assert np.allclose(y, [[0.55, 0.8, 1.05]], rtol=1e-6, atol=1e-6)         # tests/keras/test_dcn.py:16-23
    cfg = {"cross_full": {}, "cross_diag": dict(diag_scale=0.5), "cross_nobias": {},
           "cross_lowrank": {}, "cross_lowrank_diag": dict(diag_scale=0.25)}
    for tag, kw in cfg.items():
        args = dict(w=g[f"{tag}_w"] if f"{tag}_w" in g else None, b=g[f"{tag}_b"] if f"{tag}_b" in g else None,
                    u=g[f"{tag}_u"] if f"{tag}_u" in g else None, v=g[f"{tag}_v"] if f"{tag}_v" in g else None)
        y, _ = R.cross(g[f"{tag}_x0"], g[f"{tag}_x"], dtype=dt, **args, **kw)
        assert np.allclose(y, g[f"{tag}_y"], **tol(name)), tag
        y, _ = R.cross(g[f"{tag}_x0"], None, dtype=dt, **args, **kw)
        assert np.allclose(y, g[f"{tag}_y_xnone"], **tol(name)), tag
    with pytest.raises(ValueError) as e:
        R.cross(np.zeros((2, 13)), np.zeros((2, 12)), w=np.zeros((13, 13)))
    assert str(e.value) == str(g["cross_dim_error"])


def test_golden_sbcnm(gold):
    name, dt, g = gold
    for k in (3, 5, 10, 15, 30):
        ol, oy, _ = R.hard_negative_mining(g["hnm_logits"], g["hnm_labels"], k)
        assert ol.shape == g[f"hnm_k{k}_logits"].shape
        # order inside a row is unspecified (sorted=False): compare as multisets of (logit, label)
        a = np.sort(ol + 10 * oy, axis=1)
        b = np.sort(g[f"hnm_k{k}_logits"] + 10 * g[f"hnm_k{k}_labels"], axis=1)
        assert np.allclose(a, b)
    assert np.allclose(R.remove_accidental_negative(g["ran_logits"], g["ran_labels"], g["ran_ids"]), g["ran_out"])
    eye = np.eye(12, dtype=dt)
    assert np.allclose(R.remove_accidental_negative(g["ran2_logits"], eye, g["ran2_ids"]), g["ran2_out"])
    assert np.allclose(R.sampling_probability_correction(g["ran2_logits"], g["spc_p"]), g["spc_out"], **tol(name))
    for tag in ("ret_a", "ret_b", "ret_c"):
        tau = float(g[f"{tag}_tau"])
        w = g[f"{tag}_w"] if f"{tag}_w" in g else None
        loss, _, _ = R.retrieval_loss(g[f"{tag}_q"], g[f"{tag}_c"], w, None, None, None if tau < 0 else tau, None, dt)
        assert np.allclose(loss, g[f"{tag}_loss"], rtol=1e-5 if name == "f32" else 1e-12)


def test_golden_dnn(gold):
    name, dt, g = gold
    assert list(g["dnn_acts"]) == ["relu", "relu", "None"]                    # dnn.py:17-29
    y, _ = R.dnn(g["dnn_x"], [g["dnn_w0"], g["dnn_w1"], g["dnn_w2"]], [g["dnn_b0"], g["dnn_b1"], g["dnn_b2"]], "relu", dt)
    assert np.allclose(y, g["dnn_y"], **tol(name))


def _ids_from_feats(g, order):
    cols = []
    for k in order:
        v = g[f"feat_{k}"][:, 0]
        if k == "gender":
            ids = np.asarray([{"F": 0, "M": 1}.get(x, -1) for x in v.tolist()], np.int64)
        else:
            n = {"user_id": 50, "movie_id": 40, "age": 7}[k]
            v = v.astype(np.int64)
            ids = np.where((v >= 0) & (v < n), v, -1)
        cols.append(ids)
    return np.stack(cols, axis=1)


def _lin_from_dense_kernel(g, prefix, sparse_order):
    """The reference's Dense(1) kernel over the name-sorted multi-hot, split per column."""
    sizes = {"user_id": 50, "movie_id": 40, "gender": 2, "age": 7}
    out, o = {}, 0
    for nm in sparse_order:
        key = str(nm).replace("_indicator", "")
        out[key] = g[f"{prefix}_lin_kernel"][o:o + sizes[key], 0]
        o += sizes[key]
    return out


def test_golden_keras_models(gold):
    """FactorizationMachine.call / DeepFM.call: slot order = inputs.items() order (deepfm.py:39),
    first-order kernel rows follow DenseFeatures' name-sorted column order."""
    name, dt, g = gold
    order = [str(k) for k in g["cols_order_inputs"]]
    ids = _ids_from_feats(g, order)
    lin = _lin_from_dense_kernel(g, "kfm", g["kfm_sparse_order"])
    tables = [g[f"kfm_table_{k}"] for k in order]
    logit, _ = R.fm_logit(tables, [lin[k] for k in order], g["kfm_lin_bias"][0], ids, dt)
    assert np.allclose(R.sigmoid(logit, dt), g["kfm_prob"], **tol(name))
    lin = _lin_from_dense_kernel(g, "kdfm", g["kfm_sparse_order"])
    tables = [g[f"kdfm_table_{k}"] for k in order]
    prob, _, _, _ = R.deepfm_forward(tables, [lin[k] for k in order], g["kdfm_lin_bias"][0],
                                     [g["kdfm_w0"], g["kdfm_w1"], g["kdfm_w2"]],
                                     [g["kdfm_b0"], g["kdfm_b1"], g["kdfm_b2"]], ids, "relu", dt)
    assert np.allclose(prob, g["kdfm_prob"], **tol(name))


def test_golden_estimator_models(gold):
    """Estimator FM / DeepFM: slot order = embedding_columns list order (estimator fm.py:48)."""
    name, dt, g = gold
    order = [str(k).replace("_embedding", "") for k in g["efm_emb_order"]]
    ids = _ids_from_feats(g, order)
    tables = [g[f"efm_table_{k}"] for k in order]
    lins = [g[f"efm_lin_{k}_indicator"] for k in order]
    logit, _ = R.fm_logit(tables, lins, g["efm_lin_bias"][0], ids, dt)
    assert np.allclose(logit, g["efm_logit"], **tol(name))
    prob, _, _, _ = R.deepfm_forward(tables, lins, g["efm_lin_bias"][0], [g["edfm_w0"], g["edfm_w1"], g["edfm_w2"]],
                                     [g["edfm_b0"], g["edfm_b1"], g["edfm_b2"]], ids, "relu", dt)
    assert np.allclose(prob, g["edfm_prob"], **tol(name))


# ---- the reference's own tests, on the oracle ------------------------------------------------
def test_ref_fm_layer_numpy_formula():
    """tests/keras/test_fm.py:17-26."""
    rng = np.random.default_rng(0)
    e = rng.normal(size=(10, 5, 5)).astype(np.float32)
    x_sum = np.sum(e, axis=1)
    expected = 0.5 * np.sum(np.power(x_sum, 2) - np.sum(np.power(e, 2), axis=1), axis=1, keepdims=True)
    assert np.allclose(R.fm_second_order(e), expected, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("k", [3, 5, 10, 15])
def test_ref_hard_negative_mining_properties(k):
    """tests/keras/test_sbcnm.py:16-41."""
    rng = np.random.RandomState(42)
    logits = rng.uniform(size=(2, 20)).astype(np.float32)
    labels = rng.permutation(np.eye(2, 20).T).T.astype(np.float32)
    ol, oy, _ = R.hard_negative_mining(logits, labels, k)
    assert ol.shape[-1] == k + 1
    assert np.allclose((ol * oy).sum(-1), (logits * labels).sum(-1))
    logits = logits + labels * 1000.0
    ol, oy, _ = R.hard_negative_mining(logits, labels, k)
    assert np.allclose(np.sort(logits, axis=1)[:, -k - 1:], np.sort(ol))


def test_ref_remove_accidental_negative_property():
    """tests/keras/test_sbcnm.py:43-55."""
    rng = np.random.RandomState(42)
    logits = rng.uniform(size=(2, 4)).astype(np.float32)
    labels = rng.permutation(np.eye(2, 4).T).T.astype(np.float32)
    ident = rng.randint(0, 3, size=4)
    out = R.remove_accidental_negative(logits, labels, ident)
    assert np.allclose((out * labels).sum(1), (logits * labels).sum(1))


# ---- analytic backward vs torch autograd (float64, CPU) ----------------------------------------
def test_backward_fm_cross_dense_retrieval_vs_autograd():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((9, 6, 8))
    g = rng.standard_normal((9, 1))
    xt = torch.tensor(x, requires_grad=True)
    y = 0.5 * ((xt.sum(1) ** 2) - (xt ** 2).sum(1)).sum(1, keepdim=True)
    y.backward(torch.tensor(g))
    assert np.allclose(R.fm_second_order_grad(x, g, np.float64), xt.grad.numpy())

    for r in (0, 3):
        B, d = 7, 10
        x0, xx, gg = rng.standard_normal((B, d)), rng.standard_normal((B, d)), rng.standard_normal((B, d))
        b = rng.standard_normal(d)
        w = rng.standard_normal((d, d)) if r == 0 else None
        u = rng.standard_normal((d, r)) if r else None
        v = rng.standard_normal((r, d)) if r else None
        T = lambda a: None if a is None else torch.tensor(a, requires_grad=True)
        x0t, xt2, wt, ut, vt, bt = T(x0), T(xx), T(w), T(u), T(v), T(b)
        prod = (xt2 @ wt if r == 0 else (xt2 @ ut) @ vt) + bt + 0.3 * xt2
        (x0t * prod + xt2).backward(torch.tensor(gg))
        gr = R.cross_grad(x0, xx, gg, w, u, v, b, 0.3, np.float64)
        assert np.allclose(gr["gx0"], x0t.grad.numpy()) and np.allclose(gr["gx"], xt2.grad.numpy())
        assert np.allclose(gr["gb"], bt.grad.numpy())
        if r == 0:
            assert np.allclose(gr["gw"], wt.grad.numpy())
        else:
            assert np.allclose(gr["gu"], ut.grad.numpy()) and np.allclose(gr["gv"], vt.grad.numpy())

    for act in (None, "relu", "sigmoid", "tanh"):
        xx, w, b, gy = rng.standard_normal((6, 5)), rng.standard_normal((5, 4)), rng.standard_normal(4), rng.standard_normal((6, 4))
        xt2, wt, bt = torch.tensor(xx, requires_grad=True), torch.tensor(w, requires_grad=True), torch.tensor(b, requires_grad=True)
        z = xt2 @ wt + bt
        yt = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}.get(act, lambda a: a)(z)
        yt.backward(torch.tensor(gy))
        gx, gw, gb = R.dense_grad(xx, w, R.dense(xx, w, b, act, np.float64), gy, act, np.float64)
        assert np.allclose(gx, xt2.grad.numpy()) and np.allclose(gw, wt.grad.numpy()) and np.allclose(gb, bt.grad.numpy())

    nq, D = 11, 6
    q, c = rng.standard_normal((nq, D)), rng.standard_normal((nq, D))
    w, p, ids = rng.uniform(0.5, 2, nq), rng.uniform(0.1, 1, nq), rng.integers(0, 4, nq)
    qt, ct = torch.tensor(q, requires_grad=True), torch.tensor(c, requires_grad=True)
    s = qt @ ct.T - torch.log(torch.tensor(p))
    dup = torch.tensor((ids[:, None] == ids[None, :]).astype(np.float64) - np.eye(nq))
    s = (s + dup * float(R.MIN_FLOAT)) / 0.7
    loss = (torch.tensor(w) * (torch.logsumexp(s, 1) - torch.diagonal(s))).sum()
    loss.backward()
    ref_loss, _, _ = R.retrieval_loss(q, c, w, p, ids, 0.7, None, np.float64)
    assert np.allclose(ref_loss, loss.item())
    gq, gc = R.retrieval_grad(q, c, w, p, ids, 0.7, np.float64)
    assert np.allclose(gq, qt.grad.numpy()) and np.allclose(gc, ct.grad.numpy())


def test_embed_fm_grad_vs_autograd():
    rng = np.random.default_rng(5)
    rows, D, B = [7, 5, 9], 4, 40
    tables = [rng.standard_normal((r, D)) for r in rows]
    lins = [rng.standard_normal(r) for r in rows]
    ids = np.stack([rng.integers(-1, r + 1, B) for r in rows], axis=1)
    tt = [torch.tensor(t, requires_grad=True) for t in tables]
    lt = [torch.tensor(l, requires_grad=True) for l in lins]
    bias = torch.tensor(0.2, requires_grad=True, dtype=torch.float64)
    embs, lin = [], bias.expand(B)
    for s, r in enumerate(rows):
        col = torch.tensor(ids[:, s])
        ok = (col >= 0) & (col < r)
        safe = torch.where(ok, col, torch.zeros_like(col))
        embs.append(tt[s][safe] * ok[:, None])
        lin = lin + lt[s][safe] * ok
    st = torch.stack(embs, 1)
    logit = lin + 0.5 * ((st.sum(1) ** 2) - (st ** 2).sum(1)).sum(1)
    gl, gs = rng.standard_normal(B), rng.standard_normal((B, len(rows), D))
    ((logit * torch.tensor(gl)).sum() + (st * torch.tensor(gs)).sum()).backward()
    ref_logit, stack = R.fm_logit(tables, lins, 0.2, ids, np.float64)
    assert np.allclose(ref_logit.reshape(-1), logit.detach().numpy())
    gts, gls, gb = R.embed_fm_grad(rows, ids, stack, gl, gs, np.float64)
    for s in range(len(rows)):
        assert np.allclose(gts[s], tt[s].grad.numpy()) and np.allclose(gls[s], lt[s].grad.numpy())
    assert np.allclose(gb, bias.grad.item())
