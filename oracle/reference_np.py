"""ORACLE -- test infrastructure only.  Nothing under deep_recommenders_b200/ may import this.

CPU restatement, in numpy, of the arithmetic on the hot path of
LongmaoTeamTf/deep_recommenders (paths below are relative to the reference repo root).
The reference is pure Python over TensorFlow and TensorFlow is not installed in this
environment (nor on the GPU box), so the reference itself cannot execute; each function
below follows the cited reference lines op for op, with TensorFlow's documented op semantics
(tf.reduce_sum / tf.pow / Dense = x @ W[in,out] + b / embedding lookup with OOV -> zeros /
CategoricalCrossentropy(from_logits, SUM)) filled in.

How it is pinned (see oracle/README.md and tests/test_oracle.py):
  * the reference's own known-answer and property tests are ported onto it
    (tests/keras/test_dcn.py:16-23, test_fm.py:17-26, test_sbcnm.py:16-55,
     tests/estimator/test_fm.py:18-26);
  * the reference's *own source files* are executed under a numpy-backed `tensorflow`
    stand-in (oracle/tf_shim) to generate tests/golden/*.npz, and the oracle must reproduce
    those bit-for-bit in float64 / to 1e-6 in float32;
  * every analytic backward here is checked against torch-CPU autograd of the forward.
Parity UNPINNED by the reference (it has no test for them): embedding-lookup values, all
gradients, Retrieval.call, SamplingProbabilityCorrection.  For those the pin is the shim run
and the float64 twin only.

Every function takes ``dtype`` (np.float32 = the parity target, np.float64 = error budget).
"""
from __future__ import annotations

import numpy as np

MAX_FLOAT = np.finfo(np.float32).max / 100.0     # keras/models/retrieval/sbcnm.py:9
MIN_FLOAT = np.finfo(np.float32).min / 100.0     # keras/models/retrieval/sbcnm.py:10


# ------------------------------------------------------------------------------------------
# Row E: embedding lookup.  keras fm.py:47-51,57-61 / deepfm.py:25-28,39-43 build one
# DenseFeatures(embedding_column) per key; estimator fm.py:48-52 uses input_layer.  For a
# single-valued id this is a row gather; id -1 (OOV) or out of range -> zero vector.
# ------------------------------------------------------------------------------------------
def embedding_lookup(table: np.ndarray, ids: np.ndarray) -> np.ndarray:
    ids = np.asarray(ids, dtype=np.int64)
    ok = (ids >= 0) & (ids < table.shape[0])
    out = np.zeros(ids.shape + (table.shape[1],), dtype=table.dtype)
    out[ok] = table[ids[ok]]
    return out


def stack_embeddings(tables, ids: np.ndarray) -> np.ndarray:
    """tf.stack([lookup_s(ids[:, s]) for s], axis=1) -> [B,S,D]   (deepfm.py:39-44)."""
    return np.stack([embedding_lookup(t, ids[:, s]) for s, t in enumerate(tables)], axis=1)


# Row L: first-order term.  keras fm.py:16-20,26: Dense(1)(multi_hot) with a [sum N, 1] kernel
# == sum_s w_s[id_s] + b (OOV -> all-zero indicator row -> contributes 0).
def linear_term(lin_weights, bias, ids: np.ndarray, dtype=np.float32) -> np.ndarray:
    ids = np.asarray(ids, dtype=np.int64)
    out = np.zeros((ids.shape[0],), dtype=dtype)
    for s, w in enumerate(lin_weights):
        col = ids[:, s]
        ok = (col >= 0) & (col < w.shape[0])
        v = np.zeros(col.shape, dtype=dtype)
        v[ok] = w[col[ok]].astype(dtype)
        out = out + v
    return (out + dtype(bias)).reshape(-1, 1)


# Row F: FM second order.  keras fm.py:28-35 (tf.pow(.,2)) == estimator fm.py:22-26 (tf.square).
def fm_second_order(x: np.ndarray, dtype=np.float32) -> np.ndarray:
    if x.ndim != 3:
        raise ValueError("The rank of `x` should be 3. Got rank = {}.".format(x.ndim))   # estimator fm.py:19-20
    x = x.astype(dtype)
    x_sum = np.sum(x, axis=1)                                  # fm.py:28
    x_square_sum = np.sum(np.power(x, 2), axis=1)              # fm.py:29
    return (dtype(0.5) * np.sum(np.power(x_sum, 2) - x_square_sum, axis=1, keepdims=True)).astype(dtype)  # :31-35


def fm_second_order_grad(x: np.ndarray, g: np.ndarray, dtype=np.float32) -> np.ndarray:
    """d/dx of the above times upstream g [B,1]:  g * (sum_s x - x)."""
    x = x.astype(dtype)
    return (g.reshape(-1, 1, 1).astype(dtype) * (np.sum(x, axis=1, keepdims=True) - x)).astype(dtype)


def fm_logit(tables, lin_weights, bias, ids, dtype=np.float32):
    """FM.call(sparse, stack) = linear + interaction (keras fm.py:37; estimator fm.py:56)."""
    stack = stack_embeddings(tables, ids)
    return linear_term(lin_weights, bias, ids, dtype) + fm_second_order(stack, dtype), stack


def sigmoid(z, dtype=np.float32):
    z = z.astype(dtype)
    return (1.0 / (1.0 + np.exp(-z))).astype(dtype)


# Row D: Dense / DNN tower.  keras deepfm.py:30-34 ; estimator dnn.py:17-29.
def act(z, name):
    if name in (None, "linear"):
        return z
    if name == "relu":
        return np.maximum(z, 0)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    if name == "tanh":
        return np.tanh(z)
    raise ValueError(name)


def dense(x, w, b=None, activation=None, dtype=np.float32):
    z = x.astype(dtype) @ w.astype(dtype)
    if b is not None:
        z = z + b.astype(dtype)
    return act(z, activation).astype(dtype)


def dnn(x, weights, biases, activation="relu", dtype=np.float32):
    """[Dense(u, act) ...] + [Dense(last)] : activation on all but the last layer."""
    h = x
    hs = [x]
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = dense(h, w, b, activation if i < len(weights) - 1 else None, dtype)
        hs.append(h)
    return h, hs


def dense_grad(x, w, y, gy, activation=None, dtype=np.float32):
    """(gx, gw, gb) of y = act(x@w+b) given upstream gy."""
    x, w, y, gy = (a.astype(dtype) for a in (x, w, y, gy))
    if activation == "relu":
        gz = gy * (y > 0)
    elif activation == "sigmoid":
        gz = gy * y * (1 - y)
    elif activation == "tanh":
        gz = gy * (1 - y * y)
    else:
        gz = gy
    return gz @ w.T, x.T @ gz, gz.sum(axis=0)


# Model assembly (row M).  keras deepfm.py:36-47.
def deepfm_forward(tables, lin_weights, bias, dnn_w, dnn_b, ids, activation="relu", dtype=np.float32):
    fm_out, stack = fm_logit(tables, lin_weights, bias, ids, dtype)
    concat = stack.reshape(stack.shape[0], -1)                 # tf.concat(embeddings, axis=1)  deepfm.py:45
    deep, hs = dnn(concat, dnn_w, dnn_b, activation, dtype)    # deepfm.py:30-34,46
    logits = fm_out + deep
    return sigmoid(logits, dtype), logits, stack, hs           # deepfm.py:47


# Row X: Cross.  keras dcn.py:70-88.
def cross(x0, x=None, w=None, b=None, u=None, v=None, diag_scale=0.0, dtype=np.float32):
    if x is None:
        x = x0                                                  # dcn.py:72-73
    if x0.shape[-1] != x.shape[-1]:
        raise ValueError("`x0` and `x` dim mismatch. Got `x0` dim = {} and `x` dim = {}".format(
            x0.shape[-1], x.shape[-1]))                         # dcn.py:75-78
    x0 = x0.astype(dtype)
    x = x.astype(dtype)
    if w is not None:
        prod = x @ w.astype(dtype)                              # dcn.py:80-81
    else:
        prod = (x @ u.astype(dtype)) @ v.astype(dtype)          # dcn.py:83
    if b is not None:
        prod = prod + b.astype(dtype)
    if diag_scale:
        prod = prod + dtype(diag_scale) * x                     # dcn.py:85-86
    return (x0 * prod + x).astype(dtype), prod                  # dcn.py:88


def cross_grad(x0, x, g, w=None, u=None, v=None, b=None, diag_scale=0.0, dtype=np.float32):
    """Analytic backward of `cross`.  Returns dict of gradients."""
    x0, x, g = (a.astype(dtype) for a in (x0, x, g))
    _, prod = cross(x0, x, w, b, u, v, diag_scale, dtype)
    h = g * x0
    out = {"gx0": g * prod, "gb": h.sum(axis=0)}
    if w is not None:
        out["gw"] = x.T @ h
        out["gx"] = h @ w.astype(dtype).T + dtype(diag_scale) * h + g
    else:
        xu = x @ u.astype(dtype)
        t = h @ v.astype(dtype).T
        out["gv"] = xu.T @ h
        out["gu"] = x.T @ t
        out["gx"] = t @ u.astype(dtype).T + dtype(diag_scale) * h + g
    return out


# Rows R, R1, R2, R3: two-tower task.  keras/models/retrieval/sbcnm.py.
def gather_elements_along_row(data, column_indices):           # sbcnm.py:15-30
    assert data.shape[0] == column_indices.shape[0]
    return np.take_along_axis(data, column_indices, axis=1)


def hard_negative_mining(logits, labels, num_hard_negatives):  # sbcnm.py:33-49
    num_sampled = min(num_hard_negatives + 1, logits.shape[1])
    masked = logits + labels * np.float32(MAX_FLOAT)
    # tf.nn.top_k(sorted=False): any order; we use descending value, ties -> lower index
    idx = np.argsort(-masked, axis=1, kind="stable")[:, :num_sampled]
    return gather_elements_along_row(logits, idx), gather_elements_along_row(labels, idx), idx


def remove_accidental_negative(logits, labels, identifiers):   # sbcnm.py:52-75
    identifiers = np.asarray(identifiers).reshape(-1, 1)
    positive_indices = np.argmax(labels, axis=1)
    positive_identifier = identifiers[positive_indices]        # [B,1]
    duplicate = (positive_identifier == identifiers.T).astype(labels.dtype)
    duplicate = duplicate - labels
    return logits + duplicate * np.float32(MIN_FLOAT)


def sampling_probability_correction(logits, p):                # sbcnm.py:78-86
    return logits - np.log(p).astype(logits.dtype)


def categorical_crossentropy_sum(labels, scores, sample_weight=None, dtype=np.float32):
    """tf.keras.losses.CategoricalCrossentropy(from_logits=True, reduction=SUM) (sbcnm.py:100-102)."""
    s = scores.astype(dtype)
    m = s.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(s - m).sum(axis=1, keepdims=True))
    per_row = -(labels.astype(dtype) * (s - lse)).sum(axis=1)
    if sample_weight is not None:
        per_row = per_row * np.asarray(sample_weight, dtype=dtype).reshape(-1)
    return dtype(per_row.sum())


def retrieval_loss(q, c, sample_weight=None, candidate_sampling_probability=None, candidate_ids=None,
                   temperature=None, num_hard_negatives=None, dtype=np.float32):
    """Retrieval.call (sbcnm.py:120-151) with the *intended* semantics of its three optional
    branches (:136-146 are broken in the reference; the helper layers define the intent)."""
    scores = q.astype(dtype) @ c.astype(dtype).T                # :129
    labels = np.eye(scores.shape[0], scores.shape[1], dtype=dtype)   # :134
    if candidate_sampling_probability is not None:
        scores = sampling_probability_correction(scores, np.asarray(candidate_sampling_probability, dtype=dtype))
    if candidate_ids is not None:
        scores = remove_accidental_negative(scores, labels, candidate_ids).astype(dtype)
    if num_hard_negatives is not None:
        scores, labels, _ = hard_negative_mining(scores, labels, num_hard_negatives)
    if temperature is not None:
        scores = scores / dtype(temperature)                    # :148-149
    return categorical_crossentropy_sum(labels, scores, sample_weight, dtype), scores, labels


def retrieval_grad(q, c, sample_weight=None, candidate_sampling_probability=None, candidate_ids=None,
                   temperature=None, dtype=np.float64):
    """(gq, gc) of the no-hard-negative loss: G = w_i (softmax - eye) / tau ; gq = G c ; gc = G^T q."""
    _, scores, labels = retrieval_loss(q, c, sample_weight, candidate_sampling_probability, candidate_ids,
                                       temperature, None, dtype)
    s = scores.astype(dtype)
    p = np.exp(s - s.max(axis=1, keepdims=True))
    p = p / p.sum(axis=1, keepdims=True)
    G = p - labels
    if sample_weight is not None:
        G = G * np.asarray(sample_weight, dtype=dtype).reshape(-1, 1)
    if temperature is not None:
        G = G / dtype(temperature)
    return G @ c.astype(dtype), G.T @ q.astype(dtype)


# Backward of the fused embedding + linear + FM block (SURVEY.md 8a rows E/L/F backward).
def embed_fm_grad(tables_rows, ids, stack, g_logit, g_stack, dtype=np.float64):
    """Dense gradients TF autodiff + IndexedSlices densification would give.
    Returns (grad_tables list, grad_lin list, grad_bias)."""
    ids = np.asarray(ids, dtype=np.int64)
    B, S, D = stack.shape
    st = stack.astype(dtype)
    gl = np.zeros((B,), dtype) if g_logit is None else np.asarray(g_logit, dtype).reshape(B)
    dE = gl[:, None, None] * (st.sum(axis=1, keepdims=True) - st)
    if g_stack is not None:
        dE = dE + np.asarray(g_stack, dtype).reshape(B, S, D)
    gts, gls = [], []
    for s, rows in enumerate(tables_rows):
        gt = np.zeros((rows, D), dtype)
        glin = np.zeros((rows,), dtype)
        col = ids[:, s]
        ok = (col >= 0) & (col < rows)
        np.add.at(gt, col[ok], dE[ok, s, :])
        np.add.at(glin, col[ok], gl[ok])
        gts.append(gt)
        gls.append(glin)
    return gts, gls, dtype(gl.sum())


# ------------------------------------------------------------------------------------------
# SURVEY.md 8(f) "next" rows.  The arithmetic of these lives in TensorFlow (un-vendored), so each
# function restates TF's documented op; parity is UNPINNED by the reference except where a
# reference test is cited.
# ------------------------------------------------------------------------------------------
def embedding_bag(table, flat_ids, row_splits, combiner="mean", dtype=np.float32):
    """tf.nn.safe_embedding_lookup_sparse(table, ids, combiner=...) for a ragged slot (embedding_column over a
    VarLenFeature, datasets/movielens.py:122; default combiner "mean"): ids < 0 (OOV) are pruned, a bag with no
    valid id gives the zero vector.  Ids >= rows cannot occur in TF (hash / vocab columns are in range by
    construction); here they are pruned like OOV.  Sequential accumulation, then a true division."""
    flat_ids = np.asarray(flat_ids, dtype=np.int64)
    row_splits = np.asarray(row_splits, dtype=np.int64)
    B, D = row_splits.size - 1, table.shape[1]
    out = np.zeros((B, D), dtype=dtype)
    for b in range(B):
        acc, cnt = np.zeros((D,), dtype=dtype), 0
        for j in range(row_splits[b], row_splits[b + 1]):
            i = flat_ids[j]
            if 0 <= i < table.shape[0]:
                acc = acc + table[i].astype(dtype)
                cnt += 1
        if cnt:
            den = {"sum": 1.0, "mean": float(cnt), "sqrtn": float(np.sqrt(dtype(cnt)))}[combiner]
            out[b] = acc / dtype(den)
    return out


def embedding_bag_grad(rows, flat_ids, row_splits, g_out, combiner="mean", dtype=np.float64):
    """Dense gradient of embedding_bag w.r.t. the table: every valid id of bag b receives g_out[b] / den(b)."""
    flat_ids = np.asarray(flat_ids, dtype=np.int64)
    row_splits = np.asarray(row_splits, dtype=np.int64)
    B, D = g_out.shape
    gt = np.zeros((rows, D), dtype=dtype)
    for b in range(B):
        seg = flat_ids[row_splits[b]:row_splits[b + 1]]
        seg = seg[(seg >= 0) & (seg < rows)]
        if seg.size == 0:
            continue
        den = {"sum": 1.0, "mean": float(seg.size), "sqrtn": float(np.sqrt(seg.size))}[combiner]
        np.add.at(gt, seg, np.asarray(g_out[b], dtype) / den)
    return gt


def _f32_beta(b):
    """The hyper-parameters reach TensorFlow's kernels as float32 tensors: 0.999 is 0.99900001287..., which moves
    1 - beta2 by 1.3e-5 relative -- above the 1e-5 parity bar, so the oracle must round them the same way."""
    return float(np.float32(b))


def adam_lr_t(lr, step, beta1=0.9, beta2=0.999):
    """Keras optimizer_v2 Adam / tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), t = 1, 2, ...
    (betas as float32 values; evaluated in float64 here, in float32 by TF: the result is the same to ~1e-5 relative
    for small t, where 1 - b2^t cancels)."""
    b1, b2 = _f32_beta(beta1), _f32_beta(beta2)
    return float(np.float32(lr)) * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)


def adam_dense(p, g, m, v, lr_t, beta1=0.9, beta2=0.999, eps=1e-7, dtype=np.float32):
    """TensorFlow's ApplyAdam functor (the optimizer both reference examples use:
    examples/train_deepfm_on_movielens_keras.py:44, examples/train_fm_on_movielens_estimator.py:51):
        m += (g - m)(1 - b1);  v += (g*g - v)(1 - b2);  p -= lr_t * m / (sqrt(v) + eps)
    with (1 - b) formed in float32 like the functor's `T(1) - beta()`.
    Returns (p, m, v).  TF applies it densely to embedding variables too (module docstring of optim.cu)."""
    p, g, m, v = (np.asarray(a, dtype=dtype) for a in (p, g, m, v))
    omb1 = dtype(np.float32(1.0) - np.float32(beta1))
    omb2 = dtype(np.float32(1.0) - np.float32(beta2))
    m = m + (g - m) * omb1
    v = v + (g * g - v) * omb2
    p = p - (m * dtype(np.float32(lr_t))) / (np.sqrt(v) + dtype(np.float32(eps)))
    return p, m, v


def adam_rows_lazy(p, g_rows, m, v, touched, lr_t, beta1=0.9, beta2=0.999, eps=1e-7, dtype=np.float32):
    """Row-sparse ("lazy") Adam: the ApplyAdam functor on the `touched` rows only, with the row gradient already
    summed over duplicate ids; untouched rows keep p, m and v.  This is tfa.optimizers.LazyAdam's semantics,
    NOT tf.keras.optimizers.Adam's (which decays m, v of every row every step): an opt-in deviation for tables
    whose dense pass would dominate the step."""
    p, m, v = (np.array(a, dtype=dtype, copy=True) for a in (p, m, v))
    t = np.unique(np.asarray(touched, dtype=np.int64))
    t = t[(t >= 0) & (t < p.shape[0])]
    p[t], m[t], v[t] = adam_dense(p[t], np.asarray(g_rows, dtype)[t], m[t], v[t], lr_t, beta1, beta2, eps, dtype)
    return p, m, v


def top_k(scores, k):
    """tf.math.top_k(scores, k): values descending; equal values -> lower index first.  -> (values, int32 idx)."""
    scores = np.asarray(scores)
    if k > scores.shape[1]:
        raise ValueError("input must have at least k columns")
    idx = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, idx, axis=1), idx.astype(np.int32)


def take_long_axis(arr, indices):                              # factorized_top_k.py:26-41
    return np.take_along_axis(np.asarray(arr), np.asarray(indices), axis=1)


def exclude(scores, identifiers, exclude_ids, k):              # factorized_top_k.py:44-67
    scores, identifiers, exclude_ids = np.asarray(scores), np.asarray(identifiers), np.asarray(exclude_ids)
    isin = (identifiers[:, :, None] == exclude_ids[:, None, :]).any(-1)
    adjusted = scores - isin.astype(np.float32) * np.float32(1.0e5)
    k = min(k, scores.shape[1])
    _, idx = top_k(adjusted, k)
    return take_long_axis(scores, idx), take_long_axis(identifiers, idx)


def brute_force_topk(queries, candidates, identifiers, k, dtype=np.float32):   # factorized_top_k.py:322-334
    scores = queries.astype(dtype) @ candidates.astype(dtype).T
    vals, idx = top_k(scores, k)
    ident = np.arange(candidates.shape[0]) if identifiers is None else np.asarray(identifiers)
    return vals, ident[idx]


def streaming_topk(queries, candidate_batches, identifier_batches, k, handle_incomplete_batches=True,
                   dtype=np.float32):
    """Streaming.call (factorized_top_k.py:180-262): per-batch top-k (:199-211), then the reduction that
    re-selects top-k over [state | batch top-k] (:213-230); candidates are numbered by a running counter when
    no identifiers are given (:240-249)."""
    nq = queries.shape[0]
    st_s = np.zeros((nq, 0), dtype=dtype)
    st_i = None
    counter = 0
    for bi, cand in enumerate(candidate_batches):
        n = cand.shape[0]
        ident = np.arange(counter, counter + n, dtype=np.int32) if identifier_batches is None \
            else np.asarray(identifier_batches[bi])
        counter += n
        scores = queries.astype(dtype) @ cand.astype(dtype).T
        k_ = min(k, n) if handle_incomplete_batches else k
        s, idx = top_k(scores, k_)
        xi = ident[idx]
        if st_i is None:
            st_i = np.zeros((nq, 0), dtype=xi.dtype)
        js, ji = np.concatenate([st_s, s], axis=1), np.concatenate([st_i, xi], axis=1)
        k_ = min(k, js.shape[1]) if handle_incomplete_batches else k
        st_s, idx = top_k(js, k_)
        st_i = np.take_along_axis(ji, idx, axis=1)
    return st_s, st_i


def topk_categorical_accuracy(y_true, y_pred, k):
    """tf.keras.metrics.TopKCategoricalAccuracy(k) on one batch = mean(in_top_k(y_pred, argmax(y_true), k));
    tf.math.in_top_k counts a target as inside when fewer than k predictions are STRICTLY larger."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    tgt = y_true.argmax(axis=1)
    tv = np.take_along_axis(y_pred, tgt[:, None], axis=1)
    return float(((y_pred > tv).sum(axis=1) < k).mean())


def factorized_topk_metric(queries, true_candidates, top_k_scores, ks):       # factorized_top_k.py:487-511
    pos = (queries * true_candidates).sum(axis=1, keepdims=True)
    y_true = np.concatenate([np.ones_like(pos), np.zeros_like(top_k_scores)], axis=1)
    y_pred = np.concatenate([pos, top_k_scores], axis=1)
    return [topk_categorical_accuracy(y_true, y_pred, k) for k in ks]
