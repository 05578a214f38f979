"""torch.autograd.Function wrappers over the C-ABI (include/deeprec_b200.h).

PyTorch is plumbing here: it owns device memory and streams and gives the layers an autograd
graph.  All arithmetic happens in libdeeprec_b200.so.  Every function raises if its inputs are
not CUDA fp32 tensors -- there is no CPU or eager-PyTorch fallback on this path.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check

ACT_CODES = {None: 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3}


def act_code(act) -> int:
    if callable(act):
        act = getattr(act, "__name__", str(act))
    if act not in ACT_CODES:
        raise ValueError(f"unsupported activation {act!r}; supported: {sorted(k for k in ACT_CODES if k)}")
    return ACT_CODES[act]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise _lib.DeepRecError(f"{name} is on {t.device}: the hot path runs on CUDA only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _ids(t: torch.Tensor, name: str = "ids") -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.DeepRecError(f"{name} is on {t.device}: the hot path runs on CUDA only (no CPU fallback)")
    if t.dtype not in (torch.int64, torch.int32):
        raise TypeError(f"{name}: expected int64 or int32 ids, got {t.dtype}")
    return t.contiguous()


# ------------------------------------------------------------------------------------------
# Rows E + L + F: fused multi-slot gather + linear + FM
# ------------------------------------------------------------------------------------------
class EmbedFM(torch.autograd.Function):
    """stack, logit = EmbedFM(ids, weight_arena, linear_arena, bias, meta).

    `meta` carries the device pointer arrays (table_ptrs, lin_ptrs, rows) of the collection.
    Backward either scatter-adds into dense .grad buffers (meta.sparse_lr is None) or applies
    the fused sparse SGD update in place (meta.sparse_lr = lr) and returns no table gradient.
    """

    @staticmethod
    def forward(ctx, ids, weight, linear, bias, meta, want_logit: bool):
        lib = _lib.load()
        ids = _ids(ids)
        B, S = ids.shape
        D = meta.dim
        stack = torch.empty((B, S, D), device=weight.device, dtype=torch.float32)
        sum_e = torch.empty((B, D), device=weight.device, dtype=torch.float32) if want_logit else None
        logit = torch.empty((B,), device=weight.device, dtype=torch.float32) if want_logit else None
        tp, lp, rows = meta.pointers(weight, linear)
        has_lin = want_logit and meta.with_linear
        check(lib.dr_embed_fm_fwd(tp.data_ptr(), lp.data_ptr() if has_lin else None,
                                  rows.data_ptr(), ids.data_ptr(), ids.element_size(),
                                  _ptr(bias) if want_logit else None, B, S, D, meta.row_stride, meta.lin_stride,
                                  meta.flags if has_lin else 0, stack.data_ptr(), _ptr(sum_e), _ptr(logit), _stream()), "dr_embed_fm_fwd")
        ctx.meta = meta
        ctx.want_logit = want_logit
        ctx.has_lin = has_lin
        ctx.has_bias = bias is not None
        ctx.save_for_backward(ids, stack, sum_e, weight, linear, bias)
        if want_logit:
            return stack, logit
        dummy = stack.new_zeros((0,))
        ctx.mark_non_differentiable(dummy)
        return stack, dummy

    @staticmethod
    def backward(ctx, g_stack, g_logit):
        lib = _lib.load()
        ids, stack, sum_e, weight, linear, bias = ctx.saved_tensors
        meta = ctx.meta
        B, S = ids.shape
        D = meta.dim
        want_logit = ctx.want_logit
        g_stack = None if g_stack is None else _f32(g_stack, "g_stack")
        g_logit = _f32(g_logit, "g_logit") if (want_logit and g_logit is not None) else None
        if g_stack is None and g_logit is None:
            return None, None, None, None, None, None
        fused = meta.sparse_lr is not None
        if fused:
            gw, gl, gb = weight, linear, bias
            scale = -float(meta.sparse_lr)
            tp, lp, rows = meta.pointers(weight, linear)
        else:
            gw = torch.zeros_like(weight)
            gl = torch.zeros_like(linear) if (linear is not None and g_logit is not None) else None
            gb = torch.zeros_like(bias) if (bias is not None and g_logit is not None) else None
            scale = 1.0
            tp, lp, rows = meta.pointers(gw, gl, cache=False)
        lin_grad = ctx.has_lin and g_logit is not None
        with torch.no_grad():
            check(lib.dr_embed_fm_bwd(ids.data_ptr(), ids.element_size(), rows.data_ptr(),
                                      stack.data_ptr(), _ptr(sum_e), _ptr(g_logit), _ptr(g_stack),
                                      B, S, D, meta.row_stride, meta.lin_stride,
                                      meta.flags if lin_grad else 0, tp.data_ptr(),
                                      lp.data_ptr() if lin_grad else None,
                                      _ptr(gb) if g_logit is not None else None, scale, _stream()),
                  "dr_embed_fm_bwd")
        if fused:
            return None, None, None, None, None, None
        return None, gw, gl, gb, None, None


class Gather(torch.autograd.Function):
    """rows = table[ids] for one table (two-tower user / item tower); OOV id -> zero row."""

    @staticmethod
    def forward(ctx, table, ids, sparse_lr):
        lib = _lib.load()
        table = _f32(table, "table")
        ids = _ids(ids)
        n = ids.numel()
        D = table.shape[1]
        out = torch.empty((n, D), device=table.device, dtype=torch.float32)
        check(lib.dr_gather_fwd(table.data_ptr(), table.shape[0], ids.data_ptr(), ids.element_size(), n, D,
                                out.data_ptr(), _stream()), "dr_gather_fwd")
        ctx.save_for_backward(table, ids)
        ctx.sparse_lr = sparse_lr
        return out.view(*ids.shape, D)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        table, ids = ctx.saved_tensors
        g = _f32(g, "g").view(-1, table.shape[1])
        fused = ctx.sparse_lr is not None
        tgt = table if fused else torch.zeros_like(table)
        with torch.no_grad():
            check(lib.dr_scatter_add(tgt.data_ptr(), table.shape[0], ids.data_ptr(), ids.element_size(),
                                     ids.numel(), table.shape[1], g.data_ptr(),
                                     -float(ctx.sparse_lr) if fused else 1.0, _stream()), "dr_scatter_add")
        return (None if fused else tgt), None, None


# ------------------------------------------------------------------------------------------
# Row F standalone
# ------------------------------------------------------------------------------------------
class FMInteraction(torch.autograd.Function):
    """0.5 * sum_d((sum_s x)^2 - sum_s x^2), keepdims -> [B, 1]."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x = _f32(x, "x")
        if x.dim() != 3:
            raise ValueError("The rank of `x` should be 3. Got rank = {}.".format(x.dim()))
        B, S, D = x.shape
        out = torch.empty((B,), device=x.device, dtype=torch.float32)
        check(lib.dr_fm_fwd(x.data_ptr(), B, S, D, out.data_ptr(), _stream()), "dr_fm_fwd")
        ctx.save_for_backward(x)
        return out.view(B, 1)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        B, S, D = x.shape
        g = _f32(g, "g").reshape(B)
        gx = torch.empty_like(x)
        check(lib.dr_fm_bwd(x.data_ptr(), g.data_ptr(), B, S, D, gx.data_ptr(), _stream()), "dr_fm_bwd")
        return gx


# ------------------------------------------------------------------------------------------
# Row D: Dense
# ------------------------------------------------------------------------------------------
class DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act: int):
        lib = _lib.load()
        x = _f32(x, "x")
        w = _f32(w, "kernel")
        b = None if b is None else _f32(b, "bias")
        lead = x.shape[:-1]
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        if w.shape[0] != K:
            raise ValueError(f"Dense: input dim {K} does not match kernel {tuple(w.shape)}")
        N = w.shape[1]
        _lib.ensure_gemm_workspace(M, K, N, x.device)
        y = torch.empty((M, N), device=x.device, dtype=torch.float32)
        check(lib.dr_dense_fwd(x2.data_ptr(), w.data_ptr(), _ptr(b), M, K, N, act, y.data_ptr(), _stream()),
              "dr_dense_fwd")
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.save_for_backward(x2, w, y)
        ctx.lead = lead
        return y.view(*lead, N)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, w, y = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[1]
        gy = _f32(gy, "gy").reshape(M, N)
        need_x = ctx.needs_input_grad[0]
        gz_ws = torch.empty_like(gy) if ctx.act != 0 else None
        gx = torch.empty((M, K), device=gy.device, dtype=torch.float32) if need_x else None
        gw = torch.empty_like(w)
        gb = torch.empty((N,), device=gy.device, dtype=torch.float32) if ctx.has_bias else None
        check(lib.dr_dense_bwd(x2.data_ptr(), w.data_ptr(), y.data_ptr(), gy.data_ptr(), M, K, N, ctx.act,
                               _ptr(gz_ws), _ptr(gx), gw.data_ptr(), _ptr(gb), _stream()), "dr_dense_bwd")
        return (gx.view(*ctx.lead, K) if need_x else None), gw, gb, None


# ------------------------------------------------------------------------------------------
# Row X: Cross
# ------------------------------------------------------------------------------------------
class CrossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, x, w, uk, vk, b, alpha: float, same: bool):
        lib = _lib.load()
        x0 = _f32(x0, "x0")
        x = x0 if same else _f32(x, "x")
        d = x0.shape[-1]
        lead = x0.shape[:-1]
        x0_2 = x0.reshape(-1, d)
        x_2 = x.reshape(-1, d)
        B = x0_2.shape[0]
        r = 0 if w is not None else uk.shape[1]
        _lib.ensure_gemm_workspace(B, d, d, x0.device)
        y = torch.empty((B, d), device=x0.device, dtype=torch.float32)
        u = torch.empty((B, d), device=x0.device, dtype=torch.float32)
        xu = torch.empty((B, r), device=x0.device, dtype=torch.float32) if r else None
        check(lib.dr_cross_fwd(x0_2.data_ptr(), x_2.data_ptr(), _ptr(w), _ptr(uk), _ptr(vk), _ptr(b),
                               float(alpha), B, d, r, _ptr(xu), u.data_ptr(), y.data_ptr(), _stream()),
              "dr_cross_fwd")
        ctx.alpha = float(alpha)
        ctx.same = same
        ctx.r = r
        ctx.lead = lead
        ctx.save_for_backward(x0_2, x_2, w, uk, vk, b, u, xu)
        return y.view(*lead, d)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x0, x, w, uk, vk, b, u, xu = ctx.saved_tensors
        B, d = x0.shape
        r = ctx.r
        g = _f32(g, "g").reshape(B, d)
        dev = g.device
        h = torch.empty((B, d), device=dev, dtype=torch.float32)
        t = torch.empty((B, r), device=dev, dtype=torch.float32) if r else None
        gx0 = torch.empty((B, d), device=dev, dtype=torch.float32)
        gx = torch.empty((B, d), device=dev, dtype=torch.float32)
        gw = torch.empty_like(w) if w is not None else None
        guk = torch.empty_like(uk) if uk is not None else None
        gvk = torch.empty_like(vk) if vk is not None else None
        gb = torch.empty_like(b) if b is not None else None
        check(lib.dr_cross_bwd(x0.data_ptr(), x.data_ptr(), _ptr(w), _ptr(uk), _ptr(vk), ctx.alpha,
                               u.data_ptr(), _ptr(xu), g.data_ptr(), B, d, r, h.data_ptr(), _ptr(t),
                               gx0.data_ptr(), gx.data_ptr(), _ptr(gw), _ptr(guk), _ptr(gvk), _ptr(gb),
                               _stream()), "dr_cross_bwd")
        if ctx.same:
            gx0 = gx0 + gx
            gx = None
        else:
            gx = gx.view(*ctx.lead, d)
        return gx0.view(*ctx.lead, d), gx, gw, guk, gvk, gb, None, None


# ------------------------------------------------------------------------------------------
# Row R: in-batch softmax
# ------------------------------------------------------------------------------------------
# Retrieval losses with at least this many scores (nq * nc) run the tensor-core form (score blocks materialised in a
# scratch buffer of at most SOFTMAX_TC_WS_BYTES); smaller ones the fused FFMA kernels that never materialise scores.
SOFTMAX_TC_MIN_SCORES = 1 << 22
SOFTMAX_TC_WS_BYTES = 2 << 30


class InBatchSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, c, sample_weight, sampling_prob, cand_ids, inv_tau: float):
        lib = _lib.load()
        q = _f32(q, "query_embeddings")
        c = _f32(c, "candidate_embeddings")
        if q.dim() != 2 or c.dim() != 2 or q.shape[1] != c.shape[1]:
            raise ValueError(f"Retrieval: embeddings must be [nq,D] and [nc,D], got {tuple(q.shape)} {tuple(c.shape)}")
        nq, D = q.shape
        nc = c.shape[0]
        w = None if sample_weight is None else _f32(sample_weight, "sample_weight").reshape(nq)
        p = None if sampling_prob is None else _f32(sampling_prob, "candidate_sampling_probability").reshape(nc)
        ids = None
        if cand_ids is not None:
            ids = cand_ids.to(torch.int64).contiguous().reshape(nc)
        lse = torch.empty((nq,), device=q.device, dtype=torch.float32)
        loss = torch.empty((1,), device=q.device, dtype=torch.float32)
        ws = None
        if nq * nc >= SOFTMAX_TC_MIN_SCORES and D >= 32:
            # tensor-core form: score blocks through the tcgen05 GEMM, kept in a scratch buffer for the backward
            rows = min(nq, max(128, SOFTMAX_TC_WS_BYTES // (4 * nc) // 128 * 128))
            ws = torch.empty((rows, nc), device=q.device, dtype=torch.float32)
            check(lib.dr_inbatch_softmax_fwd_ws(q.data_ptr(), c.data_ptr(), _ptr(w), _ptr(p), _ptr(ids), float(inv_tau),
                                                nq, nc, D, ws.data_ptr(), rows, lse.data_ptr(), loss.data_ptr(), _stream()),
                  "dr_inbatch_softmax_fwd_ws")
        else:
            check(lib.dr_inbatch_softmax_fwd(q.data_ptr(), c.data_ptr(), _ptr(w), _ptr(p), _ptr(ids), float(inv_tau),
                                             nq, nc, D, lse.data_ptr(), loss.data_ptr(), _stream()),
                  "dr_inbatch_softmax_fwd")
        ctx.inv_tau = float(inv_tau)
        ctx.ws = ws
        ctx.save_for_backward(q, c, w, p, ids, lse)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gloss):
        lib = _lib.load()
        q, c, w, p, ids, lse = ctx.saved_tensors
        nq, D = q.shape
        nc = c.shape[0]
        gl = _f32(gloss, "gloss").reshape(1)
        gq = torch.empty_like(q)
        gc = torch.empty_like(c)
        ws = ctx.ws
        if ws is not None:
            ctx.ws = None          # consumed (overwritten with the gradient factor) by this call
            check(lib.dr_inbatch_softmax_bwd_ws(q.data_ptr(), c.data_ptr(), _ptr(w), _ptr(p), _ptr(ids), ctx.inv_tau,
                                                nq, nc, D, lse.data_ptr(), gl.data_ptr(), ws.data_ptr(), ws.shape[0], 1,
                                                gq.data_ptr(), gc.data_ptr(), _stream()), "dr_inbatch_softmax_bwd_ws")
        else:
            check(lib.dr_inbatch_softmax_bwd(q.data_ptr(), c.data_ptr(), _ptr(w), _ptr(p), _ptr(ids), ctx.inv_tau,
                                             nq, nc, D, lse.data_ptr(), gl.data_ptr(), gq.data_ptr(), gc.data_ptr(),
                                             _stream()), "dr_inbatch_softmax_bwd")
        return gq, gc, None, None, None, None


def scores(q, c, sampling_prob=None, cand_ids=None) -> torch.Tensor:
    """Materialised [nq, nc] logits with R1/R2 corrections (no temperature); not differentiable."""
    lib = _lib.load()
    q = _f32(q, "q")
    c = _f32(c, "c")
    nq, D = q.shape
    nc = c.shape[0]
    p = None if sampling_prob is None else _f32(sampling_prob, "p").reshape(nc)
    ids = None if cand_ids is None else cand_ids.to(torch.int64).contiguous().reshape(nc)
    _lib.ensure_gemm_workspace(nq, D, nc, q.device)
    out = torch.empty((nq, nc), device=q.device, dtype=torch.float32)
    check(lib.dr_scores_fwd(q.data_ptr(), c.data_ptr(), _ptr(p), _ptr(ids), nq, nc, D, out.data_ptr(), _stream()),
          "dr_scores_fwd")
    return out


def hard_negative_topk(logits: torch.Tensor, k: int):
    lib = _lib.load()
    logits = _f32(logits, "logits")
    nq, nc = logits.shape
    out_l = torch.empty((nq, k), device=logits.device, dtype=torch.float32)
    out_y = torch.empty((nq, k), device=logits.device, dtype=torch.float32)
    out_i = torch.empty((nq, k), device=logits.device, dtype=torch.int32)
    check(lib.dr_hard_negative_topk(logits.data_ptr(), nq, nc, k, out_l.data_ptr(), out_y.data_ptr(),
                                    out_i.data_ptr(), _stream()), "dr_hard_negative_topk")
    return out_l, out_y, out_i


def bce_with_logits(z: torch.Tensor, y: torch.Tensor, z_add: torch.Tensor = None):
    """(loss[1], grad_z[B], prob[B]) of mean binary cross-entropy on logits z (+ z_add), one kernel."""
    lib = _lib.load()
    z = _f32(z, "logits").reshape(-1)
    z_add = None if z_add is None else _f32(z_add, "z_add").reshape(-1)
    y = _f32(y, "labels").reshape(-1)
    B = z.numel()
    prob = torch.empty_like(z)
    gz = torch.empty_like(z)
    loss = torch.empty((1,), device=z.device, dtype=torch.float32)
    check(lib.dr_bce_logits_fwd_bwd(z.data_ptr(), _ptr(z_add), y.data_ptr(), B, prob.data_ptr(), loss.data_ptr(), gz.data_ptr(),
                                    _stream()), "dr_bce_logits_fwd_bwd")
    return loss, gz, prob


def sgd_step_(p: torch.Tensor, g: torch.Tensor, lr: float) -> None:
    lib = _lib.load()
    check(lib.dr_sgd_step(p.data_ptr(), g.data_ptr(), p.numel(), float(lr), _stream()), "dr_sgd_step")


# ------------------------------------------------------------------------------------------
# SURVEY 8(f) "next" rows: Adam, id pipeline on the device, multi-valued slots, row top-k
# ------------------------------------------------------------------------------------------
COMBINER_CODES = {"sum": 0, "mean": 1, "sqrtn": 2}


def adam_step_(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr_t: float,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7, zero_grad: bool = True,
               lr_t_dev: torch.Tensor = None) -> None:
    """In-place dense Adam over a flat fp32 buffer (TensorFlow's ApplyAdam arithmetic; lr_t carries the bias
    correction, or lr_t_dev -- a 1-element CUDA float kept by `AdamClock` -- overrides it)."""
    lib = _lib.load()
    for name, t in (("p", p), ("g", g), ("m", m), ("v", v)):
        _f32(t, name)
        if not t.is_contiguous() or t.numel() != p.numel():
            raise ValueError(f"adam_step_: {name} must be contiguous with {p.numel()} elements")
    check(lib.dr_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr_t),
                           float(beta1), float(beta2), float(eps), int(bool(zero_grad)), _ptr(lr_t_dev), _stream()),
          "dr_adam_step")


class AdamClock:
    """Device-resident step counter t and bias-corrected rate lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
    (Keras optimizer_v2 Adam._prepare_local), advanced by one kernel so a captured train step needs no host scalar."""

    def __init__(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7, device="cuda",
                 history: int = 0):
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(beta1), float(beta2), float(eps)
        self.step = torch.zeros((1,), dtype=torch.int64, device=device)
        self.lr_t = torch.zeros((1,), dtype=torch.float32, device=device)
        # ring of the last `history` bias-corrected rates (dr_adam_advance_hist): what the row-sparse TF-equal Adam
        # replays the skipped steps of a row with
        self.lr_hist = torch.zeros((history,), dtype=torch.float32, device=device) if history else None

    def advance(self) -> None:
        lib = _lib.load()
        if self.lr_hist is not None:
            check(lib.dr_adam_advance_hist(self.step.data_ptr(), self.lr, self.beta1, self.beta2, self.lr_t.data_ptr(),
                                           self.lr_hist.data_ptr(), self.lr_hist.numel(), _stream()), "dr_adam_advance_hist")
            return
        check(lib.dr_adam_advance(self.step.data_ptr(), self.lr, self.beta1, self.beta2, self.lr_t.data_ptr(),
                                  _stream()), "dr_adam_advance")


def lazy_adam_rows_(ids: torch.Tensor, rows: torch.Tensor, slot_offsets: torch.Tensor, D: int, row_stride: int,
                    flags: int, p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
                    stamp: torch.Tensor, clock: AdamClock, lin=None) -> None:
    """Row-sparse Adam over the rows `ids` [B, S] touch (see dr_lazy_adam_rows).  lin = (p, g, m, v) first-order
    arrays for the split layout, None when the weight rides in the row."""
    lib = _lib.load()
    ids = _ids(ids)
    B, S = ids.shape
    if stamp.dtype != torch.int32 or not stamp.is_cuda:
        raise TypeError("lazy_adam_rows_: stamp must be a CUDA int32 tensor [total rows]")
    lp = [None] * 4 if lin is None else [_f32(t, "lin").data_ptr() for t in lin]
    check(lib.dr_lazy_adam_rows(ids.data_ptr(), ids.element_size(), B, S, int(D), rows.data_ptr(),
                                slot_offsets.data_ptr(), p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                int(row_stride), int(flags), lp[0], lp[1], lp[2], lp[3], stamp.data_ptr(),
                                clock.step.data_ptr(), clock.lr_t.data_ptr(), clock.beta1, clock.beta2, clock.eps,
                                _stream()), "dr_lazy_adam_rows")


def hash_bucket_i64(values: torch.Tensor, num_buckets: int) -> torch.Tensor:
    """tf.strings.to_hash_bucket_fast(tf.as_string(values), num_buckets) for a CUDA int64 tensor."""
    lib = _lib.load()
    if not values.is_cuda:
        raise _lib.DeepRecError("hash_bucket_i64: values must be a CUDA tensor (host features use hashing.hash_bucket)")
    v = values.to(torch.int64).contiguous()
    out = torch.empty_like(v)
    check(lib.dr_hash_bucket_i64(v.data_ptr(), v.numel(), int(num_buckets), out.data_ptr(), _stream()),
          "dr_hash_bucket_i64")
    return out


def hash_bucket_bytes(data: torch.Tensor, offsets: torch.Tensor, num_buckets: int) -> torch.Tensor:
    """Strings shipped as one uint8 CUDA buffer + int64 offsets [n+1] -> bucket ids [n]."""
    lib = _lib.load()
    if not (data.is_cuda and offsets.is_cuda):
        raise _lib.DeepRecError("hash_bucket_bytes: data / offsets must be CUDA tensors")
    if data.dtype != torch.uint8 or offsets.dtype != torch.int64:
        raise TypeError("hash_bucket_bytes: data must be uint8 and offsets int64")
    data, offsets = data.contiguous(), offsets.contiguous()
    n = offsets.numel() - 1
    out = torch.empty((max(n, 0),), device=data.device, dtype=torch.int64)
    check(lib.dr_hash_bucket_bytes(data.data_ptr(), offsets.data_ptr(), n, int(num_buckets), out.data_ptr(), _stream()),
          "dr_hash_bucket_bytes")
    return out


def vocab_lookup_i64(values: torch.Tensor, keys_sorted: torch.Tensor, vocab_index: torch.Tensor,
                     default_id: int = -1) -> torch.Tensor:
    lib = _lib.load()
    if not values.is_cuda:
        raise _lib.DeepRecError("vocab_lookup_i64: values must be a CUDA tensor")
    v = values.to(torch.int64).contiguous()
    out = torch.empty_like(v)
    check(lib.dr_vocab_lookup_i64(v.data_ptr(), v.numel(), keys_sorted.data_ptr(), vocab_index.data_ptr(),
                                  keys_sorted.numel(), int(default_id), out.data_ptr(), _stream()),
          "dr_vocab_lookup_i64")
    return out


def _bag_fwd(lib, table_ptr, rows, row_stride, ids, splits, B, D, combiner, out_ptr, out_stride):
    check(lib.dr_embed_bag_fwd(table_ptr, rows, row_stride, ids.data_ptr(), ids.element_size(), splits.data_ptr(),
                               B, D, COMBINER_CODES[combiner], out_ptr, out_stride, _stream()), "dr_embed_bag_fwd")


def _bag_bwd(lib, ids, splits, B, D, combiner, g_ptr, g_stride, rows, row_stride, grad_ptr, scale):
    check(lib.dr_embed_bag_bwd(ids.data_ptr(), ids.element_size(), splits.data_ptr(), B, D, COMBINER_CODES[combiner],
                               g_ptr, g_stride, rows, row_stride, grad_ptr, float(scale), _stream()),
          "dr_embed_bag_bwd")


def _splits(row_splits: torch.Tensor) -> torch.Tensor:
    if not isinstance(row_splits, torch.Tensor) or row_splits.dtype != torch.int64 or not row_splits.is_cuda:
        raise TypeError("row_splits must be a CUDA int64 tensor [B+1]")
    return row_splits.contiguous()


class EmbedBag(torch.autograd.Function):
    """out[B, D] = combine over the valid ids of each ragged bag (multi-valued slot); see dr_embed_bag_fwd.

    `table` is a contiguous [rows, row_stride] fp32 tensor; the D embedding columns start at `col_offset`
    (row_stride == D, col_offset == 0 for a plain table).  sparse_lr: None -> dense gradient for `table`;
    a float -> fused sparse SGD in backward.
    """

    @staticmethod
    def forward(ctx, table, ids, row_splits, D: int, combiner: str, sparse_lr, col_offset: int = 0):
        lib = _lib.load()
        table = _f32(table, "table")
        ids = _ids(ids)
        row_splits = _splits(row_splits)
        B = row_splits.numel() - 1
        rows, stride = table.shape
        if col_offset + D > stride:
            raise ValueError(f"EmbedBag: columns [{col_offset}, {col_offset + D}) exceed the row pitch {stride}")
        out = torch.empty((B, D), device=table.device, dtype=torch.float32)
        _bag_fwd(lib, table.data_ptr() + 4 * col_offset, rows, stride, ids, row_splits, B, D, combiner, out.data_ptr(), D)
        ctx.save_for_backward(table, ids, row_splits)
        ctx.D, ctx.combiner, ctx.sparse_lr, ctx.col_offset = D, combiner, sparse_lr, col_offset
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        table, ids, row_splits = ctx.saved_tensors
        g = _f32(g, "g")
        fused = ctx.sparse_lr is not None
        tgt = table if fused else torch.zeros_like(table)
        B = row_splits.numel() - 1
        with torch.no_grad():
            _bag_bwd(lib, ids, row_splits, B, ctx.D, ctx.combiner, g.data_ptr(), ctx.D, table.shape[0], table.shape[1],
                     tgt.data_ptr() + 4 * ctx.col_offset, -float(ctx.sparse_lr) if fused else 1.0)
        return (None if fused else tgt), None, None, None, None, None, None


class EmbedFMMixed(torch.autograd.Function):
    """EmbedFM for a collection in which some slots are multi-valued.

    Each ragged slot r is first reduced into a per-batch scratch "table" of B rows in the collection's own row
    layout (embedding columns = mean of the bag's valid rows, first-order entry = SUM of their weights: the
    indicator column is a multi-hot count vector); the fused gather + FM kernel then reads slot r from that
    scratch with ids 0..B-1, so the whole model is still ONE dr_embed_fm_fwd launch.  Backward mirrors it: the
    fused backward accumulates slot r's per-example gradient in a scratch grad table, dr_embed_bag_bwd spreads
    it over the bag's rows (embedding / count, first-order x 1).

    ids [B, S]: columns of ragged slots are ignored.  bags: {slot: (flat_ids, row_splits[B+1])}.
    """

    @staticmethod
    def forward(ctx, ids, weight, linear, bias, meta, bags):
        lib = _lib.load()
        if not meta.with_linear or bias is None:
            raise ValueError("EmbedFMMixed needs a collection with the first-order term (with_linear=True)")
        ids = _ids(ids).clone()
        B, S = ids.shape
        D, RS = meta.dim, meta.row_stride
        dev = weight.device
        fused_rows = meta.layout == "fused"
        tp, lp, rows = meta.pointers(weight, linear)
        tp, lp, rows = tp.clone(), lp.clone(), rows.clone()
        slots = sorted(bags)
        scratch = torch.zeros((len(slots), B, RS), device=dev, dtype=torch.float32)
        scratch_lin = None if fused_rows else torch.zeros((len(slots), B), device=dev, dtype=torch.float32)
        ar = torch.arange(B, device=dev, dtype=ids.dtype)
        saved_bags = []
        for k, s in enumerate(slots):
            bid, bsp = bags[s]
            bid, bsp = _ids(bid, "bag ids"), _splits(bsp)
            if bsp.numel() != B + 1:
                raise ValueError(f"slot {s}: row_splits has {bsp.numel()} entries, expected B+1 = {B + 1}")
            off, n_rows = int(meta._offsets[s]), meta.rows_list[s]
            base = weight.data_ptr() + 4 * off * RS
            _bag_fwd(lib, base, n_rows, RS, bid, bsp, B, D, "mean", scratch[k].data_ptr(), RS)
            if fused_rows:
                _bag_fwd(lib, base + 4 * D, n_rows, RS, bid, bsp, B, 1, "sum", scratch[k].data_ptr() + 4 * D, RS)
                lin_ptr = scratch[k].data_ptr() + 4 * D
            else:
                _bag_fwd(lib, linear.data_ptr() + 4 * off, n_rows, 1, bid, bsp, B, 1, "sum",
                         scratch_lin[k].data_ptr(), 1)
                lin_ptr = scratch_lin[k].data_ptr()
            tp[s] = scratch[k].data_ptr()
            lp[s] = lin_ptr
            rows[s] = B
            ids[:, s] = ar
            saved_bags.append((s, bid, bsp))
        stack = torch.empty((B, S, D), device=dev, dtype=torch.float32)
        sum_e = torch.empty((B, D), device=dev, dtype=torch.float32)
        logit = torch.empty((B,), device=dev, dtype=torch.float32)
        check(lib.dr_embed_fm_fwd(tp.data_ptr(), lp.data_ptr(), rows.data_ptr(), ids.data_ptr(), ids.element_size(),
                                  _ptr(bias), B, S, D, RS, meta.lin_stride, meta.flags, stack.data_ptr(),
                                  sum_e.data_ptr(), logit.data_ptr(), _stream()), "dr_embed_fm_fwd")
        ctx.meta, ctx.saved_bags = meta, saved_bags
        ctx.save_for_backward(ids, stack, sum_e, weight, linear, bias, rows)
        return stack, logit

    @staticmethod
    def backward(ctx, g_stack, g_logit):
        lib = _lib.load()
        ids, stack, sum_e, weight, linear, bias, rows = ctx.saved_tensors
        meta = ctx.meta
        B, S = ids.shape
        D, RS = meta.dim, meta.row_stride
        dev = weight.device
        fused_rows = meta.layout == "fused"
        g_stack = None if g_stack is None else _f32(g_stack, "g_stack")
        g_logit = None if g_logit is None else _f32(g_logit, "g_logit")
        if g_stack is None and g_logit is None:
            return None, None, None, None, None, None
        if g_logit is None:
            g_logit = torch.zeros((B,), device=dev, dtype=torch.float32)
        fused_sgd = meta.sparse_lr is not None
        if fused_sgd:
            gw, gl, gb, scale = weight, linear, bias, -float(meta.sparse_lr)
        else:
            gw = torch.zeros_like(weight)
            gl = None if linear is None else torch.zeros_like(linear)
            gb = None if bias is None else torch.zeros_like(bias)
            scale = 1.0
        tp, lp, _ = meta.pointers(gw, gl, cache=False)
        tp, lp = tp.clone(), lp.clone()
        nb = len(ctx.saved_bags)
        gscratch = torch.zeros((nb, B, RS), device=dev, dtype=torch.float32)
        gscratch_lin = None if fused_rows else torch.zeros((nb, B), device=dev, dtype=torch.float32)
        for k, (s, _, _) in enumerate(ctx.saved_bags):
            tp[s] = gscratch[k].data_ptr()
            lp[s] = gscratch[k].data_ptr() + 4 * D if fused_rows else gscratch_lin[k].data_ptr()
        with torch.no_grad():
            check(lib.dr_embed_fm_bwd(ids.data_ptr(), ids.element_size(), rows.data_ptr(), stack.data_ptr(),
                                      sum_e.data_ptr(), g_logit.data_ptr(), _ptr(g_stack), B, S, D, RS,
                                      meta.lin_stride, meta.flags, tp.data_ptr(), lp.data_ptr(), _ptr(gb), scale,
                                      _stream()), "dr_embed_fm_bwd")
            for k, (s, bid, bsp) in enumerate(ctx.saved_bags):
                off, n_rows = int(meta._offsets[s]), meta.rows_list[s]
                base = gw.data_ptr() + 4 * off * RS
                # the scratch already carries `scale`: spread it with scale = 1
                _bag_bwd(lib, bid, bsp, B, D, "mean", gscratch[k].data_ptr(), RS, n_rows, RS, base, 1.0)
                if fused_rows:
                    _bag_bwd(lib, bid, bsp, B, 1, "sum", gscratch[k].data_ptr() + 4 * D, RS, n_rows, RS, base + 4 * D, 1.0)
                else:
                    _bag_bwd(lib, bid, bsp, B, 1, "sum", gscratch_lin[k].data_ptr(), 1, n_rows, 1,
                             gl.data_ptr() + 4 * off, 1.0)
        if fused_sgd:
            return None, None, None, None, None, None
        return None, gw, gl, gb, None, None


def topk_rows(scores_: torch.Tensor, k: int):
    """tf.math.top_k(scores, k) on a CUDA [nq, nc] matrix: (values [nq,k] descending, int32 indices [nq,k])."""
    lib = _lib.load()
    s = _f32(scores_, "scores")
    if s.dim() != 2:
        raise ValueError(f"topk_rows: expected a 2-D score matrix, got shape {tuple(s.shape)}")
    nq, nc = s.shape
    if k > nc or k < 1:
        raise ValueError(f"input must have at least k columns. Had {nc}, needed {k}")
    vals = torch.empty((nq, k), device=s.device, dtype=torch.float32)
    idx = torch.empty((nq, k), device=s.device, dtype=torch.int32)
    check(lib.dr_topk_rows(s.data_ptr(), nq, nc, nc, k, vals.data_ptr(), idx.data_ptr(), _stream()), "dr_topk_rows")
    return vals, idx


def _idx32(indices: torch.Tensor) -> torch.Tensor:
    if not indices.is_cuda:
        raise _lib.DeepRecError("indices must be a CUDA tensor (no CPU fallback)")
    if indices.dtype not in (torch.int32, torch.int64):
        raise TypeError(f"indices: expected int32 / int64, got {indices.dtype}")
    return indices.to(torch.int32).contiguous()


def take_long_axis(arr: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """out[i, j] = arr[i, indices[i, j]] (factorized_top_k.py:26-41) for any 4- or 8-byte dtype; a 1-D `arr` is one
    shared row: tf.gather(identifiers, indices)."""
    lib = _lib.load()
    if not arr.is_cuda:
        raise _lib.DeepRecError("take_long_axis: arr must be a CUDA tensor (no CPU fallback)")
    if arr.element_size() not in (4, 8):
        raise TypeError(f"take_long_axis: 4- or 8-byte elements only, got {arr.dtype}")
    idx = _idx32(indices)
    if idx.dim() != 2:
        raise ValueError(f"take_long_axis: indices must be 2-D, got shape {tuple(idx.shape)}")
    arr = arr.contiguous()
    nq, k = idx.shape
    if arr.dim() == 1:
        ncols, ld = arr.shape[0], 0
    elif arr.dim() == 2 and arr.shape[0] == nq:
        ncols, ld = arr.shape[1], arr.shape[1]
    else:
        raise ValueError(f"take_long_axis: arr {tuple(arr.shape)} does not match indices {tuple(idx.shape)}")
    out = torch.empty((nq, k), device=arr.device, dtype=arr.dtype)
    if nq and k:
        check(lib.dr_take_long_axis(arr.data_ptr(), arr.element_size(), nq, ncols, ld, idx.data_ptr(), k, out.data_ptr(),
                                    _stream()), "dr_take_long_axis")
    return out


def exclude_adjust(scores_: torch.Tensor, identifiers: torch.Tensor, exclude: torch.Tensor,
                   penalty: float = 1.0e5) -> torch.Tensor:
    """scores - isin(identifiers, exclude) * 1e5 (factorized_top_k.py:58-62), integer identifiers."""
    lib = _lib.load()
    s = _f32(scores_, "scores")
    ident = _ids(identifiers, "identifiers").to(torch.int64).contiguous()
    ex = _ids(exclude, "exclude").to(torch.int64).contiguous()
    if s.dim() != 2 or ident.shape != s.shape or ex.dim() != 2 or ex.shape[0] != s.shape[0]:
        raise ValueError(f"exclude_adjust: scores {tuple(s.shape)}, identifiers {tuple(ident.shape)}, "
                         f"exclude {tuple(ex.shape)} do not line up")
    out = torch.empty_like(s)
    check(lib.dr_exclude_adjust(s.data_ptr(), ident.data_ptr(), ex.data_ptr(), s.shape[0], s.shape[1], ex.shape[1],
                                float(penalty), out.data_ptr(), _stream()), "dr_exclude_adjust")
    return out


def rowwise_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[n, 1] = reduce_sum(a * b, axis=1, keepdims=True)."""
    lib = _lib.load()
    a, b = _f32(a, "a"), _f32(b, "b")
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError(f"rowwise_dot: shapes {tuple(a.shape)} vs {tuple(b.shape)}")
    out = torch.empty((a.shape[0], 1), device=a.device, dtype=torch.float32)
    check(lib.dr_rowwise_dot(a.data_ptr(), b.data_ptr(), a.shape[0], a.shape[1], out.data_ptr(), _stream()),
          "dr_rowwise_dot")
    return out


def column_rank(positive: torch.Tensor, others: torch.Tensor) -> torch.Tensor:
    """rank[i] = number of entries of others[i, :] strictly above positive[i] (int32)."""
    lib = _lib.load()
    p, o = _f32(positive, "positive").reshape(-1), _f32(others, "others")
    if o.dim() != 2 or o.shape[0] != p.shape[0]:
        raise ValueError(f"column_rank: positive {tuple(p.shape)} vs others {tuple(o.shape)}")
    rank = torch.empty((p.shape[0],), device=p.device, dtype=torch.int32)
    check(lib.dr_column_rank(p.data_ptr(), o.data_ptr(), p.shape[0], o.shape[1], o.shape[1], rank.data_ptr(), _stream()),
          "dr_column_rank")
    return rank
