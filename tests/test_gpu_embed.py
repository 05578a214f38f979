"""Rows E + L + F on the GPU through the C-ABI, against the numpy oracle (oracle/reference_np.py).

Bars (BASELINE.json north_star): gathered rows bit-exact; interaction / logit values within
1e-5 relative in fp32, measured against the float64 oracle with a cancellation-aware scale
(SURVEY.md section 7 hard part 4): |a - b| <= 1e-5 * (|linear| + 0.5*(sum (sum e)^2 + sum e^2)) + 1e-6.
"""
import numpy as np
import pytest
import torch

from oracle import reference_np as R

pytestmark = pytest.mark.gpu


def make_problem(B, rows, D, seed=0, oov_frac=0.0, id_dtype=np.int64, zipf=False):
    rng = np.random.default_rng(seed)
    S = len(rows)
    tables = [(rng.standard_normal((r, D)) / np.sqrt(D)).astype(np.float32) for r in rows]
    lins = [(rng.standard_normal((r,)) * 0.1).astype(np.float32) for r in rows]
    bias = np.float32(0.3)
    if zipf:
        ids = np.stack([np.minimum(rng.zipf(1.2, size=B) - 1, r - 1) for r in rows], axis=1)
    else:
        ids = np.stack([rng.integers(0, r, size=B) for r in rows], axis=1)
    if oov_frac > 0:
        mask = rng.random(ids.shape) < oov_frac
        ids = np.where(mask, np.where(rng.random(ids.shape) < 0.5, -1, 10 ** 7), ids)
    return tables, lins, bias, ids.astype(id_dtype)


def to_collection(tables, lins, bias, sparse_lr=None, layout="fused"):
    from deep_recommenders_b200.embedding import EmbeddingCollection
    if layout == "fused" and tables[0].shape[1] > 28:
        layout = "split"
    coll = EmbeddingCollection([t.shape[0] for t in tables], tables[0].shape[1], device="cuda", init="empty",
                               sparse_lr=sparse_lr, layout=layout)
    with torch.no_grad():
        coll.emb_view().copy_(torch.from_numpy(np.concatenate(tables, 0)))
        coll.lin_view().copy_(torch.from_numpy(np.concatenate(lins, 0)))
        coll.bias.fill_(float(bias))
    return coll


def logit_scale(tables, lins, bias, ids):
    st = R.stack_embeddings(tables, ids).astype(np.float64)
    lin = np.abs(R.linear_term(lins, 0.0, ids, np.float64)).reshape(-1) + abs(float(bias))
    return lin + 0.5 * ((st.sum(1) ** 2).sum(1) + (st ** 2).sum((1, 2)))


CASES = [
    # B, rows, D
    (1, [7], 4),
    (5, [11, 3], 8),
    (64, [100] * 6, 16),
    (257, [50, 60, 70, 2, 7, 21], 16),          # ragged tail, MovieLens-like cardinalities
    (1000, [1000] * 26, 16),
    (513, [300] * 26, 32),
    (130, [97] * 5, 64),
    (77, [41] * 3, 128),
    (99, [64] * 4, 12),                          # D/4 = 3 chunks: not a power of two
    (33, [19] * 2, 20),
]


@pytest.mark.parametrize("B,rows,D", CASES)
@pytest.mark.parametrize("id_dtype,layout", [(np.int64, "fused"), (np.int32, "fused"), (np.int64, "split")])
def test_forward_parity(B, rows, D, id_dtype, layout):
    tables, lins, bias, ids = make_problem(B, rows, D, seed=B + D, oov_frac=0.1, id_dtype=id_dtype)
    coll = to_collection(tables, lins, bias, layout=layout)
    with torch.no_grad():
        stack, logit = coll(torch.from_numpy(ids).cuda())
    torch.cuda.synchronize()
    ref_stack = R.stack_embeddings(tables, ids)
    # a gather is a copy: bit-exact
    assert np.array_equal(stack.cpu().numpy().view(np.uint32), ref_stack.view(np.uint32))
    ref_logit, _ = R.fm_logit(tables, lins, bias, ids, np.float64)
    err = np.abs(logit.cpu().numpy().astype(np.float64) - ref_logit.reshape(-1))
    tol = 1e-5 * logit_scale(tables, lins, bias, ids) + 1e-6
    assert (err <= tol).all(), f"max err {err.max()} tol {tol[np.argmax(err)]}"
    # fp32 oracle agrees to the reference's own assertAllClose default (rtol=atol=1e-6 * scale)
    ref32, _ = R.fm_logit(tables, lins, bias, ids, np.float32)
    assert np.allclose(logit.cpu().numpy(), ref32.reshape(-1), rtol=1e-5, atol=float(tol.max()))


def test_forward_empty_batch():
    tables, lins, bias, ids = make_problem(4, [10, 10], 16)
    coll = to_collection(tables, lins, bias)
    with torch.no_grad():
        stack, logit = coll(torch.zeros((0, 2), dtype=torch.int64, device="cuda"))
    assert stack.shape == (0, 2, 16) and logit.shape == (0,)


def test_all_oov_gives_bias_only():
    tables, lins, bias, _ = make_problem(8, [10, 10, 10], 16)
    coll = to_collection(tables, lins, bias)
    ids = torch.full((8, 3), -1, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        stack, logit = coll(ids)
    assert float(stack.abs().max()) == 0.0
    assert np.allclose(logit.cpu().numpy(), float(bias))


def test_rejects_bad_arguments():
    from deep_recommenders_b200.embedding import EmbeddingCollection
    with pytest.raises(ValueError):
        EmbeddingCollection([10], 6, device="cuda")            # D % 4 != 0
    coll = EmbeddingCollection([10, 10], 16, device="cuda")
    with pytest.raises(ValueError):
        coll(torch.zeros((4, 3), dtype=torch.int64, device="cuda"))   # wrong slot count
    with pytest.raises(TypeError):
        coll(torch.zeros((4, 2), dtype=torch.float32, device="cuda"))
    from deep_recommenders_b200._lib import DeepRecError
    with pytest.raises(DeepRecError):
        coll(torch.zeros((4, 2), dtype=torch.int64))           # CPU tensor: no fallback


BWD_CASES = [
    (5, [11, 3], 8),
    (257, [50, 60, 70, 2, 7, 21], 16),
    (1000, [1000] * 26, 16),
    (300, [40] * 26, 32),          # heavy duplicates (40 rows, 300 examples)
    (130, [97] * 5, 64),
    (77, [41] * 3, 128),
    (99, [64] * 4, 12),
]


@pytest.mark.parametrize("B,rows,D", BWD_CASES)
@pytest.mark.parametrize("mode,agg,layout", [(0, 0, "fused"), (0, 0, "split"), (1, 1, "fused"), (1, 0, "fused"),
                                             (1, 1, "split")])
def test_backward_parity(B, rows, D, mode, agg, layout):
    """mode 0 = slot-parallel kernel (default), mode 1 = example-parallel kernel (+- warp aggregation)."""
    from deep_recommenders_b200 import _lib
    _lib.tune("embed_bwd_mode", mode)
    _lib.tune("embed_bwd_agg", agg)
    try:
        tables, lins, bias, ids = make_problem(B, rows, D, seed=7 * B + D, oov_frac=0.05)
        coll = to_collection(tables, lins, bias, layout=layout)
        rng = np.random.default_rng(1)
        g_logit = rng.standard_normal(B).astype(np.float32)
        g_stack = rng.standard_normal((B, len(rows), D)).astype(np.float32)
        stack, logit = coll(torch.from_numpy(ids).cuda())
        loss = (logit * torch.from_numpy(g_logit).cuda()).sum() + (stack * torch.from_numpy(g_stack).cuda()).sum()
        loss.backward()
        torch.cuda.synchronize()
        ref_stack = R.stack_embeddings(tables, ids)
        gts, gls, gb = R.embed_fm_grad([t.shape[0] for t in tables], ids, ref_stack, g_logit, g_stack, np.float64)
        # atomics sum duplicates in arbitrary order: tolerance against the float64 sum of |terms|
        st64 = ref_stack.astype(np.float64)
        absdE = np.abs(g_logit)[:, None, None] * (np.abs(st64.sum(1, keepdims=True)) + np.abs(st64)) + np.abs(g_stack)
        scale_t, _, _ = R.embed_fm_grad([t.shape[0] for t in tables], ids, ref_stack * 0, None, absdE, np.float64)
        cgw, cgl, cgb = coll.grads()
        gw = cgw.cpu().numpy().astype(np.float64)
        ref = np.concatenate(gts, 0)
        sc = np.concatenate(scale_t, 0)
        assert (np.abs(gw - ref) <= 1e-5 * sc + 1e-7).all(), np.abs(gw - ref).max()
        gl = cgl.cpu().numpy().astype(np.float64)
        refl = np.concatenate(gls, 0)
        _, scl, _ = R.embed_fm_grad([t.shape[0] for t in tables], ids, ref_stack, np.abs(g_logit), None, np.float64)
        assert (np.abs(gl - refl) <= 1e-5 * np.concatenate(scl, 0) + 1e-7).all()
        assert abs(float(cgb) - gb) <= 1e-5 * np.abs(g_logit).sum() + 1e-7
        if coll.layout == "fused":
            assert float(coll.weight.grad[:, D + 1:].abs().max()) == 0.0      # pad lanes untouched
    finally:
        _lib.tune("embed_bwd_agg", 1)
        _lib.tune("embed_bwd_mode", 0)


def test_backward_same_id_1000_times():
    """The reference's own smoke tests feed the SAME id 1000x (tests/keras/test_fm.py:89-92)."""
    tables, lins, bias, _ = make_problem(4, [100, 100], 16, seed=3)
    ids = np.stack([np.full(1000, 1), np.full(1000, 2)], axis=1).astype(np.int64)
    coll = to_collection(tables, lins, bias)
    stack, logit = coll(torch.from_numpy(ids).cuda())
    logit.sum().backward()
    torch.cuda.synchronize()
    ref_stack = R.stack_embeddings(tables, ids)
    gts, gls, gb = R.embed_fm_grad([100, 100], ids, ref_stack, np.ones(1000), None, np.float64)
    cgw, cgl, cgb = coll.grads()
    gw = cgw.cpu().numpy()
    ref = np.concatenate(gts, 0)
    assert np.allclose(gw, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    assert np.allclose(cgl.cpu().numpy(), np.concatenate(gls, 0), rtol=1e-6)
    assert abs(float(cgb) - 1000.0) < 1e-3


def test_fused_sparse_sgd_equals_dense_sgd():
    tables, lins, bias, ids = make_problem(200, [30] * 6, 16, seed=5)
    lr = 0.05
    dense = to_collection(tables, lins, bias)
    fused = to_collection(tables, lins, bias, sparse_lr=lr)
    idt = torch.from_numpy(ids).cuda()
    gvec = torch.from_numpy(np.random.default_rng(2).standard_normal(200).astype(np.float32)).cuda()
    for coll in (dense, fused):
        stack, logit = coll(idt)
        ((logit * gvec).sum() + 0.5 * (stack * stack).sum()).backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        gw, gl, gb = dense.grads()
        want_w = dense.emb_view() - lr * gw
        want_l = dense.lin_view() - lr * gl
        want_b = dense.bias - lr * gb
    assert fused.weight.grad is None
    assert torch.allclose(fused.emb_view(), want_w, rtol=1e-5, atol=1e-6)
    assert torch.allclose(fused.lin_view(), want_l, rtol=1e-5, atol=1e-6)
    assert torch.allclose(fused.bias, want_b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,rows,D", [(1, 5, 4), (1000, 300, 64), (4097, 1000, 16), (333, 50, 128)])
def test_single_table_gather_scatter(n, rows, D):
    from deep_recommenders_b200 import ops
    rng = np.random.default_rng(n)
    table = rng.standard_normal((rows, D)).astype(np.float32)
    ids = rng.integers(-1, rows + 1, size=n).astype(np.int64)        # includes -1 and rows (OOV)
    t = torch.from_numpy(table).cuda().requires_grad_(True)
    out = ops.Gather.apply(t, torch.from_numpy(ids).cuda(), None)
    assert np.array_equal(out.detach().cpu().numpy(), R.embedding_lookup(table, ids))
    g = rng.standard_normal((n, D)).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    ref = np.zeros((rows, D), np.float64)
    ok = (ids >= 0) & (ids < rows)
    np.add.at(ref, ids[ok], g[ok].astype(np.float64))
    sc = np.zeros((rows, D), np.float64)
    np.add.at(sc, ids[ok], np.abs(g[ok]).astype(np.float64))
    assert (np.abs(t.grad.cpu().numpy() - ref) <= 1e-5 * sc + 1e-7).all()


def test_full_size_c2_properties():
    """BASELINE config C2 shape (B=65536, S=26, D=16, 1M-row tables): size-independent properties.
    The stack must equal an independent torch gather bit-for-bit, and the fused logit must equal
    the standalone FM kernel applied to that stack plus the gathered linear weights."""
    from deep_recommenders_b200 import ops
    from deep_recommenders_b200.embedding import EmbeddingCollection
    B, S, D, rows = 65536, 26, 16, 1_000_000
    coll = EmbeddingCollection([rows] * S, D, device="cuda", seed=1)
    with torch.no_grad():
        coll.lin_view().normal_(0, 0.1)
        coll.bias.fill_(0.25)
    gen = torch.Generator(device="cuda").manual_seed(0)
    ids = torch.randint(0, rows, (B, S), device="cuda", generator=gen)
    with torch.no_grad():
        stack, logit = coll(ids)
    offs = (torch.arange(S, device="cuda") * rows).view(1, S)
    flat = (ids + offs).view(-1)
    picked = coll.weight.detach().index_select(0, flat)
    want = picked[:, :D].contiguous().view(B, S, D)
    assert torch.equal(stack, want)
    lin = picked[:, D].view(B, S).double().sum(1)
    st = want.double()
    fm64 = 0.5 * ((st.sum(1) ** 2).sum(1) - (st ** 2).sum((1, 2)))
    ref = lin + fm64 + 0.25
    scale = lin.abs() + 0.5 * ((st.sum(1) ** 2).sum(1) + (st ** 2).sum((1, 2))) + 0.25
    assert ((logit.double() - ref).abs() <= 1e-5 * scale + 1e-6).all()
    fm_k = ops.FMInteraction.apply(want).view(-1).double()
    assert ((fm_k - fm64).abs() <= 1e-5 * scale + 1e-6).all()
    # checksum-of-checksums: scatter-add of ones over the batch conserves the count
    g = torch.ones((B, S, D), device="cuda")
    gw = torch.zeros_like(coll.weight)
    tp, lp, rws = coll.pointers(gw, None, cache=False)
    from deep_recommenders_b200 import _lib
    _lib.check(_lib.load().dr_embed_fm_bwd(ids.data_ptr(), 8, rws.data_ptr(), None, None, None, g.data_ptr(),
                                           B, S, D, coll.row_stride, coll.lin_stride, 0, tp.data_ptr(), None, None, 1.0,
                                           torch.cuda.current_stream().cuda_stream), "bwd")
    assert float(gw.sum()) == float(B * S * D)


@pytest.mark.parametrize("B,rows", [(1, [7]), (5, [11, 3]), (64, [100] * 6), (257, [50, 60, 70, 2, 7, 21]), (1000, [1000] * 26),
                                    (123, [40] * 32)])
@pytest.mark.parametrize("id_dtype", [np.int64, np.int32])
def test_forward_tma_staged_variant_matches(B, rows, id_dtype):
    """dr_embed_fm_fwd_tma (rows staged through TMA tile::gather4 into shared memory, opt-in): same stack bit for bit,
    same logit / sum_e within the oracle tolerance, OOV ids -> zero rows via the TMA out-of-bounds fill."""
    from deep_recommenders_b200 import _lib
    lib = _lib.load()
    D = 16
    tables, lins, bias, ids = make_problem(B, rows, D, seed=B + len(rows), oov_frac=0.1, id_dtype=id_dtype)
    coll = to_collection(tables, lins, bias, layout="fused")
    idt = torch.from_numpy(ids).cuda()
    S = len(rows)
    stack = torch.full((B, S, D), float("nan"), device="cuda")
    sum_e = torch.full((B, D), float("nan"), device="cuda")
    logit = torch.full((B,), float("nan"), device="cuda")
    _lib.check(lib.dr_embed_fm_fwd_tma(coll.weight.data_ptr(), coll.total_rows, coll._offsets.data_ptr(), coll._rows.data_ptr(),
                                       idt.data_ptr(), idt.element_size(), coll.bias.data_ptr(), B, S, D, coll.row_stride,
                                       stack.data_ptr(), sum_e.data_ptr(), logit.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "dr_embed_fm_fwd_tma")
    torch.cuda.synchronize()
    ref_stack = R.stack_embeddings(tables, ids)
    assert np.array_equal(stack.cpu().numpy().view(np.uint32), ref_stack.view(np.uint32))
    ref_logit, _ = R.fm_logit(tables, lins, bias, ids, np.float64)
    err = np.abs(logit.cpu().numpy().astype(np.float64) - ref_logit.reshape(-1))
    tol = 1e-5 * logit_scale(tables, lins, bias, ids) + 1e-6
    assert (err <= tol).all(), f"max err {err.max()}"
    ssum = ref_stack.astype(np.float64).sum(1)
    assert (np.abs(sum_e.cpu().numpy() - ssum) <= 1e-5 * np.abs(ref_stack.astype(np.float64)).sum(1) + 1e-7).all()
    # unsupported shapes are refused, not mis-served
    rc = lib.dr_embed_fm_fwd_tma(coll.weight.data_ptr(), coll.total_rows, coll._offsets.data_ptr(), coll._rows.data_ptr(),
                                 idt.data_ptr(), idt.element_size(), coll.bias.data_ptr(), B, S, 32, coll.row_stride,
                                 stack.data_ptr(), sum_e.data_ptr(), logit.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc < 0
