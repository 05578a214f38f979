"""Row-sharded embedding tables + data-parallel tower: the N>1 path.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Global batch = world * B,
each rank trains on its own B examples:

  fwd  ids [B,S] --dr_shard_bucket_ids--> padded per-owner local row ids
       all-to-all (ids out)            -> every owner gathers its rows  (dr_gather_fwd, fused
                                          rows [emb | w | pad]: the first-order weight rides along)
       all-to-all (vectors back)       -> dr_embed_fm_fwd with table = receive buffer, ids = inv:
                                          un-permute + stack + first-order + FM in ONE kernel
       DNN tower (replicated) -> BCE
  bwd  tower backward -> dr_embed_fm_bwd packs per-lookup gradient rows into the send buffer
       all-to-all (gradients out)      -> owners apply the row-sparse SGD update (dr_scatter_add)
       all-reduce of the flat tower gradient (+ the FM bias gradient), then dr_sgd_step.

The exchanges are equal-split (fixed capacity, -1 padded) so there is no host round trip for
split sizes; a STICKY overflow flag set by the bucket kernel is checked by train_step_host() / check_overflow().
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, ops, shard_plan
from ._lib import check
from .keras.layers.base import Dense


class ShardedEmbedding:
    """This rank's shard of the global row space, fused rows [emb | w | pad].

    exchange="nccl": rows of D+4 floats in ordinary device memory, moved by all-to-all.
    exchange="p2p" : rows padded to whole 128-B lines in SYMMETRIC memory (every rank maps every
                     other rank's shard), read / updated in place over NVLink by the fused kernels.
    """

    def __init__(self, rows: Sequence[int], dim: int, rank: int, world: int, device, seed: Optional[int] = None,
                 exchange: str = "nccl", group=None):
        self.rows_list = [int(r) for r in rows]
        self.dim, self.rank, self.world = int(dim), rank, world
        self.total_rows = sum(self.rows_list)
        self.local_rows = shard_plan.local_rows(self.total_rows, rank, world)
        max_rows = shard_plan.local_rows(self.total_rows, 0, world)       # same size on every rank
        # narrow rows (D <= 28): [emb | w | pad] = ONE 128-B line, the weight rides in the row (flags = LIN_IN_ROW).
        # wide rows (BASELINE C3-C5, D = 32 ... 128): rows of exactly D floats (whole lines, no padding) and the
        # first-order weights as a [max_rows] array trailing the shard in the same symmetric allocation.
        self.lin_in_row = exchange == "nccl" or self.dim + 1 <= 32
        if exchange == "nccl":
            if dim + 4 > 128:
                raise NotImplementedError("exchange='nccl' moves fused rows of D + 4 <= 128 floats; use exchange='p2p' "
                                          "for wider rows")
            self.vdim = self.dim + 4
        else:
            self.vdim = 32 if self.lin_in_row else self.dim
        self.flags = 1 if self.lin_in_row else 0
        self.lin_offset = 0 if self.lin_in_row else max_rows * self.vdim         # floats from the shard base
        self.device = device
        self.handle = None
        self.peer_ptrs = None
        std = 1.0 / self.dim ** 0.5
        if exchange == "p2p":
            import torch.distributed._symmetric_memory as symm_mem
            numel = max_rows * self.vdim + (0 if self.lin_in_row else (max_rows + 3) // 4 * 4)
            buf = symm_mem.empty((numel,), dtype=torch.float32, device=device)
            self.handle = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            buf.zero_()
            w = buf[:max_rows * self.vdim].view(max_rows, self.vdim)[:self.local_rows]
            self._buf = buf
            self.lin = None if self.lin_in_row else buf[self.lin_offset:self.lin_offset + max_rows][:self.local_rows]
            self.peer_ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        else:
            w = torch.zeros((self.local_rows, self.vdim), dtype=torch.float32, device=device)
            self.lin = None
        gen = torch.Generator(device=device).manual_seed((seed or 0) * 1000 + rank)
        torch.nn.init.trunc_normal_(w[:, :self.dim], 0.0, std, -2 * std, 2 * std, generator=gen)
        self.weight = w
        self.slot_offsets = torch.tensor(shard_plan.slot_offsets(self.rows_list), dtype=torch.int64, device=device)
        self.rows = torch.tensor(self.rows_list, dtype=torch.int64, device=device)

    def lin_view(self) -> torch.Tensor:
        """[local_rows] first-order weights of this rank's shard (a column of the fused rows or the trailing array)."""
        return self.weight[:, self.dim] if self.lin_in_row else self.lin


class ShardedDeepFMTrainStep:
    def __init__(self, columns, dim: int, dnn_units: Sequence[int], batch_size: int, lr: float = 0.01,
                 seed: int = 0, device=None, group=None, use_graph: bool = True, exchange: str = "p2p",
                 dw_first: bool = False):
        if not dist.is_initialized():
            raise RuntimeError("ShardedDeepFMTrainStep needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise _lib.DeepRecError("the sharded hot path runs on CUDA only (no CPU fallback)")
        self.dev = dev
        rows = [c.num_buckets for c in columns]
        self.B, self.S, self.D = int(batch_size), len(rows), int(dim)
        self.lr = float(lr)
        if exchange not in ("p2p", "nccl"):
            raise ValueError(f"exchange must be 'p2p' or 'nccl', got {exchange!r}")
        self.exchange = exchange
        self.dw_first = bool(dw_first)
        self.emb = ShardedEmbedding(rows, dim, self.rank, self.world, dev, seed, exchange, group)
        B, S, D, G = self.B, self.S, self.D, self.world
        self.n = B * S
        self.cap = shard_plan.capacity(self.n, G)
        V = self.emb.vdim
        f = dict(device=dev, dtype=torch.float32)
        # tower: identical initialisation on every rank (same seed), kept in one flat buffer; slot 0 = FM bias
        layers: List[Dense] = [Dense(u, activation="relu", seed=seed + i) for i, u in enumerate(dnn_units)]
        layers.append(Dense(1, seed=seed + len(dnn_units)))
        in_dim = S * D
        for l in layers:
            l.build((B, in_dim), device=dev)
            in_dim = l.units
        self.layers = layers
        # every parameter starts on a 16-byte boundary (the tensor-core path splits operands with float4)
        r4 = lambda n: (n + 3) // 4 * 4
        total = 4 + sum(r4(l.kernel.numel()) + r4(l.bias.numel()) for l in layers)
        self.flat = torch.zeros(total, **f)
        self.gflat = torch.zeros(total, **f)
        self.bias, self.g_bias = self.flat[0:1], self.gflat[0:1]
        self.w, self.b, self.gw, self.gb = [], [], [], []
        o = 4
        with torch.no_grad():
            for l in layers:
                for src, dst_p, dst_g in ((l.kernel, self.w, self.gw), (l.bias, self.b, self.gb)):
                    nel = src.numel()
                    self.flat[o:o + nel].copy_(src.reshape(-1))
                    src.data = self.flat[o:o + nel].view_as(src)
                    dst_p.append(src.data)
                    dst_g.append(self.gflat[o:o + nel].view_as(src))
                    o += r4(nel)
        # static buffers
        # warm-up ids spread over the tables (all-zero ids would all hit one owner)
        self.ids = torch.stack([torch.randint(0, max(1, r), (B,), device=dev) for r in rows], dim=1).contiguous()
        self.labels = torch.zeros((B,), **f)
        if exchange == "nccl":
            self.send_counts = torch.zeros((G,), device=dev, dtype=torch.int64)
            self.send_ids = torch.empty((G * self.cap,), device=dev, dtype=torch.int64)
            self.recv_ids = torch.empty((G * self.cap,), device=dev, dtype=torch.int64)
            self.inv = torch.empty((self.n,), device=dev, dtype=torch.int32)
            self.overflow = torch.zeros((1,), device=dev, dtype=torch.int32)
            self.vec_send = torch.empty((G * self.cap, V), **f)      # rows gathered for the requesters
            self.vec_recv = torch.empty((G * self.cap, V), **f)      # rows this rank asked for
            self.grad_send = torch.empty((G * self.cap, V), **f)
            self.grad_recv = torch.empty((G * self.cap, V), **f)
        self.stack = torch.empty((B, S, D), **f)
        self.sum_e = torch.empty((B, D), **f)
        self.fm_logit = torch.empty((B,), **f)
        self.acts = [torch.empty((B, l.units), **f) for l in layers]
        self.g_acts = [torch.empty((B, l.units), **f) for l in layers]
        self.gz_ws = [torch.empty((B, l.units), **f) if l._act != 0 else None for l in layers]
        self.g_stack = torch.empty((B, S, D), **f)
        self.loss = torch.zeros((1,), **f)
        self.prob = torch.empty((B,), **f)
        if exchange == "nccl":
            # "identity tables": the S slots all read the receive buffer (table base = vec_recv, rows = G*cap)
            base = self.vec_recv.data_ptr()
            self.tp = torch.full((S,), base, device=dev, dtype=torch.int64)
            self.lp = torch.full((S,), base + D * 4, device=dev, dtype=torch.int64)
            gbase = self.grad_send.data_ptr()
            self.gtp = torch.full((S,), gbase, device=dev, dtype=torch.int64)
            self.glp = torch.full((S,), gbase + D * 4, device=dev, dtype=torch.int64)
            self.trows = torch.full((S,), G * self.cap, device=dev, dtype=torch.int64)
        kmax = max([S * D] + [l.units for l in layers])
        _lib.ensure_gemm_workspace(B, kmax, kmax, dev)
        widths = [S * D] + [l.units for l in layers]
        cache_bytes = 8 * (B * (widths[0] + 2 * sum(widths[1:])) + sum(a * b for a, b in zip(widths[:-1], widths[1:])))
        _lib.set_workspace(max(cache_bytes * 5 // 4 + (1 << 20), _lib._workspace.numel() if _lib._workspace is not None else 0), dev)
        self.graph = None
        self.use_graph = use_graph
        self.graph_error = None
        self.launches_per_step = None
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._side_stream = torch.cuda.Stream(device=dev)
        self._staged = None

    # ---- checkpoint: one file per rank + metadata on rank 0; loadable by any number of ranks (checkpoint.py) -------
    def get_config(self) -> dict:
        return {"class": "ShardedDeepFMTrainStep", "rows": list(self.emb.rows_list), "dim": self.D,
                "dnn_units": [l.units for l in self.layers[:-1]], "batch_size": self.B, "lr": self.lr,
                "exchange": self.exchange, "row_width": self.emb.vdim, "lin_in_row": bool(self.emb.lin_in_row)}

    def save(self, prefix: str) -> None:
        from . import checkpoint
        torch.cuda.synchronize()
        arrays = {"weight": self.emb.weight}
        if not self.emb.lin_in_row:
            arrays["lin"] = self.emb.lin
        # one id per save, the same on every rank and in the meta file, so that files of two saves cannot be mixed
        import time
        sid = torch.tensor([time.time_ns() // 1000 if self.rank == 0 else 0], dtype=torch.int64, device=self.dev)
        dist.broadcast(sid, src=0 if self.group is None else dist.get_global_rank(self.group, 0), group=self.group)
        save_id = int(sid.item())
        checkpoint.save_rows(prefix, self.rank, self.world, self.emb.total_rows, arrays, save_id=save_id)
        dist.barrier(group=self.group)          # every shard is on disk ...
        if self.rank == 0:                      # ... before the meta file, the commit marker, is written
            checkpoint.save_meta(prefix, self.get_config(), {"flat": self.flat}, world=self.world, save_id=save_id)
            checkpoint.remove_stale_shards(prefix, self.world)
        dist.barrier(group=self.group)

    def load(self, prefix: str) -> None:
        """Restore from a checkpoint of the same model written by ANY world size (rows are re-sharded while loading)."""
        from . import checkpoint
        meta = checkpoint.load_meta(prefix)
        cfg, mine = meta["config"], self.get_config()
        for k in ("rows", "dim", "dnn_units", "row_width", "lin_in_row"):
            if cfg[k] != mine[k]:
                raise ValueError(f"checkpoint {prefix!r} was written for {k}={cfg[k]}, this model has {k}={mine[k]}")
        got = checkpoint.load_rows(prefix, self.rank, self.world)
        with torch.no_grad():
            self.emb.weight.copy_(got["weight"].to(self.dev))
            if not self.emb.lin_in_row:
                self.emb.lin.copy_(got["lin"].to(self.dev))
            self.flat.copy_(meta["replicated"]["flat"].to(self.dev))
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _a2a(self, out, inp):
        dist.all_to_all_single(out, inp, group=self.group)

    # The step is a chain  compute segment -> collective -> compute segment ...  The compute segments
    # are captured as CUDA graphs; the NCCL collectives are issued between the replays.
    def _segments_p2p(self, mark):
        """exchange="p2p": the fused gather+FM kernel reads every owner's shard directly over NVLink
        and the backward's vector atomics land directly in the owner's shard -- no id exchange, no
        owner-side gather/scatter kernels, no all-to-all.  The tower all-reduce separates the step's
        remote reads from its remote updates; a symmetric-memory barrier ends the step."""
        lib = self.lib
        B, S, D, G, V = self.B, self.S, self.D, self.world, self.emb.vdim

        def st():
            return torch.cuda.current_stream().cuda_stream

        def seg_gather():
            check(lib.dr_embed_fm_fwd_sharded(self.emb.peer_ptrs.data_ptr(), G, self.emb.slot_offsets.data_ptr(),
                                              self.emb.rows.data_ptr(), self.ids.data_ptr(), 8, self.bias.data_ptr(),
                                              B, S, D, V, self.emb.flags, self.emb.lin_offset, self.stack.data_ptr(),
                                              self.sum_e.data_ptr(), self.fm_logit.data_ptr(), st()),
                  "dr_embed_fm_fwd_sharded")
            mark("embed_fm_fwd_p2p")

        def col_reads_done():
            # every rank has finished reading every shard: from here on remote updates may land (this barrier, not the
            # tower all-reduce, is what orders a step's remote reads before its remote updates -- so the update can
            # overlap the layer-0 weight gradient and the all-reduce instead of waiting for them)
            self.emb.handle.barrier(channel=1)
            mark("barrier_reads")

        def seg_tower():
            check(lib.dr_gemm_plane_cache(1), "dr_gemm_plane_cache")
            x, K = self.stack, S * D
            L = len(self.layers)
            gz = self.g_acts[-1]
            head = (L >= 2 and self.layers[-1].units == 1 and self.layers[-1]._act == 0 and self.layers[-2].units <= 256)
            for i, l in enumerate(self.layers[:L - 1] if head else self.layers):
                check(lib.dr_dense_fwd(x.data_ptr(), self.w[i].data_ptr(), self.b[i].data_ptr(), B, K, l.units,
                                       l._act, self.acts[i].data_ptr(), st()), "dr_dense_fwd")
                x, K = self.acts[i], l.units
                mark(f"dense_fwd_{i}")
            if head:      # final Dense(1) + BCE + their backward: one kernel (see DeepFMTrainStep._enqueue)
                check(lib.dr_dense_head_bce_fwd_bwd(self.acts[L - 2].data_ptr(), self.w[L - 1].data_ptr(),
                                                    self.b[L - 1].data_ptr(), self.fm_logit.data_ptr(),
                                                    self.labels.data_ptr(), B, self.layers[L - 2].units,
                                                    self.layers[L - 2]._act, self.acts[L - 1].data_ptr(),
                                                    self.prob.data_ptr(), self.loss.data_ptr(), gz.data_ptr(),
                                                    self.g_acts[L - 2].data_ptr(), self.gw[L - 1].data_ptr(),
                                                    self.gb[L - 1].data_ptr(), self.gb[L - 2].data_ptr(), st()),
                      "dr_dense_head_bce_fwd_bwd")
                mark("head_bce")
            else:
                check(lib.dr_bce_logits_fwd_bwd(self.acts[-1].data_ptr(), self.fm_logit.data_ptr(),
                                                self.labels.data_ptr(), B, self.prob.data_ptr(), self.loss.data_ptr(),
                                                gz.data_ptr(), st()), "dr_bce")
                mark("bce")
            first = L - 2 if head else L - 1
            for i in range(first, 0, -1):      # chained backward (dr_dense_bwd_chain): see DeepFMTrainStep._enqueue
                l = self.layers[i]
                top = i == L - 1
                have_gb = head and i == L - 2
                xin, Kin, gx = self.acts[i - 1], self.layers[i - 1].units, self.g_acts[i - 1]
                check(lib.dr_dense_bwd_chain(xin.data_ptr(), self.w[i].data_ptr(),
                                             self.acts[i].data_ptr() if top else None, self.g_acts[i].data_ptr(), B, Kin,
                                             l.units, l._act if top else 0, ops._ptr(self.gz_ws[i]) if top else None,
                                             gx.data_ptr(), self.gw[i].data_ptr(),
                                             None if have_gb else self.gb[i].data_ptr(), xin.data_ptr(),
                                             self.layers[i - 1]._act, st()), "dr_dense_bwd_chain")
                mark(f"dense_bwd_{i}")
            l = self.layers[0]      # layer 0: the input gradient first (the embedding update needs it) ...
            top0 = L == 1
            gb0 = None if (head and L == 2) else self.gb[0].data_ptr()
            check(lib.dr_dense_bwd(self.stack.data_ptr(), self.w[0].data_ptr(), self.acts[0].data_ptr() if top0 else None,
                                   self.g_acts[0].data_ptr(), B, S * D, l.units, l._act if top0 else 0,
                                   ops._ptr(self.gz_ws[0]) if top0 else None,
                                   self.g_stack.data_ptr(), None, gb0, st()), "dr_dense_bwd(dx)")
            mark("dense_bwd_0_dx")

        def seg_update():           # replayed on the SIDE stream: remote vector atomics into the owners' shards
            gz = self.g_acts[-1]
            check(lib.dr_embed_fm_bwd_sharded(self.emb.peer_ptrs.data_ptr(), G, self.emb.slot_offsets.data_ptr(),
                                              self.emb.rows.data_ptr(), self.ids.data_ptr(), 8, self.stack.data_ptr(),
                                              self.sum_e.data_ptr(), gz.data_ptr(), self.g_stack.data_ptr(), B, S, D,
                                              V, self.emb.flags, self.emb.lin_offset, None, -self.lr / G, st()),
                  "dr_embed_fm_bwd_sharded")

        def seg_dw():               # ... meanwhile, on the main stream: layer-0 weight gradient + FM bias gradient
            l = self.layers[0]
            gz0 = (self.gz_ws[0] if l._act != 0 else self.g_acts[0]) if len(self.layers) == 1 else self.g_acts[0]
            check(lib.dr_dense_bwd(self.stack.data_ptr(), self.w[0].data_ptr(), None, gz0.data_ptr(), B, S * D, l.units,
                                   0, None, None, self.gw[0].data_ptr(), None, st()), "dr_dense_bwd(dw)")
            check(lib.dr_gemm_plane_cache(0), "dr_gemm_plane_cache")
            mark("dense_bwd_0_dw")
            torch.sum(self.g_acts[-1].view(-1), dim=0, keepdim=True, out=self.g_bias)      # FM bias gradient (tiny)
            mark("bias_grad")

        def col_allreduce():
            dist.all_reduce(self.gflat, group=self.group)
            mark("allreduce_dense")

        def seg_sgd():
            check(lib.dr_sgd_step(self.flat.data_ptr(), self.gflat.data_ptr(), self.flat.numel(), self.lr / G, st()),
                  "dr_sgd_step")
            mark("sgd")

        def col_barrier():
            mark("embed_fm_bwd_p2p_tail")               # what was left of the side-stream update after the join
            self.emb.handle.barrier(channel=0)          # all remote atomics of this step have been issued
            mark("barrier")

        if self.dw_first:
            # the persistent weight-gradient GEMM is submitted BEFORE the remote update (both depend only on the layer-0
            # input gradient): it takes its one CTA per SM and the update's CTAs fill what is left (knob tc_dw_share)
            return [(seg_gather, "graph"), (col_reads_done, "eager"), (seg_tower, "graph"), (None, "fork"),
                    (seg_dw, "graph"), (seg_update, "graph_side_fork"), (col_allreduce, "eager"), (seg_sgd, "graph"),
                    (None, "join_side"), (col_barrier, "eager")]
        return [(seg_gather, "graph"), (col_reads_done, "eager"), (seg_tower, "graph"), (seg_update, "graph_side"),
                (seg_dw, "graph"), (col_allreduce, "eager"), (seg_sgd, "graph"), (None, "join_side"),
                (col_barrier, "eager")]

    def _segments(self, mark):
        if self.exchange == "p2p":
            return self._segments_p2p(mark)
        lib = self.lib
        B, S, D, G, V = self.B, self.S, self.D, self.world, self.emb.vdim

        def st():
            return torch.cuda.current_stream().cuda_stream

        def seg_bucket():
            check(lib.dr_shard_bucket_ids(self.ids.data_ptr(), 8, self.n, S, self.emb.slot_offsets.data_ptr(),
                                          self.emb.rows.data_ptr(), G, self.cap, self.send_counts.data_ptr(),
                                          self.send_ids.data_ptr(), self.inv.data_ptr(), self.overflow.data_ptr(), st()),
                  "dr_shard_bucket_ids")
            mark("bucket")

        def col_ids():
            self._a2a(self.recv_ids, self.send_ids)
            mark("a2a_ids")

        def seg_gather():
            check(lib.dr_gather_fwd(self.emb.weight.data_ptr(), self.emb.local_rows, self.recv_ids.data_ptr(), 8,
                                    G * self.cap, V, self.vec_send.data_ptr(), st()), "dr_gather_fwd")
            mark("owner_gather")

        def col_vec():
            self._a2a(self.vec_recv, self.vec_send)
            mark("a2a_vectors")

        def seg_main():
            check(lib.dr_gemm_plane_cache(1), "dr_gemm_plane_cache")
            check(lib.dr_embed_fm_fwd(self.tp.data_ptr(), self.lp.data_ptr(), self.trows.data_ptr(),
                                      self.inv.data_ptr(), 4, self.bias.data_ptr(), B, S, D, V, V, 1,
                                      self.stack.data_ptr(), self.sum_e.data_ptr(), self.fm_logit.data_ptr(), st()),
                  "dr_embed_fm_fwd")
            mark("embed_fm_fwd")
            x, K = self.stack, S * D
            for i, l in enumerate(self.layers):
                check(lib.dr_dense_fwd(x.data_ptr(), self.w[i].data_ptr(), self.b[i].data_ptr(), B, K, l.units,
                                       l._act, self.acts[i].data_ptr(), st()), "dr_dense_fwd")
                x, K = self.acts[i], l.units
                mark(f"dense_fwd_{i}")
            gz = self.g_acts[-1]
            check(lib.dr_bce_logits_fwd_bwd(self.acts[-1].data_ptr(), self.fm_logit.data_ptr(),
                                            self.labels.data_ptr(), B, self.prob.data_ptr(), self.loss.data_ptr(),
                                            gz.data_ptr(), st()), "dr_bce")
            mark("bce")
            for i in range(len(self.layers) - 1, -1, -1):
                l = self.layers[i]
                xin = self.stack if i == 0 else self.acts[i - 1]
                Kin = S * D if i == 0 else self.layers[i - 1].units
                gx = self.g_stack if i == 0 else self.g_acts[i - 1]
                check(lib.dr_dense_bwd(xin.data_ptr(), self.w[i].data_ptr(), self.acts[i].data_ptr(),
                                       self.g_acts[i].data_ptr(), B, Kin, l.units, l._act, ops._ptr(self.gz_ws[i]),
                                       gx.data_ptr(), self.gw[i].data_ptr(), self.gb[i].data_ptr(), st()),
                      "dr_dense_bwd")
                mark(f"dense_bwd_{i}")
            check(lib.dr_gemm_plane_cache(0), "dr_gemm_plane_cache")
            # pack per-lookup gradient rows [dE | g_logit | 0 0 0] into the padded send buffer
            self.grad_send.zero_()
            self.g_bias.zero_()
            check(lib.dr_embed_fm_bwd(self.inv.data_ptr(), 4, self.trows.data_ptr(), self.stack.data_ptr(),
                                      self.sum_e.data_ptr(), gz.data_ptr(), self.g_stack.data_ptr(), B, S, D, V, V, 1,
                                      self.gtp.data_ptr(), self.glp.data_ptr(), self.g_bias.data_ptr(), 1.0, st()),
                  "dr_embed_fm_bwd")
            mark("embed_fm_bwd_pack")

        def col_grads():
            self._a2a(self.grad_recv, self.grad_send)
            mark("a2a_grads")

        def seg_scatter():
            # owners: row-sparse SGD on the fused rows (mean over the GLOBAL batch: 1/world)
            check(lib.dr_scatter_add(self.emb.weight.data_ptr(), self.emb.local_rows, self.recv_ids.data_ptr(), 8,
                                     G * self.cap, V, self.grad_recv.data_ptr(), -self.lr / G, st()), "dr_scatter_add")
            mark("owner_scatter_sgd")

        def col_allreduce():
            dist.all_reduce(self.gflat, group=self.group)
            mark("allreduce_dense")

        def seg_sgd():
            check(lib.dr_sgd_step(self.flat.data_ptr(), self.gflat.data_ptr(), self.flat.numel(), self.lr / G, st()),
                  "dr_sgd_step")
            mark("sgd")

        return [(seg_bucket, "graph"), (col_ids, "eager"), (seg_gather, "graph"), (col_vec, "eager"), (seg_main, "graph"),
                (col_grads, "eager"), (seg_scatter, "graph"), (col_allreduce, "eager"), (seg_sgd, "graph")]

    # segment kinds: "graph" = compute on the main stream (captured as a CUDA graph), "eager" = a collective issued
    # between replays, "graph_side" = compute replayed on the side stream after it has waited for the main stream
    # (runs concurrently with what follows), "join_side" = the main stream waits for the side stream.
    def _run_plan(self, plan):
        main = torch.cuda.current_stream()
        for fn, kind in plan:
            if kind == "graph_side":
                self._side_stream.wait_stream(main)
                with torch.cuda.stream(self._side_stream):
                    fn()
            elif kind == "fork":
                fork = torch.cuda.Event()
                fork.record(main)
            elif kind == "graph_side_fork":
                self._side_stream.wait_event(fork)
                with torch.cuda.stream(self._side_stream):
                    fn()
            elif kind == "join_side":
                main.wait_stream(self._side_stream)
            else:
                fn()

    def _enqueue(self, mark=None):
        mark = mark or (lambda label: None)
        mark("start")
        self._run_plan(self._segments(mark))

    def _persistent_state(self):
        """Every tensor a step mutates and a later step reads (this rank's shard incl. trailing first-order weights,
        the flat tower)."""
        return [self.emb._buf if self.exchange == "p2p" else self.emb.weight, self.flat]

    def capture(self):
        """Warm up eagerly, then record the compute segments into CUDA graphs (collectives stay eager NCCL calls:
        capturing the NCCL all-to-all itself dead-locked on this torch/NCCL build).

        Side-effect free: the warm-up is a real step on the static warm-up ids (with remote atomics into the peers'
        shards), so every rank snapshots its shard and the tower first and restores them once ALL ranks have finished
        the warm-up (barrier on both sides of the restore)."""
        n0 = _lib.launch_count()
        with torch.no_grad():
            state = self._persistent_state()
            saved = [t.clone() for t in state]
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)          # every peer's remote updates of the warm-up step have landed
            for t, keep in zip(state, saved):
                t.copy_(keep)
            del saved
            torch.cuda.synchronize()
            dist.barrier(group=self.group)          # nobody starts a real step before every shard is restored
        self.launches_per_step = _lib.launch_count() - n0
        self.check_overflow()
        if self.use_graph:
            plan = []
            for fn, kind in self._segments(lambda label: None):
                if kind in ("graph", "graph_side", "graph_side_fork"):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        fn()
                    plan.append((g.replay, kind))
                else:
                    plan.append((fn, kind))
            self.graph = plan
        return self

    def check_overflow(self):
        """exchange='nccl' only: raise if ANY bucket call since the last check ran out of send slots (the flag is
        sticky on the device; it is cleared here after it has been read).  Called by capture() and by
        train_step_host() (which synchronises on the loss anyway); callers of the asynchronous step() / run() call it
        whenever they synchronise."""
        if self.exchange == "p2p":
            return
        if int(self.overflow.item()) != 0:
            self.overflow.zero_()
            raise _lib.DeepRecError(f"shard exchange overflow: a destination needed more than cap={self.cap} slots "
                                    "(low-cardinality or skewed slot): lookups were dropped in a step since the last "
                                    "check; rebuild the trainer with a larger shard_plan.capacity slack")

    def run(self):
        if self.graph is not None:
            self._run_plan(self.graph)
        else:
            self._enqueue()

    def step(self, ids, labels):
        self.ids.copy_(ids, non_blocking=True)
        self.labels.copy_(labels.reshape(-1), non_blocking=True)
        self.run()
        return self.loss

    def stage_host(self, ids_host, labels_host):
        if not hasattr(self, "_spare"):
            self._spare = (torch.empty_like(self.ids), torch.empty_like(self.labels))
        with torch.cuda.stream(self._copy_stream):
            self._spare[0].copy_(ids_host, non_blocking=True)
            self._spare[1].copy_(labels_host.reshape(-1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._staged = ev

    def train_step_host(self, ids_host, labels_host, next_ids_host=None, next_labels_host=None) -> float:
        cur = torch.cuda.current_stream()
        if self._staged is None:
            self.stage_host(ids_host, labels_host)
        cur.wait_event(self._staged)
        self.ids.copy_(self._spare[0], non_blocking=True)
        self.labels.copy_(self._spare[1], non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        self._staged = None
        if next_ids_host is not None:
            self._copy_stream.wait_event(done)
            self.stage_host(next_ids_host, next_labels_host)
        self.run()
        loss = float(self.loss.item())
        self.check_overflow()
        return loss

    def fit_host(self, batches, depth: int = 2):
        """Pipelined epoch loop over HOST batches (training.fit_host): H2D of the next batch and the D2H read of every
        step's loss stay off the critical path; the overflow flag of the NCCL exchange is checked once at the end."""
        from .training import fit_host
        losses = fit_host(self, batches, depth)
        self.check_overflow()
        return losses

    def time_embed_fwd(self, ids_pool, iters: int = 30) -> float:
        """Mean duration (ms) of this rank's fused gather+FM forward alone (p2p: (G-1)/G of the rows are
        read from peers over NVLink inside the same kernel).  nccl exchange: times the local
        identity-table kernel only."""
        lib, st = self.lib, torch.cuda.current_stream().cuda_stream
        B, S, D, G, V = self.B, self.S, self.D, self.world, self.emb.vdim

        def launch(k):
            if self.exchange == "p2p":
                ids = ids_pool[k % len(ids_pool)]
                check(lib.dr_embed_fm_fwd_sharded(self.emb.peer_ptrs.data_ptr(), G, self.emb.slot_offsets.data_ptr(),
                                                  self.emb.rows.data_ptr(), ids.data_ptr(), 8, self.bias.data_ptr(),
                                                  B, S, D, V, self.emb.flags, self.emb.lin_offset,
                                                  self.stack.data_ptr(), self.sum_e.data_ptr(),
                                                  self.fm_logit.data_ptr(), st), "dr_embed_fm_fwd_sharded")
            else:
                check(lib.dr_embed_fm_fwd(self.tp.data_ptr(), self.lp.data_ptr(), self.trows.data_ptr(),
                                          self.inv.data_ptr(), 4, self.bias.data_ptr(), B, S, D, V, V, 1,
                                          self.stack.data_ptr(), self.sum_e.data_ptr(), self.fm_logit.data_ptr(), st),
                      "dr_embed_fm_fwd")

        for k in range(5):
            launch(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(iters):
            launch(k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def profile_kernels(self, ids_pool, labels_pool, iters: int = 10):
        runs = []
        for it in range(iters + 2):      # back-to-back, no host sync in between (steady-state clocks / caches)
            self.ids.copy_(ids_pool[it % len(ids_pool)], non_blocking=True)
            self.labels.copy_(labels_pool[it % len(labels_pool)].reshape(-1), non_blocking=True)
            evs = []

            def mark(label, evs=evs):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs.append((label, e))

            self._enqueue(mark)
            runs.append(evs)
        torch.cuda.synchronize()
        sums, count = {}, 0
        for evs in runs[2:]:
            count += 1
            for (l0, e0), (l1, e1) in zip(evs[:-1], evs[1:]):
                sums[l1] = sums.get(l1, 0.0) + e0.elapsed_time(e1)
        shares = {k: v / count for k, v in sums.items()}
        return {"embed_fm_fwd_ms": shares.get("embed_fm_fwd", shares.get("embed_fm_fwd_p2p"))}, shares
