"""Sharded-table checkpoints on the CPU: a checkpoint written by W ranks must load into any number of
ranks with every global row landing at (row mod world, row div world) -- file I/O and index arithmetic only."""
import numpy as np
import pytest
import torch

from deep_recommenders_b200 import checkpoint, shard_plan


def _global(total, V, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((total, V), generator=g), torch.randn((total,), generator=g)


@pytest.mark.parametrize("total", [1, 7, 64, 1003])
@pytest.mark.parametrize("ws", [1, 2, 3, 8])
def test_resharding_round_trip(tmp_path, total, ws):
    W, L = _global(total, 5, seed=total + ws)
    prefix = str(tmp_path / "ckpt")
    own = lambda r, w: torch.arange(r, total, w) if r < total else torch.empty((0,), dtype=torch.int64)
    for r in range(ws):
        rows = own(r, ws)
        checkpoint.save_rows(prefix, r, ws, total, {"weight": W[rows], "lin": L[rows]})
    assert checkpoint.saved_world(prefix) == ws
    for wn in (1, 2, 4, 5):
        for q in range(wn):
            got = checkpoint.load_rows(prefix, q, wn)
            rows = own(q, wn)
            assert got["weight"].shape[0] == shard_plan.local_rows(total, q, wn)
            assert torch.equal(got["weight"], W[rows]) and torch.equal(got["lin"], L[rows])
    only = checkpoint.load_rows(prefix, 0, 1, names=["lin"])
    assert list(only) == ["lin"] and torch.equal(only["lin"], L)


def test_meta_and_error_paths(tmp_path):
    prefix = str(tmp_path / "m")
    cfg = {"rows": [3, 4], "dim": 16, "dnn_units": [8]}
    checkpoint.save_meta(prefix, cfg, {"flat": torch.arange(10.0)})
    blob = checkpoint.load_meta(prefix)
    assert blob["config"] == cfg and torch.equal(blob["replicated"]["flat"], torch.arange(10.0))
    with pytest.raises(ValueError, match="local rows"):
        checkpoint.save_rows(prefix, 0, 2, 7, {"weight": torch.zeros(3, 2)})          # rank 0 of 2 owns 4 of 7 rows
    with pytest.raises(FileNotFoundError):
        checkpoint.load_rows(str(tmp_path / "absent"), 0, 1)
    W, _ = _global(6, 2)
    checkpoint.save_rows(prefix, 0, 2, 6, {"weight": W[0::2]})
    with pytest.raises(FileNotFoundError):                                            # shard 1 of 2 missing
        checkpoint.load_rows(prefix, 0, 1)


def test_save_id_and_meta_commit_marker(tmp_path):
    """Round-1 advisor finding: shards and the meta file carry one save id; a mix of two saves is refused; the meta file
    names the writer's world, so stale shards of another world size do not make the checkpoint ambiguous."""
    total = 10
    W, _ = _global(total, 3)
    prefix = str(tmp_path / "c")
    for r in range(2):
        checkpoint.save_rows(prefix, r, 2, total, {"weight": W[r::2]}, save_id=11)
    checkpoint.save_meta(prefix, {"rows": [total]}, {"flat": torch.zeros(2)}, world=2, save_id=11)
    assert torch.equal(checkpoint.load_rows(prefix, 0, 1)["weight"], W)
    # a later save at world 3 crashed after writing its shards but before the meta file: the committed save still loads
    for r in range(3):
        checkpoint.save_rows(prefix, r, 3, total, {"weight": W[r::3] + 1}, save_id=12)
    assert checkpoint.saved_world(prefix) == 2
    assert torch.equal(checkpoint.load_rows(prefix, 0, 1)["weight"], W)
    # a later save at the SAME world crashed between the shards: shard 0 is new, shard 1 old -> refused
    checkpoint.save_rows(prefix, 0, 2, total, {"weight": W[0::2] + 2}, save_id=13)
    with pytest.raises(ValueError, match="different saves"):
        checkpoint.load_rows(prefix, 0, 1)
    # completing that save (shard 1, then the meta file last) makes it the checkpoint; stale world-3 shards are removed
    checkpoint.save_rows(prefix, 1, 2, total, {"weight": W[1::2] + 2}, save_id=13)
    checkpoint.save_meta(prefix, {"rows": [total]}, {"flat": torch.zeros(2)}, world=2, save_id=13)
    assert checkpoint.remove_stale_shards(prefix, 2) == 3
    assert torch.equal(checkpoint.load_rows(prefix, 1, 2)["weight"], (W + 2)[1::2])
