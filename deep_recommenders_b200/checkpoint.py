"""Checkpoints of row-sharded embedding tables ("sharded-table save/load with get_config
round-trip"; the reference's round trips are tests/keras/test_fm.py:44-65 -- SavedModel of a single-process model).

One file per rank, `<prefix>.shard<rank>-of-<world>.pt`, holding that rank's rows of the global row space
(owner = row mod world, local index = row div world: shard_plan.py) plus a small metadata dict; rank 0 also writes
`<prefix>.meta.pt` with the model config and the replicated (tower) parameters -- LAST, after every shard is on disk: it
is the commit marker of a save and names the writer's world size and the save's id, which every shard carries too, so a
crash between the per-rank renames (old and new shards side by side) is detected at load time instead of yielding
inconsistent tables.  A checkpoint written by W ranks can
be loaded by ANY number of ranks: `load_rows` re-shards on the fly by reading the saved shards and picking the rows
the new rank owns.  File I/O and index arithmetic only -- no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch

from . import shard_plan

FORMAT = 1


def shard_path(prefix: str, rank: int, world: int) -> str:
    return f"{prefix}.shard{rank:03d}-of-{world:03d}.pt"


def save_rows(prefix: str, rank: int, world: int, total_rows: int, arrays: Dict[str, torch.Tensor],
              save_id: int = 0) -> str:
    """Write this rank's rows.  `arrays`: name -> tensor whose first dimension is local_rows(total_rows, rank, world)
    (e.g. {"weight": [local, V], "lin": [local]}); tensors are moved to the CPU for writing.  `save_id`: one number per
    save, identical on every rank and in the meta file (load_rows refuses a mix)."""
    n = shard_plan.local_rows(total_rows, rank, world)
    out = {}
    for k, t in arrays.items():
        if t.shape[0] != n:
            raise ValueError(f"{k}: first dimension {t.shape[0]} != local rows {n} of rank {rank}/{world}")
        out[k] = t.detach().to("cpu").contiguous()
    path = shard_path(prefix, rank, world)
    tmp = path + ".tmp"
    torch.save({"format": FORMAT, "rank": rank, "world": world, "total_rows": int(total_rows), "save_id": int(save_id),
                "arrays": out}, tmp)
    os.replace(tmp, path)           # a crashed writer never leaves a half-written shard under the final name
    return path


def _meta_or_none(prefix: str) -> Optional[dict]:
    path = prefix + ".meta.pt"
    if not os.path.exists(path):
        return None
    return torch.load(path, map_location="cpu", weights_only=True)


def saved_world(prefix: str) -> int:
    """World size of the writer: from the meta file (the commit marker) when it records one, else from the one shard-0
    file name (checkpoints written shard by shard without a meta file, older checkpoints)."""
    meta = _meta_or_none(prefix)
    if meta is not None and meta.get("world") is not None:
        return int(meta["world"])
    d = os.path.dirname(prefix) or "."
    base = os.path.basename(prefix) + ".shard000-of-"
    hits = [f for f in os.listdir(d) if f.startswith(base) and f.endswith(".pt")]
    if len(hits) != 1:
        raise FileNotFoundError(f"no unique checkpoint shard 0 for prefix {prefix!r} (found {hits})")
    return int(hits[0][len(base):len(base) + 3])


def load_rows(prefix: str, rank: int, world: int, names: Optional[Sequence[str]] = None) -> Dict[str, torch.Tensor]:
    """The rows `rank` of `world` owns, assembled from a checkpoint written by any number of ranks (CPU tensors)."""
    ws = saved_world(prefix)
    meta = _meta_or_none(prefix)
    save_id = None if meta is None or meta.get("save_id") is None else int(meta["save_id"])
    out: Dict[str, torch.Tensor] = {}
    total = None
    mine = None
    # same world as the writer: this rank's rows are exactly saved shard `rank` (one file, no re-sharding);
    # otherwise every saved shard holds some of them (memory-mapped: only the needed rows are paged in)
    for s in ([rank] if ws == world else range(ws)):
        blob = torch.load(shard_path(prefix, s, ws), map_location="cpu", weights_only=True, mmap=(ws != world))
        if blob.get("format") != FORMAT or blob["rank"] != s or blob["world"] != ws:
            raise ValueError(f"{shard_path(prefix, s, ws)}: not a shard {s} of {ws} in format {FORMAT}")
        sid = int(blob.get("save_id", 0))
        if save_id is None:
            save_id = sid
        elif sid != save_id:
            raise ValueError(f"{shard_path(prefix, s, ws)} belongs to save {sid}, the checkpoint {prefix!r} to save "
                             f"{save_id}: files of different saves are mixed (a writer crashed between shards?)")
        if total is None:
            total = int(blob["total_rows"])
            mine = (torch.arange(rank, total, world, dtype=torch.int64) if rank < total
                    else torch.empty((0,), dtype=torch.int64))                    # global rows this rank owns
            assert mine.numel() == shard_plan.local_rows(total, rank, world)
        elif int(blob["total_rows"]) != total:
            raise ValueError("shards disagree on total_rows")
        sel = (mine % ws) == s                                                    # ... that live in saved shard s
        src = mine[sel] // ws
        dst = torch.nonzero(sel, as_tuple=False).reshape(-1)
        for k, t in blob["arrays"].items():
            if names is not None and k not in names:
                continue
            if k not in out:
                out[k] = torch.empty((mine.numel(),) + tuple(t.shape[1:]), dtype=t.dtype)
            out[k][dst] = t[src]
    return out


def save_meta(prefix: str, config: dict, replicated: Dict[str, torch.Tensor], world: Optional[int] = None,
              save_id: Optional[int] = None) -> str:
    """Rank 0, after every rank's save_rows has returned (barrier): config + replicated parameters + the writer's world
    size and save id.  Written last, it commits the save."""
    path = prefix + ".meta.pt"
    tmp = path + ".tmp"
    torch.save({"format": FORMAT, "config": config, "world": None if world is None else int(world),
                "save_id": None if save_id is None else int(save_id),
                "replicated": {k: v.detach().to("cpu").contiguous() for k, v in replicated.items()}}, tmp)
    os.replace(tmp, path)
    return path


def remove_stale_shards(prefix: str, world: int) -> int:
    """Delete shard files of this prefix written by OTHER world sizes (an earlier save at another scale): they would make
    the name-based `saved_world` ambiguous and waste space.  Returns how many were removed."""
    d = os.path.dirname(prefix) or "."
    base = os.path.basename(prefix) + ".shard"
    keep = f"-of-{world:03d}.pt"
    n = 0
    for f in os.listdir(d):
        if f.startswith(base) and f.endswith(".pt") and "-of-" in f and not f.endswith(keep):
            os.remove(os.path.join(d, f))
            n += 1
    return n


def load_meta(prefix: str) -> dict:
    blob = torch.load(prefix + ".meta.pt", map_location="cpu", weights_only=True)
    if blob.get("format") != FORMAT:
        raise ValueError(f"{prefix}.meta.pt: unknown checkpoint format {blob.get('format')}")
    return blob
