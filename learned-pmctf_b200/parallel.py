"""GOP-sharded data parallelism.

Nothing crosses a GOP boundary in the reference (test_pMCTF_flex.py:131-141: frames are read per GOP, the
dpb is reset per stage, frames_coded / frames_orig are per-GOP lists), so the work items
(q_index, sequence, gop_idx) are partitioned statically over one process per GPU with NO data-path
collective.  The only exchange is one all_gather of the fixed-shape fp64 per-frame statistics at the
end (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests)."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the default process group when
    WORLD_SIZE > 1.  One process per GPU: LOCAL_RANK selects the device."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def work_items(q_indices: Sequence[int], n_sequences: int, gops_per_sequence: int) -> List[Tuple[int, int, int]]:
    """Flattened (q_index, sequence, gop_idx) list in the reference's loop order (test_pMCTF_flex.py:455-498, :131)."""
    return [(q, s, g) for q in q_indices for s in range(n_sequences) for g in range(gops_per_sequence)]


def shard(items: Sequence, rank: int, world: int) -> List:
    """Static round-robin partition: rank r owns items r, r+world, ...  Every item is owned exactly once and the
    load differs by at most one item between ranks."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(items[rank::world])


def max_items_per_rank(n_items: int, world: int) -> int:
    return (n_items + world - 1) // world


def gather_stats(local: torch.Tensor, n_items: int, rank: int, world: int) -> torch.Tensor:
    """local: fp64 [items_of_this_rank, frames, fields] -> [n_items, frames, fields] on every rank, rows restored
    to the global item order of `shard`.  One all_gather of a fixed-shape (padded) tensor."""
    if world == 1:
        return local
    cap = max_items_per_rank(n_items, world)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.size(0)] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    out = out.view(world, cap, *local.shape[1:])
    rows = [out[i % world, i // world] for i in range(n_items)]
    return torch.stack(rows)


def max_over_ranks(value: float, device) -> float:
    """Device-side MAX all-reduce of a timing (ms)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
