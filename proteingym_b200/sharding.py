"""Multi-GPU plumbing for the scorer: assays are independent, so ranks never exchange activations.
One process per GPU; weights are read/built on rank 0 and broadcast once, assays are assigned by LPT on an analytic
cost, per-assay score vectors are gathered once at the end. Works on any torch.distributed backend (NCCL on the GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

import torch


def rank_world():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def barrier_if_distributed():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def atomic_write(path: str, write_fn):
    """Write through a temporary file + os.replace so that a concurrent reader never sees a truncated file."""
    import os
    tmp = f"{path}.tmp.{os.getpid()}"
    write_fn(tmp)
    os.replace(tmp, path)


def assay_cost(L: int, layers: int, d: int, ffn: int, n_positions: int | None = None, window: int = 1024) -> float:
    """Algorithmic FLOPs of one masked-marginal assay: P * F_fwd(T), T = min(L+2, window)."""
    T = min(L + 2, window)
    P = L if n_positions is None else n_positions
    return float(P) * (layers * (2.0 * T * (4 * d * d + 2 * d * ffn) + 4.0 * T * T * d) + 2.0 * T * (d * d + d * 33))


def lpt_assign(costs, world: int):
    """Longest-processing-time-first: returns ``assignment[rank] = [assay indices]`` (deterministic; ties by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return out


def position_chunk(P: int, world: int, rank: int):
    """Secondary partitioning: the P masked positions of ONE assay in contiguous chunks of ceil(P / world) rows.
    Returns (lo, hi, chunk). Equal chunk size keeps the all-gather a single fixed-size collective; the last ranks may get fewer
    (or no) rows. Rows are computed independently of how they are batched, so scores are bit-identical for any world size."""
    chunk = (P + world - 1) // world if P > 0 else 0
    lo = min(P, rank * chunk)
    return lo, min(P, lo + chunk), chunk


def all_gather_rows(full: torch.Tensor, chunk: int, group=None) -> torch.Tensor:
    """``full`` [world * chunk, V]: this rank has written its rows at [rank * chunk, ...). After the call every rank holds every
    chunk (the only data-path collective of the position-partitioned mode: <= 135 KB per assay). Returns ``full``."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if chunk == 0 or world == 1:
        return full
    views = [full[r * chunk:(r + 1) * chunk] for r in range(world)]
    try:
        dist.all_gather_into_tensor(full, views[rank].clone(), group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        dist.all_gather(views, views[rank].clone(), group=group)
    return full


def broadcast_state(state: dict | None, src: int = 0, device=None) -> dict:
    """Rank ``src`` holds ``state`` (name -> fp32 tensor); every rank returns an identical copy on ``device``."""
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        meta = [[(k, tuple(v.shape)) for k, v in sorted(state.items())]]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for k, shape in meta[0]:
        if rank == src:
            t = state[k].to(device=device, dtype=torch.float32).contiguous()
        else:
            t = torch.empty(shape, dtype=torch.float32, device=device)
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def gather_scores(local: dict, dst: int = 0, device=None):
    """``local``: assay index -> 1-D float32 score tensor computed on this rank. Returns the merged dict on ``dst``
    (None elsewhere). One size exchange + one padded gather."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    keys = sorted(local)
    sizes = [None] * world
    dist.all_gather_object(sizes, [(k, int(local[k].numel())) for k in keys])
    tot = [sum(n for _, n in s) for s in sizes]
    width = max(tot + [1])
    buf = torch.zeros(width, dtype=torch.float32, device=device)
    if keys:
        buf[:tot[rank]] = torch.cat([local[k].to(device=device, dtype=torch.float32).reshape(-1) for k in keys])
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for r in range(world):
        off = 0
        for k, n in sizes[r]:
            merged[k] = bufs[r][off:off + n].cpu().numpy()
            off += n
    return merged
