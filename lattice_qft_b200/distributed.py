"""Multi-GPU plumbing over torch.distributed (one process per GPU).

Two ways to use several GPUs, as in SURVEY.md §8(e):
  * ONE lattice as t-slabs: `slab_engine()` gives every rank an Engine that owns T/world planes; the
    halo exchange (ncclSend/Recv or peer stores, overlapped with the interior update) lives inside
    liblqft.  torch.distributed is only the bootstrap: it broadcasts the NCCL id.
  * independent chains ("replicas only", like sim.rs:90-93): `chains_of_rank()` deals the chains of
    a coupling scan round-robin to the ranks; nothing is exchanged but the final results.

The helpers that only shuffle host arrays (`scatter_slab`, `gather_slabs`, `exchange_halos`) work on
any backend (gloo on CPU in the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import engine as E
from .engine import Engine


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def broadcast_bytes(payload: Optional[bytes], n: int, src: int = 0, device=None) -> bytes:
    """Broadcast `n` bytes from rank `src` (used for the 128-byte NCCL unique id)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(n, dtype=torch.uint8, device=device or "cpu")
    if dist.get_rank() == src:
        buf.copy_(torch.tensor(list(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().tolist())


def slab_engine(x: int, y: int, t: int, n_chains: int = 1, device: int = 0, flags: int = 0) -> Engine:
    """Engine of this rank's t-slab of one global x*y*t lattice (all ranks call it collectively)."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return Engine(x, y, t, n_chains=n_chains, device=device, flags=flags)
    on_gpu = dist.get_backend() == "nccl"
    nccl_id = broadcast_bytes(E.comm_unique_id() if rank == 0 else None, 128, 0,
                              torch.device("cuda", device) if on_gpu else None)
    return Engine(x, y, t, n_chains=n_chains, device=device, world_size=ws, rank=rank, nccl_id=nccl_id, flags=flags)


def chains_of_rank(n_chains: int, rank: int, world_size: int) -> List[int]:
    """Round-robin deal of independent chains to ranks (coupling / Wilson scans)."""
    return list(range(rank, n_chains, world_size))


def replica_engine(x: int, y: int, t: int, params: Sequence[Sequence], device: int = 0, flags: int = 0):
    """Engine holding this rank's share of a scan of independent chains — sim.rs:90-93 / action_sim.rs:60-71 hand the
    scan points to rayon workers; here `params[c] = (beta, seed, wilson_width, wilson_height)` of chain c goes to rank
    c mod world_size.  Nothing is exchanged between the ranks.  Returns (engine, global ids of its chains) or
    (None, []) for a rank without work."""
    rank, ws = world()
    mine = chains_of_rank(len(params), rank, ws)
    if not mine:
        return None, []
    eng = Engine(x, y, t, n_chains=len(mine), device=device, flags=flags)
    for slot, c in enumerate(mine):
        beta, seed, *wil = params[c]
        eng.set_chain(slot, beta, seed, *(wil if wil else (0, 0)))
    return eng, mine


def gather_replicas(per_chain: np.ndarray, mine: Sequence[int], n_chains: int) -> np.ndarray:
    """Rows of `per_chain` (this rank's chains, in slot order) put at their global chain ids, summed over the ranks:
    every rank gets the [n_chains, ...] table of the whole scan."""
    import torch
    import torch.distributed as dist
    per_chain = np.asarray(per_chain, dtype=np.float64)
    full = np.zeros((n_chains,) + per_chain.shape[1:], dtype=np.float64)
    for slot, c in enumerate(mine):
        full[c] = per_chain[slot]
    if world()[1] > 1:
        on_gpu = dist.get_backend() == "nccl"
        tt = torch.from_numpy(full).cuda() if on_gpu else torch.from_numpy(full)
        dist.all_reduce(tt)
        full = tt.cpu().numpy()
    return full


def scatter_slab(values: np.ndarray, x: int, y: int, t: int, rank: int, world_size: int) -> np.ndarray:
    """This rank's slab of a global lexicographic field (t is the slowest index: a contiguous range)."""
    t0, nt = E.slab(t, world_size, rank)
    plane = x * y
    return np.ascontiguousarray(values.reshape(-1)[t0 * plane:(t0 + nt) * plane])


def gather_slabs(local: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
    """Concatenate the ranks' slabs in rank order on `dst` (host arrays, any backend)."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return local.copy()
    mine = torch.from_numpy(np.ascontiguousarray(local))
    parts = [torch.empty_like(mine) for _ in range(ws)] if rank == dst else None
    dist.gather(mine, parts, dst=dst)
    return np.concatenate([p.numpy() for p in parts]) if rank == dst else None


def allreduce_sum(values: Sequence[float]) -> np.ndarray:
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64)
    if world()[1] > 1:
        dist.all_reduce(t)
    return t.numpy()


def exchange_halos(slab_with_ghosts: np.ndarray, x: int, y: int, nt: int) -> None:
    """Host-array form of liblqft's halo schedule: my first owned plane becomes the lower neighbour's
    upper ghost, my last owned plane the upper neighbour's lower ghost (periodic ring of ranks).
    `slab_with_ghosts` holds nt + 2 planes; in place."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    plane = x * y
    v = slab_with_ghosts.reshape(nt + 2, plane)
    if ws == 1:
        v[0], v[nt + 1] = v[nt].copy(), v[1].copy()
        return
    lo, hi = (rank - 1) % ws, (rank + 1) % ws
    send_lo, send_hi = torch.from_numpy(v[1].copy()), torch.from_numpy(v[nt].copy())
    recv_hi, recv_lo = torch.empty_like(send_lo), torch.empty_like(send_lo)
    # same pairing as exchange_halo_nccl: sends (lo, hi) then receives (hi, lo); with two ranks both
    # neighbours are the same peer and messages match in posting order.
    ops = [dist.P2POp(dist.isend, send_lo, lo, tag=0), dist.P2POp(dist.isend, send_hi, hi, tag=1),
           dist.P2POp(dist.irecv, recv_hi, hi, tag=0), dist.P2POp(dist.irecv, recv_lo, lo, tag=1)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    v[nt + 1] = recv_hi.numpy()
    v[0] = recv_lo.numpy()
