"""N > 1 host logic on CPU: world_size-2 `gloo` processes run the slab decomposition with the oracle as
the per-rank compute (test double) and the product's slab plan / halo schedule / gather helpers; the
result must equal the single-lattice oracle bit for bit (RNG keyed by the global site index)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import seeded_field

X, Y, T = 8, 6, 8
BETA, SEED, SWEEPS = 0.27, 97, 3
WILSON = (5, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from lattice_qft_b200 import distributed as D
    from lattice_qft_b200 import engine as E
    from oracle import oracle as O
    full = seeded_field(X * Y * T)
    t0, nt = E.slab(T, ws, rank)
    local = D.scatter_slab(full, X, Y, T, rank, ws)
    plane = X * Y
    buf = np.zeros((nt + 2) * plane, dtype=np.int32)
    buf[plane:(nt + 1) * plane] = local
    D.exchange_halos(buf, X, Y, nt)
    acc = 0
    for s in range(SWEEPS):
        for colour in (0, 1):
            acc += O.half_sweep_slab(X, Y, T, t0, nt, buf, colour, BETA, SEED, s, WILSON)
            D.exchange_halos(buf, X, Y, nt)
    total_acc = D.allreduce_sum([acc])[0]
    gathered = D.gather_slabs(buf[plane:(nt + 1) * plane])
    # replicas: chains dealt round-robin, nothing exchanged
    mine = D.chains_of_rank(7, rank, ws)
    assert mine == list(range(rank, 7, ws))
    # ... but the results: each rank's rows land at their global chain ids (oracle energies as the per-chain payload)
    lat = O.Lattice([X, Y, T])
    rows = np.array([[O.integer_energy(lat, seeded_field(X * Y * T, seed=c))[0], c] for c in mine], dtype=np.float64).reshape(len(mine), 2)
    table = D.gather_replicas(rows, mine, 7)
    assert table.shape == (7, 2) and (table[:, 1] == np.arange(7)).all()
    assert all(table[c, 0] == O.integer_energy(lat, seeded_field(X * Y * T, seed=c))[0] for c in range(7))
    # the NCCL-id bootstrap path works on any backend
    got = D.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    assert got == bytes(range(128))
    if rank == 0:
        np.save(out_path, np.concatenate([gathered, np.array([int(total_acc)], dtype=np.int32)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 4])
def test_slab_decomposition_invariance_gloo(tmp_path, ws):
    from oracle import oracle as O
    out = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    got = np.load(out)
    lat = O.Lattice([X, Y, T])
    want = seeded_field(X * Y * T)
    wil = O.Wilson(lat, *WILSON)
    acc = sum(O.sweep(lat, want, BETA, SEED, s, order=1, wilson=wil) for s in range(SWEEPS))
    assert (got[:-1] == want).all()
    assert got[-1] == acc


def test_scatter_and_plan_consistency():
    from lattice_qft_b200 import distributed as D
    from lattice_qft_b200 import engine as E
    full = np.arange(X * Y * T, dtype=np.int32)
    parts = [D.scatter_slab(full, X, Y, T, r, 4) for r in range(4)]
    assert (np.concatenate(parts) == full).all()
    assert [E.slab(T, 4, r) for r in range(4)] == [(0, 2), (2, 2), (4, 2), (6, 2)]
