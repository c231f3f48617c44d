"""Worker of tests/test_gpu_multi.py: run under torchrun with one rank per GPU.

Checks that ONE lattice decomposed into t-slabs over the ranks gives, bit for bit, the result of the same
lattice on a single GPU (and of the CPU oracle at the small size): sweeps, run schedule, energy / action,
slice moments, correlator, normalize."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lattice_qft_b200 import Engine  # noqa: E402
from lattice_qft_b200 import distributed as D  # noqa: E402
from lattice_qft_b200 import engine as E  # noqa: E402


def gather_field(eng, local):
    rank, ws = D.world()
    mine = torch.from_numpy(local).cuda()
    parts = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(parts, mine)
    return torch.cat(parts).cpu().numpy()


def check(size, beta, seed, wilson, n_sweeps, flags=0):
    rank, ws = D.world()
    X, Y, T = size
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    full = np.random.default_rng(1234).integers(-128, 128, size=X * Y * T, dtype=np.int32)
    eng = D.slab_engine(X, Y, T, device=dev, flags=flags)
    eng.set_chain(0, beta, seed, *wilson)
    eng.upload(0, D.scatter_slab(full, X, Y, T, rank, ws))
    acc = eng.sweep(0, n_sweeps, want_accepted=True)[0]
    e_obs, e_int, _ = eng.run(12, first_step=0, sweep_base=n_sweeps, energy_every=4, normalize_every=10)
    g_obs, g_int, _ = eng.run(24, first_step=0, sweep_base=n_sweeps + 12, energy_every=2, normalize_every=4)  # graph replay
    e_now = eng.energy()
    s_now = eng.action()
    s1, s2 = eng.slice_moments(0)
    corr = eng.correlator(0, reps=10, sweep_index=5)
    got = gather_field(eng, eng.download(0))
    acc_t = torch.tensor([acc], dtype=torch.int64, device="cuda")
    dist.all_reduce(acc_t)
    eng.close()
    if rank == 0:
        ref = Engine(X, Y, T, device=dev, flags=flags)
        ref.set_chain(0, beta, seed, *wilson)
        ref.upload(0, full)
        racc = ref.sweep(0, n_sweeps, want_accepted=True)[0]
        r_obs, r_int, _ = ref.run(12, first_step=0, sweep_base=n_sweeps, energy_every=4, normalize_every=10)
        q_obs, q_int, _ = ref.run(24, first_step=0, sweep_base=n_sweeps + 12, energy_every=2, normalize_every=4)
        assert (q_int == g_int).all() and (q_obs == g_obs).all()
        assert int(acc_t.item()) == racc, (int(acc_t.item()), racc)
        assert (r_int == e_int).all() and (r_obs == e_obs).all()
        assert (ref.energy()[0] == e_now[0]).all() and (ref.action()[0] == s_now[0]).all()
        r1, r2 = ref.slice_moments(0)
        assert (r1 == s1).all() and (r2 == s2).all()
        assert (ref.correlator(0, reps=10, sweep_index=5) == corr).all()
        want = ref.download(0)
        assert (want == got).all(), f"{int((want != got).sum())} sites differ for {size} ws={ws}"
        ref.close()
        print(f"ok {size} ws={ws} flags={flags} accepted={racc}", flush=True)
    dist.barrier()


def check_replicas(size, n_chains, n_steps):
    """Replicas only (sim.rs:90-93): the chains of a coupling / Wilson scan dealt round-robin to the ranks give the
    energies of ONE handle holding all of them."""
    rank, ws = D.world()
    X, Y, T = size
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    params = [(0.20 + 0.01 * c, 100 + c, (5 * c) % X, (3 * c) % T) for c in range(n_chains)]
    eng, mine = D.replica_engine(X, Y, T, params, device=dev)
    e_int = np.zeros((0, D.E.run_measurements(0, n_steps, 2)))
    if eng is not None:
        eng.init_hot()
        _, e_int, _ = eng.run(n_steps, energy_every=2, normalize_every=10)
        eng.close()
    table = D.gather_replicas(e_int, mine, n_chains)
    if rank == 0:
        ref = Engine(X, Y, T, n_chains=n_chains, device=dev)
        for c, (beta, seed, w, h) in enumerate(params):
            ref.set_chain(c, beta, seed, w, h)
        ref.init_hot()
        _, r_int, _ = ref.run(n_steps, energy_every=2, normalize_every=10)
        ref.close()
        assert (table == r_int.astype(np.float64)).all(), "replica energies differ from the single handle"
        print(f"ok replicas {size} chains={n_chains} ws={ws}", flush=True)
    dist.barrier()


def main():
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    ws = dist.get_world_size()
    check((16, 16, 8 * ws), 0.25, 77, (0, 0), 4)
    check((8, 6, 4 * ws), 0.3, 5, (5, 3), 5)
    check((64, 32, 4 * ws), 0.22, 9, (21, 6), 3)
    check((256, 16, 4 * ws), 0.25, 11, (0, 0), 3)           # 16-bit kernel on slabs: pushes over peer memory, ghost tasks
    check((256, 16, 4 * ws), 0.25, 11, (0, 0), 3, flags=256)  # LQFT_FLAG_NO_NARROW: i32 kernels, halos by ncclSend/Recv
    check((256, 4, 40 * ws), 0.27, 13, (200, 50), 5)        # narrow kernel, several plane groups per slab, Wilson sheet
    check((256, 16, 4 * ws), 0.25, 11, (0, 0), 3, flags=2)  # scalar kernels on slabs
    check((512, 8, 2 * ws), 0.25, 12, (100, 3), 2)
    check_replicas((64, 16, 16), 2 * ws + 1, 12)             # uneven deal: rank 0 holds one chain more
    check_replicas((16, 16, 16), max(ws - 1, 1), 10)         # fewer chains than ranks: one rank has no work
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
