#!/usr/bin/env python
"""bench.py — site updates/sec of the red-black Metropolis sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 : the throughput micro-config of SURVEY.md §8(d): one 256^3 i32 chain, beta = 0.25,
        pre-thermalised by 200 sweeps from the seeded hot start.
N > 1 : the same 256 x 256 cross-section, 256 t-planes per GPU (lattice 256 x 256 x 256N),
        slab-decomposed along t with halo exchange: weak scaling of the headline config.
A "step" is SWEEPS_PER_STEP full sweeps of the whole lattice.  `value` is measured with the field
resident in HBM; `e2e` goes through the C ABI with host buffers (upload from pinned memory, the
Simulation::run loop body with energy every 10th sweep, download of field + energies).

--impl reference times the CPU restatement of the reference's own data path (oracle/, "port": the
reference is Rust and cannot be built here) on all host cores, one independent chain per core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

X = Y = 256
T_PER_GPU = 256
BETA = 0.25
SEED = 0xC0FFEE
THERMALISE = 200
SWEEPS_PER_STEP = 1000
CPU_SAMPLE_T = 32        # the CPU legs sweep a 256 x 256 x 32 slice per core (same row / plane strides as 256^3)
CPU_SAMPLE_SWEEPS = 24   # ~2 s wall on 16 cores: 20-40 core-seconds of CPU work
BYTES_PER_UPDATE = 8  # one i32 read + one i32 write per site update
METRIC = "site updates/sec"
UNIT = "site-updates/s"


def measured_traffic(lattice, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh).get("x".join(str(v) for v in lattice) + ":" + kernel)
        return (t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]) if t else (None, None)
    except Exception:
        return None, None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled (NVML, every 10 ms) while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[0].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception as exc:  # NVML missing: report it, never fake a clock
            self.reasons.add(f"nvml unavailable: {exc}")

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def cpu_baseline_sample(threads: int, sweeps: int, x: int, y: int, t: int):
    """Oracle port of the reference data path, one independent chain per thread."""
    from oracle import oracle as O
    lat = O.Lattice([x, y, t])
    secs, _ = O.bench_chains(lat, BETA, SEED, sweeps, 10, threads)
    updates = float(threads) * sweeps * lat.n_sites
    return updates / secs, secs


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    # bounded sample of the 256^3 workload: each step = one lexicographic sweep of a 256 x 256 x 32
    # slab-sized lattice per core (same row/plane strides as 256^3; the full 805 MB neighbour table of
    # 256^3 would only add build time), all cores busy, like sim.rs's rayon pool.
    t_sample = CPU_SAMPLE_T
    lat = O.Lattice([X, Y, t_sample])
    per_step = CPU_SAMPLE_SWEEPS  # sweeps per step and core; the clock covers the sweep loops only (orc_bench_chains)
    for _ in range(args.warmup):
        O.bench_chains(lat, BETA, SEED, per_step, 10, cores)
    t0 = time.time()
    secs = 0.0
    for _ in range(args.steps):
        s, _ = O.bench_chains(lat, BETA, SEED, per_step, 10, cores)
        secs += s
    updates = float(cores) * args.steps * per_step * lat.n_sites
    value = updates / secs
    sample = (f"{cores} independent chains x {args.steps} steps of {per_step} lexicographic sweeps of {X}x{Y}x{t_sample} per core "
              f"(neighbour-table i32 path of algorithm.rs:117-157; field and tables built outside the clock), "
              f"wall {time.time() - t0:.1f}s")
    cfg = workload_config(args.gpus)
    cfg["workload"] = (f"reference arm: {X}x{Y}x{t_sample} periodic lattice per core, {cores} independent chains (sim.rs:90-93), "
                       f"i32 heights, beta={BETA}, lexicographic Metropolis, {per_step} sweeps per step; a bounded sample of "
                       f"the {X}x{Y}x{T_PER_GPU * args.gpus} workload of the b200 arm")
    cfg["lattice_per_core"] = [X, Y, t_sample]
    cfg["sweeps_per_step"] = per_step
    for k in ("decomposition", "l2", "storage"):  # properties of the b200 arm's run, not of this one
        cfg.pop(k, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def other_configs(device: int):
    """The other BASELINE.json configs on one GPU, each timed briefly with CUDA events through the C ABI
    (device-resident, synthetic hot starts): updates/s and measurements/s.  Parity for these shapes is in
    tests/; these are throughput notes beside the headline line, not the headline."""
    import torch
    from lattice_qft_b200 import Engine

    def timed(eng, fn, reps):
        stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", device))
        fn()
        eng.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        eng.synchronize()
        return a.elapsed_time(b) * 1e-3 / reps

    out = {}
    # config 2: 64^3 coupling scan, 64 independent chains batched in one handle, energy every 10th sweep
    eng = Engine(64, 64, 64, n_chains=64, device=device)
    for c in range(64):
        eng.set_chain(c, 0.20 + 0.005 * c, 0xC0FFEE + c)
    eng.init_hot()
    eng.run(200, energy_every=0)
    sec = timed(eng, lambda: eng.run(200, sweep_base=200, energy_every=10), 3)
    out["coupling_scan_64x64^3"] = {"site_updates_per_s": 64 * 64**3 * 200 / sec,
                                    "energy_measurements_per_s": 64 * 20 / sec, "sweeps_per_call": 200}
    eng.close()
    # config 4: 64^3 Wilson sweep, widths {10,21,32,42,53} x heights {16,32,64} + one free chain, energy every sweep
    shapes = [(w, h) for h in (16, 32, 64) for w in (10, 21, 32, 42, 53)] + [(0, 0)]
    eng = Engine(64, 64, 64, n_chains=len(shapes), device=device)
    for c, (w, h) in enumerate(shapes):
        eng.set_chain(c, 0.26, 0xBEEF + c, w, h)
    eng.init_hot()
    eng.run(100, energy_every=0)
    sec = timed(eng, lambda: eng.run(100, sweep_base=100, energy_every=1), 3)
    out["wilson_scan_16x64^3"] = {"site_updates_per_s": len(shapes) * 64**3 * 100 / sec,
                                  "energy_measurements_per_s": len(shapes) * 100 / sec}
    eng.close()
    # config 3: 128^3 correlator, 100 reference sites per measurement (reference work: 100 * N site visits each):
    # (a) one lqft_correlator call per measurement through the ABI, (b) sim.rs:78's workload — correlator every 10th
    # sweep — as ONE lqft_run_observables call, (c) the same call measuring after EVERY sweep, whose extra time over a
    # measurement-free run is the device cost per measurement
    eng = Engine(128, 128, 128, device=device)
    eng.set_chain(0, 0.25, 0xABCD)
    eng.init_hot()
    eng.run(100, energy_every=0)
    sec = timed(eng, lambda: eng.correlator(0, reps=100, sweep_index=7), 50)
    out["correlator_128^3_100_sites"] = {"measurements_per_s": 1.0 / sec,
                                         "reference_equivalent_site_visits_per_s": 100 * 128**3 / sec,
                                         "what": "lqft_correlator: one launch (moments + reference heights + O(R*T) epilogue "
                                                 "on the device), one copy, one synchronisation"}
    sec0 = timed(eng, lambda: eng.run(200, sweep_base=100, energy_every=0), 3)
    out["correlator_128^3_100_sites"]["site_updates_per_s"] = 128**3 * 200 / sec0
    sec10 = timed(eng, lambda: eng.run_observables(200, sweep_base=100, correlator_every=10, correlator_reps=100), 3)
    sec1 = timed(eng, lambda: eng.run_observables(200, sweep_base=100, correlator_every=1, correlator_reps=100), 3)
    out["correlator_128^3_100_sites"]["sim_rs_workload"] = {
        "what": "sim.rs:78: sweep + correlator(100 reps) every 10th sweep + normalize every 10th, one lqft_run_observables call",
        "site_updates_per_s": 128**3 * 200 / sec10, "measurements_per_s": 20 / sec10,
        "in_run_measurements_per_s": 200 / max(sec1 - sec0, 1e-9),
        "in_run_us_per_measurement": 1e6 * (sec1 - sec0) / 200}
    eng.close()
    # config 3 as SURVEY 8(d) states it: the three couplings 0.22, 0.25, 0.28, which sim.rs:90-93 runs side by side
    # under rayon — here three chains of one handle, every half-sweep one launch over all of them
    eng = Engine(128, 128, 128, n_chains=3, device=device)
    for c, b in enumerate((0.22, 0.25, 0.28)):
        eng.set_chain(c, b, 0xABCD + c)
    eng.init_hot()
    eng.run(100, energy_every=0)
    sec10 = timed(eng, lambda: eng.run_observables(200, sweep_base=100, correlator_every=10, correlator_reps=100), 3)
    out["correlator_128^3_100_sites"]["sim_rs_workload_3_couplings"] = {
        "what": "the same call on three chains (beta 0.22, 0.25, 0.28) in one handle: sim.rs:90-93's par_iter over the couplings",
        "site_updates_per_s": 3 * 128**3 * 200 / sec10, "measurements_per_s": 3 * 20 / sec10}
    eng.close()
    # config 1: the reference's own run — 16^3, one chain per coupling of INVESTIGATE_ARY9 (lib.rs:50), energy every
    # 10th sweep; small enough to live in shared memory: the whole schedule is one launch (kernels_resident.cuh)
    for name, n_chains in (("reference_case_16^3_1_chain", 1), ("reference_run_16^3_10_couplings", 10)):
        eng = Engine(16, 16, 16, n_chains=n_chains, device=device)
        for c in range(n_chains):
            eng.set_chain(c, 0.20 + 0.01 * c, 1 + c)
        eng.init_hot()
        sec = timed(eng, lambda: eng.run(4000, energy_every=10), 2)
        out[name] = {"site_updates_per_s": n_chains * 4096 * 4000 / sec, "energy_measurements_per_s": n_chains * 400 / sec,
                     "launches_per_call": 1}
        eng.close()
    # energy / action measurement at the headline size: through the ABI (launch + copy + synchronise per call) and
    # inside a run (device cost: a run measuring after every sweep minus a measurement-free run)
    eng = Engine(256, 256, 256, device=device)
    eng.set_chain(0, BETA, SEED)
    eng.init_hot()
    eng.run(50, energy_every=0)
    sec = timed(eng, lambda: eng.energy(), 50)
    out["energy_256^3"] = {"measurements_per_s": 1.0 / sec, "GB_per_s_of_4B_per_site": 4 * 256**3 / sec / 1e9}
    sec0 = timed(eng, lambda: eng.run(100, sweep_base=50, energy_every=0), 3)
    sec1 = timed(eng, lambda: eng.run(100, sweep_base=50, energy_every=1), 3)
    per = max(sec1 - sec0, 1e-9) / 100
    out["energy_256^3"]["in_run_measurements_per_s"] = 1.0 / per
    out["energy_256^3"]["in_run_GB_per_s_of_4B_per_site"] = 4 * 256**3 / per / 1e9
    eng.close()
    return out


def fresh_nccl_id(world: int, rank: int):
    """A new ncclUniqueId for one more slab-decomposed handle, broadcast from rank 0."""
    import torch
    import torch.distributed as dist
    from lattice_qft_b200 import engine as E
    if world == 1:
        return None
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.tensor(list(E.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def parity_check(world: int, rank: int, device: int):
    """Before anything is timed: the slab-decomposed path of THIS launch (N ranks, peer-memory deep halo, graphs)
    against one GPU holding the whole lattice — 256 x 64 x 16N, 7 sweeps with acceptance counts, a 12-step run with
    energies and normalisation — post-sweep fields, counts and energies bit for bit on rank 0.  N = 1 checks the
    same kernels with the rank as its own neighbour (LQFT_FLAG_SELF_HALO)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lattice_qft_b200 import Engine, _lib
    x, y, t = 256, 64, 16 * world
    beta, seed = 0.25, 0x5EED + world
    flags = _lib.FLAG_SELF_HALO if world == 1 else 0
    eng = Engine(x, y, t, device=device, world_size=world, rank=rank, nccl_id=fresh_nccl_id(world, rank), flags=flags)
    eng.set_chain(0, beta, seed, 100, t // 2)
    eng.init_hot()
    acc = eng.sweep(0, 7, want_accepted=True)[0]
    e_obs, e_int, _ = eng.run(12, sweep_base=7, energy_every=3, normalize_every=10)
    mine = torch.from_numpy(eng.download(0)).cuda()
    plan = eng.plan_info()
    eng.close()
    acc_t = torch.tensor([acc], dtype=torch.int64, device="cuda")
    if world > 1:
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        dist.all_reduce(acc_t)
    else:
        parts = [mine]
    ok, detail = True, ""
    if rank == 0:
        whole = torch.cat(parts).cpu().numpy()
        ref = Engine(x, y, t, device=device, flags=_lib.FLAG_NO_NARROW)
        ref.set_chain(0, beta, seed, 100, t // 2)
        ref.init_hot()
        racc = ref.sweep(0, 7, want_accepted=True)[0]
        re_obs, re_int, _ = ref.run(12, sweep_base=7, energy_every=3, normalize_every=10)
        want = ref.download(0)
        ref.close()
        bad = int((whole != want).sum())
        ok = bad == 0 and int(acc_t.item()) == racc and bool((e_int == re_int).all()) and bool((e_obs == re_obs).all())
        if not ok:
            detail = f"{bad} sites differ, accepted {int(acc_t.item())} vs {racc}, energies {e_int.tolist()} vs {re_int.tolist()}"
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.broadcast(flag, 0)
    res = {"ranks": world, "ok": bool(flag.item()), "lattice": [x, y, t], "path": plan.get("path"),
           "ghost_depth": plan.get("ghost_depth"), "p2p": plan.get("p2p"),
           "what": "N-rank slab run == 1-GPU i32 run of the whole lattice: field, acceptance counts, in-run energies"}
    if detail:
        res["detail"] = detail
    return res


def config5_leg(world: int, rank: int, device: int, barrier):
    """BASELINE.json config 5 as a weak-scaling series: 1024 x 1024 x 128 planes per GPU (N = 8: 1024^3), 100 timed
    sweeps on the 16-bit kernel after warm-up, energy summed over the ranks.  Beside it, on the same GPUs in the same
    run, every rank sweeps a periodic 1024 x 1024 x 128 lattice of its own (no halo): the N = 1 rate the slab rate
    is divided by."""
    import torch
    import torch.distributed as dist
    from lattice_qft_b200 import Engine
    x = y = 1024
    tl, sweeps = 128, 100

    def rate(eng, sites_per_rank):
        stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", device))
        eng.init_hot()
        eng.sweep(0, 50)
        eng.sweep(50, sweeps)  # same call shape as the timed one: graphs instantiated
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        eng.sweep(50 + sweeps, sweeps)
        b.record(stream)
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e = float(eng.energy()[1][0])
        plan = eng.plan_info()
        segs = eng.narrow_segments()
        return sites_per_rank * sweeps / (float(ms.item()) * 1e-3), e, plan, segs

    solo = Engine(x, y, tl, device=device)
    solo.set_chain(0, BETA, SEED + rank)
    solo_rate, _, _, _ = rate(solo, x * y * tl)
    solo.close()
    out = {"lattice": [x, y, tl * world], "planes_per_gpu": tl, "sweeps_timed": sweeps,
           "single_gpu_periodic_slab_updates_per_s": solo_rate}
    if world == 1:
        out["site_updates_per_s"] = solo_rate
        out["per_gpu_updates_per_s"] = solo_rate
        out["efficiency_vs_single_gpu_slab"] = 1.0
        peak, _ = measured_peak_gbs()
        out["frac_hbm_roofline_8B_per_update"] = solo_rate * 8 / 1e9 / peak
        out["frac_hbm_roofline_kernel_bytes_6B"] = solo_rate * 6 / 1e9 / peak
        traffic, src = measured_traffic([x, y, tl], "k_sweep_n16")
        out["dram_bytes_per_half_sweep_ncu"] = traffic
        out["dram_bytes_source"] = src
        out["kernel_bytes_per_half_sweep_6B"] = 6 * x * y * tl // 2
        out["kernel_bytes_note"] = ("the 512 MiB field does not fit L2: a half-sweep reads both colours and writes one, 6 B per "
                                    "update; `dram_bytes_per_half_sweep_ncu` is what one launch moved under ncu")
        if traffic:
            out["dram_GBps_at_measured_rate"] = traffic * (solo_rate / (x * y * tl / 2)) / 1e9
            out["frac_hbm_peak_measured_traffic"] = out["dram_GBps_at_measured_rate"] / peak
        return out
    eng = Engine(x, y, tl * world, device=device, world_size=world, rank=rank, nccl_id=fresh_nccl_id(world, rank))
    eng.set_chain(0, BETA, SEED)
    per_rank, e, plan, segs = rate(eng, x * y * tl)
    eng.close()
    out["site_updates_per_s"] = per_rank * world
    out["per_gpu_updates_per_s"] = per_rank
    out["efficiency_vs_single_gpu_slab"] = per_rank / solo_rate
    out["energy_all_ranks"] = e
    out["ghost_depth"] = plan.get("ghost_depth")
    out["narrow_segments"] = segs
    return out


def replicas_leg(world: int, rank: int, device: int, barrier):
    """The reference's own parallelism (sim.rs:90-93, action_sim.rs:60-71: one chain per rayon worker) over N GPUs:
    a coupling scan of 64 chains of 64^3 per GPU, dealt round-robin (distributed.replica_engine), nothing exchanged.
    Timed on the device, max over the ranks; the energies of the whole scan are gathered once at the end."""
    import torch
    import torch.distributed as dist
    from lattice_qft_b200 import distributed as D
    n = 64 * world
    params = [(0.20 + 0.005 * (c % 64), 0xC0FFEE + c, 0, 0) for c in range(n)]
    eng, mine = D.replica_engine(64, 64, 64, params, device=device)
    eng.init_hot()
    eng.run(200, energy_every=10)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", device))
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    reps = 3
    for _ in range(reps):
        e_obs, _, _ = eng.run(200, sweep_base=200, energy_every=10)
    b.record(stream)
    barrier()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    table = D.gather_replicas(e_obs, mine, n)
    eng.close()
    rate = n * 64.0**3 * 200 * reps / (float(ms.item()) * 1e-3)
    return {"chains": n, "chains_per_gpu": 64, "lattice_per_chain": [64, 64, 64], "site_updates_per_s": rate,
            "per_gpu_updates_per_s": rate / world, "energies_gathered": int(table.shape[0] * table.shape[1]),
            "what": "replicas only: chains dealt round-robin to the ranks, no data-path collective"}


def workload_config(n_gpus: int):
    t = T_PER_GPU * n_gpus
    return {
        "workload": f"{X}x{Y}x{t} periodic lattice, 1 chain, i32 heights, beta={BETA}, red-black Metropolis, "
                    f"{SWEEPS_PER_STEP} sweeps per step, pre-thermalised {THERMALISE} sweeps from seeded hot start",
        "lattice": [X, Y, t], "sweeps_per_step": SWEEPS_PER_STEP, "beta": BETA,
        "decomposition": "none" if n_gpus == 1 else f"t-slabs of {T_PER_GPU} planes per GPU; 8 ghost planes per side, redundantly "
                         "updated for 8 half-sweeps; the last two half-sweeps of such a period push the boundary planes into "
                         "the neighbours' other ghost set over peer memory (NVLink), the next kernel's prologue flips the sets",
        "l2": "L2 flushed (256 MiB write) between timed steps; the 32 MiB/GPU field itself is L2-resident by nature",
        "storage": "the device field is stored as u16 (heights biased around the chain's normalisation offset; exact integers, "
                   "bit-identical to the i32 kernels and the oracle); it moves to i32 arrays only if a height leaves the "
                   "+-1023 guard band; host buffers at the ABI are i32",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE.json configs")
    ap.add_argument("--no-parity", action="store_true", help="skip the slab-vs-single-GPU bit-exactness check (experiments)")
    ap.add_argument("--flags", type=int, default=0, help="LQFT_FLAG_* for experiments")
    ap.add_argument("--sweeps", type=int, default=None, help="sweeps per step override (experiments)")
    ap.add_argument("--lattice", default=None, help="X,Y,Tpergpu override for experiments (not the headline config)")
    args = ap.parse_args()
    global X, Y, T_PER_GPU, SWEEPS_PER_STEP
    if args.lattice:
        X, Y, T_PER_GPU = (int(v) for v in args.lattice.split(","))
    if args.sweeps:
        SWEEPS_PER_STEP = args.sweeps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    from lattice_qft_b200 import Engine
    from lattice_qft_b200 import engine as E

    if not torch.cuda.is_available() or E.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: lattice_qft_b200 has no CPU fallback")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus}")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.tensor(list(E.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        nccl_id = bytes(buf.cpu().tolist())

    parity = None
    if not args.lattice and not args.no_parity:
        parity = parity_check(world, rank, local_rank)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "parity_n": parity, "error": "slab run differs from the "
                                  "single-GPU run; nothing was timed"}), flush=True)
            raise SystemExit(3)

    t_global = T_PER_GPU * world
    eng = Engine(X, Y, t_global, n_chains=1, device=local_rank, world_size=world, rank=rank, nccl_id=nccl_id,
                 flags=args.flags)
    eng.set_chain(0, BETA, SEED)
    eng.init_hot()
    eng.sweep(0, THERMALISE)
    eng.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n_local = eng.n_slab
    sweep_ctr = THERMALISE

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    # ---------------- device-resident value -------------------------------------------------------
    for _ in range(args.warmup):
        eng.sweep(sweep_ctr, SWEEPS_PER_STEP)
        sweep_ctr += SWEEPS_PER_STEP
    launches0 = eng.kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    for a, b in ev:
        flush_l2()
        a.record(stream)
        eng.sweep(sweep_ctr, SWEEPS_PER_STEP)
        b.record(stream)
        sweep_ctr += SWEEPS_PER_STEP
    barrier()
    clocks = sampler.stop()
    launches = eng.kernel_launches() - launches0
    narrow_segments = eng.narrow_segments()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    total_updates = float(X) * Y * t_global * SWEEPS_PER_STEP * args.steps
    value = total_updates / (ms * 1e-3)

    # ---------------- end to end through the C ABI with host buffers ---------------------------------
    host_in = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    host_out = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    np_in, np_out = host_in.numpy(), host_out.numpy()
    eng.download(0, np_in)
    e2e_steps = max(3, min(args.steps, 5))

    def e2e_step(base):
        eng.upload(0, np_in)                                   # H2D from pinned memory
        e_obs, _, _ = eng.run(SWEEPS_PER_STEP, first_step=0, sweep_base=base, energy_every=10, normalize_every=10)
        eng.download(0, np_out)                                # D2H of the step's result
        return e_obs

    for i in range(2):
        e2e_step(sweep_ctr)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e_obs = e2e_step(sweep_ctr + i * SWEEPS_PER_STEP)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(X) * Y * t_global * SWEEPS_PER_STEP * e2e_steps / float(te.item())
    h2d = n_local * 4
    d2h = n_local * 4 + int(e_obs.size) * 8

    eng.close()
    other = None
    if not args.lattice and not args.no_configs:
        c5 = config5_leg(world, rank, local_rank, barrier)
        other = other_configs(local_rank) if world == 1 else {}
        other["config5_1024x1024x128g"] = c5
        if world > 1:
            other["coupling_scan_replicas_64x64^3_per_gpu"] = replicas_leg(world, rank, local_rank, barrier)
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_sweep_launches = 2 * SWEEPS_PER_STEP * args.steps  # per rank, dominant kernel: half-sweep
        per_launch_s = (ms * 1e-3) / n_sweep_launches
        alg_bytes_per_launch = BYTES_PER_UPDATE * (n_local / 2)
        achieved = alg_bytes_per_launch / per_launch_s / 1e9
        # which kernel family did the sweeps: the 16-bit storage (kernels_narrow.cuh) or the i32 fallback kernel
        kernel = "k_sweep_n16" if narrow_segments > 0 else "k_sweep_vec4"
        moved = (6 if narrow_segments > 0 else 12) * (n_local / 2)
        traffic, traffic_src = measured_traffic([X, Y, T_PER_GPU], kernel) if world == 1 else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i32", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "what": "lqft_upload(pinned host field) + lqft_run(SWEEPS_PER_STEP steps: sweep, energy every 10th, "
                            "normalize every 10th) + lqft_download(field) per step, wall clock"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": f"{kernel}: half-sweep (one colour) of the red-black Metropolis sweep",
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "algorithmic_bytes_note": "SURVEY.md 8(d): 8 B per update = one i32 read + one i32 write, the "
                                                   "reference's storage type; `achieved` and `frac` use this figure",
                         "kernel_bytes_per_launch": moved,
                         "kernel_bytes_note": "what a red-black half-sweep on two colour arrays must move: both colours "
                                              "read, the own colour written = 6 B per update on the 16-bit storage (12 B on "
                                              "i32); frac_kernel_bytes = that / time / peak.  At 256^3 the 32 MiB field is L2 "
                                              "resident (`traffic`: the DRAM bytes of one cold launch); the 1024 x 1024 x 128 "
                                              "leg under `configs` is the HBM-bound case",
                         "frac_kernel_bytes": moved / per_launch_s / 1e9 / peak,
                         "avg_launch_us": per_launch_s * 1e6},
        }
        if parity is not None:
            line["parity_n"] = parity
        if other:
            line["configs"] = other
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, secs = cpu_baseline_sample(cores, CPU_SAMPLE_SWEEPS, X, Y, CPU_SAMPLE_T)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{cores} independent chains x {CPU_SAMPLE_SWEEPS} lexicographic sweeps of {X}x{Y}x{CPU_SAMPLE_T} "
                          f"({secs:.1f}s wall, {cores * secs:.0f} core-seconds); C restatement of the reference's "
                          "neighbour-table i32 path (oracle/lqft_oracle.c), hot start, energy every 10th sweep"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
