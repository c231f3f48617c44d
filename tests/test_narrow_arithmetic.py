"""The packed 16-bit arithmetic of the narrow sweep kernel (lattice_qft_b200/csrc/kernels_narrow.cuh), restated in
numpy and compared with the plain integer decision of the spec over the whole range the host
admits: heights within 1023 of the centre at entry, up to 1000 sweeps later (|h - centre| <= 2023), Wilson
offsets, every table length up to 256.  No device needed: this pins the arithmetic, the GPU tests pin the kernel."""
import numpy as np
import pytest

from lattice_qft_b200 import engine as E

BIAS = 2048
C = 0x60006000


def s16(x):
    """Low 16 bits as a signed value (what add.s16x2 / min.s16x2 see)."""
    x = np.asarray(x, dtype=np.int64) & 0xFFFF
    return np.where(x >= 0x8000, x - 0x10000, x)


def packed_index(v_lo, v_hi, nb_lo, nb_hi, off_lo, off_hi, kc):
    """u = 6 v + C - sum (+ packed offsets) in 32-bit wrap-around arithmetic, then add.s16x2 + min.s16x2.relu."""
    m32 = 0xFFFFFFFF
    v = (v_lo + (v_hi << 16)) & m32
    ns = np.zeros_like(v)
    for a, b in zip(nb_lo, nb_hi):
        ns = (ns + a + (b << 16)) & m32               # plain 32-bit adds of packed pairs
    u = (6 * v + C - ns) & m32
    u = (u + (off_lo & m32) + ((off_hi << 16) & m32)) & m32   # (uint32)lo + ((uint32)hi << 16), as n16_wilson builds it
    c_off = (kc - 24576) & 0xFFFF
    lo = np.clip(s16((u & 0xFFFF) + c_off), None, 2 * kc)      # add wraps, min, then relu
    hi = np.clip(s16((u >> 16) + c_off), None, 2 * kc)
    return np.maximum(lo, 0), np.maximum(hi, 0)


@pytest.mark.parametrize("beta", [0.25, 0.05, 0.7])
def test_packed_pair_arithmetic_equals_the_plain_decision_index(beta):
    thr = E.threshold_table(beta)
    assert len(thr) <= 256                       # what narrow_geometry_ok admits
    kc = max(len(thr) - 1, 3)
    rng = np.random.default_rng(int(beta * 1000))
    n = 200_000
    lim = 1023 + 1000                            # guard band + one segment of sweeps
    centre = int(rng.integers(-50_000, 50_000))

    def heights(extreme):
        h = rng.integers(-lim, lim + 1, size=n, dtype=np.int64)
        if extreme:
            h[: n // 4] = rng.choice([-lim, lim], size=n // 4)
        return h + centre

    for extreme in (False, True):
        v = [heights(extreme), heights(extreme)]
        nb = [[heights(extreme) for _ in range(6)] for _ in range(2)]
        off = [rng.integers(-1, 2, size=n, dtype=np.int64) for _ in range(2)]
        b = lambda h: h - centre + BIAS          # noqa: E731  (k_narrow)
        assert all(((b(x) >= 0) & (b(x) < 4096)).all() for x in v + nb[0] + nb[1])
        i_lo, i_hi = packed_index(b(v[0]), b(v[1]), [b(x) for x in nb[0]], [b(x) for x in nb[1]], off[0], off[1], kc)
        for half, idx in ((0, i_lo), (1, i_hi)):
            d = 6 * v[half] - sum(nb[half]) + off[half]
            assert (idx == np.clip(d, -kc, kc) + kc).all()


def test_interleaved_table_equals_two_sided_rule():
    """Entry (i << 1) | coin of build_interleaved: d = i - kc, coin 1 proposes +1 (m = d), coin 0 proposes -1 (m = -d);
    accept iff r <= thr[clamp(m + 3)] | coin << 31 — the rule every other kernel uses (metropolis_delta)."""
    thr = E.threshold_table(0.25)
    n_thr = len(thr)
    kc = max(n_thr - 1, 3)
    rng = np.random.default_rng(3)
    r = rng.integers(0, 2**32, size=100_000, dtype=np.uint64)
    d = rng.integers(-kc - 5, kc + 6, size=r.size)
    coin = (r >> 31).astype(np.int64)
    i = np.clip(d, -kc, kc) + kc
    entry = (i << 1) | coin                                   # __funnelshift_l(r, i, 1)
    up = entry & 1
    m = np.where(up == 1, (entry >> 1) - kc, kc - (entry >> 1))
    t = thr[np.clip(m + 3, 0, n_thr - 1)].astype(np.uint64) | (up.astype(np.uint64) << 31)
    got = np.where(r <= t, np.where(up == 1, 1, -1), 0)
    s = np.where(coin == 1, 1, -1)
    u = (r & 0x7FFFFFFF).astype(np.int64)
    want = np.where(u <= thr[np.clip(s * np.clip(d, -kc, kc) + 3, 0, n_thr - 1)], s, 0)
    assert (got == want).all()


def test_deep_halo_wedges_tile_every_half_sweep():
    """launch_sweep's split of the half-sweeps that run underneath an exchange: interior + two wedges cover exactly
    the planes [-extra, Tl + extra), the wedges are even-sized, start at the parity the single-range launch has,
    and never touch planes the exchange is still reading ([0, depth) and [Tl - depth, Tl)) before it has ended."""
    for Tl in (24, 40, 256):
        for depth in (2, 4, 8):
            for overlap in (1, 2, 3):
                if Tl < 2 * (depth + overlap + 1):
                    continue
                for k in range(overlap):
                    extra = depth - 1 - k
                    if extra < 0:
                        break
                    a = depth + k + 1
                    interior = set(range(a, Tl - a))
                    lo = set(range(-extra, -extra + extra + a))
                    hi = set(range(Tl - a, Tl - a + extra + a))
                    assert interior | lo | hi == set(range(-extra, Tl + extra))
                    assert not (interior & lo) and not (interior & hi) and not (lo & hi)
                    assert (extra + a) % 2 == 0 and (Tl - 2 * a) % 2 == 0
                    assert ((Tl - a) - (extra + a) - (-extra)) % 2 == 0      # second range continues the parity
                    assert not interior & (set(range(depth)) | set(range(Tl - depth, Tl)))
                    # what the k-th interior reads (t-neighbours) was completed by the (k-1)-th interior
                    if k:
                        prev = set(range(a - 1, Tl - a + 1))
                        assert {p - 1 for p in interior} | {p + 1 for p in interior} <= prev
