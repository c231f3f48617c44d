"""Mirror of src/lattice.rs: lattice geometry (index order, neighbour order).

The reference tabulates neighbours (`Lattice::values`, lattice.rs:65-80, 48 B/site).  Here the
table is never built: rows are computed on demand through the host helpers of liblqft
(lqft_host_neighbours) — the device kernels compute them in registers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

from ._lib import lib


class LatticeCoords:
    """lattice.rs:9-43."""

    def __init__(self, array: Sequence[int]):
        self._c = [int(v) for v in array]

    @staticmethod
    def new(array: Sequence[int]) -> "LatticeCoords":
        return LatticeCoords(array)

    def cleanup(self, size: Sequence[int]) -> "LatticeCoords":
        return LatticeCoords([c % s for c, s in zip(self._c, size)])

    def move_one(self, size: Sequence[int], direction: int) -> "LatticeCoords":
        d = len(self._c)
        c = list(self._c)
        axis = direction % d
        c[axis] = c[axis] + 1 if direction < d else c[axis] + (size[axis] - 1)
        return LatticeCoords(c).cleanup(size)

    def into_array(self) -> List[int]:
        return list(self._c)


class Lattice:
    """Lattice<D, SIZE> (lattice.rs:53-123).  `values[index]` is the neighbour row
    [+dir_0 .. +dir_{D-1}, -dir_0 .. -dir_{D-1}] (lattice.rs:49-51)."""

    def __init__(self, size: Sequence[int]):
        self.size = [int(s) for s in size]
        self.d = len(self.size)
        self.n_sites = 1
        for s in self.size:
            self.n_sites *= s
        self.values = _NeighbourRows(self)

    @staticmethod
    def new(size: Sequence[int]) -> "Lattice":
        return Lattice(size)

    def get_neighbours_array(self, index: int) -> List[int]:
        if self.d == 3:
            out = (C.c_uint64 * 6)()
            lib().lqft_host_neighbours(self.size[0], self.size[1], self.size[2], index, out)
            return [int(v) for v in out]
        coords = self.calc_coords_from_index(index)
        return [self.calc_index_from_coords(coords.move_one(self.size, direction))
                for direction in range(2 * self.d)]

    def pos_neighbours_array(self, index: int) -> List[int]:
        return self.get_neighbours_array(index)[: self.d]

    def calc_index_from_coords(self, coords) -> int:
        c = coords.into_array() if isinstance(coords, LatticeCoords) else list(coords)
        if self.d == 3:
            return int(lib().lqft_host_index(self.size[0], self.size[1], self.size[2],
                                             c[0] % self.size[0], c[1] % self.size[1], c[2] % self.size[2]))
        total, stride = 0, 1
        for i in range(self.d):
            total += (c[i] % self.size[i]) * stride
            stride *= self.size[i]
        return total

    def calc_coords_from_index(self, index: int) -> LatticeCoords:
        if self.d == 3:
            out = (C.c_uint32 * 3)()
            lib().lqft_host_coords(self.size[0], self.size[1], self.size[2], index, out)
            return LatticeCoords([int(v) for v in out])
        coords, stride = [], 1
        for i in range(self.d):
            coords.append((index // stride) % self.size[i])
            stride *= self.size[i]
        return LatticeCoords(coords)

    def get_random_lattice_point_index(self, rng) -> int:
        """lattice.rs:120-122; `rng` is an algorithm.SweepRng (replaces ThreadRng)."""
        return rng.gen_site(self.n_sites)


class _NeighbourRows:
    """Indexable stand-in for `Lattice::values: Vec<[usize; 2D]>`."""

    def __init__(self, lattice: Lattice):
        self._l = lattice

    def __len__(self):
        return self._l.n_sites

    def __getitem__(self, index: int) -> List[int]:
        if not 0 <= index < self._l.n_sites:
            raise IndexError(index)
        return self._l.get_neighbours_array(index)

    def __iter__(self):
        for i in range(self._l.n_sites):
            yield self._l.get_neighbours_array(i)


class Lattice3d(Lattice):
    """Lattice3d<MAX_X, MAX_Y, MAX_T> (lattice.rs:127-161)."""

    def __init__(self, max_x: int, max_y: int, max_t: int):
        super().__init__([max_x, max_y, max_t])

    @staticmethod
    def new(max_x: int, max_y: int, max_t: int) -> "Lattice3d":  # type: ignore[override]
        return Lattice3d(max_x, max_y, max_t)

    def to_lattice2d(self) -> "Lattice2d":
        return Lattice2d(self.size[0], self.size[1])


class Lattice2d(Lattice):
    """Lattice2d<MAX_X, MAX_Y> (lattice.rs:165-189)."""

    def __init__(self, max_x: int, max_y: int):
        super().__init__([max_x, max_y])
