"""Mirror of src/fields.rs for the hot path: Field, WilsonField, the Action and HeightField traits.

A `Field` owns a host copy of the heights (lexicographic, like `Field::values`) and, lazily, a
device-resident copy inside a liblqft handle.  Whole-lattice work (sweeps, energy, action,
normalize) runs on the device; the two copies are reconciled on demand, so code written against
the reference API (`field.values[...]`, `field.integer_energy()`, `algo.field_sweep(&mut field..)`)
keeps its meaning.  Single-site accessors (`local_action`, `bond_difference`, ...) are O(1) host
arithmetic on the synced values; they are conveniences of the trait, not the hot path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from ._lib import lib
from .engine import Engine, pick_site
from . import _lib
from .lattice import Lattice, LatticeCoords


class Field:
    """Field<'a, i32, 3, SIZE> (fields.rs:17-68)."""

    def __init__(self, lattice: Lattice, values: Optional[np.ndarray] = None, device: int = 0):
        self.lattice = lattice
        self._device = device
        n = lattice.n_sites
        self._host = np.zeros(n, dtype=np.int32) if values is None else np.array(values, dtype=np.int32).reshape(-1)
        if self._host.size != n:
            raise AssertionError(f"values.len() = {self._host.size} != SIZE = {n}")  # fields.rs:64
        self._host_valid = True
        self._dev_valid = False
        self._engine: Optional[Engine] = None
        self._params = None  # (beta, seed, w, h) currently set on the engine

    # -- constructors (fields.rs:32-67) -------------------------------------------------------
    @staticmethod
    def new(lattice: Lattice, device: int = 0) -> "Field":
        return Field(lattice, None, device)

    @staticmethod
    def random(lattice: Lattice, seed: int = 0, device: int = 0) -> "Field":
        """Field::<i8>::random widened to i32 (computation.rs:225-226): uniform in [-128,127],
        generated on the device from stream LQFT_STREAM_HOTSTART of `seed`."""
        f = Field(lattice, None, device)
        eng = f._need_engine()
        f._set_params(1.0, seed, 0, 0)
        eng.init_hot()
        f._dev_valid, f._host_valid = True, False
        return f

    @staticmethod
    def from_field(field: "Field") -> "Field":
        return Field(field.lattice, field.values.copy(), field._device)

    @staticmethod
    def from_values(lattice: Lattice, values: Sequence[int], device: int = 0) -> "Field":
        return Field(lattice, np.asarray(values), device)

    # -- host/device reconciliation -----------------------------------------------------------
    def _need_engine(self) -> Engine:
        if self._engine is None:
            if self.lattice.d != 3:
                raise NotImplementedError("the device path holds 3-dimensional lattices (Lattice3d) only")
            x, y, t = self.lattice.size
            self._engine = Engine(x, y, t, n_chains=1, device=self._device)
        return self._engine

    def _set_params(self, beta: float, seed: int, w: int, h: int):
        p = (float(beta), int(seed), int(w), int(h))
        if self._params != p:
            self._need_engine().set_chain(0, *p)
            self._params = p

    def _to_device(self) -> Engine:
        eng = self._need_engine()
        if self._params is None:
            self._set_params(1.0, 0, 0, 0)
        if not self._dev_valid:
            eng.upload(0, self._host)
            self._dev_valid = True
        return eng

    def _device_changed(self):
        self._dev_valid, self._host_valid = True, False

    @property
    def values(self) -> np.ndarray:
        """Read-only view of `Field::values` (downloads if the device copy is newer)."""
        if not self._host_valid:
            self._host.flags.writeable = True
            self._engine.download(0, self._host)
            self._host_valid = True
        self._host.flags.writeable = False
        return self._host

    @values.setter
    def values(self, new_values):
        v = np.array(new_values, dtype=np.int32).reshape(-1)
        if v.size != self.lattice.n_sites:
            raise AssertionError("wrong number of values")
        self._host = v
        self._host_valid, self._dev_valid = True, False

    def set_value(self, index: int, value: int):
        _ = self.values
        self._host.flags.writeable = True
        self._host[index] = value
        self._dev_valid = False

    # -- Action trait (fields.rs:651-768): whole-lattice reductions on the device --------------
    def integer_action(self) -> int:
        return int(self._to_device().action()[0][0])

    def integer_energy(self) -> int:
        """fields.rs:728-737, exact in int64 (the reference accumulates in i32, quirk Q5)."""
        return int(self._to_device().energy()[0][0])

    def float_action(self) -> float:
        return float(self.integer_action())

    def float_energy(self) -> float:
        return float(self.integer_energy())

    def action_observable(self, temp: float) -> float:
        return temp * self.float_action()

    def energy_observable(self, temp: float) -> float:
        return temp * self.float_energy()

    # -- Action trait: single-site accessors (host arithmetic) ---------------------------------
    def get_value(self, index: int) -> int:
        return int(self.values[index])

    def get_value_from_coords(self, coords: LatticeCoords) -> int:
        return int(self.values[self.lattice.calc_index_from_coords(coords)])

    def assumed_bond_difference(self, index: int, direction: int, value: int) -> int:
        return int(value) - int(self.values[self.lattice.get_neighbours_array(index)[direction]])

    def bond_difference(self, index: int, direction: int) -> int:
        return self.assumed_bond_difference(index, direction, self.get_value(index))

    def assumed_bond_action(self, index: int, direction: int, value: int) -> int:
        d = self.assumed_bond_difference(index, direction, value)
        return d * d

    def bond_action(self, index: int, direction: int) -> int:
        d = self.bond_difference(index, direction)
        return d * d

    def assumed_local_action(self, index: int, value: int) -> int:
        v = self.values
        return int(sum((int(value) - int(v[n])) ** 2 for n in self.lattice.get_neighbours_array(index)))

    def local_action(self, index: int) -> int:
        return self.assumed_local_action(index, self.get_value(index))

    # -- HeightField trait (fields.rs:604-633) --------------------------------------------------
    def shift_values(self, shift: int):
        self._to_device().shift_values(0, int(shift))
        self._device_changed()

    def normalize_random(self, rng=None):
        """Subtract the height of one uniformly chosen site (fields.rs:621-625).  The reference
        uses a private ThreadRng; here the pick comes from `rng` (algorithm.SweepRng) or from
        stream LQFT_STREAM_NORMALIZE at counter 0 of seed 0."""
        eng = self._to_device()
        site = rng.gen_site(self.lattice.n_sites) if rng is not None else pick_site(
            0, _lib.STREAM_NORMALIZE, 0, 0, self.lattice.n_sites)
        eng.normalize_site(0, site)
        self._device_changed()


class WilsonField:
    """WilsonField<'a, i32, SIZE> (fields.rs:528-568, 770-846): a Field plus the sheet of +y bonds
    leaving (x, 0, t), x < width, t < height, whose differences are shifted by +1."""

    def __init__(self, field: Field, width: int, height: int):
        if field.lattice.d != 3:
            raise NotImplementedError("WilsonField is 3-dimensional (fields.rs:530)")
        self.field = field
        self.width, self.height = int(width), int(height)
        self.spin = True  # fields.rs:565

    @staticmethod
    def new(lattice: Lattice, width: int, height: int) -> "WilsonField":
        return WilsonField(Field.new(lattice), width, height)

    @staticmethod
    def from_field(field: Field, width: int, height: int) -> "WilsonField":
        return WilsonField(field, width, height)

    def is_active(self, index: int, direction: int) -> bool:
        """LinksField::is_active through BondsFieldNew::get_value (fields.rs:282-291, 396-398)."""
        lat = self.field.lattice
        if direction % 6 >= 3:
            index = lat.get_neighbours_array(index)[direction]
            direction -= 3
        if direction != 1:
            return False
        x, y, t = lat.calc_coords_from_index(index).into_array()
        return y == 0 and x < self.width and t < self.height

    def wilson_offset(self, index: int) -> int:
        x, y, t = self.field.lattice.calc_coords_from_index(index).into_array()
        return int(lib().lqft_host_wilson_offset(self.field.lattice.size[1], x, y, t, self.width, self.height))

    # -- Action trait, single-site accessors (fields.rs:805-845) ---------------------------------
    def get_value(self, index: int) -> int:
        return self.field.get_value(index)

    def get_value_from_coords(self, coords: LatticeCoords) -> int:
        return self.field.get_value_from_coords(coords)

    def assumed_bond_difference(self, index: int, direction: int, value: int) -> int:
        d = self.field.assumed_bond_difference(index, direction, value)
        if self.is_active(index, direction):
            modifier = {1: 1, 4: -1}.get(direction, 0)
            d += modifier * (1 if self.spin else -1)
        return d

    def bond_difference(self, index: int, direction: int) -> int:
        return self.assumed_bond_difference(index, direction, self.get_value(index))

    def assumed_bond_action(self, index: int, direction: int, value: int) -> int:
        d = self.assumed_bond_difference(index, direction, value)
        return d * d

    def bond_action(self, index: int, direction: int) -> int:
        d = self.bond_difference(index, direction)
        return d * d

    def assumed_local_action(self, index: int, value: int) -> int:
        return sum(self.assumed_bond_action(index, direction, value) for direction in range(6))

    def local_action(self, index: int) -> int:
        return self.assumed_local_action(index, self.get_value(index))

    def integer_energy(self) -> int:
        raise NotImplementedError(
            "Wilson-aware whole-lattice energy is only reached from WilsonTest (computation.rs:494), "
            "which has no public constructor; WilsonSim measures `wilson.field` (computation.rs:335)")

    # -- HeightField (fields.rs:635-645) -----------------------------------------------------------
    def normalize_random(self, rng=None):
        self.field.normalize_random(rng)

    def shift_values(self, shift: int):
        self.field.shift_values(shift)
