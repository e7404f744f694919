"""Bank-conflict-aware placement of the factor blocks in the per-instance shared-memory slice.

The access schedule of the factorisation and the triangular solves is static: which lane touches
which double in which instruction follows from the task lists the topology compiler emits.  The
shared-memory data pipe is the busiest unit of the sweep kernel (``profiles/smem_pipe.json``) and a
quarter of its wavefronts are bank-conflict replays, so the compiler can choose the slot of every 3x3
block to minimise them.

Model (matches ncu's per-line ``L1 Wavefronts Shared`` / ``Excessive`` for ``okin_factor`` to two
digits): a warp-wide 64-bit access is served one half-warp at a time; within a half-warp two lanes
conflict when they touch *different* doubles whose addresses are equal modulo 16 (32 four-byte
banks); the wavefronts of a half-warp are the largest number of distinct doubles in any one of the 16
bank pairs.
"""

from __future__ import annotations

import numpy as np


class AccessTrace:
    """Flat record of (instruction, lane, address) triples; an address is either absolute or a
    (block, inner offset) pair inside the factor storage whose slot is still free to choose."""

    def __init__(self, lb_base: int):
        self.lb_base = lb_base
        self.inst, self.lane, self.blk, self.off = [], [], [], []
        self.n_inst = 0

    def access(self, lanes, refs, offsets=(0,)) -> None:
        """One instruction per entry of ``offsets``: lane ``lanes[i]`` reads ``refs[i] + offset``.
        ``refs[i]`` = ``("LB", offset_in_factor_storage)`` or ``("ABS", absolute_offset)``."""
        for extra in offsets:
            for lane, (kind, off) in zip(lanes, refs):
                self.inst.append(self.n_inst)
                self.lane.append(lane)
                if kind == "LB":
                    self.blk.append(off // 9)           # rows and blocks never straddle a block boundary
                    self.off.append(off % 9 + extra)
                else:
                    self.blk.append(-1)
                    self.off.append(off + extra)
            self.n_inst += 1

    def freeze(self) -> None:
        self.inst = np.asarray(self.inst, dtype=np.int64)
        self.blk = np.asarray(self.blk, dtype=np.int64)
        self.off = np.asarray(self.off, dtype=np.int64)
        half = np.asarray(self.lane, dtype=np.int64) // 16
        self.key = self.inst * 2 + half
        self.n_keys = 2 * self.n_inst
        self.in_lb = self.blk >= 0

    def wavefronts(self, slot: np.ndarray) -> int:
        addr = np.where(self.in_lb, self.lb_base + 9 * slot[np.maximum(self.blk, 0)] + self.off, self.off)
        pairs = np.unique(self.key * (1 << 20) + addr)          # distinct doubles per (instruction, half)
        key, bank = pairs >> 20, (pairs & ((1 << 20) - 1)) % 16
        counts = np.bincount(key * 16 + bank, minlength=16 * self.n_keys).reshape(self.n_keys, 16)
        return int(counts.max(axis=1).sum())

    def ideal(self) -> int:
        return int(np.unique(self.key).size)


def tune_block_slots(trace: AccessTrace, n_blocks: int, iterations: int = 1500, seed: int = 0) -> tuple:
    """Hill climbing over pairwise slot swaps.  Returns ``(slot_of_block, wavefronts_before,
    wavefronts_after, ideal)``; deterministic for a given trace."""
    trace.freeze()
    rng = np.random.default_rng(seed)
    slot = np.arange(n_blocks, dtype=np.int64)
    before = best = trace.wavefronts(slot)
    for _ in range(iterations):
        i, j = rng.choice(n_blocks, 2, replace=False)
        slot[i], slot[j] = slot[j], slot[i]
        cost = trace.wavefronts(slot)
        if cost <= best:
            best = cost
        else:
            slot[i], slot[j] = slot[j], slot[i]
    return slot, before, best, trace.ideal()


class UnitTrace:
    """Like ``AccessTrace`` for storage made of variable-size units (the row-gradient vectors of one
    constraint row): a record is (instruction, lane, unit, offset inside the unit) and the address of
    a unit follows from the order the units are laid out in."""

    def __init__(self, base: int, sizes):
        self.base, self.sizes = base, np.asarray(sizes, dtype=np.int64)
        self.inst, self.lane, self.unit, self.off = [], [], [], []
        self.n_inst = 0

    def access(self, lanes, refs, offsets=(0,)) -> None:
        """``refs[i] = (unit, offset inside the unit)``; one instruction per entry of ``offsets``."""
        for extra in offsets:
            for lane, (unit, off) in zip(lanes, refs):
                self.inst.append(self.n_inst)
                self.lane.append(lane)
                self.unit.append(unit)
                self.off.append(off + extra)
            self.n_inst += 1

    def freeze(self) -> None:
        self.inst = np.asarray(self.inst, dtype=np.int64)
        self.unit = np.asarray(self.unit, dtype=np.int64)
        self.off = np.asarray(self.off, dtype=np.int64)
        self.key = self.inst * 2 + np.asarray(self.lane, dtype=np.int64) // 16
        self.n_keys = 2 * self.n_inst

    def offsets(self, order: np.ndarray) -> np.ndarray:
        """Start offset of every unit when laid out in ``order``."""
        start = np.zeros(len(order), dtype=np.int64)
        start[order] = np.concatenate([[0], np.cumsum(self.sizes[order])[:-1]])
        return start

    def wavefronts(self, order: np.ndarray) -> int:
        addr = self.base + self.offsets(order)[self.unit] + self.off
        pairs = np.unique(self.key * (1 << 20) + addr)
        key, bank = pairs >> 20, (pairs & ((1 << 20) - 1)) % 16
        counts = np.bincount(key * 16 + bank, minlength=16 * self.n_keys).reshape(self.n_keys, 16)
        return int(counts.max(axis=1).sum())

    def ideal(self) -> int:
        return int(np.unique(self.key).size)


def tune_unit_order(trace: UnitTrace, iterations: int = 1500, seed: int = 1) -> tuple:
    """Hill climbing over pairwise swaps in the layout order of the units.  Returns ``(order,
    wavefronts_before, wavefronts_after, ideal)``."""
    trace.freeze()
    rng = np.random.default_rng(seed)
    n = len(trace.sizes)
    order = np.arange(n, dtype=np.int64)
    before = best = trace.wavefronts(order)
    for _ in range(iterations):
        i, j = rng.choice(n, 2, replace=False)
        order[i], order[j] = order[j], order[i]
        cost = trace.wavefronts(order)
        if cost <= best:
            best = cost
        else:
            order[i], order[j] = order[j], order[i]
    return order, before, best, trace.ideal()
