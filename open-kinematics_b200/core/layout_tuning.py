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
