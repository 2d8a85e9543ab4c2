"""Sharding of one IQ stream across GPUs.

The hot path has no exchange step: every candidate start g0 is decided from the quantizer bits
q[g0 .. g0+PKL), and every q bit from the running sum of ONE reference block.  A rank that owns
the detection blocks [first_block, last_block) therefore only needs to be fed from
`first_fed_block = first_block - ceil(PKL/BS) - 1`: the extra leading block supplies real
lead-in samples (decode.go:165) to the first block whose bits matter, and the ceil(PKL/BS)
blocks supply the Quantized history (decode.go:166) of the first owned block.  Shards are cut
on block boundaries because the running sum restarts at every block (decode.go:232-236).
Results are concatenated on the host in block order; no collective touches the data path.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    first_block: int       # first owned detection block (global index)
    last_block: int        # one past the last owned block
    first_fed_block: int   # first block whose IQ bytes the rank must be fed (halo start)

    @property
    def halo_blocks(self) -> int:
        return self.first_block - self.first_fed_block

    def owns_block(self, block: int) -> bool:
        return self.first_block <= block < self.last_block

    def owns_start(self, g0: int, block_size: int, buffer_length: int) -> bool:
        """Does this rank report the candidate whose first preamble bit is stream bit g0?
        (block = floor((g0 + BUF) / BS) - 1, SURVEY.md section 2.1 'global view')."""
        return self.owns_block((g0 + buffer_length) // block_size - 1)

    def to_global(self, local_block: int) -> int:
        return local_block + self.first_fed_block

    def keep(self, local_block: int) -> bool:
        return self.owns_block(self.to_global(local_block))


def halo_blocks(block_size: int, packet_length: int) -> int:
    return -(-packet_length // block_size) + 1


def plan(total_blocks: int, nranks: int, block_size: int, packet_length: int) -> list[ShardPlan]:
    """Contiguous, nearly equal shards of [0, total_blocks)."""
    if nranks < 1 or total_blocks < 0:
        raise ValueError("bad shard request")
    halo = halo_blocks(block_size, packet_length)
    out, base, rem, lo = [], total_blocks // nranks, total_blocks % nranks, 0
    for r in range(nranks):
        hi = lo + base + (1 if r < rem else 0)
        out.append(ShardPlan(r, lo, hi, max(0, lo - halo)))
        lo = hi
    return out
