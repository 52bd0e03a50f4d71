"""Tile-index-range sharding.  Tiles are independent, so rank g of G takes the
contiguous range [g*N/G, (g+1)*N/G) and writes its fixed-size records straight into the same slice
of the output; order is restored by construction and there is no collective on the data path."""
from __future__ import annotations

from typing import Tuple


def tile_range(rank: int, world: int, n: int) -> Tuple[int, int]:
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError("tile_range: need 0 <= rank < world and n >= 0")
    return (n * rank) // world, (n * (rank + 1)) // world
