"""Flat, tile-aligned key arenas (the B200 replacement of per-key pinned ``comm_buf_`` staging buffers).

Parity: key bookkeeping of ``KVStoreDist`` — ``EncodeDefaultKey`` (``src/kvstore/kvstore_dist.h:721-761``): small keys go to one
server ``(key*9973) % num_servers``, arrays ≥ ``MXNET_KVSTORE_BIGARRAY_BOUND`` are partitioned across all servers; MultiGPS global
sharding (``kvstore_dist_server.h:1770-1810``).  Here every key owns a contiguous, 1024-float-aligned range of ONE flat arena so that
a whole step's push/pull is a single kernel launch; tiles (1024 floats) are the unit of ownership:

* local tier  : tile ``t`` is reduced by party member ``t % party_size`` (round-robin shards inside the party);
* global tier : tile ``t`` of a *big* key (≥ bigarray bound) is owned by global server ``t % num_gs`` (partitioned), every tile of a
  *small* key by global server ``(key*9973) % num_gs`` — the reference's two rules, applied per tile.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

TILE = 1024
BIGARRAY_BOUND_DEFAULT = 1000000


@dataclass
class KeySlot:
    key: int
    shape: tuple
    numel: int
    offset: int      # floats from arena start (multiple of TILE)
    tiles: int
    priority: int = 0
    lr_mult: float = 1.0
    wd_mult: float = 1.0


@dataclass
class ArenaLayout:
    slots: list = field(default_factory=list)
    total: int = 0

    @staticmethod
    def build(keys_shapes, mults=None):
        lay = ArenaLayout()
        off = 0
        for i, (k, shape) in enumerate(keys_shapes):
            n = int(np.prod(shape)) if len(shape) else 1
            tiles = max(1, (n + TILE - 1) // TILE)
            lm, wm = (1.0, 1.0) if mults is None else mults[i]
            lay.slots.append(KeySlot(k, tuple(shape), n, off, tiles, 0, lm, wm))
            off += tiles * TILE
        lay.total = off
        return lay

    @property
    def num_tiles(self):
        return self.total // TILE

    def tile_key(self):
        out = np.empty(self.num_tiles, dtype=np.int32)
        for i, s in enumerate(self.slots):
            out[s.offset // TILE: s.offset // TILE + s.tiles] = i
        return out

    def key_tiles(self):
        return np.array([s.tiles for s in self.slots], dtype=np.int32)

    def tile_mult(self):
        out = np.ones((self.num_tiles, 2), dtype=np.float32)
        for s in self.slots:
            out[s.offset // TILE: s.offset // TILE + s.tiles] = (s.lr_mult, s.wd_mult)
        return out

    def global_owner_index(self, num_gs, bigarray_bound=BIGARRAY_BOUND_DEFAULT):
        """Per-tile index into the global-server list (reference rules, see module docstring)."""
        out = np.zeros(self.num_tiles, dtype=np.int32)
        for s in self.slots:
            t0 = s.offset // TILE
            if s.numel >= bigarray_bound and num_gs > 1:
                out[t0:t0 + s.tiles] = (np.arange(s.tiles) % num_gs)
            else:
                out[t0:t0 + s.tiles] = (int(s.key) * 9973) % num_gs
        return out

    def view(self, arena, i):
        s = self.slots[i]
        return arena[s.offset:s.offset + s.numel].view(s.shape if len(s.shape) else (1,))
