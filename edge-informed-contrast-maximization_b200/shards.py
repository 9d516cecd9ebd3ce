"""Window shards: the on-disk / wire format of STAGED windows for multi-GPU sweeps.

One shard file holds consecutive windows of one sequence in exactly the form ``loss_func`` / ``Plan.set_window`` take them: rectified
int16 pixel coordinates, float64 times normalised to the evaluation interval, the float64 edge maps and their reference times - what
the reference produces per sample with ``get_sample`` (src/dataloaders/dsec_loader.py:285-349: rectification, fixed-N window) followed
by ``stage_datasample`` (src/experiments/e00/exp_mgr.py:283-327: time normalisation, edge extraction) and then uploads as jnp arrays.
Writing the shard once (``eincm_b200.dataloaders`` does those steps on the device) takes h5py, OpenCV and the CPU out of the sweep:
a rank maps the file, and a window is three contiguous byte ranges that go to the device as they are.

Layout (little-endian; every section starts on a 64-byte boundary so that it can be read straight into pinned memory)::

    header   64 B   magic 'EINCMSH1' | u32 version | u32 H | u32 W | u32 tile | u32 n_windows | u32 r_max | u32 flags | u32 0
                    | u64 index_offset | u64 file_bytes
    payload  per window:  xs i16[n] | ys i16[n] | ts f64[n] | tile_counts u32[tiles_y * tiles_x] (FLAG_TILE_MAJOR)
                          | edges f64[R][H][W] (FLAG_EDGES)
    index    per window:  u64 payload_offset | u64 payload_bytes | u64 n_events | i64 t_start_us | i64 t_end_us
                          | i64 n_event_deficiency | u32 R | u32 crc32(payload) | f64 edge_ts[r_max]

The index is written last (a writer appends windows without knowing their number) and the header is patched on ``close``: a shard whose
header has ``index_offset == 0`` was not closed and is refused.

FLAG_TILE_MAJOR ("pre-tiled"): the events of a window are stored in 16 x 16 SOURCE-TILE-major order (row-major over the tiles, time order
kept inside a tile) with the per-tile counts next to them.  The objective does not depend on the order of the events - the image of warped
events is a sum of exactly rounded fixed-point votes, bit-identical under any permutation (tests/test_gpu_shards.py) - and this is the
order the library's staging sorts them into anyway (csrc/k_prep.cuh), so its scatter writes become sequential.  It also makes the EVENT
SPLIT a byte-range read: ``event_range_for_rank`` cuts the window at tile boundaries from the prefix sums of ``tile_counts``, so a rank
of ``EventSplitObjective`` reads one contiguous range per array and its events cover a compact band of the sensor.

This module is file plumbing (numpy, no CUDA): the compute stays behind ``Plan``; ``ShardReader.stage`` is the one call that touches it.
"""
import os
import struct
import zlib
from typing import NamedTuple, Optional, Sequence, Tuple

import numpy as np

__all__ = ['ShardWriter', 'ShardReader', 'ShardWindow', 'ShardError', 'tile_major_order', 'MAGIC', 'VERSION', 'TILE',
           'FLAG_EDGES', 'FLAG_TILE_MAJOR']

MAGIC = b'EINCMSH1'
VERSION = 1
TILE = 16                  # kTile of csrc/k_prep.cuh
ALIGN = 64
FLAG_EDGES = 1
FLAG_TILE_MAJOR = 2

_HEADER = struct.Struct('<8sIIIIIIIIQQ')                # 56 bytes, padded to 64
_HEADER_BYTES = 64
_INDEX_FIXED = struct.Struct('<QQQqqqII')               # 56 bytes, then r_max doubles


class ShardError(ValueError):
    """A malformed, truncated, unclosed or corrupted shard, or a window that cannot be stored."""


def _pad(n: int) -> int:
    return (-n) % ALIGN


def tile_major_order(xs, ys, sensor_size: Tuple[int, int], tile: int = TILE):
    """``(order, tile_counts)``: the stable permutation that puts the events into source-tile-major order (tiles row-major, the original -
    time - order inside a tile) and the number of events of every tile, ``[tiles_y * tiles_x]`` uint32."""
    H, W = int(sensor_size[0]), int(sensor_size[1])
    tiles_y, tiles_x = -(-H // tile), -(-W // tile)
    xs, ys = np.asarray(xs), np.asarray(ys)
    key = (ys.astype(np.int64) // tile) * tiles_x + (xs.astype(np.int64) // tile)
    order = np.argsort(key, kind='stable')
    counts = np.bincount(key, minlength=tiles_y * tiles_x).astype(np.uint32)
    return order, counts


class ShardWindow(NamedTuple):
    xs: np.ndarray                       # int16 [n]
    ys: np.ndarray                       # int16 [n]
    ts: np.ndarray                       # float64 [n]
    edges: Optional[np.ndarray]          # float64 [R][H][W] or None (shard written without FLAG_EDGES)
    edge_ts: np.ndarray                  # float64 [R]
    tile_counts: Optional[np.ndarray]    # uint32 [tiles_y * tiles_x] or None (not FLAG_TILE_MAJOR)
    t_start_us: int
    t_end_us: int
    n_event_deficiency: int

    def args(self):
        """Positional operands of ``loss_func`` after theta (reference src/eincm/losses.py:108-114)."""
        return self.xs, self.ys, self.ts, self.edges, self.edge_ts


class ShardWriter:
    """Appends staged windows to ``path``.  Use as a context manager, or call ``close`` (the index and the final header are written there)."""

    def __init__(self, path: str, sensor_size: Tuple[int, int], r_max: int = 8, store_edges: bool = True, tile_major: bool = True):
        self.H, self.W = int(sensor_size[0]), int(sensor_size[1])
        if not (0 < self.H < 32768 and 0 < self.W < 32768):
            raise ShardError(f'sensor size {sensor_size} does not fit int16 pixel coordinates')
        if not 1 <= int(r_max) <= 8:
            raise ShardError('r_max must be 1..8 (EINCM_MAX_REFS)')
        self.r_max = int(r_max)
        self.flags = (FLAG_EDGES if store_edges else 0) | (FLAG_TILE_MAJOR if tile_major else 0)
        self.path = path
        self._f = open(path, 'wb')
        self._index = []
        self._f.write(self._header(0, 0, 0))
        self._pos = _HEADER_BYTES

    def _header(self, n_windows: int, index_offset: int, file_bytes: int) -> bytes:
        h = _HEADER.pack(MAGIC, VERSION, self.H, self.W, TILE, n_windows, self.r_max, self.flags, 0, index_offset, file_bytes)
        return h + b'\0' * (_HEADER_BYTES - len(h))

    def _section(self, a: np.ndarray, crc: int) -> int:
        b = np.ascontiguousarray(a).tobytes()
        b += b'\0' * _pad(len(b))
        self._f.write(b)
        self._pos += len(b)
        return zlib.crc32(b, crc)

    def add_window(self, xs, ys, ts, edges=None, edge_ts: Sequence[float] = (), t_start_us: int = 0, t_end_us: int = 0,
                   n_event_deficiency: int = 0) -> int:
        """Stores one window; returns its index in the shard.  ``xs, ys`` must lie inside the sensor (rectified, cropped events)."""
        if self._f is None:
            raise ShardError('the shard is closed')
        xs, ys = np.asarray(xs), np.asarray(ys)
        ts = np.asarray(ts, dtype=np.float64)
        n = int(xs.shape[0])
        if xs.ndim != 1 or ys.shape != xs.shape or ts.shape != xs.shape:
            raise ShardError('xs, ys, ts must be one-dimensional and of the same length')
        if n and (int(xs.min()) < 0 or int(xs.max()) >= self.W or int(ys.min()) < 0 or int(ys.max()) >= self.H):
            raise ShardError('event coordinates outside the sensor: rectify / crop before writing the shard')
        edge_ts = np.asarray(edge_ts, dtype=np.float64).reshape(-1)
        R = int(edge_ts.shape[0])
        if not 1 <= R <= self.r_max:
            raise ShardError(f'a window needs 1..{self.r_max} reference times, got {R}')
        if self.flags & FLAG_EDGES:
            if edges is None:
                raise ShardError('this shard stores edge maps: edges is required')
            edges = np.asarray(edges, dtype=np.float64)
            if edges.shape != (R, self.H, self.W):
                raise ShardError(f'edges must have shape {(R, self.H, self.W)}, got {edges.shape}')
        xs, ys = xs.astype(np.int16), ys.astype(np.int16)
        counts = None
        if self.flags & FLAG_TILE_MAJOR:
            order, counts = tile_major_order(xs, ys, (self.H, self.W))
            xs, ys, ts = xs[order], ys[order], ts[order]
        start, crc = self._pos, 0
        crc = self._section(xs, crc)
        crc = self._section(ys, crc)
        crc = self._section(ts, crc)
        if counts is not None:
            crc = self._section(counts, crc)
        if self.flags & FLAG_EDGES:
            crc = self._section(edges, crc)
        ets = np.zeros(self.r_max, np.float64)
        ets[:R] = edge_ts
        self._index.append(_INDEX_FIXED.pack(start, self._pos - start, n, int(t_start_us), int(t_end_us), int(n_event_deficiency), R,
                                             crc & 0xffffffff) + ets.tobytes())
        return len(self._index) - 1

    def close(self):
        if self._f is None:
            return
        index_offset = self._pos
        for rec in self._index:
            self._f.write(rec)
            self._pos += len(rec)
        self._f.seek(0)
        self._f.write(self._header(len(self._index), index_offset, self._pos))
        self._f.close()
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        else:                                   # leave the header unpatched: the reader refuses the file
            self._f.close()
            self._f = None
        return False


class ShardReader:
    """Maps a shard; ``window(i)`` returns views into the mapping (no copy).  ``verify=True`` checks every payload CRC on open."""

    def __init__(self, path: str, verify: bool = False):
        self.path = path
        size = os.path.getsize(path)
        if size < _HEADER_BYTES:
            raise ShardError(f'{path}: too short for a shard header')
        self._m = np.memmap(path, dtype=np.uint8, mode='r')
        magic, version, H, W, tile, n_windows, r_max, flags, _, index_offset, file_bytes = _HEADER.unpack_from(self._m[:_HEADER.size].tobytes())
        if magic != MAGIC:
            raise ShardError(f'{path}: not a window shard (magic {magic!r})')
        if version != VERSION or tile != TILE:
            raise ShardError(f'{path}: version {version} / tile {tile} not supported (this reader: {VERSION} / {TILE})')
        if index_offset == 0:
            raise ShardError(f'{path}: the shard was not closed by its writer')
        self.sensor_size = (int(H), int(W))
        self.n_windows, self.r_max, self.flags = int(n_windows), int(r_max), int(flags)
        self._rec = _INDEX_FIXED.size + 8 * self.r_max
        if file_bytes != size or index_offset + self._rec * self.n_windows != size or not 1 <= self.r_max <= 8:
            raise ShardError(f'{path}: truncated or inconsistent (header says {file_bytes} bytes, file has {size})')
        self._index = []
        for i in range(self.n_windows):
            o = int(index_offset) + i * self._rec
            fixed = _INDEX_FIXED.unpack_from(self._m[o:o + _INDEX_FIXED.size].tobytes())
            ets = np.frombuffer(self._m[o + _INDEX_FIXED.size:o + self._rec].tobytes(), dtype=np.float64)
            off, nbytes, n, R = fixed[0], fixed[1], fixed[2], fixed[6]
            if off % ALIGN or off < _HEADER_BYTES or off + nbytes > index_offset or not 1 <= R <= self.r_max or nbytes != self._payload_bytes(n, R):
                raise ShardError(f'{path}: index record {i} is inconsistent')
            self._index.append(fixed + (ets[:R].copy(),))
        if verify:
            for i in range(self.n_windows):
                self.verify(i)

    @property
    def tiles(self) -> Tuple[int, int]:
        return -(-self.sensor_size[0] // TILE), -(-self.sensor_size[1] // TILE)

    @property
    def tile_major(self) -> bool:
        return bool(self.flags & FLAG_TILE_MAJOR)

    def __len__(self):
        return self.n_windows

    def _payload_bytes(self, n: int, R: int) -> int:
        H, W = self.sensor_size
        b = 2 * (2 * n + _pad(2 * n)) + 8 * n + _pad(8 * n)
        if self.flags & FLAG_TILE_MAJOR:
            nt = self.tiles[0] * self.tiles[1]
            b += 4 * nt + _pad(4 * nt)
        if self.flags & FLAG_EDGES:
            b += 8 * R * H * W + _pad(8 * R * H * W)
        return b

    def n_events(self, i: int) -> int:
        return int(self._index[i][2])

    def payload_range(self, i: int) -> Tuple[int, int]:
        """``(offset, bytes)`` of window ``i`` in the file: what a rank that does not map the file has to read."""
        return int(self._index[i][0]), int(self._index[i][1])

    def verify(self, i: int):
        off, nbytes = self.payload_range(i)
        if zlib.crc32(self._m[off:off + nbytes].tobytes()) & 0xffffffff != self._index[i][7]:
            raise ShardError(f'{self.path}: payload of window {i} does not match its checksum')

    def window(self, i: int) -> ShardWindow:
        if not 0 <= i < self.n_windows:
            raise IndexError(f'window {i} of a shard with {self.n_windows}')
        off, _, n, t0, t1, deficiency, R, _, ets = self._index[i]
        H, W = self.sensor_size
        n = int(n)

        def take(dtype, count):
            nonlocal off
            nbytes = np.dtype(dtype).itemsize * count
            a = self._m[off:off + nbytes].view(dtype)
            off += nbytes + _pad(nbytes)
            return a

        xs, ys, ts = take(np.int16, n), take(np.int16, n), take(np.float64, n)
        counts = take(np.uint32, self.tiles[0] * self.tiles[1]) if self.flags & FLAG_TILE_MAJOR else None
        edges = take(np.float64, R * H * W).reshape(R, H, W) if self.flags & FLAG_EDGES else None
        return ShardWindow(xs, ys, ts, edges, ets, counts, int(t0), int(t1), int(deficiency))

    # -- multi-GPU sweeps ---------------------------------------------------------------------------------------------------------------
    def windows_for_rank(self, rank: int, world: int) -> range:
        """Window sharding WITH the handover (reference src/eincm/solver.py:302-347 chains consecutive windows): a contiguous block per rank."""
        return range((self.n_windows * rank) // world, (self.n_windows * (rank + 1)) // world)

    def event_range_for_rank(self, i: int, rank: int, world: int) -> Tuple[int, int]:
        """Event split of window ``i``: ``[a, b)`` of its event arrays for ``rank``.  Pre-tiled shards cut at tile boundaries (balanced on the
        prefix sums of the tile counts: whole source tiles per rank); others fall back to equal contiguous slabs (``parallel.split_events``)."""
        n = self.n_events(i)
        if not self.flags & FLAG_TILE_MAJOR:
            return (n * rank) // world, (n * (rank + 1)) // world
        ends = np.cumsum(self.window(i).tile_counts, dtype=np.int64)

        def cut(r):
            if r <= 0:
                return 0
            if r >= world:
                return n
            target = (n * r) // world
            k = int(np.searchsorted(ends, target, side='left'))        # first tile boundary at or behind the target
            return int(ends[min(k, len(ends) - 1)])

        return cut(rank), cut(rank + 1)

    def stage(self, plan, i: int, event_range: Optional[Tuple[int, int]] = None, stream=None) -> ShardWindow:
        """``plan.set_window`` with window ``i`` (or the ``event_range`` of it a rank of the event split owns)."""
        if not self.flags & FLAG_EDGES:
            raise ShardError('this shard stores no edge maps: stage the events and pass the edges to plan.set_window yourself')
        w = self.window(i)
        a, b = (0, len(w.xs)) if event_range is None else event_range
        plan.set_window(w.xs[a:b], w.ys[a:b], w.ts[a:b], w.edges, w.edge_ts, stream=stream)
        return w
