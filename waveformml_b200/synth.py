"""Synthetic PROSPECT-style events (SURVEY.md 8d / BASELINE.md section 2).

Shapes and dtypes follow what the reference's datasets hand to collate_fn:
coords int32 [N,3] = (x, y, event) and int16 waveform samples, two PMTs x n_samples per hit
(src/datasets/H5CompoundTypes.py:105-120, src/datasets/HDF5Dataset.py:282-302); the grid is
14 x 11 segments (src/models/SPConvNet.py:51).  There is no HDF5 in the timed loop.
"""
import numpy as np

GRID_X, GRID_Y = 14, 11
N_ADC_BITS = 14
MAX_RANGE_INV = 1.0 / (2 ** N_ADC_BITS - 1)  # src/datasets/HDF5Dataset.py:15-17

_STEPS = np.array([(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)], dtype=np.int64)


def make_events(n_events, n_samples=150, seed=1234, full_grid=False, n_classes=3, event_offset=0):
    """Returns dict(coords int32 [N,3] (x,y,evt), wave int16 [N,2*n_samples], labels int64 [B],
    z float32 [N]).  Rows are sorted by (event, x, y).

    Sparse mode: hits/event m = clip(1 + Poisson(2), 1, 10); first hit uniform on the grid, the
    rest grown by random 8-neighbour steps from an existing hit (a track-like cluster).
    full_grid: every one of the 154 cells is hit (the high-occupancy sweep, config C5).
    """
    rng = np.random.default_rng(seed)
    rows = []
    if full_grid:
        xs, ys = np.meshgrid(np.arange(GRID_X), np.arange(GRID_Y), indexing="ij")
        cell = np.stack([xs.ravel(), ys.ravel()], axis=1)
        for e in range(n_events):
            ev = np.full((cell.shape[0], 1), e, dtype=np.int64)
            rows.append(np.concatenate([cell, ev], axis=1))
    else:
        mult = np.clip(1 + rng.poisson(2.0, size=n_events), 1, 10)
        for e in range(n_events):
            hits = {(int(rng.integers(GRID_X)), int(rng.integers(GRID_Y)))}
            order = list(hits)
            tries = 0
            while len(order) < mult[e] and tries < 200:
                tries += 1
                bx, by = order[int(rng.integers(len(order)))]
                dx, dy = _STEPS[int(rng.integers(8))]
                c = (bx + int(dx), by + int(dy))
                if 0 <= c[0] < GRID_X and 0 <= c[1] < GRID_Y and c not in hits:
                    hits.add(c)
                    order.append(c)
            cells = np.array(sorted(hits), dtype=np.int64)
            ev = np.full((cells.shape[0], 1), e, dtype=np.int64)
            rows.append(np.concatenate([cells, ev], axis=1))
    coords = np.concatenate(rows, axis=0).astype(np.int32)
    coords[:, 2] += event_offset
    n = coords.shape[0]
    wave = rng.integers(0, 2 ** N_ADC_BITS, size=(n, 2 * n_samples), dtype=np.int16)
    labels = rng.integers(0, n_classes, size=n_events).astype(np.int64)
    z = rng.random(n, dtype=np.float32)
    return {"coords": coords, "wave": wave, "labels": labels, "z": z}


def make_events_3d(n_events, n_samples=16, seed=1234, n_classes=3, occupancy=0.3):
    """Voxel form of the same events for net_type "3DConvolution" (src/models/SPConvNet.py:42-49: spatial size
    [14, 11, n_samples], coordinate columns permuted [3, 0, 1, 2]): every hit of make_events() contributes the
    time samples where its waveform is "above threshold" (a contiguous pulse of random start and length plus
    random isolated samples at rate `occupancy`/4), one row per (x, y, t) with the two PMT values as features.
    Returns dict(coords int32 [N,4] = (x, y, t, evt) sorted by (evt, x, y, t), wave int16 [N,2], labels int64 [B])."""
    base = make_events(n_events, n_samples=1, seed=seed, n_classes=n_classes)
    rng = np.random.default_rng(seed + 7)
    rows = []
    for x, y, e in base["coords"]:
        start = int(rng.integers(0, n_samples))
        length = 1 + int(rng.integers(0, max(1, int(occupancy * n_samples))))
        on = np.zeros(n_samples, dtype=bool)
        on[start:start + length] = True
        on |= rng.random(n_samples) < occupancy / 4
        for t in np.nonzero(on)[0]:
            rows.append((x, y, t, e))
    coords = np.array(rows, dtype=np.int32).reshape(-1, 4)
    wave = rng.integers(0, 2 ** N_ADC_BITS, size=(coords.shape[0], 2), dtype=np.int16)
    return {"coords": coords, "wave": wave, "labels": base["labels"]}
