"""Event reader / writer over the reference's pulse files, on top of io/h5lite.py.

  WAVEFORM_PAIR_CAL      the on-disk record of `*WaveformPairSim.h5` / `*WaveformPairCal.h5`
                         (src/datasets/H5CompoundTypes.py:105-120: evt i8, t f8, dt/z/E/PSD f4, PE f4[2], coord i4[3],
                         waveform i2[130], EZ f4[2], PID i4; 324 bytes); `pair_record_dtype(ns)` for other lengths
  H5Input                src/datasets/HDF5IO.py:25-79: setup_table / next_chunk(nrows, preserve_event) -- chunks of
                         rows that never split an event
  H5Output               src/datasets/HDF5IO.py:82-112: create_table / add_rows / close (gzip-9, chunks of 1024 rows),
                         the writer side of src/datasets/PredictionWriter.py:73-104
  PulseFiles             src/datasets/HDF5Dataset.py:136-217, 225-347, 377-403: directories of files -> items with an
                         event range each; item i -> ([coords (x, y, event) int32, waveform], labels) with the
                         reference's slicing by event id, label-from-directory rule and dtype contract.  The waveform
                         stays int16 (the 1/16383 normalisation is done by the GPU batcher, wfsp_batch_pack);
                         `normalized()` gives the reference's float view for comparison.
  feed(step, files, ...) pinned double-buffered host batches -> GraphTrainStep.prefetch: the H2D copy of batch i+1
                         overlaps the replay of batch i.

No h5py: h5lite reads the subset of HDF5 these files use.  Semantics are pinned by running the reference's OWN
H5Input.next_chunk and HDF5Dataset.__getitem__ over h5lite-backed files (tests/golden/make_h5_fixture.py); the byte
format itself is pinned only as far as h5lite is (see its docstring)."""
import fnmatch
import os
import re

import numpy as np

from . import h5lite

N_ADC_BITS = 14
MAX_RANGE_INV = 1.0 / (2 ** N_ADC_BITS - 1)  # src/datasets/HDF5Dataset.py:14-17


def pair_record_dtype(n_samples=65, with_z=False):
    """WaveformPairCal record (H5CompoundTypes.py:105-120) for 2*n_samples waveform samples per hit."""
    fields = [("evt", "<i8"), ("t", "<f8"), ("dt", "<f4"), ("z", "<f4"), ("E", "<f4"), ("PSD", "<f4"), ("PE", "<f4", (2,)),
              ("coord", "<i4", (3,)), ("waveform", "<i2", (2 * n_samples,)), ("EZ", "<f4", (2,)), ("PID", "<i4")]
    return np.dtype(fields)


WAVEFORM_PAIR_CAL = pair_record_dtype(65)
assert WAVEFORM_PAIR_CAL.itemsize == 324

_SORT_RE = re.compile(r"_(\d+)")


def _sort_key(path):
    """HDF5Dataset._sort_pattern (HDF5Dataset.py:20-25): files ordered by the first `_<number>` in their path."""
    nums = _SORT_RE.findall(str(path))
    return (0, int(nums[0]), "") if nums else (1, 0, str(path))


class H5Input:
    """Sequential reader of one table, `next_chunk` keeps events whole (HDF5IO.py:55-79)."""

    def __init__(self, path):
        self.path = path
        self.h5f = h5lite.File(path)
        self.table, self.table_name, self.table_length = None, "", 0
        self.event_index_name, self.event_index_coord = "", None
        self.current_index = -1

    def close(self):
        self.h5f.close()

    def setup_table(self, name, data_type=None, event_index_name="coord", event_index_coord=None, base="/"):
        self.table = self.h5f[base + name]
        if data_type is not None and np.dtype(data_type) != self.table.dtype:
            raise h5lite.H5Error("table %s holds %r, expected %r" % (name, self.table.dtype, np.dtype(data_type)))
        self.record_type, self.record_length = self.table.dtype, self.table.dtype.itemsize
        self.table_name, self.table_length = name, len(self.table)
        self.event_index_name, self.event_index_coord = event_index_name, event_index_coord
        self.current_index = -1

    def get_event_number(self, row):
        v = row[self.event_index_name]
        return v if self.event_index_coord is None else v[self.event_index_coord]

    def next_chunk(self, nrows=2048, preserve_event=True):
        """Same contract as the reference: rows [i, i + nrows) extended to the end of the last event; the final chunk
        is whatever remains (returned when i + nrows >= length); then None once, then the reader rewinds.  The
        reference extends one row at a time through h5py; here the event ids of a look-ahead window are read at once."""
        if self.table is None:
            raise RuntimeError("No table opened!")
        if self.current_index == -2:
            self.current_index = -1
            return None
        if self.current_index == -1:
            self.current_index = 0
        ci, n = self.current_index, self.table_length
        if ci + nrows >= n:
            self.current_index = -2
            return self.table[ci:n]
        end = ci + nrows
        if preserve_event:
            last = self.get_event_number(self.table[end - 1])
            while True:
                look = self.table[end:min(n, end + max(64, nrows // 8))]
                ids = look[self.event_index_name]
                if self.event_index_coord is not None:
                    ids = ids[:, self.event_index_coord]
                same = np.nonzero(ids != last)[0]
                if same.size:
                    end += int(same[0])
                    break
                end += len(look)
                if end >= n:
                    break
        data = self.table[ci:end]
        self.current_index = -2 if (preserve_event and end >= n) else end
        return data


class H5Output:
    """create_table / add_rows / close_table with the reference's defaults (gzip 9, chunks of 1024 rows).  Rows are
    buffered and the file is written at close(): the flat h5lite writer has no resizing."""

    def __init__(self, path):
        self.path, self.tables, self.table_index, self._attrs, self._opts = path, {}, {}, {}, {}

    def create_table(self, name, shape, data_type, compression="gzip", maxshape=(None,), compression_opts=9,
                     chunks=(1024,), **kwargs):
        self.tables[name] = np.zeros(shape, dtype=data_type)
        self.table_index[name] = 0
        self._attrs[name] = {}
        self._opts[name] = (int(chunks[0]) if chunks else None, compression_opts if compression == "gzip" else None)

    def add_rows(self, name, rows):
        i = self.table_index[name]
        self.tables[name][i:i + rows.shape[0]] = rows
        self.table_index[name] = i + rows.shape[0]

    def set_attr(self, name, key, value):
        self._attrs[name][key] = np.asarray(value)

    def flush(self, table=None):
        pass

    def close(self):
        w = h5lite.Writer(self.path)
        for name, arr in self.tables.items():
            chunks, gz = self._opts[name]
            if arr.shape[0] == 0:
                chunks, gz = None, None
            w.create_dataset(name, arr, chunks=chunks, gzip=gz if chunks else None, attrs=self._attrs[name])
        w.close()


class PulseFiles:
    """Directories of pulse files -> items; item i is one file's event range.

    Mirrors HDF5Dataset (non-group mode, event_based): one directory per class (label = directory index when
    label_name is None, HDF5Dataset.py:311-316), files sorted by `_sort_key`, interleaved across directories so every
    class reaches `events_per_dir` together (:160-176), the last file of a directory truncated to the remaining
    events (:386-389), `nevents` attribute = events in a file (:377-383)."""

    def __init__(self, dirs, file_pattern="*WaveformPairSim.h5", data_name="WaveformPairs", coord_name="coord",
                 feat_name="waveform", events_per_dir=1 << 62, label_name=None, label_map=None, cache_size=1):
        self.dirs = [os.path.normpath(os.path.abspath(d)) for d in dirs]
        self.data_name, self.coord_name, self.feat_name, self.label_name = data_name, coord_name, feat_name, label_name
        self.label_map = {int(k): v for k, v in (label_map or {}).items()} or None
        self.events_per_dir, self.cache_size = events_per_dir, cache_size
        self.items, self._cache = [], {}
        per_dir = []
        for d in self.dirs:
            if not os.path.isdir(d):
                raise RuntimeError("{0} is not a valid directory.".format(d))
            files = sorted((os.path.join(d, f) for f in os.listdir(d) if fnmatch.fnmatch(f, file_pattern)), key=_sort_key)
            if not files:
                raise RuntimeError("No hdf5 datasets found")
            per_dir.append(files)
        n_events = [0] * len(self.dirs)
        if len(per_dir) == 1:
            ordered = list(per_dir[0])
        else:
            tally, ordered = [0] * len(per_dir), []
            while sum(len(f) for f in per_dir) > 0 and any(t < events_per_dir and len(f) > 0 for t, f in zip(tally, per_dir)):
                for i, fs in enumerate(per_dir):
                    while fs and tally[i] < events_per_dir:
                        ordered.append(fs.pop(0))
                        tally[i] += self._nevents(ordered[-1])
                        if not tally[i] < max(tally):
                            break
        for path in ordered:
            di = self.dirs.index(os.path.dirname(path))
            if n_events[di] >= events_per_dir:
                continue
            nfile = self._nevents(path)
            n = min(nfile, events_per_dir - n_events[di])
            n_events[di] += n
            self.items.append({"file_path": path, "n_events": int(nfile), "event_range": [0, int(n) - 1], "dir_index": di})
        self.n_events = n_events

    def _nevents(self, path):
        with h5lite.File(path) as f:
            return int(np.asarray(f[self.data_name].attrs["nevents"]).reshape(-1)[0])

    def __len__(self):
        return len(self.items)

    def _table(self, path):
        if path not in self._cache:
            with h5lite.File(path) as f:
                rec = f[self.data_name][()]
            if len(self._cache) >= self.cache_size:
                self._cache.pop(next(iter(self._cache)))
            self._cache[path] = rec
        return self._cache[path]

    def __getitem__(self, index):
        """([coords int32 [n,3] = (x, y, event), waveform int16 [n, 2*ns]], labels) -- numpy arrays of the rows whose
        event id lies in the item's event range (HDF5Dataset._concat_range: rows up to the FIRST row of event
        `range[1] + 1`, from the first row of event `range[0]`)."""
        di = self.items[index]
        rec = self._table(di["file_path"])
        coords = rec[self.coord_name]
        ev = coords[:, 2]
        lo, hi = di["event_range"]
        first, second = 0, len(ev)
        if hi + 1 < di["n_events"]:
            second = int(np.nonzero(ev == hi + 1)[0][0])
        if lo > 0:
            first = int(np.nonzero(ev == lo)[0][0])
        c = np.ascontiguousarray(coords[first:second]).astype(np.int32, copy=False)
        v = np.ascontiguousarray(rec[self.feat_name][first:second])
        if self.label_name is None:
            y = np.full((hi + 1 - lo,), di["dir_index"], dtype=np.int64)
        else:
            y = np.array(rec[self.label_name][first:second])
            if self.label_map is not None:
                for k, val in self.label_map.items():
                    y[y == k] = val
            y = y.astype(np.int64) if y.dtype == np.int32 else y.astype(np.float32)
        return [c, v], y

    @staticmethod
    def normalized(vals):
        """The reference's float features: int16 -> float32, * 1/(2^14 - 1) (HDF5Dataset.py:227,290,345-346)."""
        return vals.astype(np.float32) * np.float32(MAX_RANGE_INV)


def collate(samples):
    """collate_fn (src/engineering/PSDDataModule.py:10-20) on host arrays: running event offset added to coords[:, 2]
    of every item after the first; returns (coords, waveform, labels, item_rows, item_events)."""
    coords, waves, labels, rows, evs = [], [], [], [0], []
    off = 0
    for (c, v), y in samples:
        c = c.copy()
        if off:
            c[:, 2] += off
        off += len(y)
        coords.append(c)
        waves.append(v)
        labels.append(y)
        rows.append(rows[-1] + c.shape[0])
        evs.append(len(y))
    return np.concatenate(coords), np.concatenate(waves), np.concatenate(labels), rows, evs


def event_batches(files, events_per_batch, rezero=True):
    """Yields (coords (x, y, event-in-batch) int32, waveform int16, labels int64) batches of exactly
    `events_per_batch` whole events cut from the items of `files` in order (the reference uses one item per batch,
    dataloader batch_size 1; fixed-size batches are what a captured graph replays).  The tail that does not fill a
    batch is dropped."""
    pend_c, pend_w, pend_y, have = [], [], [], 0
    for i in range(len(files)):
        (c, w), y = files[i]
        first_ev = files.items[i]["event_range"][0]
        ev = c[:, 2] - first_ev  # 0-based inside the item
        # row offset of every event of the item (ids are contiguous and rows sorted by event)
        starts = np.searchsorted(ev, np.arange(len(y) + 1), side="left")
        e0 = 0
        while e0 < len(y):
            take = min(len(y) - e0, events_per_batch - have)
            r0, r1 = int(starts[e0]), int(starts[e0 + take])
            cc = c[r0:r1].copy()
            cc[:, 2] = ev[r0:r1] - e0 + have
            pend_c.append(cc)
            pend_w.append(w[r0:r1])
            pend_y.append(y[e0:e0 + take])
            have += take
            e0 += take
            if have == events_per_batch:
                yield np.concatenate(pend_c), np.concatenate(pend_w), np.concatenate(pend_y)
                pend_c, pend_w, pend_y, have = [], [], [], 0


def feed(step, batches, on_loss=None):
    """Drives a harness.GraphTrainStep from host batches with double-buffered pinned staging: while the graph replays
    batch i, batch i+1 is copied into pinned memory and sent to the device on the copy stream.  Returns the number of
    steps.  `batches` yields (coords, waveform, labels) numpy arrays; on_loss(i, loss) is called with the loss of step i
    (a host read per step: leave it None for throughput)."""
    import torch
    cap, B = step.row_capacity, step.batch_size
    pins = [(torch.empty((cap, 3), dtype=torch.int32).pin_memory(),
             torch.empty((cap, step.wave.shape[1]), dtype=step.wave.dtype).pin_memory(),
             torch.empty((B,), dtype=step.target.dtype).pin_memory()) for _ in range(3)]

    copied = [None, None, None]  # event after the H2D copy that last read pinned set k: the host must not overwrite it earlier

    def stage(k, batch):
        c, w, y = batch
        n = c.shape[0]
        if n > cap:
            raise ValueError("batch has %d rows, graph capacity is %d" % (n, cap))
        pc, pw, py = pins[k % 3]
        if copied[k % 3] is not None:
            copied[k % 3].synchronize()
        pc[:n].copy_(torch.from_numpy(np.ascontiguousarray(c)))
        pw[:n].copy_(torch.from_numpy(np.ascontiguousarray(w)))
        py[:len(y)].copy_(torch.from_numpy(np.ascontiguousarray(y)).to(py.dtype))
        step.prefetch(pc[:n], pw[:n], py[:len(y)])
        copied[k % 3] = step.pending_ready()

    it = iter(batches)
    nxt = next(it, None)
    k = 0
    if nxt is None:
        return 0
    stage(0, nxt)
    while True:
        out = step.run()
        nxt = next(it, None)
        if nxt is not None:
            stage(k + 1, nxt)
        if on_loss is not None:
            on_loss(k, float(out.item()))
        k += 1
        if nxt is None:
            break
    torch.cuda.synchronize()
    return k
