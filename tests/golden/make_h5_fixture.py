"""Generates tests/golden/h5/* and tests/golden/h5_reader_golden.npz by running the REFERENCE's own reader code --
src/datasets/HDF5IO.py (H5Input.next_chunk) and src/datasets/HDF5Dataset.py (HDF5Dataset.__init__ / __getitem__ /
_concat_range) -- unmodified over pulse files, with `h5py` replaced by a thin adapter over waveformml_b200.io.h5lite
(h5py / libhdf5 do not exist in this image).  Run in the build container only:

    python tests/golden/make_h5_fixture.py

What this pins: the chunking that keeps events whole, the file ordering / events_per_dir truncation, the event-range
slicing, the label rule and the dtype contract of the reference's data path (row f2).  What it does NOT pin: the HDF5
byte format (that is tests/test_h5lite.py against a file written by the real HDF5 library)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from waveformml_b200.io import events, h5lite  # noqa: E402

# ---- stand-ins for modules the reference imports but this image lacks (none of them is on the code path used here)
h5py = types.ModuleType("h5py")


class _File(h5lite.File):
    def __init__(self, path, mode="r", **kw):
        assert mode == "r"
        super().__init__(path)


h5py.File = _File
for _n in ("h5t", "h5f", "h5d", "h5s", "Datatype"):
    setattr(h5py, _n, None)
sys.modules["h5py"] = h5py
sys.modules["git"] = types.ModuleType("git")
pl = types.ModuleType("pytorch_lightning")
plp = types.ModuleType("pytorch_lightning.plugins")
plp.DDPPlugin = object
pl.plugins = plp
sys.modules["pytorch_lightning"], sys.modules["pytorch_lightning.plugins"] = pl, plp

import torch  # noqa: E402
from src.datasets.HDF5Dataset import HDF5Dataset  # noqa: E402
from src.datasets.HDF5IO import H5Input  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "h5")


def make_file(path, n_events, seed, ns=65, **kw):
    """A `*WaveformPairSim.h5` file: WaveformPairs table of WaveformPairCal records + nevents attribute."""
    sys.path.insert(0, ROOT)
    from waveformml_b200.synth import make_events
    ev = make_events(n_events, n_samples=ns, seed=seed)
    n = ev["coords"].shape[0]
    rng = np.random.default_rng(seed + 1)
    rec = np.zeros(n, dtype=events.pair_record_dtype(ns))
    rec["coord"], rec["waveform"] = ev["coords"], ev["wave"]
    rec["evt"] = ev["coords"][:, 2]
    rec["t"] = np.cumsum(rng.random(n))
    for k in ("dt", "z", "E", "PSD"):
        rec[k] = rng.random(n).astype(np.float32)
    rec["PE"], rec["EZ"] = rng.random((n, 2)).astype(np.float32), rng.random((n, 2)).astype(np.float32)
    rec["PID"] = rng.integers(1, 7, size=n).astype(np.int32)
    w = h5lite.Writer(path)
    w.create_dataset("WaveformPairs", rec, attrs={"nevents": np.array([n_events], dtype=np.int64)}, **kw)
    w.close()
    return rec


def main():
    for sub in ("Gamma", "Electron"):
        os.makedirs(os.path.join(OUT, sub), exist_ok=True)
    specs = [("Gamma", 0, 23, dict(chunks=16, gzip=9)), ("Gamma", 1, 17, dict()),
             ("Electron", 0, 19, dict(chunks=1024, gzip=4, shuffle=True)), ("Electron", 1, 31, dict(chunks=8))]
    for sub, i, nev, kw in specs:
        make_file(os.path.join(OUT, sub, "s_%d_WaveformPairSim.h5" % i), nev, 100 * (i + 1) + len(sub), **kw)
    gold = {}
    # ---- H5Input.next_chunk (reference code) over one file, several chunk sizes
    path = os.path.join(OUT, "Electron", "s_1_WaveformPairSim.h5")
    for nrows in (5, 16, 50, 1000):
        inp = H5Input(path)
        inp.setup_table("WaveformPairs", events.pair_record_dtype(65), "coord", event_index_coord=2)
        lens, sums = [], []
        for _ in range(2):  # two passes: the reader rewinds after returning None once
            while True:
                d = inp.next_chunk(nrows)
                if d is None:
                    lens.append(-1)
                    break
                lens.append(len(d))
                sums.append(int(d["waveform"].astype(np.int64).sum()) + int(d["coord"].astype(np.int64).sum()))
        gold["chunks_%d_len" % nrows] = np.array(lens)
        gold["chunks_%d_sum" % nrows] = np.array(sums)
        inp.close()
    # ---- HDF5Dataset (reference code): two class directories, truncated to 30 events per directory
    dirs = [os.path.join(OUT, "Gamma"), os.path.join(OUT, "Electron")]
    for tag, per_dir, label_name in (("a", 30, None), ("b", 1000, None), ("c", 30, "PID")):
        ds = HDF5Dataset(dirs, "*WaveformPairSim.h5", "WaveformPairs", "coord", "waveform", per_dir, torch.device("cpu"),
                         label_name=label_name, normalize=True)
        gold[tag + "_n_items"] = np.array([len(ds)])
        gold[tag + "_files"] = np.array([os.path.relpath(di["file_path"], OUT) for di in ds.info["data_info"]])
        gold[tag + "_ranges"] = np.array([di["event_range"] + [di["n_events"], di["dir_index"]] for di in ds.info["data_info"]])
        for i in range(len(ds)):
            (c, v), y = ds[i]
            gold["%s_%d_coords" % (tag, i)] = c.numpy()
            gold["%s_%d_vals" % (tag, i)] = v.numpy()
            gold["%s_%d_y" % (tag, i)] = y.numpy()
    # ---- an item that starts in the middle of a file (event_range[0] > 0, as PulseDataset's splits produce)
    ds = HDF5Dataset(dirs[:1], "*WaveformPairSim.h5", "WaveformPairs", "coord", "waveform", 1000, torch.device("cpu"),
                     normalize=True)
    ds.info["data_info"][0]["event_range"] = [7, 15]
    (c, v), y = ds[0]
    gold["mid_coords"], gold["mid_vals"], gold["mid_y"] = c.numpy(), v.numpy(), y.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "h5_reader_golden.npz"), **gold)
    print("wrote", len(gold), "arrays;", sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(OUT) for f in fs),
          "bytes of h5 fixtures")


if __name__ == "__main__":
    main()
