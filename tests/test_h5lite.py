"""Row f2 (HDF5 event reader / on-disk contract) on the CPU: the minimal HDF5 implementation against a file written by
the real HDF5 library, writer -> reader round trips over every layout / filter the reference's files use, and the
event-level semantics (chunks that keep events whole, file ordering, event-range slicing, labels, dtypes) against
goldens produced by the REFERENCE's own H5Input / HDF5Dataset code (tests/golden/make_h5_fixture.py)."""
import os

import numpy as np
import pytest

from waveformml_b200.io import events, h5lite

HERE = os.path.dirname(os.path.abspath(__file__))
H5DIR = os.path.join(HERE, "golden", "h5")
GOLD = np.load(os.path.join(HERE, "golden", "h5_reader_golden.npz"))


def test_reads_a_file_written_by_the_real_hdf5_library():
    """MATLAB 7.3 = HDF5 with a 512-byte user block (superblock 0, symbol-table group, v1 object header with an
    attribute, contiguous float64 dataset): exercises the base-address handling and every structure of an old-style file."""
    f = h5lite.File(os.path.join(H5DIR, "testhdf5_7.4_GLNX86.mat"))
    assert f.keys() == ["testdouble"] and f._base == 512
    ds = f["testdouble"]
    assert ds.dtype == np.dtype("<f8") and ds.shape == (9, 1)
    assert bytes(ds.attrs["MATLAB_class"]) == b"double"
    np.testing.assert_array_equal(ds[()][:, 0], np.arange(9) * (np.pi / 4))  # 0 : pi/4 : 2*pi
    try:  # the same variable through scipy's independent reader of the v5 sibling file
        import scipy.io
        sib = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testdouble_7.4_GLNX86.mat")
        if os.path.exists(sib):
            np.testing.assert_array_equal(ds[()][:, 0], scipy.io.loadmat(sib)["testdouble"][0])
    except ImportError:
        pass


def _records(n, ns=65, seed=0):
    rng = np.random.default_rng(seed)
    dt = events.pair_record_dtype(ns)
    a = np.zeros(n, dtype=dt)
    for k in dt.names:
        a[k] = rng.integers(-300, 1000, size=a[k].shape).astype(a[k].dtype)
    return a


@pytest.mark.parametrize("kw", [dict(), dict(chunks=1024, gzip=9), dict(chunks=64, gzip=4, shuffle=True), dict(chunks=7),
                                dict(chunks=3, gzip=1)], ids=["contiguous", "gzip9-1024", "shuffle-gzip", "chunks7", "deep-btree"])
def test_writer_reader_round_trip(tmp_path, kw):
    a = _records(4099)
    assert a.dtype.itemsize == 324  # H5CompoundTypes.py:119
    p = str(tmp_path / "t.h5")
    w = h5lite.Writer(p)
    w.create_dataset("WaveformPairs", a, attrs={"nevents": np.array([77], dtype=np.int64)}, **kw)
    w.create_dataset("labels", np.arange(10, dtype=np.int32))
    w.create_dataset("empty", np.zeros((0,), dtype=a.dtype), chunks=16, gzip=1)
    w.close()
    with h5lite.File(p) as f:
        assert sorted(f.keys()) == ["WaveformPairs", "empty", "labels"]
        ds = f["WaveformPairs"]
        assert ds.dtype == a.dtype and ds.shape == (4099,) and int(ds.attrs["nevents"][0]) == 77
        assert ds.chunks == ((kw["chunks"],) if "chunks" in kw else None)
        assert np.array_equal(ds[()].view(np.uint8), a.view(np.uint8))
        np.testing.assert_array_equal(ds[1000:3001]["coord"], a["coord"][1000:3001])  # a range that crosses chunks
        np.testing.assert_array_equal(ds["waveform"], a["waveform"])
        assert ds[4098]["evt"] == a[4098]["evt"] and ds[-1]["evt"] == a[-1]["evt"]
        np.testing.assert_array_equal(f["labels"][()], np.arange(10))
        assert f["empty"].shape == (0,) and len(f["empty"][()]) == 0


def test_datatype_messages_all_three_compound_encodings():
    """Encodings 1 (member dimensions inline), 2 (array class members) and 3 (unpadded names, minimal offsets) of the
    same record parse to one numpy dtype; encoding 2 is what the writer emits."""
    import struct
    dt = np.dtype([("coord", "<i4", (3,)), ("waveform", "<i2", (4,)), ("PID", "<i4")])
    v2 = h5lite.encode_datatype(dt)
    assert h5lite.parse_datatype(v2)[0] == dt

    def fixed(size, signed=True):
        return struct.pack("<BBBBI", (1 << 4) | 0, 8 if signed else 0, 0, 0, size) + struct.pack("<HH", 0, 8 * size)

    def name(s, pad):
        b = s.encode() + b"\0"
        return b + b"\0" * ((-len(b)) % 8 if pad else 0)
    # version 1: name padded, offset, rank, 3 reserved, perm(4), reserved(4), 4 dims, member type
    v1 = struct.pack("<BBBBI", (1 << 4) | 6, 3, 0, 0, dt.itemsize)
    for nm, off, rank, dims, size in (("coord", 0, 1, (3, 0, 0, 0), 4), ("waveform", 12, 1, (4, 0, 0, 0), 2), ("PID", 20, 0, (0,) * 4, 4)):
        v1 += name(nm, True) + struct.pack("<IB3xII4I", off, rank, 0, 0, *dims) + fixed(size)
    assert h5lite.parse_datatype(v1)[0] == dt
    # version 3: name unpadded, 1-byte offsets (itemsize < 256), array class version 3 (no permutation)
    v3 = struct.pack("<BBBBI", (3 << 4) | 6, 3, 0, 0, dt.itemsize)
    arr = lambda n, size: struct.pack("<BBBBI", (3 << 4) | 10, 0, 0, 0, n * size) + struct.pack("<BI", 1, n) + fixed(size)  # noqa: E731
    v3 += name("coord", False) + bytes([0]) + arr(3, 4) + name("waveform", False) + bytes([12]) + arr(4, 2)
    v3 += name("PID", False) + bytes([20]) + fixed(4)
    assert h5lite.parse_datatype(v3)[0] == dt


def _open(path, nrows=None):
    inp = events.H5Input(path)
    inp.setup_table("WaveformPairs", events.pair_record_dtype(65), "coord", event_index_coord=2)
    return inp


@pytest.mark.parametrize("nrows", [5, 16, 50, 1000])
def test_next_chunk_matches_the_reference_reader(nrows):
    """src/datasets/HDF5IO.py:55-79 run unmodified produced these chunk lengths / checksums: events stay whole, the
    last chunk is the remainder, then None once, then the reader starts over."""
    inp = _open(os.path.join(H5DIR, "Electron", "s_1_WaveformPairSim.h5"))
    lens, sums = [], []
    for _ in range(2):
        while True:
            d = inp.next_chunk(nrows)
            if d is None:
                lens.append(-1)
                break
            lens.append(len(d))
            sums.append(int(d["waveform"].astype(np.int64).sum()) + int(d["coord"].astype(np.int64).sum()))
            ev = d["coord"][:, 2]
            assert np.all(np.diff(ev) >= 0)
    np.testing.assert_array_equal(lens, GOLD["chunks_%d_len" % nrows])
    np.testing.assert_array_equal(sums, GOLD["chunks_%d_sum" % nrows])
    # no event is split across chunks
    inp2 = _open(os.path.join(H5DIR, "Electron", "s_1_WaveformPairSim.h5"))
    seen = set()
    while True:
        d = inp2.next_chunk(nrows)
        if d is None:
            break
        ids = set(np.unique(d["coord"][:, 2]).tolist())
        assert not (ids & seen)
        seen |= ids
    assert seen == set(range(31))


@pytest.mark.parametrize("tag,per_dir,label_name", [("a", 30, None), ("b", 1000, None), ("c", 30, "PID")])
def test_pulse_files_match_the_reference_dataset(tag, per_dir, label_name):
    """HDF5Dataset.__init__ / __getitem__ / _concat_range (HDF5Dataset.py:136-217, 225-347) run unmodified produced
    the goldens: same items in the same order, same truncation, same rows, labels and dtypes.  The reference hands
    out float32 features already multiplied by 1/16383; here the waveform stays int16 for the GPU batcher and
    `normalized()` reproduces the reference's values bit for bit."""
    dirs = [os.path.join(H5DIR, "Gamma"), os.path.join(H5DIR, "Electron")]
    pf = events.PulseFiles(dirs, events_per_dir=per_dir, label_name=label_name)
    assert len(pf) == int(GOLD[tag + "_n_items"][0])
    assert [os.path.relpath(it["file_path"], H5DIR) for it in pf.items] == list(GOLD[tag + "_files"])
    np.testing.assert_array_equal(
        [it["event_range"] + [it["n_events"], it["dir_index"]] for it in pf.items], GOLD[tag + "_ranges"])
    for i in range(len(pf)):
        (c, v), y = pf[i]
        gc, gv, gy = GOLD["%s_%d_coords" % (tag, i)], GOLD["%s_%d_vals" % (tag, i)], GOLD["%s_%d_y" % (tag, i)]
        assert c.dtype == np.int32 and v.dtype == np.int16
        np.testing.assert_array_equal(c, gc)
        assert gv.dtype == np.float32
        np.testing.assert_array_equal(events.PulseFiles.normalized(v), gv)
        assert y.dtype == gy.dtype
        np.testing.assert_array_equal(y, gy)


def test_item_starting_mid_file():
    pf = events.PulseFiles([os.path.join(H5DIR, "Gamma")])
    pf.items[0]["event_range"] = [7, 15]
    (c, v), y = pf[0]
    np.testing.assert_array_equal(c, GOLD["mid_coords"])
    np.testing.assert_array_equal(events.PulseFiles.normalized(v), GOLD["mid_vals"])
    np.testing.assert_array_equal(y, GOLD["mid_y"])
    assert c[0, 2] == 7 and c[-1, 2] == 15 and len(y) == 9


def test_collate_and_fixed_size_batches():
    dirs = [os.path.join(H5DIR, "Gamma"), os.path.join(H5DIR, "Electron")]
    pf = events.PulseFiles(dirs)
    samples = [pf[i] for i in range(len(pf))]
    coords, wave, labels, rows, evs = events.collate(samples)
    total = sum(len(s[1]) for s in samples)
    assert rows[-1] == coords.shape[0] == wave.shape[0] and labels.shape[0] == total == 23 + 17 + 19 + 31
    assert np.all(np.diff(coords[:, 2]) >= 0) and coords[-1, 2] == total - 1  # running event offset (PSDDataModule.py:10-20)
    got = list(events.event_batches(pf, 16))
    assert len(got) == total // 16
    seen_rows = 0
    for c, w, y in got:
        assert y.shape == (16,) and c[0, 2] == 0 and c[-1, 2] == 15 and np.all(np.diff(c[:, 2]) >= 0)
        assert set(np.unique(c[:, 2]).tolist()) == set(range(16))  # every event of the synthetic files has >= 1 hit
        seen_rows += c.shape[0]
    # the batches are the collated stream cut at event boundaries
    np.testing.assert_array_equal(np.concatenate([w for _, w, _ in got]), wave[:seen_rows])
    np.testing.assert_array_equal(np.concatenate([y for _, _, y in got]), labels[:16 * len(got)])


def test_h5output_writes_what_h5input_reads(tmp_path):
    """The inference writer path (PredictionWriter.write_predictions, PredictionWriter.py:73-104): chunks from
    next_chunk are converted and appended with add_rows; the result holds the same rows."""
    src = os.path.join(H5DIR, "Gamma", "s_0_WaveformPairSim.h5")
    inp = _open(src)
    out = events.H5Output(str(tmp_path / "pred.h5"))
    out.create_table("WaveformPairs", (inp.table_length,), inp.record_type)
    out.set_attr("WaveformPairs", "nevents", inp.table.attrs["nevents"])
    while True:
        d = inp.next_chunk(20)
        if d is None:
            break
        d = d.copy()
        d["PID"] = 7  # "swap_values": overwrite one field with the model's output
        out.add_rows("WaveformPairs", d)
    out.close()
    with h5lite.File(str(tmp_path / "pred.h5")) as f, h5lite.File(src) as g:
        a, b = f["WaveformPairs"][()], g["WaveformPairs"][()]
        assert f["WaveformPairs"].chunks == (1024,) and int(f["WaveformPairs"].attrs["nevents"][0]) == 23
        np.testing.assert_array_equal(a["waveform"], b["waveform"])
        assert np.all(a["PID"] == 7)
