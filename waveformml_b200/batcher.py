"""Sparse-tensor batcher on the GPU (wfsp_batch_pack).

Mirrors collate_fn (src/engineering/PSDDataModule.py:10-20), the int16 -> float normalisation of
HDF5Dataset._concat_range (src/datasets/HDF5Dataset.py:282-302,345-346) and the batch-first permute
of SPConvNet.forward (src/models/SPConvNet.py:63-64) in one pass over the rows."""
import torch

from . import _lib
from .synth import MAX_RANGE_INV


def item_tables(item_rows, item_events, device):
    """Device copies of the per-item row offsets and running event offsets collate_fn needs."""
    rows = torch.as_tensor(item_rows, dtype=torch.int64)
    evs = torch.as_tensor(item_events, dtype=torch.int64)
    offs = torch.zeros_like(evs)
    if evs.numel() > 1:
        offs[1:] = torch.cumsum(evs, 0)[:-1]  # running offset; item 0 is left untouched
    return rows.to(device, non_blocking=True), offs.to(device, non_blocking=True), evs.numel()


def pack_batch(coords_xye, wave, item_rows=None, item_events=None, scale=MAX_RANGE_INV, out_dtype=torch.float32,
               n_rows=None, tables=None, feats_stream=None):
    """coords_xye int32 [N,3] = (x, y, event id local to its item) and wave int16|f32 [N,C], both on
    the GPU, items concatenated.  item_rows [n_items+1] row offsets and item_events [n_items] events
    per item (host lists / tensors); omitted = a single item.  Returns (indices int32 [N,3] =
    (event, x, y), features [N,C] = wave * scale).

    Graph path: n_rows = int32 device scalar with the live row count (the inputs are capacity-sized
    static buffers) and tables = item_tables(...) prepared once, so nothing is copied from the host
    inside a captured region.  feats_stream: run the waveform half on that stream (it must already be ordered after
    the inputs; the caller joins it before reading `features`) -- the indices, which the rulebooks wait for, are then
    not queued behind the much larger waveform conversion."""
    lib = _lib.load()
    _lib.require_cuda(coords_xye, wave)
    dev = wave.device
    coords_xye, wave = coords_xye.contiguous(), wave.contiguous()
    if coords_xye.dtype != torch.int32:
        raise RuntimeError("coords must be int32")
    n, c = wave.shape
    if tables is None:
        if item_rows is None:
            item_rows, item_events = [0, n], [0]
        tables = item_tables(item_rows, item_events, dev)
    rows_d, offs_d, n_items = tables
    wdt = {torch.int16: _lib.I16, torch.float32: _lib.F32}[wave.dtype]
    odt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[out_dtype]
    indices = torch.empty((n, 3), dtype=torch.int32, device=dev)
    # bf16 output = the tensor-core operand format (include/wfsp.h section 6): row pitch rounded up to 8 channels,
    # padding zero -- the fused sparse stack then reads the batcher's output directly, no fp32 copy and no cast pass
    pitch = (c + 7) // 8 * 8 if out_dtype == torch.bfloat16 else c
    alloc = torch.empty if (pitch == c or c % 4 == 0) else torch.zeros  # the vector kernel zeroes the padding itself
    feats = alloc((n, pitch), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        if feats_stream is None:
            _lib.check(lib.wfsp_batch_pack(_lib.ptr(coords_xye), _lib.ptr(wave), wdt, n, _lib.ptr(n_rows), c,
                                           _lib.ptr(rows_d), _lib.ptr(offs_d), n_items, float(scale), _lib.ptr(indices),
                                           _lib.ptr(feats), odt, pitch, _lib.stream()))
        else:
            _lib.check(lib.wfsp_batch_pack(_lib.ptr(coords_xye), None, wdt, n, _lib.ptr(n_rows), c,
                                           _lib.ptr(rows_d), _lib.ptr(offs_d), n_items, float(scale), _lib.ptr(indices),
                                           None, odt, pitch, _lib.stream()))
            with torch.cuda.stream(feats_stream):
                _lib.check(lib.wfsp_batch_pack(None, _lib.ptr(wave), wdt, n, _lib.ptr(n_rows), c,
                                               _lib.ptr(rows_d), _lib.ptr(offs_d), n_items, float(scale), None,
                                               _lib.ptr(feats), odt, pitch, _lib.stream()))
    return indices, feats


def collate_fn(batch, scale=None):
    """Same call shape as the reference's collate_fn: batch = list of ([coords, feats], labels) with
    GPU tensors; returns [coords (x, y, global event), feats], labels.  With `scale` the features
    are int16 waveforms that still need normalising (the fused path); without, they are copied."""
    coords = torch.cat([b[0][0] for b in batch])
    wave = torch.cat([b[0][1] for b in batch])
    labels = torch.cat([b[1] for b in batch])
    rows, evs, r = [0], [], 0
    for b in batch:
        r += b[0][0].shape[0]
        rows.append(r)
        evs.append(b[1].shape[0])
    indices, feats = pack_batch(coords, wave, rows, evs, 1.0 if scale is None else scale)
    return [indices[:, [1, 2, 0]].contiguous(), feats], labels
