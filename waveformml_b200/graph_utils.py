"""Host-side mirror of src/utils/GraphUtils.py:7-40 (window_edges) over wfsp_window_edges: same signature and
result (edge order included), computed on the GPU.  There is no CPU path."""
import torch

from . import _lib


def window_edges(coo, batch, max_dist=1, self_loops=True):
    """coo: int64 [N, 2] (x, y); batch: int64 [N] event id per hit (hits of one event contiguous, as the
    reference's collate produces them).  Returns edge_index int64 [2, E] on the same device."""
    assert coo.dtype == torch.int64 and batch.dtype == torch.int64
    _lib.require_cuda(coo, batch)
    lib = _lib.load()
    dev = coo.device
    n = coo.shape[0]
    xy = coo.transpose(0, 1).contiguous()
    batch = batch.contiguous()
    ws_bytes = lib.wfsp_window_edges_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        args = (int(max_dist) + 1, n, _lib.ptr(xy[0]), _lib.ptr(xy[1]), _lib.ptr(batch), int(bool(self_loops)))
        _lib.check(lib.wfsp_window_edges(*args, None, None, 0, _lib.ptr(count), _lib.ptr(ws), ws_bytes, _lib.stream()))
        e = int(count.item())  # one readback: the output shape is data dependent (the reference over-allocates)
        edge_index = torch.empty((2, e), dtype=torch.int64, device=dev)
        if e:
            _lib.check(lib.wfsp_window_edges(*args, _lib.ptr(edge_index[0]), _lib.ptr(edge_index[1]), e, _lib.ptr(count),
                                             _lib.ptr(ws), ws_bytes, _lib.stream()))
    return edge_index
