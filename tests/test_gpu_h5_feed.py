"""Row f2 end to end on the GPU: pulse files -> PulseFiles / event_batches -> pinned double-buffered staging
(events.feed -> GraphTrainStep.prefetch) -> captured training step; and the a2 dtype contract through the file path:
the GPU batcher's features equal the float32 values the REFERENCE's HDF5Dataset hands out (golden from its own code)."""
import copy
import os

import numpy as np
import pytest
import torch

from waveformml_b200 import batcher, harness, spconv, stacks
from waveformml_b200.io import events

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
H5DIR = os.path.join(HERE, "golden", "h5")
DIRS = [os.path.join(H5DIR, "Gamma"), os.path.join(H5DIR, "Electron")]


def test_batcher_features_equal_reference_dataset_values(cuda_device):
    gold = np.load(os.path.join(HERE, "golden", "h5_reader_golden.npz"))
    pf = events.PulseFiles(DIRS, events_per_dir=30)
    for i in range(len(pf)):
        (c, v), _ = pf[i]
        idx, feats = batcher.pack_batch(torch.from_numpy(c).to(cuda_device), torch.from_numpy(v).to(cuda_device))
        assert torch.equal(feats.cpu(), torch.from_numpy(gold["a_%d_vals" % i]))  # int16 -> f32 * 1/16383, bit exact
        assert torch.equal(idx.cpu(), torch.from_numpy(gold["a_%d_coords" % i])[:, [2, 0, 1]])  # SPConvNet.py:63-64 permute


def test_feed_from_files_tracks_eager_training(cuda_device):
    spconv.set_math_mode("fp32")
    try:
        B = 16
        torch.manual_seed(4)
        m1 = stacks.PSDClassifier(n_samples=65, n_classes=2).to(cuda_device).train()
        m2 = copy.deepcopy(m1)
        s1 = harness.TrainStep(m1, "psd", lr=0.01, momentum=0.9)
        s2 = harness.GraphTrainStep(m2, "psd", B, B * 10, 130, lr=0.01, momentum=0.9)
        pf = events.PulseFiles(DIRS)
        batches = list(events.event_batches(pf, B))
        assert len(batches) == 5
        c0, w0, y0 = (torch.from_numpy(a).to(cuda_device) for a in batches[0])
        s2.load(c0, w0, y0)
        s2.capture()
        graph_losses = []
        n = events.feed(s2, iter(batches), on_loss=lambda i, v: graph_losses.append(v))
        assert n == 5
        for (c, w, y), lg in zip(batches, graph_losses):
            idx, feats = batcher.pack_batch(torch.from_numpy(c).to(cuda_device), torch.from_numpy(w).to(cuda_device))
            le = float(s1.step(idx, feats, torch.from_numpy(y).to(cuda_device), B))
            assert abs(le - lg) < 2e-3 * max(abs(le), 1e-3), (le, lg)
        for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            rel = float((a.detach() - b.detach()).norm() / a.detach().norm())
            assert rel < 1e-3, (k, rel)
        # throughput form: no loss readback, same result
        assert events.feed(s2, iter(batches)) == 5
        assert s2.duplicate_inputs() is False
    finally:
        spconv.set_math_mode("bf16")
