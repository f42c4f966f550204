"""Experiment: how fast is one C2 training step when the launch sequence is replayed from a CUDA
graph (exact shapes, rulebooks prebuilt)?  Gives the GPU-bound floor that a sync-free static-shape
execution mode could reach, next to the eager step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from waveformml_b200 import batcher, harness, spconv, stacks  # noqa: E402
from waveformml_b200.spconv import ops  # noqa: E402
from waveformml_b200.synth import make_events  # noqa: E402


def main(B):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = stacks.PSDClassifier().to(dev).train()
    step = harness.TrainStep(model, "psd")
    ev = make_events(B, n_samples=150, seed=1234)
    coords, wave = torch.from_numpy(ev["coords"]).to(dev), torch.from_numpy(ev["wave"]).to(dev)
    labels = torch.from_numpy(ev["labels"]).to(dev)

    def eager():
        idx, feats = batcher.pack_batch(coords, wave)
        return step.step(idx, feats, labels, B)

    for _ in range(5):
        eager()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        eager()
    torch.cuda.synchronize()
    print("B=%d eager  %.3f ms/step" % (B, (time.perf_counter() - t0) / 20 * 1e3))

    # prebuild the rulebooks, then hand them out in order while capturing
    built = []
    real_build = ops.build_rulebook

    def recording(*a, **k):
        rb = real_build(*a, **k)
        built.append(rb)
        return rb

    ops.build_rulebook = recording
    eager()
    torch.cuda.synchronize()
    queue = []
    ops.build_rulebook = lambda *a, **k: queue.pop(0)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            queue[:] = list(built)
            eager()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    queue[:] = list(built)
    idx, feats = batcher.pack_batch(coords, wave)
    with torch.cuda.graph(g):
        loss = step.step(idx, feats, labels, B)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print("B=%d graph  %.3f ms/step (loss %.4f)" % (B, a.elapsed_time(b) / 50, float(loss)))
    ops.build_rulebook = real_build


if __name__ == "__main__":
    for B in (64, 1024):
        main(B)
