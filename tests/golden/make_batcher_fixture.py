"""Generates tests/golden/batcher_golden.npz by running the REFERENCE's own collate_fn
(/root/reference/src/engineering/PSDDataModule.py:10-20) and its normalisation constant
(src/datasets/HDF5Dataset.py:15-17) on small seeded items.  Build container only.

    python tests/golden/make_batcher_fixture.py

pytorch_lightning and src.utils.util are absent / heavy here, so they are stubbed before the
reference module is imported; collate_fn itself runs unmodified.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)

pl = types.ModuleType("pytorch_lightning")
pl.LightningDataModule = object
sys.modules["pytorch_lightning"] = pl
for name in ("src", "src.utils", "src.utils.util"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["src.utils.util"].DictionaryUtility = object
sys.modules["src.utils.util"].ModuleUtility = object
spec = importlib.util.spec_from_file_location("ref_psd_datamodule", os.path.join(REF, "src/engineering/PSDDataModule.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

from waveformml_b200.synth import make_events  # noqa: E402

MAX_RANGE_INV = 1.0 / (2 ** 14 - 1)  # src/datasets/HDF5Dataset.py:15-17

items, raw = [], {}
for i, (nev, seed) in enumerate([(5, 10), (1, 11), (7, 12), (3, 13)]):
    ev = make_events(nev, n_samples=4, seed=seed)
    coords = torch.from_numpy(ev["coords"].copy())            # (x, y, item-local event), int32
    wave = torch.from_numpy(ev["wave"].copy())
    feats = wave.type(torch.float32)
    feats *= MAX_RANGE_INV                                     # HDF5Dataset.py:345-346
    labels = torch.from_numpy(ev["labels"].copy())
    raw["coords%d" % i], raw["wave%d" % i], raw["labels%d" % i] = ev["coords"], ev["wave"], ev["labels"]
    items.append(([coords, feats], labels))

(c, f), y = mod.collate_fn(items)
np.savez(os.path.join(ROOT, "tests/golden/batcher_golden.npz"), n_items=len(items), out_coords=c.numpy(),
         out_feats=f.numpy(), out_labels=y.numpy(), **raw)
print("collate_fn ->", tuple(c.shape), tuple(f.shape), tuple(y.shape), "batch size", int(c[-1, -1]) + 1)
