"""Input feed (show_edit_tell_b200/feed.py) against the reference's own Dataset / collate code, AST-extracted and run on
synthetic files in the reference's on-disk layout.  The live comparison runs where /root/reference exists; the format
checks run everywhere."""
import ast
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_extract as RX
from show_edit_tell_b200 import feed


def _write_tree(root, n_img=4, V=50, seed=0):
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "caption data"))
    os.makedirs(os.path.join(root, "data", "cocobu_att"))
    os.makedirs(os.path.join(root, "data", "cocobu_fc"))
    names = ["COCO_%06d.jpg" % i for i in range(n_img)]
    caps, lens = [], []
    for _ in range(n_img * 5):
        L = int(rng.randint(5, 21))
        caps.append([V - 2] + rng.randint(1, V - 3, size=L - 2).tolist() + [V - 1] + [0] * (20 - L))
        lens.append(L)
    util, dets = {}, []
    for i, nm in enumerate(names):
        pl = int(rng.randint(3, 19))
        util[nm] = {"caption": "x", "encoded_previous_caption": rng.randint(1, V - 3, size=pl).tolist() + [0] * (18 - pl),
                    "previous_caption_length": [pl], "image_ids": 1000 + i}
        dets.append(["v" if i % 2 else "t", i // 2])
        n = int(rng.randint(10, 101))
        np.savez(os.path.join(root, "data", "cocobu_att", "%d.npz" % (1000 + i)), feat=rng.rand(n, 2048).astype(np.float32))
        np.save(os.path.join(root, "data", "cocobu_fc", "%d.npy" % (1000 + i)), rng.rand(2048).astype(np.float32))
    for fn, obj in (("TRAIN_CAPTIONS_coco.json", caps), ("TRAIN_CAPLENS_coco.json", lens), ("TRAIN_names_coco.json", names),
                    ("CAPUTIL_train.json", util), ("TRAIN_GENOME_DETS_coco.json", dets)):
        with open(os.path.join(root, "caption data", fn), "w") as f:
            json.dump(obj, f)
    feats = {"t": rng.rand((n_img + 1) // 2, 36, 2048).astype(np.float32), "v": rng.rand(n_img // 2 + 1, 36, 2048).astype(np.float32)}
    return feats


def _ref_namespace(rel_path, names, extra_ns):
    with open(os.path.join(RX.REFERENCE_ROOT, rel_path)) as f:
        tree = ast.parse(f.read())
    body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    ns = {"torch": torch, "np": np, "os": os, "json": json, "Dataset": torch.utils.data.Dataset}
    ns.update(extra_ns)
    exec(compile(ast.Module(body=body, type_ignores=[]), rel_path, "exec"), ns)
    return ns


def test_formats_and_prefetcher(tmp_path):
    root = str(tmp_path)
    _write_tree(root)
    ds = feed.AdaptiveTrainSet(root)
    assert len(ds) == 20
    batch = feed.collate_adaptive([ds[i] for i in (0, 7, 13)], root=root, pin=False)
    images, mean = batch[0], batch[1]
    assert images.shape == (3, 100, 2048) and images.dtype == torch.float32 and mean.shape == (3, 2048)
    n0 = np.load(os.path.join(root, "data", "cocobu_att", "1000.npz"))["feat"].shape[0]
    assert bool((images[0, n0:] == 0).all()) and bool((images[0, :n0].abs().sum(1) > 0).all())
    loader = [feed.collate_adaptive([ds[i], ds[i + 1]], root=root, pin=False) for i in range(0, 10, 2)]
    got = list(feed.DevicePrefetcher(loader, "cpu"))
    assert len(got) == 5 and all(torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) for a, b in zip(got, loader))


@pytest.mark.skipif(not RX.reference_available(), reason="reference tree not present")
def test_datasets_match_reference_classes(tmp_path, monkeypatch):
    root = str(tmp_path)
    feats = _write_tree(root, seed=3)
    monkeypatch.chdir(root)                                # the reference opens relative paths
    # adaptive: Dataset items and collate_fn_train (adaptive_features/editnet_adaptive.py:17-80)
    ns = _ref_namespace("adaptive_features/editnet_adaptive.py", {"COCOTrainDataset", "collate_fn_train"}, {})
    ref_ds, mine = ns["COCOTrainDataset"](), feed.AdaptiveTrainSet(root)
    assert len(ref_ds) == len(mine)
    idx = [0, 3, 8, 11, 19]
    for i in idx:
        assert all(torch.equal(a, b) for a, b in zip(ref_ds[i], mine[i]))
    ref_b = ns["collate_fn_train"]([ref_ds[i] for i in idx])
    my_b = feed.collate_adaptive([mine[i] for i in idx], root=root, pin=False)
    assert torch.equal(ref_b[0].float(), my_b[0]) and torch.equal(ref_b[1].float(), my_b[1])   # cast of :574-575
    assert all(torch.equal(a, b) for a, b in zip(ref_b[2:], my_b[2:]))
    # fixed 36-region set (editnet.py:19-77), h5py replaced by arrays of the same layout
    fake = types.ModuleType("h5py")
    fake.File = lambda path, mode="r": {"image_features": feats["t" if "train36" in path else "v"]}
    ns2 = _ref_namespace("editnet.py", {"COCOTrainDataset"}, {"h5py": fake})
    ref2, mine2 = ns2["COCOTrainDataset"](), feed.FixedRegionTrainSet(root, feats["t"], feats["v"])
    for i in idx:
        assert all(torch.equal(a, b) for a, b in zip(ref2[i], mine2[i]))
