"""Input feed for the decode path: the reference's on-disk formats -> pinned host batches ->
device, with the host-to-device copy of batch i+1 overlapping the compute of batch i.

Datasets mirror the reference's `Dataset` classes item for item (same files, same fields, same order):
  * `FixedRegionTrainSet`  — `COCOTrainDataset` of editnet.py:19-77: 36x2048 bottom-up features from the two HDF5
    files written by bottom-up_features/tsv.py:56-68 (`image_features (I,36,2048) float32`); any array-like indexed by
    image row works (h5py dataset, np.memmap, ndarray), which keeps this module free of an h5py dependency;
  * `AdaptiveTrainSet` + `collate_adaptive` — `COCOTrainDataset` / `collate_fn_train` of
    adaptive_features/editnet_adaptive.py:17-80: per-image `cocobu_att/<id>.npz['feat'] (n,2048)` zero-padded to 100
    regions and `cocobu_fc/<id>.npy (2048,)` (adaptive_features/make_bu_data.py:49-50).  The reference pads into float64
    and casts on the device (:66-67, :574-575); here the batch is built in float32 directly (bit-identical after the cast:
    the files hold float32) and in pinned memory.
`DevicePrefetcher` wraps any iterable of host batches (tuples of tensors).
"""
import json
import os

import numpy as np
import torch


class _CaptionTables:
    """the four JSON tables both train sets read (editnet.py:30-43; preprocess_caps.py:126-149,
    preprocess_existing_caps.py:25-32)"""

    def __init__(self, root, split="TRAIN", util="train"):
        cd = os.path.join(root, "caption data")
        with open(os.path.join(cd, "%s_CAPTIONS_coco.json" % split)) as j:
            self.captions = json.load(j)
        with open(os.path.join(cd, "%s_CAPLENS_coco.json" % split)) as j:
            self.caplens = json.load(j)
        with open(os.path.join(cd, "%s_names_coco.json" % split)) as j:
            self.names = json.load(j)
        with open(os.path.join(cd, "CAPUTIL_%s.json" % util)) as j:
            self.caption_util = json.load(j)
        self.cpi = 5                                     # captions per image (editnet.py:27)

    def text_fields(self, i):
        name = self.names[i // self.cpi]
        u = self.caption_util[name]
        caption = torch.LongTensor(self.captions[i])
        caplen = torch.LongTensor([self.caplens[i]])
        previous_caption = torch.LongTensor(u["encoded_previous_caption"])
        prev_caplen = torch.LongTensor(u["previous_caption_length"])
        lo = (i // self.cpi) * self.cpi
        all_captions = torch.LongTensor(self.captions[lo:lo + self.cpi])
        return name, caption, caplen, previous_caption, prev_caplen, all_captions


class FixedRegionTrainSet(torch.utils.data.Dataset):
    """items: (img (36,2048) f32, caption, caplen, previous_caption, prev_caplen, all_captions) — editnet.py:45-74"""

    def __init__(self, root, train_features, val_features):
        self.t = _CaptionTables(root)
        self.train_features, self.val_features = train_features, val_features
        with open(os.path.join(root, "caption data", "TRAIN_GENOME_DETS_coco.json")) as j:
            self.objdet = json.load(j)

    def __len__(self):
        return len(self.t.captions)

    def __getitem__(self, i):
        _, caption, caplen, prev, prev_len, all_caps = self.t.text_fields(i)
        split, row = self.objdet[i // self.t.cpi]
        src = self.val_features if split == "v" else self.train_features          # editnet.py:63-66
        img = torch.from_numpy(np.asarray(src[row], dtype=np.float32))
        return img, caption, caplen, prev, prev_len, all_caps


class AdaptiveTrainSet(torch.utils.data.Dataset):
    """items: (image_id, caption, caplen, previous_caption, prev_caplen, all_captions) — editnet_adaptive.py:39-52"""

    def __init__(self, root):
        self.t = _CaptionTables(root)

    def __len__(self):
        return len(self.t.captions)

    def __getitem__(self, i):
        name, caption, caplen, prev, prev_len, all_caps = self.t.text_fields(i)
        image_id = torch.LongTensor([self.t.caption_util[name]["image_ids"]])
        return image_id, caption, caplen, prev, prev_len, all_caps


def collate_adaptive(data, root=".", max_regions=100, feat_dim=2048, pin=None):
    """`collate_fn_train`, editnet_adaptive.py:58-80: -> (images (B,100,2048), images_mean (B,2048), captions, caplens,
    previous_captions, prev_caplens, all_captions); float32, pinned when a CUDA device is present and the call runs in the
    main process."""
    image_id, caption, caplen, previous_caption, prev_caplen, all_captions = zip(*data)
    B = len(caption)
    if pin is None:
        # pin only in the main process: a forked DataLoader worker must not touch CUDA ("Cannot re-initialize CUDA in
        # forked subprocess"); with num_workers > 0 use DataLoader(pin_memory=True) as the reference does
        pin = torch.cuda.is_available() and torch.utils.data.get_worker_info() is None
    images = torch.zeros(B, max_regions, feat_dim, pin_memory=pin)
    images_mean = torch.zeros(B, feat_dim, pin_memory=pin)
    for i, img_id in enumerate(image_id):
        att = np.load(os.path.join(root, "data", "cocobu_att", "%d.npz" % int(img_id)))["feat"]
        images[i, :att.shape[0]] = torch.from_numpy(np.ascontiguousarray(att, dtype=np.float32))
        images_mean[i] = torch.from_numpy(np.asarray(np.load(os.path.join(root, "data", "cocobu_fc", "%d.npy" % int(img_id))),
                                                     dtype=np.float32))
    return (images, images_mean, torch.stack(caption, 0), torch.stack(caplen, 0), torch.stack(previous_caption, 0),
            torch.stack(prev_caplen, 0), torch.stack(all_captions, 0))


_copy_streams = {}


def _copy_stream(device):
    """One copy stream per device for the life of the process: the caching allocator keeps a block pool per stream, so
    a fresh stream per prefetcher (one per epoch) would cudaMalloc its double buffer again each time -- measured on
    B200: 2-13 ms in front of the first step of every new iterator."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _copy_streams.get(idx)
    if st is None:
        st = _copy_streams[idx] = torch.cuda.Stream(device=device)
    return st


class DevicePrefetcher:
    """Iterate `loader` (host batches: tuples of tensors) one batch ahead: while the consumer computes on batch i, the
    tensors of batch i+1 travel host->device on a copy stream into the other half of a double buffer.  Pinned source
    tensors make the copies truly asynchronous (`DataLoader(pin_memory=True)` as in editnet.py:790-793, or
    `collate_adaptive`).  On a CPU device it degrades to a plain iterator (tests)."""

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self.host_batch = None      # the host tensors of the batch handed out last (lengths stay useful on the host)
        self.cuda = self.device.type == "cuda"
        self.stream = _copy_stream(self.device) if self.cuda else None

    def _stage(self, batch):
        if not self.cuda:
            return tuple(batch), None, tuple(batch)
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev, tuple(batch)

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev, self.host_batch = nxt
            try:
                nxt = self._stage(next(it))            # enqueue the next copy before handing out the current batch
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                for t in cur:                            # the caching allocator must not recycle these under the copy
                    if torch.is_tensor(t):
                        t.record_stream(torch.cuda.current_stream(self.device))
            yield cur

    def __len__(self):
        return len(self.loader)
