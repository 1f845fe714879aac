"""The loader's on-disk formats in front of the device ingest.

The reference's datasets (common/datasets/image_dataset.py) are Chainer datasets: a listing file names the frames
(``images.csv``: one path per line, chainer ``ImageDataset``; ``gt.csv``: tab-separated ``path<TAB>int<TAB>int...``,
``LabeledImageDataset.__init__`` :104-110), ``get_example`` decodes one PNG with PIL, resizes it with
``resize_image(image, image_size)`` (PIL LANCZOS, :16-28), scales bounding boxes with chainercv's ``resize_bbox`` (:176-180)
and returns ``image / 255`` (:98, :181-182) -- all per sample, on the host, in loader threads.

Here the same listings and PNGs are read on the host (file I/O and PNG decoding stay there), but a BATCH of decoded uint8
frames is uploaded as it is (a quarter of the float32 bytes) and ``resize_image(...) / 255`` runs on the device
(``loans_stn_ingest_u8``, bit for bit Pillow's result).  ``get_example(i)`` keeps the reference's per-sample contract (CHW
float32 in [0,1] as a CUDA tensor); ``get_batch(indices)`` is the call a training loop should use.

Not mirrored: the random augmentations (``transform_probability > 0`` needs imgaug / chainercv, absent here and outside this
path) and ``image_mode='L'``; both raise.  No CPU fallback: the frames come back as CUDA tensors.
"""
import csv
import os

import numpy as np
import torch

from loans_b200.functions.ingest import FrameIngest


def read_image_listing(path):
    """``images.csv`` as chainer's ``ImageDataset`` reads it: one path per line (surrounding white space stripped)."""
    with open(path) as f:
        return [line.strip() for line in f if line.strip()]


def read_labeled_listing(path, label_dtype=np.int32):
    """``gt.csv`` as the reference reads it (common/datasets/image_dataset.py:104-110): tab-separated, the path first, then
    the label integers (a multiple of four of them = bounding boxes ``top, left, bottom, right``)."""
    pairs = []
    with open(path) as f:
        for pair in csv.reader(f, delimiter='\t'):
            if pair:
                pairs.append((pair[0], list(map(label_dtype, pair[1:]))))
    return pairs


def decode_frame(path):
    """One file -> uint8 (H, W, 3): chainer's ``_read_image_as_array`` (``numpy.asarray(Image.open(path), float32)``, a 2-D
    image gets a channel axis), the reference's ``numpy.tile`` of single-channel frames (:77-78, :157-158), and the front of
    ``resize_image``: ``Image.fromarray(... .astype('uint8')).convert('RGB')`` (:17-21)."""
    from PIL import Image
    with Image.open(path) as f:
        image = np.asarray(f, dtype=np.float32)
    if image.ndim == 2:
        image = image[:, :, None]
    if image.shape[2] == 1:
        image = np.tile(image, (1, 1, 3))
    u8 = image.astype('uint8')
    if u8.shape[2] != 3:                                    # e.g. RGBA: convert('RGB') drops the alpha channel
        u8 = np.asarray(Image.fromarray(u8).convert('RGB'))
    return np.ascontiguousarray(u8)


def resize_bbox(bbox, in_size, out_size):
    """chainercv.transforms.resize_bbox (restated; chainercv is absent): (R, 4) boxes ``y_min, x_min, y_max, x_max`` scaled from
    in_size (H, W) to out_size, float32 in, float32 out."""
    bbox = np.array(bbox, dtype=np.float32, copy=True)
    y_scale = float(out_size[0]) / in_size[0]
    x_scale = float(out_size[1]) / in_size[1]
    bbox[:, 0] = y_scale * bbox[:, 0]
    bbox[:, 2] = y_scale * bbox[:, 2]
    bbox[:, 1] = x_scale * bbox[:, 1]
    bbox[:, 3] = x_scale * bbox[:, 3]
    return bbox


class _DeviceFrames(object):
    """Decoded frames -> float32 NCHW / 255 on the device, one prepared ``FrameIngest`` per (frame size, batch capacity)."""

    def __init__(self, image_size, device):
        if not torch.cuda.is_available():
            raise RuntimeError("the device loader needs a CUDA device (loans_b200 has no CPU fallback)")
        self.image_size = None if image_size is None else (int(image_size[0]), int(image_size[1]))
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._ingests = {}

    def _ingest(self, n, hw):
        key = (hw, self.image_size)
        op = self._ingests.get(key)
        if op is None or op.b < n:
            op = self._ingests[key] = FrameIngest(max(n, 1 if op is None else 2 * op.b), hw, self.image_size, device=self.dev)
        return op

    def __call__(self, frames):
        """list of uint8 (H, W, 3) arrays -> list of float32 (3, oH, oW) CUDA tensors (views of one batch per frame size)."""
        out = [None] * len(frames)
        groups = {}
        for i, f in enumerate(frames):
            groups.setdefault(f.shape[:2], []).append(i)
        for hw, idx in groups.items():
            host = torch.from_numpy(np.stack([frames[i] for i in idx])).pin_memory()
            dev = host.to(self.dev, non_blocking=True)
            res = self._ingest(len(idx), hw)(dev)
            for k, i in enumerate(idx):
                out[i] = res[k]
        return out


class ImageDataset(object):
    """``common/datasets/image_dataset.py`` ``ImageDataset`` (:48-98) with the frame path on the device.

    ``paths``: a listing file (one path per line) or a list of paths; ``image_size``: (H, W) or None (no resize)."""

    def __init__(self, paths, root='.', image_size=None, image_mode='RGB', transform_probability=0, device=None, **kwargs):
        if image_mode != 'RGB':
            raise NotImplementedError("only image_mode='RGB' is on the device path")
        if transform_probability > 0:
            raise NotImplementedError("the random augmentations (imgaug / chainercv) are outside this path: transform_probability must be 0")
        self._paths = read_image_listing(paths) if isinstance(paths, str) else list(paths)
        self._root = root
        self.image_size = image_size
        self._frames = _DeviceFrames(image_size, device)

    def __len__(self):
        return len(self._paths)

    def _decode(self, i):
        return decode_frame(os.path.join(self._root, self._paths[i]))

    def get_batch(self, indices):
        """float32 (B, 3, oH, oW) CUDA tensor in [0, 1] (frames of one size after the resize; a list of (3,H,W) tensors when
        image_size is None and the files differ in size)."""
        indices = list(indices)
        return self.assemble_batch(indices, [self._decode(i) for i in indices])

    def assemble_batch(self, indices, decoded):
        """The batch of ``indices`` from their already decoded uint8 frames (the loader threads of ``MultithreadIterator``)."""
        imgs = self._frames(decoded)
        if len({tuple(t.shape) for t in imgs}) == 1:
            return torch.stack(imgs)
        return imgs

    def get_example(self, i):
        return self._frames([self._decode(i)])[0]

    __getitem__ = get_example


class LabeledImageDataset(ImageDataset):
    """``LabeledImageDataset`` (:101-182): ``pairs`` is the tab-separated listing (or a list of (path, labels)); a label whose
    length is a multiple of four is a set of boxes, checked (``check_for_bad_label``) and scaled with the frame."""

    def __init__(self, pairs, root='.', label_dtype=np.int32, image_size=None, image_mode='RGB', transform_probability=0,
                 return_dummy_scores=True, device=None):
        self._label_dtype = label_dtype
        self._pairs = read_labeled_listing(pairs, label_dtype) if isinstance(pairs, str) else list(pairs)
        self.return_dummy_scores = return_dummy_scores
        super().__init__([p for p, _ in self._pairs], root=root, image_size=image_size, image_mode=image_mode,
                         transform_probability=transform_probability, device=device)

    def shrink_dataset(self, new_size):
        self._pairs = self._pairs[:new_size]
        self._paths = self._paths[:new_size]

    @staticmethod
    def check_for_bad_label(label, image_size):
        """Boxes must lie inside the frame they were annotated on, give or take a tenth of its size (reference :143-149: an
        AssertionError with the sizes in its message, raised before the boxes are scaled)."""
        h, w = image_size
        low = np.array([-0.1 * h, -0.1 * w]), np.array([1.1 * h, 1.1 * w])
        inside = (label[:, 0:2] >= low[0]).all() and (label[:, 2:4] <= low[1]).all()
        assert inside, ("Label can not be scaled correctly are you sure you created the dataset correctly, and provided the "
                        "correct sizes? Image size: %s, label: %s" % (image_size, label))

    def _label(self, i, frame_hw):
        label = np.array(self._pairs[i][1], dtype=self._label_dtype)
        if len(label.shape) > 0 and len(label) % 4 == 0:
            label = np.reshape(label, (len(label) // 4, -1))
        if self.image_size is not None:
            if len(label.shape) > 1:
                self.check_for_bad_label(label, frame_hw)
                label = resize_bbox(label.astype(np.float32), frame_hw, self.image_size)
            label = label.astype(self._label_dtype)
        return label

    def get_batch(self, indices):
        """(frames (B,3,oH,oW) CUDA float32, list of label arrays[, zeros (B,1)])."""
        indices = list(indices)
        return self.assemble_batch(indices, [self._decode(i) for i in indices])

    def assemble_batch(self, indices, decoded):
        labels = [self._label(i, f.shape[:2]) for i, f in zip(indices, decoded)]
        imgs = self._frames(decoded)
        frames = torch.stack(imgs) if len({tuple(t.shape) for t in imgs}) == 1 else imgs
        if self.return_dummy_scores:
            return frames, labels, np.zeros((len(indices), 1))
        return frames, labels

    def get_example(self, i):
        frame = self._decode(i)
        label = self._label(i, frame.shape[:2])
        image = self._frames([frame])[0]
        if self.return_dummy_scores:
            return image, label, np.zeros((1,))
        return image, label

    __getitem__ = get_example


class MultithreadIterator(object):
    """``chainer.iterators.MultithreadIterator(dataset, batch_size, repeat=True, shuffle=True, n_threads=1)`` as the reference
    drives its datasets (train_sheep_localizer.py:113-116), for the datasets above: the files of a batch are decoded by
    ``n_threads`` loader threads (PIL releases the GIL while it inflates a PNG), the NEXT batch is being decoded while the
    current one is consumed, and ``next()`` returns the batch already assembled on the device -- what
    ``concat_examples(batch, device)`` would give the updater: ``dataset.get_batch``'s result.

    Epoch bookkeeping as chainer's iterators: ``epoch``, ``is_new_epoch``, ``epoch_detail``, ``previous_epoch_detail``,
    ``reset()``; the order is a fresh ``numpy.random.permutation`` per epoch when ``shuffle``; with ``repeat`` a batch that
    reaches the end of an epoch is filled up from the next epoch's order, without it the last batch is short and the next
    call raises ``StopIteration``."""

    def __init__(self, dataset, batch_size, repeat=True, shuffle=True, n_threads=1):
        from concurrent.futures import ThreadPoolExecutor
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self._repeat, self._shuffle = bool(repeat), bool(shuffle)
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(n_threads)))
        self.reset()

    def reset(self):
        self.current_position = 0
        self.epoch = 0
        self.is_new_epoch = False
        self._previous_epoch_detail = -1.0
        n = len(self.dataset)
        self._order = np.random.permutation(n) if self._shuffle else None
        self._pending = None
        self._prefetch()

    @property
    def epoch_detail(self):
        return self.epoch + self.current_position / len(self.dataset)

    @property
    def previous_epoch_detail(self):
        return None if self._previous_epoch_detail < 0 else self._previous_epoch_detail

    def _next_indices(self):
        """Indices of the next batch and the iterator state behind it (chainer's SerialIterator arithmetic)."""
        n = len(self.dataset)
        if not self._repeat and self.epoch > 0:
            return None
        i, i_end = self.current_position, self.current_position + self.batch_size
        order = self._order
        idx = list(range(i, min(i_end, n))) if order is None else [int(v) for v in order[i:i_end]]
        state = {"epoch": self.epoch, "is_new_epoch": False, "order": order, "position": i_end}
        if i_end >= n:
            state["epoch"] += 1
            state["is_new_epoch"] = True
            state["position"] = 0
            if self._repeat:
                rest = i_end - n
                if order is not None:
                    state["order"] = np.random.permutation(n)
                if rest > 0:
                    idx += list(range(rest)) if order is None else [int(v) for v in state["order"][:rest]]
                    state["position"] = rest
        return idx, state

    def _prefetch(self):
        nxt = self._next_indices()
        if nxt is None:
            self._pending = None
            return
        idx, state = nxt
        futures = [self._pool.submit(self.dataset._decode, i) for i in idx]
        self._pending = (idx, state, futures)

    def __iter__(self):
        return self

    def __next__(self):
        if self._pending is None:
            raise StopIteration
        idx, state, futures = self._pending
        decoded = [f.result() for f in futures]
        self._previous_epoch_detail = self.epoch_detail
        self.epoch, self.is_new_epoch, self._order, self.current_position = state["epoch"], state["is_new_epoch"], state["order"], state["position"]
        self._prefetch()                                    # the loader threads start on the next batch now
        return self.dataset.assemble_batch(idx, decoded)

    next = __next__

    def finalize(self):
        self._pool.shutdown(wait=False)


__all__ = ["ImageDataset", "LabeledImageDataset", "MultithreadIterator", "read_image_listing", "read_labeled_listing", "decode_frame", "resize_bbox"]
