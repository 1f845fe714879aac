"""CPU tests of the host side of the device loader (loans_b200/datasets.py): the listing formats and the per-file decode
against the reference's own statement sequence executed with the real PIL (reference common/datasets/image_dataset.py:16-28,
:75-78, :104-110; chainer's ``_read_image_as_array`` restated).  The resize / `/ 255` part is the device's (GPU tests)."""
import os

import numpy as np
import pytest

from loans_b200 import datasets as ds


def _write_pngs(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(5)
    files = {}
    for name, mode, shape in (("rgb.png", "RGB", (20, 31, 3)), ("gray.png", "L", (17, 9)), ("rgba.png", "RGBA", (8, 12, 4)),
                              ("big.png", "RGB", (96, 128, 3))):
        arr = rng.integers(0, 256, shape, dtype=np.uint8)
        Image.fromarray(arr, mode).save(os.path.join(tmp_path, name))
        files[name] = arr
    return files


def _reference_decode(path):
    """chainer ImageDataset.get_example + the reference's tile + the front of resize_image, with the real PIL."""
    from PIL import Image
    with Image.open(path) as f:
        image = np.asarray(f, dtype=np.float32)
    if image.ndim == 2:
        image = image[:, :, np.newaxis]
    image = image.transpose(2, 0, 1)                                     # CHW, as chainer returns it
    if image.shape[0] == 1:
        image = np.tile(image, (3, 1, 1))                                # image_dataset.py:77-78
    pil_image = Image.fromarray(image.transpose(1, 2, 0).astype('uint8')).convert('RGB')      # :19-21
    return np.asarray(pil_image)


def test_decode_matches_the_reference_sequence(tmp_path):
    files = _write_pngs(str(tmp_path))
    for name in files:
        p = os.path.join(str(tmp_path), name)
        got = ds.decode_frame(p)
        assert got.dtype == np.uint8 and got.shape[2] == 3 and got.flags["C_CONTIGUOUS"]
        assert np.array_equal(got, _reference_decode(p)), name


def test_listings(tmp_path):
    il = os.path.join(str(tmp_path), "images.csv")
    with open(il, "w") as f:
        f.write("a/0001.png\n  b/0002.png  \n\n")
    assert ds.read_image_listing(il) == ["a/0001.png", "b/0002.png"]
    gl = os.path.join(str(tmp_path), "gt.csv")
    with open(gl, "w") as f:
        f.write("x.png\t10\t20\t110\t220\ny.png\t1\t2\t3\t4\t5\t6\t7\t8\nz.png\t3\n")
    pairs = ds.read_labeled_listing(gl)
    assert [p for p, _ in pairs] == ["x.png", "y.png", "z.png"]
    assert pairs[0][1] == [10, 20, 110, 220] and len(pairs[1][1]) == 8 and pairs[2][1] == [3]
    assert all(isinstance(v, np.int32) for v in pairs[0][1])


def test_resize_bbox_and_bad_labels():
    b = np.array([[10, 20, 110, 220], [0, 0, 384, 512]], np.int32)
    out = ds.resize_bbox(b.astype(np.float32), (384, 512), (224, 224))
    assert out.dtype == np.float32
    want = b.astype(np.float32).copy()
    want[:, [0, 2]] *= np.float32(224.0 / 384)
    want[:, [1, 3]] *= np.float32(224.0 / 512)
    assert np.allclose(out, want, rtol=1e-6)
    assert np.array_equal(out.astype(np.int32)[1], [0, 0, 224, 224])
    ds.LabeledImageDataset.check_for_bad_label(b, (384, 512))
    with pytest.raises(AssertionError):
        ds.LabeledImageDataset.check_for_bad_label(np.array([[0, 0, 500, 512]]), (384, 512))


def test_no_cpu_fallback_and_unsupported_modes(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        ds.ImageDataset(["a.png"], image_size=(8, 8))
    with pytest.raises(NotImplementedError):
        ds.ImageDataset(["a.png"], image_size=(8, 8), transform_probability=0.5)
    with pytest.raises(NotImplementedError):
        ds.ImageDataset(["a.png"], image_size=(8, 8), image_mode='L')


class _FakeDataset(object):
    """Stands in for the device datasets: the iterator only needs __len__, _decode and assemble_batch."""

    def __init__(self, n):
        self.n = n
        self.decoded = []

    def __len__(self):
        return self.n

    def _decode(self, i):
        self.decoded.append(i)
        return -i

    def assemble_batch(self, indices, decoded):
        assert decoded == [-i for i in indices]
        return list(indices)


def test_multithread_iterator_epoch_semantics():
    """chainer.iterators.MultithreadIterator as the reference uses it (train_sheep_localizer.py:113-116): order, epoch counters,
    the wrap of a repeating iterator into the next epoch's order, the short last batch and StopIteration without repeat."""
    it = ds.MultithreadIterator(_FakeDataset(7), 3, repeat=False, shuffle=False, n_threads=3)
    seen = [(b, it.epoch, it.is_new_epoch, round(it.epoch_detail, 4)) for b in it]
    assert seen == [([0, 1, 2], 0, False, 0.4286), ([3, 4, 5], 0, False, 0.8571), ([6], 1, True, 1.0)]
    with pytest.raises(StopIteration):
        next(it)
    it.reset()
    assert next(it) == [0, 1, 2] and it.previous_epoch_detail == 0.0
    np.random.seed(4)
    it = ds.MultithreadIterator(_FakeDataset(5), 2, repeat=True, shuffle=True, n_threads=2)
    batches = [next(it) for _ in range(5)]                                  # two epochs of five
    flat = [i for b in batches for i in b]
    assert sorted(flat[:5]) == [0, 1, 2, 3, 4] and sorted(flat[5:]) == [0, 1, 2, 3, 4]
    assert it.epoch == 2 and it.is_new_epoch and it.current_position == 0
    it = ds.MultithreadIterator(_FakeDataset(4), 3, repeat=True, shuffle=False)
    assert [next(it) for _ in range(3)] == [[0, 1, 2], [3, 0, 1], [2, 3, 0]]
    assert it.epoch == 2 and it.current_position == 1
