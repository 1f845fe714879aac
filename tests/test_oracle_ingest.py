"""CPU tests: the ingest oracle (oracle/ingest_numpy.py: Pillow's 8-bit Lanczos resampling restated + `/ 255`) against the
fixtures the real PIL produced (tests/golden/ingest_lanczos.npz, made by tests/golden/make_golden.py from the reference's
own statement sequence, common/datasets/image_dataset.py:16-28,98) and, where Pillow is importable, against PIL live."""
import os

import numpy as np
import pytest

from oracle import ingest_numpy as ig

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ingest_lanczos.npz")


def test_oracle_reproduces_the_pil_fixtures_bit_for_bit():
    g = np.load(GOLDEN)
    assert int(g["n_cases"]) >= 8
    for i in range(int(g["n_cases"])):
        p = "c%02d_" % i
        out = ig.ingest(g[p + "frames"], tuple(g[p + "size"]))
        assert out.dtype == np.float32 and np.array_equal(out, g[p + "out_u8"].astype(np.float32) / 255), i


def test_oracle_against_live_pil():
    pytest.importorskip("PIL")
    rng = np.random.default_rng(5)
    for (h, w, oh, ow) in [(33, 47, 17, 29), (20, 30, 75, 75), (224, 224, 224, 224), (90, 64, 45, 64), (7, 7, 1, 1), (2, 3, 9, 8)]:
        f = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        assert np.array_equal(ig.ingest(f, (oh, ow)), ig.pil_reference(f, (oh, ow))), (h, w, oh, ow)


def test_coefficients_are_normalised_fixed_point():
    for n_in, n_out in [(512, 224), (224, 512), (100, 100), (7, 3)]:
        bounds, kk = ig.lanczos_coeffs(n_in, n_out)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all()
        sums = kk.sum(axis=1)
        assert np.abs(sums - (1 << ig.PRECISION_BITS)).max() <= kk.shape[1]          # rounding of each tap only
    # a constant image stays constant (the taps sum to one up to that rounding: never more than one grey level off)
    f = np.full((1, 40, 50, 3), 200, np.uint8)
    out = ig.ingest(f, (13, 21)) * 255
    assert np.abs(out - 200).max() <= 1


def test_no_resampling_is_a_plain_division():
    f = np.arange(2 * 4 * 5 * 3, dtype=np.uint8).reshape(2, 4, 5, 3)
    out = ig.ingest(f)
    assert np.array_equal(out, f.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255))
