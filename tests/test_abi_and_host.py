"""CPU tests: the C-ABI library loads and exports every symbol include/loans_stn.h declares (no compute calls
without a GPU), argument validation answers before any launch, host-side helpers behave."""
import os
import re

import numpy as np
import pytest

import loans_b200
from loans_b200 import _lib
from loans_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        from loans_b200.build import build
        build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "loans_stn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(loans_stn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 10
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.loans_stn_abi_version() == _lib.ABI_VERSION
    m = re.search(r"#define LOANS_STN_ABI_VERSION (\d+)", hdr)
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_validation_happens_before_any_launch(L):
    n0 = _lib.launch_count()
    assert L.loans_stn_crop_fwd(None, None, 1.0, None, None, 1, 1, 3, 8, 8, 4, 4, 0, None) != 0
    assert b"NULL" in L.loans_stn_last_error()
    assert L.loans_stn_crop_fwd(1, 1, 1.0, 1, None, 3, 2, 3, 8, 8, 4, 4, 0, None) != 0
    assert b"multiple" in L.loans_stn_last_error()
    assert L.loans_stn_crop_fwd(1, 1, 1.0, 1, None, 2, 1, 3, 8, 8, 0, 4, 0, None) != 0
    assert L.loans_stn_crop_fwd(1, 1, 1.0, 1, None, 2, 1, 3, 8, 8, 4, 4, 7, None) != 0
    assert b"dtype" in L.loans_stn_last_error()
    assert L.loans_stn_crop_bwd(1, 1, 1.0, 1, None, None, None, None, 2, 1, 3, 8, 8, 4, 4, 0, None) != 0
    assert L.loans_stn_grid_fwd(1, 1, -1, 4, 4, None) != 0
    # empty batch is a successful no-op
    assert L.loans_stn_crop_fwd(None, None, 1.0, None, None, 0, 1, 3, 8, 8, 4, 4, 0, None) == 0
    assert _lib.launch_count() == n0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    L_ = _lib.lib()
    assert L_.loans_stn_crop_fwd(1, 1, 1.0, 1, None, 1, 1, 3, 8, 8, 4, 4, 0, None) != 0
    assert b"no CPU fallback" in L_.loans_stn_last_error()
    from loans_b200.functions import stn_crop
    with pytest.raises(RuntimeError):
        stn_crop(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 3), (4, 4))


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ("import sys; import loans_b200, loans_b200.functions, loans_b200.parallel, loans_b200.workloads; "
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; assert not bad, bad")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "loans_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "stn_oracle" not in src or f == "stn_math.cuh", f


def test_config_switch():
    assert loans_b200.config.train is True
    with loans_b200.using_config("train", False):
        assert loans_b200.config.train is False
        from loans_b200.functions.rotation_droput import draw_mask_value
        assert draw_mask_value(0.25) == 0.25
    assert loans_b200.config.train is True
    from loans_b200.functions.rotation_droput import draw_mask_value
    np.random.seed(0)
    draws = {draw_mask_value(0.5) for _ in range(32)}
    assert draws == {0.0, 1.0}
    assert draw_mask_value(0.0) == 0.0 and draw_mask_value(1.0) == 1.0
    with pytest.raises(AttributeError):
        with loans_b200.using_config("nope", 1):
            pass


def test_algorithmic_bytes_match_baseline_md():
    # BASELINE.md table, bytes per crop with / without gx
    table = {"cfg1": (1322160, 720048), "cfg2": (1126448, 524336), "cfg3": (3798276, 652548), "cfg4": (916656, 720048),
             "cfg5": (1322160, 720048)}
    for name, (with_gx, no_gx) in table.items():
        wl = W.WORKLOADS[name]
        n = wl.batch * wl.crops_per_frame
        f, b = W.algorithmic_bytes(wl, need_gx=True)
        assert (f + b) // n == with_gx, name
        f, b = W.algorithmic_bytes(wl, need_gx=False)
        assert (f + b) // n == no_gx, name


def test_workload_generator_is_seeded_and_shaped():
    wl = W.WORKLOADS["cfg4"]
    a = W.make_inputs(wl, batch=2)
    b = W.make_inputs(wl, batch=2)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["x"].shape == (2, 3, 512, 512) and a["theta"].shape == (32, 2, 3) and a["gy"].shape == (32, 3, 75, 75)
    assert np.all(a["theta"][:, 0, 1] == 0)                  # rotation_ratio 0.0 workloads are generated axis-aligned
    c = W.make_inputs(W.WORKLOADS["cfg2"], batch=4)
    assert np.any(c["theta"][:, 0, 1] != 0)                  # cfg2 has no dropout node: general affine


def test_chainer_binding_is_import_guarded():
    # chainer/cupy are absent here: the binding must import cleanly and say so when asked to install
    from loans_b200 import chainer_compat
    if chainer_compat.HAVE_CHAINER:
        pytest.skip("chainer present")
    with pytest.raises(ImportError):
        chainer_compat.install()
    src = open(chainer_compat.__file__).read()
    for name in ("loans_stn_rotation_dropout", "loans_stn_grid_fwd", "loans_stn_grid_bwd", "loans_stn_sampler_fwd",
                 "loans_stn_sampler_bwd", "loans_stn_crop_fwd", "loans_stn_crop_bwd"):
        assert name in src and name in _lib.SIGNATURES
