"""GPU test: loans_b200/chainer_compat.py EXECUTED -- install() plus the reference's three calls as the reference writes
them -- against the chainer / cupy stand-in of tests/chainer_standin (own process, see tests/chainer_binding_harness.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chainer_binding_runs_the_reference_calls():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "chainer_binding_harness.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    r = json.loads(res.stdout.strip().splitlines()[-1])
    for mode in ("full", "sampler", "off"):
        for key in (mode, mode + "_gx"):
            m = r[key]
            assert m["rois_exact"] and m["points_exact"], (key, m)              # bit-identical to the oracle in every mode
            assert m["gtheta_rel"] <= 1e-4, (key, m)
            if key.endswith("_gx"):
                assert m["gx_rel"] <= (1e-5 if mode == "off" else 2e-6), (key, m)   # "off": explicit-grid sampler, float atomics
    # launches: "full" = dropout + ONE fused kernel forward, ONE fused kernel backward (+ the grid node's and the dropout's
    # own tiny backward only when something sent a gradient to points)
    assert r["full"]["fwd_launches"] == 2 and r["full"]["bwd_launches_no_points_grad"] == 1 and r["full"]["bwd_launches"] == 3
    assert r["sampler"]["fwd_launches"] == 3 and r["off"]["fwd_launches"] == 3
    # the fast kernels are reached through the reference's own three names
    assert r["full"]["bwd_kernel"] == "stn_bwd_theta_tab_kernel" and r["sampler"]["bwd_kernel"] in ("rotation_dropout_kernel",)
    assert r["full_gx"]["bwd_kernel"] in ("stn_bwd_kernel", "stn_bwd_band_kernel/row", "stn_bwd_band_kernel/cta")
    assert r["test_mode_exact"] and r["test_mode_backward_raises"]
    assert r["float64_theta_refused"] and r["numpy_frames_refused"] and r["prepare_images_exact"]
