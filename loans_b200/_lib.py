"""ctypes loader of libloans_stn.so (the C ABI in include/loans_stn.h).

The library is built in-tree by ``loans_b200.build.build()`` (``__graft_entry__.build()`` calls it).  There is
deliberately no fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libloans_stn.so")
ABI_VERSION = 2
F32, BF16 = 0, 1
FLAG_GRAY, FLAG_UPRIGHT, FLAG_NHWC4 = 1, 2, 4
CFG_FORCE_GENERAL = 1
CFG_BAND_BACKWARD = 3
CFG_PDL = 8
# include/loans_stn_devel.h: test hooks (always there) ...
CFG_BAND_CS, CFG_BAND_ROWS, CFG_BAND_TILE_KB, CFG_BAND_VARIANT = 4, 5, 6, 7
# ... and A/B switches of a -DSTN_DEVEL build (the product build answers them with an error)
CFG_TMA_FORWARD = 2
CFG_KFRAME_ROWS = 13
CFG_THETA_FIRST = 9

_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_fl = ctypes.c_float

# name -> argtypes, exactly include/loans_stn.h
SIGNATURES = {
    "loans_stn_abi_version": [],
    "loans_stn_last_error": [],
    "loans_stn_launch_count": [],
    "loans_stn_last_kernel": [],
    "loans_stn_configure": [_i, _i],
    "loans_stn_rotation_dropout": [_vp, _fl, _vp, _i, _vp],
    "loans_stn_prepare_images": [_vp, _fl, _vp, _i, _i, _i, _i, _vp],
    "loans_stn_ingest_workspace_bytes": [_i, _i, _i, _i, _i],
    "loans_stn_ingest_prepare": [_vp, _i, _i, _i, _i, _vp],
    "loans_stn_ingest_u8": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "loans_stn_grid_fwd": [_vp, _vp, _i, _i, _i, _vp],
    "loans_stn_grid_bwd": [_vp, _vp, _i, _i, _i, _vp],
    "loans_stn_sampler_fwd": [_vp, _vp, _vp] + [_i] * 8 + [_vp],
    "loans_stn_sampler_bwd": [_vp, _vp, _vp, _vp, _vp] + [_i] * 8 + [_vp],
    "loans_stn_crop_fwd": [_vp, _vp, _fl, _vp, _vp] + [_i] * 8 + [_vp],
    "loans_stn_crop_bwd": [_vp, _vp, _fl, _vp, _vp, _vp, _vp, _vp] + [_i] * 8 + [_vp],
    "loans_stn_crop_fwd_ex": [_vp, _vp, _fl, _vp, _vp, _vp] + [_i] * 9 + [_vp],
    "loans_stn_crop_bwd_ex": [_vp, _vp, _fl, _vp, _vp, _vp, _vp, _vp, _vp] + [_i] * 9 + [_vp],
    "loans_stn_crop_fwd_corners": [_vp, _vp, _fl, _vp, _vp] + [_i] * 8 + [_vp],
    "loans_stn_crop_bwd_corners": [_vp, _vp, _fl, _vp, _vp, _vp, _vp] + [_i] * 8 + [_vp],
}


# include/loans_stn_devel.h
DEVEL_SIGNATURES = {
    "loans_stn_probe": [_i, _vp, _vp, ctypes.c_longlong, ctypes.c_longlong, _i, _vp],
}


class StnLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StnLibraryError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). loans_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in list(SIGNATURES.items()) + list(DEVEL_SIGNATURES.items()):
            fn = getattr(handle, name)           # AttributeError if the ABI lost a symbol
            fn.argtypes = argtypes
            fn.restype = _i
        handle.loans_stn_last_error.restype = ctypes.c_char_p
        handle.loans_stn_launch_count.restype = ctypes.c_ulonglong
        handle.loans_stn_last_kernel.restype = ctypes.c_char_p
        handle.loans_stn_ingest_workspace_bytes.restype = ctypes.c_longlong
        if handle.loans_stn_abi_version() != ABI_VERSION:
            raise StnLibraryError("libloans_stn.so ABI %d != expected %d" % (handle.loans_stn_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        raise StnLibraryError("%s failed: %s" % (what, lib().loans_stn_last_error().decode()))


def force_general(on):
    """Tests / A-B runs: never take the axis-aligned kernels (results are bit-identical either way)."""
    check(lib().loans_stn_configure(CFG_FORCE_GENERAL, int(bool(on))), "loans_stn_configure")


def tma_forward(on):
    """-DSTN_DEVEL builds only: the TMA-staged forward kernel for axis-aligned crops (mask01 == 0)."""
    check(lib().loans_stn_configure(CFG_TMA_FORWARD, int(bool(on))), "loans_stn_configure")


def pdl(on):
    """Programmatic dependent launch of the fused kernels (default on)."""
    check(lib().loans_stn_configure(CFG_PDL, int(bool(on))), "loans_stn_configure")


def theta_first(on):
    check(lib().loans_stn_configure(CFG_THETA_FIRST, int(bool(on))), "loans_stn_configure")


def band_backward(on):
    """Backward of axis-aligned crops through the band kernel: True = whenever it applies, False = never (always the
    general kernel), None = the library default (by shape: wide frame rows)."""
    check(lib().loans_stn_configure(CFG_BAND_BACKWARD, -1 if on is None else int(bool(on))), "loans_stn_configure")


def band_tuning(cs=0, rows=0, tile_kb=0, variant=0):
    """A/B knobs of the band kernel (0 = automatic): CTAs per crop, crop rows per band, tile budget, kernel variant."""
    for key, val in ((CFG_BAND_CS, cs), (CFG_BAND_ROWS, rows), (CFG_BAND_TILE_KB, tile_kb), (CFG_BAND_VARIANT, variant)):
        check(lib().loans_stn_configure(key, int(val)), "loans_stn_configure")


def kframe_rows(rows=0):
    """Frame rows per CTA of the several-crops-per-frame gx kernel (0 = automatic)."""
    check(lib().loans_stn_configure(CFG_KFRAME_ROWS, int(rows)), "loans_stn_configure")


def launch_count():
    return int(lib().loans_stn_launch_count())


def last_kernel():
    """'+'-separated names of the kernels this thread's last compute call launched (which one a dispatch rule picked)."""
    return lib().loans_stn_last_kernel().decode()
