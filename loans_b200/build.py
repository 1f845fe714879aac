"""In-tree build of libloans_stn.so: nvcc -gencode arch=compute_100a,code=sm_100a (see csrc/Makefile)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False, verbose=False):
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, stdout=subprocess.DEVNULL)
    res = subprocess.run(["make", "-C", csrc, "-j4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libloans_stn.so failed")
    return os.path.join(_HERE, "libloans_stn.so")
