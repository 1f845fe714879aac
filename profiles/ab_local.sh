#!/bin/bash
# Build a variant of the library HERE (nvcc cross-compiles sm_100a without a GPU) into ab_builds/<name>.so, from the working
# tree or from a git revision, with extra nvcc flags; the .so files travel to the GPU box with the tree and are timed
# against each other, interleaved on one GPU, by profiles/so_ab.py.
# usage: profiles/ab_local.sh <name> "<extra nvcc flags>" [git-rev]
set -e
name="$1"; extra="$2"; rev="$3"
root="$(cd "$(dirname "$0")/.." && pwd)"
tmp="$(mktemp -d)"
mkdir -p "$tmp/loans_b200/csrc" "$tmp/include" "$root/ab_builds"
if [ -n "$rev" ]; then
  git -C "$root" archive "$rev" loans_b200/csrc include | tar -x -C "$tmp"
else
  cp "$root"/loans_b200/csrc/*.cu "$root"/loans_b200/csrc/*.cuh "$root"/loans_b200/csrc/Makefile "$tmp/loans_b200/csrc/"
  cp "$root"/include/*.h "$tmp/include/"
fi
make -C "$tmp/loans_b200/csrc" -j8 EXTRA="$extra" >/dev/null 2>"$tmp/err.log" || { tail -20 "$tmp/err.log"; exit 1; }
grep -E "spill" "$tmp/err.log" | sort | uniq -c | sort -rn | head -3 || true
cp "$tmp/loans_b200/libloans_stn.so" "$root/ab_builds/$name.so"
rm -rf "$tmp"
echo "built ab_builds/$name.so"
