#!/bin/bash
# static SASS instruction counts per kernel of the built library (proxy for the instruction diet; no GPU needed)
cuobjdump -sass loans_b200/libloans_stn.so 2>/dev/null | awk '/Function :/ {name=$3} /^\s+\/\*[0-9a-f]{4}\*\// {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | sort -k2 | grep -E "${1:-.}"
