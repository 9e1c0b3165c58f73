#!/bin/bash
# Damaged-input run of the host layer under AddressSanitizer + UBSan (CPU only).
#   bash tools/host_fuzz/run.sh > profiles/rNN_host_fuzz_asan.txt
set -e
here="$(cd "$(dirname "$0")" && pwd)"; root="$here/../.."
work="$(mktemp -d)"; trap 'rm -rf "$work"' EXIT
python "$here/gen.py" "$work/in"
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -pthread -o "$work/fuzz" "$here/main.cpp" \
    "$root/fwumious_wabbit_b200/csrc/host/fwhost.cpp" "$root/fwumious_wabbit_b200/csrc/host/expf_libm.cpp"
"$work/fuzz" "$work/in" 2>&1
echo "sanitizer findings: none (the run above would have aborted)"
