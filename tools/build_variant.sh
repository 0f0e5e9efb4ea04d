#!/bin/bash
# usage: tools/build_variant.sh <git-commit> <name>   ->  ml-cvnets_b200/csrc/ab/<name>.so  (kernel library as of that commit; A/B timing on one box
# with CVB_LIB=ml-cvnets_b200/csrc/ab/<name>.so python tools/microbench.py ... / python bench.py ...; the ABI version must match the checkout)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
git -C "$ROOT" archive "$1" include ml-cvnets_b200/csrc | tar -x -C "$T"
python "$T/ml-cvnets_b200/csrc/build.py" > "$T/build.log" 2>&1 || { tail -20 "$T/build.log"; exit 1; }
mkdir -p "$ROOT/ml-cvnets_b200/csrc/ab"
cp "$T/ml-cvnets_b200/csrc/libcvnets_b200.so" "$ROOT/ml-cvnets_b200/csrc/ab/$2.so"
rm -rf "$T"
echo "$ROOT/ml-cvnets_b200/csrc/ab/$2.so"
