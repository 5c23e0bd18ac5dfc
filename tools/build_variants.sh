#!/bin/bash
# builds libflk.so variants for tools/ab_step.py:  tools/build_variants.sh name "-DFLAG=.. -DFLAG2=.." [name flags ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
while [ $# -gt 1 ]; do
  FLK_LIB_OUT=$PWD/build_variants/libflk_$1.so FLK_NVCC_EXTRA="$2" python -c "from examples_b200 import build; build.build(force=True)"
  echo "built build_variants/libflk_$1.so ($2)"
  shift 2
done
