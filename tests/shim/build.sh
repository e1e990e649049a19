#!/bin/sh
# Build the shim test driver against the reference headers (only where the reference tree exists).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${DEEPMD_SOURCE_DIR:-/root/reference}
CUDA=${CUDA_HOME:-/usr/local/cuda}
[ -d "$REF/source/lib/include" ] || { echo "no reference tree"; exit 0; }
mkdir -p "$HERE/_build"
g++ -O2 -std=c++17 -DGOOGLE_CUDA=1 -I "$REF/source/lib/include" -I "$CUDA/include" "$HERE/shim_driver.cc" \
    -o "$HERE/_build/shim_driver" -L "$ROOT/deepmd-kit_b200/lib" -ldeepmd_op_cuda -ldpb200 \
    -L "$ROOT/oracle/_ref" -ldeepmd_ref -L "$CUDA/lib64" -lcudart \
    -Wl,-rpath,"$ROOT/deepmd-kit_b200/lib" -Wl,-rpath,"$ROOT/oracle/_ref" -Wl,-rpath,"$CUDA/lib64"
echo built "$HERE/_build/shim_driver"
