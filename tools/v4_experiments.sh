#!/bin/bash
# timing experiments on the version-4 residual-layer kernel (GPU box): rebuilds the library with experiment switches
for defs in "WAE_V4_STAGES=3" "WAE_V4_NOXRES=1" ""; do
  echo "== defs: '$defs'"
  WAE_NVCC_DEFS="$defs" python -m wavenet_autoencoders_b200.build --force > /dev/null 2>&1
  python tools/layer_cluster_sweep.py -4 2>&1 | grep cluster=
done
