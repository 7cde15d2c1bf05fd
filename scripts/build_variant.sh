#!/bin/bash
# build_variant.sh <name> : compile the CURRENT csrc tree into variants/lib<name>.so (for A/B runs on the GPU box;
# select with BSSM_LIB_PATH=variants/lib<name>.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
python -m bayesssm_b200.build >/dev/null
cp bayesssm_b200/libbayesssm_b200.so variants/lib$1.so
echo built variants/lib$1.so
