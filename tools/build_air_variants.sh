#!/bin/bash
# builds lib/variants/libtvm_b200_<name>.so for A/B runs of the generated AIR kernels (TVM_B200_LIB selects one at run time);
# restores the default generation at the end
set -e
cd "$(dirname "$0")/../triton-vm_b200"
mkdir -p lib/variants
build() {  # name, env...
  name=$1; shift
  env "$@" python -m airgen.codegen_cuda > /dev/null
  make -j8 > /dev/null 2>&1
  cp lib/libtvm_b200.so lib/variants/libtvm_b200_$name.so
  echo "built $name: $(head -2 csrc/air_gen/air_chunks.inc | tail -1)"
}
for v in "$@"; do
  case $v in
    chunks160) build chunks160 TVM_AIR_FUSED=0 TVM_AIR_BUDGET=160 ;;
    groups220) build groups220 TVM_AIR_FUSED=1 TVM_AIR_GROUP_BUDGET=220 ;;
    minb2) build minb2 TVM_AIR_MIN_BLOCKS=2 ;;
    minb3) build minb3 TVM_AIR_MIN_BLOCKS=3 ;;
    budget60) build budget60 TVM_AIR_BUDGET=60 ;;
    split400) build split400 TVM_AIR_SPLIT=1 TVM_AIR_SPLIT_BUDGET=400 ;;
    split250) build split250 TVM_AIR_SPLIT=1 TVM_AIR_SPLIT_BUDGET=250 ;;
  esac
done
python -m airgen.codegen_cuda > /dev/null && make -j8 > /dev/null 2>&1 && echo "default restored: $(head -2 csrc/air_gen/air_chunks.inc | tail -1)"
