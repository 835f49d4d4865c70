#!/bin/bash
# builds lib/variants/libtvm_b200_<name>.so for A/B runs of the generated AIR kernels (TVM_B200_LIB selects one at run time)
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
build chunks100 TVM_AIR_FUSED=0 TVM_AIR_BUDGET=100
build chunks160 TVM_AIR_FUSED=0 TVM_AIR_BUDGET=160
build groups220 TVM_AIR_FUSED=1 TVM_AIR_BUDGET=100 TVM_AIR_GROUP_BUDGET=220
