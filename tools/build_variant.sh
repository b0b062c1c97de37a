# Build a variant of the library for A/B runs on the GPU box:
#   tools/build_variant.sh <name> [-DRB_RT_CTAS=3 ...]   ->  radiobear_b200/lib/librb_<name>.so   (RB_LIB_PATH selects it)
set -e
name=$1; shift
cd "$(dirname "$0")/../radiobear_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o ../lib/librb_$name.so alpha_kernels.cu rt_kernels.cu probe_kernels.cu capi.cu
