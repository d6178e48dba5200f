#!/bin/bash
# Builds the UNMODIFIED reference (/root/reference = INET-RC/GeoMX = MXNet 1.4.0 + modified KVStore/ps-lite) for sm_100 with CUDA 12.9 and
# installs it under baseline/_ref (git-ignored).  No reference source file is edited; everything that a 2019 code base needs to meet a 2025
# toolchain is done with build flags, a force-included prelude, an object-level fix-up and two extra translation units that #include the
# reference sources verbatim (all in baseline/refbuild/, each file documents itself).  ~45 min on 8 cores.
#
#   what is built      : libmxnet.so with USE_CUDA=1 USE_NCCL=1 USE_CUDNN=0 (the shipped GPU config has cuDNN off, make/gpu_config.mk:80)
#   what is left out   : USE_DIST_KVSTORE (ps-lite needs ZeroMQ 4.1.4 + protobuf 2.5.0, fetched by wget in 3rdparty/ps-lite/make/deps.mk:5-29 —
#                        no network), USE_OPENCV (no headers), USE_LAPACK (no liblapack) -> the in-process multi-GPU kvstores
#                        (local / device / nccl) are available, dist_sync is not.
#   BLAS               : baseline/refbuild/miniblas.c (a few CBLAS entry points, CPU paths only; the GPU path is cuBLAS)
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=${1:-/root/reference}
WORK=${WORK:-/tmp/refbuild}
DEPS=${DEPS:-/tmp/refdeps}
JOBS=${JOBS:-8}
rm -rf "$WORK"; mkdir -p "$WORK" "$DEPS/include" "$DEPS/lib"
cp -r "$SRC"/. "$WORK"/                      # /root/reference is read-only and make writes into the tree
cp "$HERE/refbuild/cblas.h" "$HERE/refbuild/cuda12_prelude.h" "$DEPS/include/"
gcc -O3 -fPIC -shared -Wl,-soname,libopenblas.so -o "$DEPS/lib/libopenblas.so" "$HERE/refbuild/miniblas.c" -I"$HERE/refbuild"
cat > "$WORK/config.mk" <<CFG
export CC = gcc
export CXX = g++
export NVCC = /usr/local/cuda/bin/nvcc
DEV = 0
DEBUG = 0
ADD_LDFLAGS = -L$DEPS/lib
ADD_CFLAGS = -I$DEPS/include
USE_CUDA = 1
USE_CUDA_PATH = /usr/local/cuda
ENABLE_CUDA_RTC = 1
USE_CUDNN = 0
USE_NCCL = 1
USE_NCCL_PATH = NONE
USE_OPENCV = 0
USE_LIBJPEG_TURBO = 0
USE_OPENMP = 1
USE_MKLDNN = 0
USE_NNPACK = 0
USE_BLAS = openblas
USE_LAPACK = 0
USE_INTEL_PATH = NONE
USE_STATIC_MKL = NONE
USE_SSE = 1
USE_DIST_KVSTORE = 0
USE_HDFS = 0
USE_S3 = 0
USE_OPERATOR_TUNING = 1
USE_GPERFTOOLS = 0
USE_JEMALLOC = 0
EXTRA_OPERATORS =
USE_CPP_PACKAGE = 0
CUDA_ARCH = -gencode arch=compute_100,code=sm_100
MXNET_PLUGINS =
NVCCFLAGS = -include $DEPS/include/cuda12_prelude.h
CFG
cd "$WORK"
# pass 1: everything that compiles as is (-k: keep going past the few files handled below)
make -k -j"$JOBS" lib/libmxnet.so > build.log 2>&1 || true
NVCC_LINE=$(make -n -W src/kvstore/kvstore_utils.cu build/src/kvstore/kvstore_utils_gpu.o 2>/dev/null | grep "nvcc -c" | head -1)
cu_cmd() { echo "$NVCC_LINE" | sed "s#-o build/src/kvstore/kvstore_utils_gpu.o#-o $2#; s#src/kvstore/kvstore_utils.cu#$1#"; }
# (a) three .cu files whose nvcc-generated host stub names cuda::std::plus<void> at global scope (ambiguous with mshadow::cuda): fix the STUB
for f in leaky_relu nn/dropout quantization/quantized_conv; do
  [ -f build/src/operator/${f}_gpu.o ] || eval "$HERE/refbuild/nvcc_stubfix.sh $(cu_cmd src/operator/${f}.cu build/src/operator/${f}_gpu.o)"
done
# (b) implicit template instantiations that gcc 13 / nvcc 12.9 no longer emit out of line: compat TUs that #include the reference files verbatim
mkdir -p compat && cp "$HERE"/refbuild/compat_elemwise_binary_op_basic.* compat/
CC_LINE=$(make -n -W src/operator/tensor/elemwise_binary_op_basic.cc build/src/operator/tensor/elemwise_binary_op_basic.o 2>/dev/null | grep "^g++" | head -1)
eval "$(echo "$CC_LINE" | sed 's#-MMD -c src/operator/tensor/elemwise_binary_op_basic.cc#-Isrc/operator/tensor -c compat/compat_elemwise_binary_op_basic.cc#')"
eval "$(cu_cmd compat/compat_elemwise_binary_op_basic.cu build/src/operator/tensor/elemwise_binary_op_basic_gpu.o) -Isrc/operator/tensor"
touch build/src/operator/tensor/elemwise_binary_op_basic.o build/src/operator/tensor/elemwise_binary_op_basic_gpu.o
make -j"$JOBS" lib/libmxnet.so > link.log 2>&1
# install: the reference's own python package through pip, then the library + BLAS shim next to it (xz: 270 MB -> 40 MB for the GPU-box snapshot)
cd "$HERE/.."
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref "$WORK/python"
cp "$DEPS/lib/libopenblas.so" baseline/_ref/mxnet/
cp "$WORK/lib/libmxnet.so" baseline/_ref/mxnet/ && strip --strip-unneeded baseline/_ref/mxnet/libmxnet.so
xz -T0 -2 -k -f baseline/_ref/mxnet/libmxnet.so
echo "reference installed under baseline/_ref (python bench.py --impl reference)"
