// Force-included (nvcc -include) in front of every .cu of the UNMODIFIED reference when it is built with CUDA 12.9.
// Two toolchain incompatibilities of MXNet 1.4 (2019) are resolved here with build flags only, no source change:
//  1. CUDA >= 11 ships libcu++ whose namespace is ::cuda.  mshadow has mshadow::cuda and the operator sources say
//     `using namespace mshadow;` and then `cuda::Reduce1D<...>` (src/operator/nn/softmax-inl.h:180) or include <cub/cub.cuh> after a
//     global using-directive, so `cuda::` becomes ambiguous.  Parsing CUB first (before any using-directive exists) and giving the
//     namespace mxnet::op::mxnet_op its own alias `cuda = mshadow::cuda` (found by unqualified lookup before the enclosing scopes) restores
//     the meaning the code had in 2019.
//  2. The bundled 3rdparty/cub (1.8) must not shadow the toolkit's CUB that thrust 2.x requires; nvcc's own include path wins.
#pragma once
#include <cub/cub.cuh>
namespace mshadow { namespace cuda {} }
namespace mxnet { namespace op { namespace mxnet_op { namespace cuda = ::mshadow::cuda; } } }
