// Build-compat translation unit (see compat_elemwise_binary_op_basic.cc): compiled INSTEAD of src/operator/tensor/elemwise_binary_op_basic.cu,
// which it includes verbatim; adds the explicit instantiations other objects link against.  nvcc 12.9 additionally gives implicitly
// instantiated host templates that launch kernels internal linkage, explicit instantiations keep external linkage.
#include "../src/operator/tensor/elemwise_binary_op_basic.cu"
namespace mxnet {
namespace op {
#define GX_INST_DNSCSRDNS_GPU(OP)                                                                                                      \
  template void ElemwiseBinaryOp::DnsCsrDnsOp<OP>(mshadow::Stream<gpu>*, const nnvm::NodeAttrs&, const OpContext&, const NDArray&,     \
                                                  const NDArray&, OpReqType, const NDArray&, const bool);
GX_INST_DNSCSRDNS_GPU(mshadow_op::plus)
GX_INST_DNSCSRDNS_GPU(mshadow_op::minus)
template void ElemwiseBinaryOp::DnsCsrCsrOp<gpu, mshadow_op::mul>(const nnvm::NodeAttrs&, const OpContext&, const NDArray&, const NDArray&,
                                                                  OpReqType, const NDArray&, const bool);
}  // namespace op
}  // namespace mxnet
