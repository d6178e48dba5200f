// Build-compat translation unit for the UNMODIFIED reference (compiled INSTEAD of src/operator/tensor/elemwise_binary_op_basic.cc, which it
// includes verbatim).  MXNet 1.4 relies on the compiler emitting out-of-line copies of a few implicitly instantiated member templates that
// other translation units reference (elemwise_binary_broadcast_op_basic.{cc,cu}).  gcc 13 at -O3 inlines them away, so the symbols the
// 2019 toolchains happened to produce are requested explicitly here.  No reference source line is changed.
#include "../src/operator/tensor/elemwise_binary_op_basic.cc"
namespace mxnet {
namespace op {
#define GX_INST_DNSCSRDNS(OP)                                                                                                          \
  template void ElemwiseBinaryOp::DnsCsrDnsOp<OP>(mshadow::Stream<cpu>*, const nnvm::NodeAttrs&, const OpContext&, const NDArray&,     \
                                                  const NDArray&, OpReqType, const NDArray&, const bool);
GX_INST_DNSCSRDNS(mshadow_op::plus)
GX_INST_DNSCSRDNS(mshadow_op::minus)
template void ElemwiseBinaryOp::DnsCsrCsrOp<cpu, mshadow_op::mul>(const nnvm::NodeAttrs&, const OpContext&, const NDArray&, const NDArray&,
                                                                  OpReqType, const NDArray&, const bool);
}  // namespace op
}  // namespace mxnet
