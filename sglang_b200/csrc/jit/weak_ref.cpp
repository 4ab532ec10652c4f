// weak_ref_tensor: a non-owning alias of a tensor's storage (same data pointer, sizes, strides).  CUDA-graph buffers are kept
// alive by the graph's private pool; Python-side references held by output lists would otherwise pin the tensors (and their
// allocator blocks) across captures.  Reference: python/sglang/kernels/aot/csrc/memory/weak_ref_tensor.cpp.
#include <torch/library.h>
#include <torch/types.h>

namespace {
at::Tensor weak_ref_tensor(const at::Tensor& t) {
  if (!t.defined() || t.numel() == 0) return t;
  return at::from_blob(t.data_ptr(), t.sizes(), t.strides(), [](void*) {}, t.options());
}
}  // namespace

TORCH_LIBRARY(sglb_jit_weak_ref, m) { m.def("weak_ref_tensor(Tensor t) -> Tensor", &weak_ref_tensor); }
