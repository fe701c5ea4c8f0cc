// Symmetric (peer-mapped) device memory over CUDA IPC: the substrate of the fused NVLink paths.
#pragma once
#include <torch/extension.h>

namespace lca {
void bind_symm(pybind11::module_& m);
}
