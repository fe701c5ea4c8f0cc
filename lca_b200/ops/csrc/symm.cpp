// Symmetric heap primitives.
//
// Every rank cudaMalloc's an identically sized slab, exports it with cudaIpcGetMemHandle, the
// handles travel through torch.distributed's store (python side, all_gather_object), and each
// rank maps every peer's slab with cudaIpcOpenMemHandle.  From then on kernels address peer HBM
// with plain ld/st/red over NVLink 5 / NVSwitch; there is no NCCL call on the fused paths.
// The reference has no equivalent (it only ever calls torch.distributed collectives,
// SURVEY.md 2.3); this replaces its NCCL communicators for intra-node groups.
#include "symm.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include <string>

#include "launchers.h"

namespace lca {

#define SYMM_OK(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    TORCH_CHECK(_e == cudaSuccess, #expr, " failed: ", cudaGetErrorString(_e));           \
  } while (0)

static int64_t symm_alloc(int64_t bytes, int64_t device) {
  c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(device));
  void* p = nullptr;
  const cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();      // out-of-memory is not sticky: clear it so the caller can fall back
    TORCH_CHECK(false, "symm_alloc: cudaMalloc(", bytes, " bytes) failed: ", cudaGetErrorString(e));
  }
  SYMM_OK(cudaMemset(p, 0, static_cast<size_t>(bytes)));
  SYMM_OK(cudaDeviceSynchronize());
  return reinterpret_cast<int64_t>(p);
}

static void symm_free(int64_t ptr) { cudaFree(reinterpret_cast<void*>(ptr)); }

static pybind11::bytes symm_export(int64_t ptr) {
  cudaIpcMemHandle_t h;
  SYMM_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(ptr)));
  return pybind11::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
}

static int64_t symm_import(const std::string& handle, int64_t device) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  SYMM_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return reinterpret_cast<int64_t>(p);
}

static void symm_unmap(int64_t ptr) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)); }

static at::Tensor symm_tensor(int64_t ptr, std::vector<int64_t> shape, at::ScalarType dtype, int64_t device) {
  auto opts = at::TensorOptions().dtype(dtype).device(at::kCUDA, static_cast<c10::DeviceIndex>(device));
  return at::from_blob(reinterpret_cast<void*>(ptr), shape, [](void*) {}, opts);
}

static bool can_access_peer(int64_t dev, int64_t peer) {
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, static_cast<int>(dev), static_cast<int>(peer)) != cudaSuccess) return false;
  return ok != 0;
}

void bind_symm(pybind11::module_& m) {
  m.def("symm_alloc", &symm_alloc, "cudaMalloc + zero a slab; returns the device pointer");
  m.def("symm_free", &symm_free);
  m.def("symm_export", &symm_export, "cudaIpcGetMemHandle -> bytes");
  m.def("symm_import", &symm_import, "cudaIpcOpenMemHandle(bytes) -> peer-mapped device pointer");
  m.def("symm_unmap", &symm_unmap);
  m.def("symm_tensor", &symm_tensor, "view raw device memory as a tensor (no ownership)");
  m.def("can_access_peer", &can_access_peer);
}

}  // namespace lca
