// C++ entry points implemented in the .cu files (kernels are compiled without torch headers).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fmha_params.h"

namespace lca {

// fmha_fwd_sm100.cu
cudaError_t launch_fmha_fwd(const FwdParams& p, int head_dim, bool bf16, int num_sms, cudaStream_t stream);


// fmha_fwd_fp8_sm100.cu (experimental)
cudaError_t launch_fmha_fwd_fp8(const FwdParams& p, int head_dim, int num_sms, cudaStream_t stream);

// fmha_bwd_sm100.cu
cudaError_t launch_fmha_bwd(const BwdParams& p, int head_dim, bool bf16, bool is_dkv, int num_sms, cudaStream_t stream);

// util_kernels.cu
cudaError_t launch_merge_out_lse(float* out_acc, float* lse_acc, const void* block_out, int block_dtype /*0 f32,1 bf16,2 f16*/,
                                 const float* block_lse, int B, int S, int H, int D, cudaStream_t stream);
cudaError_t launch_finalize_out(const float* out_acc, void* out, int out_dtype, int64_t n, cudaStream_t stream);
cudaError_t launch_flatten_lse(const float* lse_padded, float* lse_flat, const int* cu_seqlens, int B, int H, int max_s,
                               int total, cudaStream_t stream);
cudaError_t launch_unflatten_lse(const float* lse_flat, float* lse_padded, const int* cu_seqlens, int B, int H, int max_s,
                                 int total, cudaStream_t stream);
// (B, S, G, Hl, D) <-> (G, B, S, Hl, D)-style head/sequence permutes used around the NCCL all-to-all
cudaError_t launch_permute_heads_out(const void* src, void* dst, int B, int S, int G, int HlD_bytes, cudaStream_t stream);
cudaError_t launch_permute_heads_in(const void* src, void* dst, int B, int S, int G, int HlD_bytes, cudaStream_t stream);
cudaError_t launch_delta(const void* out, const void* dout, int dtype, float* delta, const float* lse, float* lse2,
                         int B, int S, int H, int D,
                         int64_t o_sb, int64_t o_ss, int64_t o_sh, int64_t do_sb, int64_t do_ss, int64_t do_sh,
                         cudaStream_t stream);

cudaError_t launch_quant_e4m3(const void* x, int dtype, uint8_t* y, float* scale, const float* ext_scale, int B, int S, int H,
                              int D, int64_t sb, int64_t ss, int64_t sh, cudaStream_t stream);
cudaError_t launch_wait_counter(const uint32_t* sig, uint32_t target, cudaStream_t stream);

// tensor-map helper (tma_host.cpp part of bindings): 4-D (D, H, S, B) 16-bit tensor, box (64,1,128,1), SWIZZLE_128B
bool encode_tmap_4d(CUtensorMap* out, const void* base, int64_t D, int64_t H, int64_t S, int64_t B, int64_t stride_h,
                    int64_t stride_s, int64_t stride_b, int box_rows, const char** err);

}  // namespace lca
