// Memory-bound helper kernels (sm_100a): online-softmax merge, fp32->16-bit finalize, varlen LSE
// flatten/unflatten, head<->sequence permutes around the NCCL all-to-all, dO.O row sums.
//
// Parity: yunchang/ring/utils.py:10-51 (_update_out_and_lse, TorchScript),
//         yunchang/ring/triton_utils.py:6-137 (Triton flatten/unflatten kernels),
//         the .contiguous() staging copies of yunchang/comm/all_to_all.py:45-49,62-65,78-100.
// All of them are pure streaming kernels: 16-byte vector accesses, grid-stride, no smem.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <cstdio>
#include <type_traits>

#include "launchers.h"

namespace lca {
namespace {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

// out_acc (B,S,H,D) fp32, lse_acc (B,H,S) fp32; one thread handles 4 consecutive d of one (b,s,h).
// The LSE is updated by the thread with d-group 0 AFTER all threads of the row have read it:
// rows are processed by D/4 consecutive threads of one warp (D/4 <= 32 and divides 32), so a
// __syncwarp between read and write is sufficient.
template <typename T>
__global__ void merge_kernel(float* __restrict__ out_acc, float* __restrict__ lse_acc,
                             const T* __restrict__ bout, const float* __restrict__ blse, int B, int S,
                             int H, int D) {
  const int dg = D >> 2;
  const int64_t nrows = static_cast<int64_t>(B) * S * H;
  const int64_t total = nrows * dg;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t start = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // all threads of a warp iterate the same number of times (total padded to warp multiple below)
  const int64_t total_pad = (total + 31) / 32 * 32;
  for (int64_t i = start; i < total_pad; i += step) {
    const bool live = i < total;
    const int64_t rowi = live ? i / dg : 0;
    const int g = live ? static_cast<int>(i - rowi * dg) : 1;
    const int h = static_cast<int>(rowi % H);
    const int64_t bs = rowi / H;
    const int s = static_cast<int>(bs % S);
    const int b = static_cast<int>(bs / S);
    const int64_t li = (static_cast<int64_t>(b) * H + h) * S + s;
    float w_old = 0.f, w_new = 0.f, lnew = 0.f;
    if (live) {
      const float la = lse_acc[li];
      const float lb = blse[li];
      const float mx = fmaxf(la, lb);
      if (mx == -INFINITY) {
        lnew = -INFINITY;
      } else {
        const float ea = __expf(la - mx), eb = __expf(lb - mx);
        const float sum = ea + eb;
        lnew = mx + __logf(sum);
        w_old = ea / sum;
        w_new = eb / sum;
      }
      float4 a = *reinterpret_cast<const float4*>(out_acc + rowi * D + g * 4);
      float bx, by, bz, bw;
      if constexpr (sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(bout + rowi * D + g * 4);
        bx = t.x; by = t.y; bz = t.z; bw = t.w;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(bout + rowi * D + g * 4);
        const T* e = reinterpret_cast<const T*>(&t);
        bx = to_f(e[0]); by = to_f(e[1]); bz = to_f(e[2]); bw = to_f(e[3]);
      }
      a.x = a.x * w_old + bx * w_new;
      a.y = a.y * w_old + by * w_new;
      a.z = a.z * w_old + bz * w_new;
      a.w = a.w * w_old + bw * w_new;
      *reinterpret_cast<float4*>(out_acc + rowi * D + g * 4) = a;
    }
    __syncwarp();
    if (live && g == 0) lse_acc[li] = lnew;
  }
}

template <typename T>
__global__ void finalize_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n4) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += step) {
    const float4 a = *reinterpret_cast<const float4*>(src + i * 4);
    if constexpr (sizeof(T) == 2) {
      T o[4];
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        o[0] = __float2bfloat16_rn(a.x); o[1] = __float2bfloat16_rn(a.y);
        o[2] = __float2bfloat16_rn(a.z); o[3] = __float2bfloat16_rn(a.w);
      } else {
        o[0] = __float2half_rn(a.x); o[1] = __float2half_rn(a.y);
        o[2] = __float2half_rn(a.z); o[3] = __float2half_rn(a.w);
      }
      *reinterpret_cast<uint2*>(dst + i * 4) = *reinterpret_cast<uint2*>(o);
    } else {
      *reinterpret_cast<float4*>(dst + i * 4) = a;
    }
  }
}

// padded (B,H,max_s) -> flat (H,total) and back; grid (ceil(max_s/128), B, H)
__global__ void flatten_lse_kernel(const float* __restrict__ pad, float* __restrict__ flat,
                                   const int* __restrict__ cu, int H, int max_s, int total) {
  const int b = blockIdx.y, h = blockIdx.z;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int beg = cu[b], len = cu[b + 1] - beg;
  if (s < len) flat[static_cast<int64_t>(h) * total + beg + s] = pad[(static_cast<int64_t>(b) * H + h) * max_s + s];
}
__global__ void unflatten_lse_kernel(const float* __restrict__ flat, float* __restrict__ pad,
                                     const int* __restrict__ cu, int H, int max_s, int total) {
  const int b = blockIdx.y, h = blockIdx.z;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int beg = cu[b], len = cu[b + 1] - beg;
  if (s < max_s)
    pad[(static_cast<int64_t>(b) * H + h) * max_s + s] = s < len ? flat[static_cast<int64_t>(h) * total + beg + s] : -INFINITY;
}

// src (B,S,G,X) -> dst (G,B,S,X), X = chunk of `xv` 16-byte vectors (or the inverse with kOut=false)
template <bool kOut>
__global__ void permute_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t BS, int G, int xv) {
  const int64_t total = BS * G * xv;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += step) {
    const int x = static_cast<int>(i % xv);
    const int64_t r = i / xv;
    const int g = static_cast<int>(r % G);
    const int64_t bs = r / G;
    const int64_t j = (static_cast<int64_t>(g) * BS + bs) * xv + x;   // index in (G,BS,X)
    if (kOut) dst[j] = src[i]; else dst[i] = src[j];
  }
}

// delta[b,h,s] = sum_d out[b,s,h,d] * dout[b,s,h,d]; one warp per (b,s,h)
template <typename T>
__global__ void delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta,
                             const float* __restrict__ lse, float* __restrict__ lse2, int B, int S, int H, int D, int64_t o_sb, int64_t o_ss, int64_t o_sh,
                             int64_t d_sb, int64_t d_ss, int64_t d_sh) {
  const int lane = threadIdx.x & 31;
  const int64_t nrows = static_cast<int64_t>(B) * S * H;
  const int64_t wstep = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < nrows; r += wstep) {
    const int h = static_cast<int>(r % H);
    const int64_t bs = r / H;
    const int s = static_cast<int>(bs % S);
    const int b = static_cast<int>(bs / S);
    const T* o = out + b * o_sb + s * o_ss + h * o_sh;
    const T* g = dout + b * d_sb + s * d_ss + h * d_sh;
    float acc = 0.f;
    for (int d = lane * 8; d < D; d += 256) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + d);
      const uint4 c = *reinterpret_cast<const uint4*>(g + d);
      const T* ea = reinterpret_cast<const T*>(&a);
      const T* ec = reinterpret_cast<const T*>(&c);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += to_f(ea[k]) * to_f(ec[k]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const int64_t idx = (static_cast<int64_t>(b) * H + h) * S + s;
      delta[idx] = acc;
      if (lse2 != nullptr) {   // log2-domain LSE for the backward kernels; rows without keys -> +inf => P = 0
        const float l = lse[idx];
        lse2[idx] = (l == -INFINITY) ? INFINITY : l * 1.4426950408889634f;
      }
    }
  }
}

// spin (device side, stream ordered) until a monotonic system-scope counter reaches `target`
__global__ void wait_counter_kernel(const uint32_t* sig, uint32_t target) {
  uint32_t v;
  unsigned long long polls = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(sig) : "memory");
    if (static_cast<int32_t>(v - target) >= 0) break;
    __nanosleep(100);
    if (++polls > (1ull << 24)) {      // ~30 s: a peer never delivered -> fail loudly instead of hanging the GPU
      printf("[lca_b200] watchdog: counter %p stuck at %u, waiting for %u\n", sig, v, target);
      __trap();
    }
  } while (true);
}

// bf16/fp16 (B,S,H,D) -> e4m3 (B,S,H,D) with one fp32 scale per (b, h, 128-row block) [or a caller-provided per-(b,h)
// scale when ext_scale != nullptr].  scale = amax / 448 so that x / scale fills the e4m3 range.  grid (nblk, H, B).
template <typename T>
__global__ void quant_e4m3_kernel(const T* __restrict__ x, uint8_t* __restrict__ y, float* __restrict__ scale,
                                  const float* __restrict__ ext_scale, int S, int H, int D, int64_t sb, int64_t ss,
                                  int64_t sh, int nblk) {
  const int blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int r0 = blk * 128;
  const int nrows = min(128, S - r0);
  const int vec_per_row = D / 8;
  const int total = nrows * vec_per_row;
  const T* base = x + b * sb + static_cast<int64_t>(r0) * ss + h * sh;
  __shared__ float red[32];
  float sc;
  if (ext_scale != nullptr) {
    sc = ext_scale[b * H + h];
  } else {
    float amax = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int r = i / vec_per_row, c = i - r * vec_per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(base + r * ss + c * 8);
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(to_f(e[k])));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x < 32) {
      float a = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, off));
      if (threadIdx.x == 0) red[0] = a;
    }
    __syncthreads();
    sc = red[0] > 0.f ? red[0] / 448.f : 1.f;
    if (threadIdx.x == 0) scale[(static_cast<int64_t>(b) * H + h) * nblk + blk] = sc;
  }
  const float inv = 1.f / sc;
  uint8_t* ybase = y + ((static_cast<int64_t>(b) * S + r0) * H + h) * D;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / vec_per_row, c = i - r * vec_per_row;
    const uint4 v = *reinterpret_cast<const uint4*>(base + r * ss + c * 8);
    const T* e = reinterpret_cast<const T*>(&v);
    uint16_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(o[k]) : "f"(to_f(e[2 * k + 1]) * inv), "f"(to_f(e[2 * k]) * inv));
    *reinterpret_cast<uint2*>(ybase + static_cast<int64_t>(r) * H * D + c * 8) = *reinterpret_cast<uint2*>(o);
  }
}

inline int grid_for(int64_t work_items, int threads) {
  int64_t g = (work_items + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

cudaError_t launch_merge_out_lse(float* out_acc, float* lse_acc, const void* block_out, int block_dtype,
                                 const float* block_lse, int B, int S, int H, int D, cudaStream_t stream) {
  if (D % 4 != 0 || D / 4 > 32 || 32 % (D / 4) != 0) return cudaErrorInvalidValue;  // one row per <= warp
  const int64_t total = static_cast<int64_t>(B) * S * H * (D / 4);
  const int threads = 256;
  const int grid = grid_for(total, threads);
  if (block_dtype == 0)
    merge_kernel<float><<<grid, threads, 0, stream>>>(out_acc, lse_acc, static_cast<const float*>(block_out), block_lse, B, S, H, D);
  else if (block_dtype == 1)
    merge_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(out_acc, lse_acc, static_cast<const __nv_bfloat16*>(block_out), block_lse, B, S, H, D);
  else
    merge_kernel<__half><<<grid, threads, 0, stream>>>(out_acc, lse_acc, static_cast<const __half*>(block_out), block_lse, B, S, H, D);
  return cudaGetLastError();
}

cudaError_t launch_finalize_out(const float* out_acc, void* out, int out_dtype, int64_t n, cudaStream_t stream) {
  const int64_t n4 = n / 4;
  const int threads = 256;
  const int grid = grid_for(n4, threads);
  if (out_dtype == 1) finalize_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(out_acc, static_cast<__nv_bfloat16*>(out), n4);
  else if (out_dtype == 2) finalize_kernel<__half><<<grid, threads, 0, stream>>>(out_acc, static_cast<__half*>(out), n4);
  else finalize_kernel<float><<<grid, threads, 0, stream>>>(out_acc, static_cast<float*>(out), n4);
  return cudaGetLastError();
}

cudaError_t launch_flatten_lse(const float* pad, float* flat, const int* cu, int B, int H, int max_s, int total, cudaStream_t stream) {
  dim3 grid((max_s + 127) / 128, B, H);
  flatten_lse_kernel<<<grid, 128, 0, stream>>>(pad, flat, cu, H, max_s, total);
  return cudaGetLastError();
}
cudaError_t launch_unflatten_lse(const float* flat, float* pad, const int* cu, int B, int H, int max_s, int total, cudaStream_t stream) {
  dim3 grid((max_s + 127) / 128, B, H);
  unflatten_lse_kernel<<<grid, 128, 0, stream>>>(flat, pad, cu, H, max_s, total);
  return cudaGetLastError();
}

cudaError_t launch_permute_heads_out(const void* src, void* dst, int B, int S, int G, int x_bytes, cudaStream_t stream) {
  if (x_bytes % 16) return cudaErrorInvalidValue;
  const int64_t BS = static_cast<int64_t>(B) * S;
  const int xv = x_bytes / 16;
  permute_kernel<true><<<grid_for(BS * G * xv, 256), 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), BS, G, xv);
  return cudaGetLastError();
}
cudaError_t launch_permute_heads_in(const void* src, void* dst, int B, int S, int G, int x_bytes, cudaStream_t stream) {
  if (x_bytes % 16) return cudaErrorInvalidValue;
  const int64_t BS = static_cast<int64_t>(B) * S;
  const int xv = x_bytes / 16;
  permute_kernel<false><<<grid_for(BS * G * xv, 256), 256, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), BS, G, xv);
  return cudaGetLastError();
}

cudaError_t launch_quant_e4m3(const void* x, int dtype, uint8_t* y, float* scale, const float* ext_scale, int B, int S, int H,
                              int D, int64_t sb, int64_t ss, int64_t sh, cudaStream_t stream) {
  if (D % 8) return cudaErrorInvalidValue;
  const int nblk = (S + 127) / 128;
  dim3 grid(nblk, H, B);
  if (dtype == 1)
    quant_e4m3_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), y, scale, ext_scale, S, H, D, sb, ss, sh, nblk);
  else
    quant_e4m3_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(x), y, scale, ext_scale, S, H, D, sb, ss, sh, nblk);
  return cudaGetLastError();
}

cudaError_t launch_wait_counter(const uint32_t* sig, uint32_t target, cudaStream_t stream) {
  wait_counter_kernel<<<1, 1, 0, stream>>>(sig, target);
  return cudaGetLastError();
}

cudaError_t launch_delta(const void* out, const void* dout, int dtype, float* delta, const float* lse, float* lse2,
                         int B, int S, int H, int D,
                         int64_t o_sb, int64_t o_ss, int64_t o_sh, int64_t d_sb, int64_t d_ss, int64_t d_sh,
                         cudaStream_t stream) {
  if (D % 8) return cudaErrorInvalidValue;
  const int64_t nrows = static_cast<int64_t>(B) * S * H;
  const int threads = 256;
  const int grid = grid_for(nrows * 32, threads);
  if (dtype == 1)
    delta_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), delta, lse, lse2, B, S, H, D, o_sb, o_ss, o_sh, d_sb, d_ss, d_sh);
  else
    delta_kernel<__half><<<grid, threads, 0, stream>>>(static_cast<const __half*>(out), static_cast<const __half*>(dout), delta, lse, lse2, B, S, H, D, o_sb, o_ss, o_sh, d_sb, d_ss, d_sh);
  return cudaGetLastError();
}

}  // namespace lca
