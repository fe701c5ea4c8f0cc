// Communication CTAs of the fused USP kernels (shared by forward and backward).
//
// For every destination sp-rank d (self first, then rotated so NVSwitch ports are evenly loaded):
//   wait until d has entered this call (ready-to-receive flag), copy my KV-like head-slices for d -- and my
//   Q-like head-slices (+ per-row statistic) if d is in my Ulysses group -- into d's staging with 16-byte
//   st.global over NVLink, then fence.sys + red.release.sys on d's arrival counters.  The compute CTAs of d
//   poll those counters with ld.acquire.sys right before the TMA loads of the matching segment.
#pragma once
#include "fmha_params.h"
#include "sm100_ptx.cuh"
#include <cstdio>

namespace lca {

// Spin-wait watchdog: a peer that never arrives (crashed rank, mismatched call sequence) must not hang the
// GPU forever -- after ~30 s of polling the kernel traps, which surfaces as a CUDA error on the host.
constexpr unsigned long long kWatchdogPolls = 1ull << 24;   // ~1-2 us per poll

static __device__ __noinline__ void spin_until_ge(const uint32_t* addr, uint32_t target, unsigned ns) {
  unsigned long long polls = 0;
  while (static_cast<int32_t>(ptx::ld_acquire_sys(addr) - target) < 0) {
    __nanosleep(ns);
    if (++polls > kWatchdogPolls) {
      printf("[lca_b200] watchdog: flag %p stuck at %u, waiting for %u (block %d)\n", addr, ptx::ld_relaxed_sys(addr),
             target, static_cast<int>(blockIdx.x));
      __trap();
    }
  }
}

struct CopyMsg {
  const unsigned char* src;
  unsigned char* dst;
  long long src_sb, src_ss, dst_sb, dst_ss;   // bytes
  int nrows, row_vecs;                        // rows per batch, 16-byte vectors per row
};

__device__ __forceinline__ void comm_copy(const CopyMsg& m, int B, int tid, int nthreads) {
  const long long total = static_cast<long long>(B) * m.nrows * m.row_vecs;
  constexpr int UNR = 4;
  for (long long base = static_cast<long long>(tid) * UNR; base < total; base += static_cast<long long>(nthreads) * UNR) {
    uint4 val[UNR];
    long long doff[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const long long i = base + j;
      doff[j] = -1;
      if (i < total) {
        const int c = static_cast<int>(i % m.row_vecs);
        const long long br = i / m.row_vecs;
        const int row = static_cast<int>(br % m.nrows);
        const int b = static_cast<int>(br / m.nrows);
        val[j] = *reinterpret_cast<const uint4*>(m.src + b * m.src_sb + row * m.src_ss + c * 16);
        doff[j] = b * m.dst_sb + row * m.dst_ss + c * 16;
      }
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j)
      if (doff[j] >= 0) *reinterpret_cast<uint4*>(m.dst + doff[j]) = val[j];
  }
}

// Copy loop of the EXPERIMENTAL push engine (comm_cta<true>): 8 x 16 B in flight per thread instead of 4 (the plain
// loop is latency-bound at ~25 GB/s per CTA), row/column split by shift when the row length is a power of two
// (it is for every head_dim x heads combination in use) instead of two 64-bit divisions per vector.
// kStrong: the destination is an NVLS *multicast* address (one store is replicated by the NVSwitch into every rank's
// slab); multimem addresses must be written with multimem.st (SASS: STG.E.128.STRONG.SYS).
template <bool kStrong>
__device__ __forceinline__ void comm_copy_fast(const CopyMsg& m, int B, int tid, int nthreads) {
  constexpr int UNR = 8;
  const unsigned rv = static_cast<unsigned>(m.row_vecs);
  const unsigned nrows = static_cast<unsigned>(m.nrows);
  const unsigned long long total = static_cast<unsigned long long>(B) * nrows * rv;
  const bool pow2 = (rv & (rv - 1u)) == 0u;
  const int sh = 31 - __clz(static_cast<int>(rv));
  const unsigned long long step = static_cast<unsigned long long>(nthreads >> 5) * (32 * UNR);
  for (unsigned long long base = static_cast<unsigned long long>(tid >> 5) * (32 * UNR) + (tid & 31); base < total;
       base += step) {
    uint4 val[UNR];
    long long doff[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const unsigned long long i = base + 32ull * j;      // a warp reads 512 contiguous bytes per j
      doff[j] = -1;
      if (i < total) {
        const unsigned long long br = pow2 ? (i >> sh) : (i / rv);
        const unsigned c = static_cast<unsigned>(i - br * rv);
        unsigned b = 0, row = static_cast<unsigned>(br);
        if (B > 1) { b = row / nrows; row -= b * nrows; }
        val[j] = *reinterpret_cast<const uint4*>(m.src + b * m.src_sb + row * m.src_ss + c * 16ll);
        doff[j] = b * m.dst_sb + row * m.dst_ss + c * 16ll;
      }
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (doff[j] < 0) continue;
      if constexpr (kStrong) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(m.dst + doff[j]),
                     "f"(__uint_as_float(val[j].x)), "f"(__uint_as_float(val[j].y)), "f"(__uint_as_float(val[j].z)),
                     "f"(__uint_as_float(val[j].w))
                     : "memory");
      } else {
        *reinterpret_cast<uint4*>(m.dst + doff[j]) = val[j];
      }
    }
  }
}

// kMc (EXPERIMENTAL; LCA_B200_FAST_PUSH=1: faster copy loops only; LCA_B200_NVLS=1 with the VMM slab: broadcast too):
// when the host put the slab's NVLS multicast address into
// peer_slab[kMaxPeers - 1], everything that goes to EVERY rank (K/V; in the owner-computes backward also Q, dO and the
// row statistics) is written ONCE to the multicast window instead of P times to unicast peers: NVLink egress / P.
template <bool kMc>
static __device__ __noinline__ void comm_cta(const CommParams& c) {
  using namespace ptx;
  const int tid = static_cast<int>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nthreads = c.n_comm * blockDim.x;
  const int me = c.r * c.U + c.u;
  const int esz = 2;
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < c.P)      // tell every peer my staging is free for this call
    st_release_sys(c.peer_sig[threadIdx.x] + kSigRTR + me, c.epoch);
  const long long row_off_kv = (static_cast<long long>(c.r) * c.U + c.u) * c.rows;
  const long long row_off_q = static_cast<long long>(c.u) * c.rows;
  bool mc = false;
  if constexpr (kMc) {
    unsigned char* mcb = c.peer_slab[kMaxPeers - 1];     // multicast base, or a small sentinel: fast unicast copies only
    // a head slice per destination exists only when heads are scattered (U > 1); the broadcast needs ONE slice for all
    mc = (reinterpret_cast<uintptr_t>(mcb) > 4096u) && (c.U == 1);
    if (mc) {
      if (static_cast<int>(threadIdx.x) < c.P) spin_until_ge(c.my_sig + kSigRTR + threadIdx.x, c.epoch, 32);
      __syncthreads();                    // every rank has entered this call: all staging buffers may be overwritten
      CopyMsg m;
      m.nrows = c.rows;
      m.row_vecs = c.Hkvl * c.D * esz / 16;
      m.dst_ss = static_cast<long long>(c.Hkvl) * c.D * esz;
      m.dst_sb = c.stage_kv_rows * m.dst_ss;
      for (int t = 0; t < c.n_kv; ++t) {
        m.src = static_cast<const unsigned char*>(c.kvt[t].src);
        m.src_sb = c.kvt[t].sb * esz; m.src_ss = c.kvt[t].ss * esz;
        m.dst = mcb + c.kvt[t].off + row_off_kv * m.dst_ss;
        comm_copy_fast<true>(m, c.B, tid, nthreads);
      }
      if (c.q_to_all) {
        m.row_vecs = c.Hl * c.D * esz / 16;
        m.dst_ss = static_cast<long long>(c.Hl) * c.D * esz;
        m.dst_sb = c.stage_kv_rows * m.dst_ss;
        for (int t = 0; t < c.n_q; ++t) {
          m.src = static_cast<const unsigned char*>(c.qt[t].src);
          m.src_sb = c.qt[t].sb * esz; m.src_ss = c.qt[t].ss * esz;
          m.dst = mcb + c.qt[t].off + row_off_kv * m.dst_ss;
          comm_copy_fast<true>(m, c.B, tid, nthreads);
        }
        for (int t = 0; t < c.n_stat; ++t) {
          CopyMsg s;
          s.nrows = c.Hl;
          s.row_vecs = c.rows * 4 / 16;
          s.src = reinterpret_cast<const unsigned char*>(c.stat[t]);
          s.src_sb = static_cast<long long>(c.H) * c.rows * 4;
          s.src_ss = static_cast<long long>(c.rows) * 4;
          s.dst = mcb + c.stat_off[t] + row_off_kv * 4;
          s.dst_sb = static_cast<long long>(c.Hl) * c.stage_kv_rows * 4;
          s.dst_ss = c.stage_kv_rows * 4;
          comm_copy_fast<true>(s, c.B, tid, nthreads);
        }
      }
      __threadfence_system();
      __syncthreads();
    }
  }
  for (int i = 0; i < c.P; ++i) {
    const int d = (me + i) % c.P;
    const int du = d % c.U, dr = d / c.U;
    if (mc) {                             // payload already broadcast: only the arrival counters remain
      if (threadIdx.x == 0) {
        red_add_release_sys(c.peer_sig[d] + kSigKV + me, 1u);
        if (dr == c.r) red_add_release_sys(c.peer_sig[d] + kSigQ + c.u, 1u);
        red_add_release_sys(c.peer_sig[d] + kSigQA + me, 1u);
      }
      continue;
    }
    if (threadIdx.x == 0) {
      spin_until_ge(c.my_sig + kSigRTR + d, c.epoch, 32);
    }
    __syncthreads();
    CopyMsg m;
    // KV-like head-slice of destination du: kv head(s) [h0, h0 + Hkvl)
    const int h0 = (c.Hkv >= c.U) ? du * c.Hkvl : (du * c.Hkv) / c.U;
    m.nrows = c.rows;
    m.row_vecs = c.Hkvl * c.D * esz / 16;
    m.dst_ss = static_cast<long long>(c.Hkvl) * c.D * esz;
    m.dst_sb = c.stage_kv_rows * m.dst_ss;
    for (int t = 0; t < c.n_kv; ++t) {
      m.src = static_cast<const unsigned char*>(c.kvt[t].src) + static_cast<long long>(h0) * c.D * esz;
      m.src_sb = c.kvt[t].sb * esz; m.src_ss = c.kvt[t].ss * esz;
      m.dst = c.peer_slab[d] + c.kvt[t].off + row_off_kv * m.dst_ss;
      if constexpr (kMc) comm_copy_fast<false>(m, c.B, tid, nthreads);
      else comm_copy(m, c.B, tid, nthreads);
    }
    const bool have_q = c.n_q > 0 || c.n_stat > 0;
    const bool send_q = have_q && (c.q_to_all || dr == c.r);
    if (send_q) {
      const long long q_rows = c.q_to_all ? c.stage_kv_rows : c.stage_q_rows;     // rows of the destination staging
      const long long q_off = c.q_to_all ? row_off_kv : row_off_q;
      m.row_vecs = c.Hl * c.D * esz / 16;
      m.dst_ss = static_cast<long long>(c.Hl) * c.D * esz;
      m.dst_sb = q_rows * m.dst_ss;
      for (int t = 0; t < c.n_q; ++t) {
        m.src = static_cast<const unsigned char*>(c.qt[t].src) + static_cast<long long>(du) * c.Hl * c.D * esz;
        m.src_sb = c.qt[t].sb * esz; m.src_ss = c.qt[t].ss * esz;
        m.dst = c.peer_slab[d] + c.qt[t].off + q_off * m.dst_ss;
        if constexpr (kMc) comm_copy_fast<false>(m, c.B, tid, nthreads);
        else comm_copy(m, c.B, tid, nthreads);
      }
      for (int t = 0; t < c.n_stat; ++t) {   // (B, H, rows) fp32 -> destination (B, Hl, q_rows) at column q_off
        CopyMsg s;
        s.nrows = c.Hl;
        s.row_vecs = c.rows * 4 / 16;
        s.src = reinterpret_cast<const unsigned char*>(c.stat[t]) + static_cast<long long>(du) * c.Hl * c.rows * 4;
        s.src_sb = static_cast<long long>(c.H) * c.rows * 4;
        s.src_ss = static_cast<long long>(c.rows) * 4;
        s.dst = c.peer_slab[d] + c.stat_off[t] + q_off * 4;
        s.dst_sb = static_cast<long long>(c.Hl) * q_rows * 4;
        s.dst_ss = q_rows * 4;
        if constexpr (kMc) comm_copy_fast<false>(s, c.B, tid, nthreads);
        else comm_copy(s, c.B, tid, nthreads);
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      red_add_release_sys(c.peer_sig[d] + kSigKV + me, 1u);
      if (dr == c.r) red_add_release_sys(c.peer_sig[d] + kSigQ + c.u, 1u);
      red_add_release_sys(c.peer_sig[d] + kSigQA + me, 1u);
    }
  }
  // my output buffer is complete once every compute rank has scattered its tiles into it
  if (blockIdx.x == 0 && threadIdx.x == 0 && c.o_target != 0) {
    spin_until_ge(c.my_sig + kSigODone, c.o_target, 64);
  }
}

__device__ __forceinline__ void wait_arrival(const uint32_t* flags, uint32_t epoch, int idx) {
  using namespace ptx;
  if (idx < 0) return;
  spin_until_ge(flags + idx, epoch, 64);
  fence_proxy_async();   // order the acquire before the async-proxy (TMA) reads that follow
}

}  // namespace lca
