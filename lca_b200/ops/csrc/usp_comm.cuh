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

// Spin-wait watchdog: a peer that never arrives (crashed rank, mismatched call sequence) must not hang the GPU forever.
// The budget is a launch parameter (CommParams::watchdog_ns, from LCA_B200_WATCHDOG_S, default 600 s; 0 = never trap):
// ranks legitimately drift by many seconds around checkpoints / evaluation / dataloader stalls.  After the budget
// the kernel prints the stuck flag and traps, which surfaces as a CUDA error on the host.
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

static __device__ __noinline__ void spin_until_ge(const uint32_t* addr, uint32_t target, unsigned ns,
                                                  unsigned long long watchdog_ns) {
  unsigned polls = 0;
  unsigned long long t0 = 0;
  while (static_cast<int32_t>(ptx::ld_acquire_sys(addr) - target) < 0) {
    __nanosleep(ns);
    if ((++polls & 1023u) == 0u && watchdog_ns != 0ull) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0ull) t0 = now;
      if (now - t0 > watchdog_ns) {
        printf("[lca_b200] watchdog: flag %p stuck at %u, waiting for %u (block %d) after %llu ms\n", addr,
               ptx::ld_relaxed_sys(addr), target, static_cast<int>(blockIdx.x), (now - t0) / 1000000ull);
        __trap();
      }
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

// Legacy scalar push engine (CommParams::push_mode == 0, LCA_B200_PUSH=scalar): every thread of the push CTAs moves
// 16-byte vectors with ld.global / st.global.  ~25 GB/s per CTA (4 x 16 B in flight per thread); kept as the
// reference implementation the bulk engine below is validated against.
static __device__ __noinline__ void comm_cta(const CommParams& c, bool wait_o) {
  using namespace ptx;
  const int tid = static_cast<int>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nthreads = c.n_comm * blockDim.x;
  const int me = c.r * c.U + c.u;
  const int esz = 2;
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < c.P)      // tell every peer my staging is free for this call
    st_release_sys(c.peer_sig[threadIdx.x] + kSigRTR + me, c.epoch);
  const long long row_off_kv = (static_cast<long long>(c.r) * c.U + c.u) * c.rows;
  const long long row_off_q = static_cast<long long>(c.u) * c.rows;
  for (int i = 0; i < c.P; ++i) {
    const int d = (me + i) % c.P;
    const int du = d % c.U, dr = d / c.U;
    if (threadIdx.x == 0) {
      spin_until_ge(c.my_sig + kSigRTR + d, c.epoch, 32, c.watchdog_ns);
    }
    __syncthreads();
    CopyMsg m;
    // KV-like head-slice of destination du: kv head(s) [h0, h0 + Hkvl)
    const int h0 = (c.Hkv >= c.U) ? du * c.Hkvl : (du * c.Hkv) / c.U;
    m.nrows = c.rows;
    m.row_vecs = c.Hkvl * c.D * esz / 16;
    m.dst_ss = static_cast<long long>(c.Hkvl) * c.D * esz;
    m.dst_sb = c.stage_kv_rows * m.dst_ss;
    for (int t = 0; t < c.n_kv && ((c.kv_dst_mask >> d) & 1u); ++t) {
      m.src = static_cast<const unsigned char*>(c.kvt[t].src) + static_cast<long long>(h0) * c.D * esz;
      m.src_sb = c.kvt[t].sb * esz; m.src_ss = c.kvt[t].ss * esz;
      m.dst = c.peer_slab[d] + c.kvt[t].off + row_off_kv * m.dst_ss;
      comm_copy(m, c.B, tid, nthreads);
    }
    const bool have_q = c.n_q > 0 || c.n_stat > 0;
    const bool send_q = have_q && (c.q_to_all || dr == c.r) && ((c.q_dst_mask >> d) & 1u);
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
        comm_copy(m, c.B, tid, nthreads);
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
        comm_copy(s, c.B, tid, nthreads);
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
  if (wait_o && blockIdx.x == 0 && threadIdx.x == 0 && c.o_target != 0) {
    spin_until_ge(c.my_sig + kSigODone, c.o_target, 64, c.watchdog_ns);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Bulk push engine (CommParams::push_mode == 1, the default): the TMA unit does the copying.
//
// One warp per push CTA drives a ring of kPushStages x 32 KiB shared-memory stages:
//   cp.async.bulk  global(local HBM) -> smem   (mbarrier complete_tx),   kPushLag chunks ahead of
//   cp.async.bulk  smem -> global(peer slab over NVLink, bulk groups).
// ~128 KiB of loads are in flight per CTA (the scalar loop above had 16 KiB), no address arithmetic per 16 bytes,
// no registers holding payload, and a handful of CTAs saturate an NVLink direction.  Contiguous rows move as ONE bulk
// operation per chunk; strided rows (Ulysses head slices) as one operation per row, issued by the 32 lanes in parallel.
// Per destination: wait ready-to-receive, stream its messages, cp.async.bulk.wait_group 0 (writes performed),
// fence, red.release.sys on the destination's arrival counters.
// In the owner-computes backward (q_to_all) K/V go to EVERY destination first and Q/dO/statistics in a second sweep:
// the dQ pass only needs remote K/V; remote Q/dO are first touched by the dK/dV pass one launch later.
constexpr int kPushStageBytes = 32768;
constexpr int kPushStages = 6;
constexpr int kPushLag = 4;
constexpr int kPushSmemBytes = 1024 + kPushStages * kPushStageBytes + 128;   // alignment slack + stages + barriers

struct BulkMsg {
  const unsigned char* src;
  unsigned char* dst;
  long long src_sb, src_ss, dst_sb, dst_ss;   // bytes
  int nrows, row_bytes;                       // rows per batch, bytes per row (multiple of 16)
  int rpc, ppr, cpb;                          // rows per chunk, pieces per row, chunks per batch
  int nchunks;
};

__device__ __forceinline__ void bulk_msg_finish(BulkMsg& m, int B) {
  if (m.row_bytes >= kPushStageBytes) {
    m.rpc = 1;
    m.ppr = (m.row_bytes + kPushStageBytes - 1) / kPushStageBytes;
    m.cpb = m.nrows * m.ppr;
  } else {
    m.rpc = kPushStageBytes / m.row_bytes;
    m.ppr = 1;
    m.cpb = (m.nrows + m.rpc - 1) / m.rpc;
  }
  m.nchunks = B * m.cpb;
}

struct BulkChunk {
  const unsigned char* src;
  unsigned char* dst;
  long long src_ss, dst_ss;
  int n, len;                                 // n rows of len bytes each (n > 1 only when len == row_bytes)
};

__device__ __forceinline__ void bulk_decode(const BulkMsg* msgs, int n_msg, int g, BulkChunk& ck) {
  int mi = 0;
  while (mi < n_msg - 1 && g >= msgs[mi].nchunks) { g -= msgs[mi].nchunks; ++mi; }
  const BulkMsg& m = msgs[mi];
  const int b = g / m.cpb;
  const int k = g - b * m.cpb;
  int r0, n, byte0, len;
  if (m.ppr > 1) {
    r0 = k / m.ppr;
    n = 1;
    byte0 = (k - r0 * m.ppr) * kPushStageBytes;
    len = min(kPushStageBytes, m.row_bytes - byte0);
  } else {
    r0 = k * m.rpc;
    n = min(m.rpc, m.nrows - r0);
    byte0 = 0;
    len = m.row_bytes;
  }
  ck.src = m.src + b * m.src_sb + r0 * m.src_ss + byte0;
  ck.dst = m.dst + b * m.dst_sb + r0 * m.dst_ss + byte0;
  ck.src_ss = m.src_ss;
  ck.dst_ss = m.dst_ss;
  ck.n = n;
  ck.len = len;
}

// state of one push CTA's pipeline; lives in registers of warp 0 (all lanes hold identical copies)
struct BulkPipe {
  uint32_t stage0, bar0;
  uint32_t loads, stores;      // running chunk counters (slot = counter % kPushStages, phase = counter / kPushStages)
};

// Streams chunks g = first, first + step, ... < total of the message list through the pipeline and drains it.
__device__ __forceinline__ void bulk_stream(BulkPipe& pp, const BulkMsg* msgs, int n_msg, int total, int first, int step,
                                            int lane) {
  using namespace ptx;
  const int n_mine = first < total ? (total - first + step - 1) / step : 0;
  for (int it = 0; it < n_mine + kPushLag; ++it) {
    if (it < n_mine) {
      bulk_wait_read<kPushStages - kPushLag - 1>();        // the store that last used this slot has read its smem
      BulkChunk ck;
      bulk_decode(msgs, n_msg, first + it * step, ck);
      const uint32_t slot = pp.loads % kPushStages;
      const uint32_t dst = pp.stage0 + slot * kPushStageBytes, bar = pp.bar0 + 8 * slot;
      if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(ck.n) * ck.len);
      __syncwarp();
      if (ck.n == 1 || ck.src_ss == ck.len) {
        if (lane == 0) bulk_g2s(dst, ck.src, static_cast<uint32_t>(ck.n) * ck.len, bar);
      } else {
        for (int r = lane; r < ck.n; r += 32) bulk_g2s(dst + r * ck.len, ck.src + r * ck.src_ss, ck.len, bar);
      }
      ++pp.loads;
    }
    if (it >= kPushLag) {
      BulkChunk ck;
      bulk_decode(msgs, n_msg, first + (it - kPushLag) * step, ck);
      const uint32_t slot = pp.stores % kPushStages;
      const uint32_t src = pp.stage0 + slot * kPushStageBytes;
      mbar_wait(pp.bar0 + 8 * slot, (pp.stores / kPushStages) & 1);
      if (ck.n == 1 || ck.dst_ss == ck.len) {
        if (lane == 0) bulk_s2g(ck.dst, src, static_cast<uint32_t>(ck.n) * ck.len);
      } else {
        for (int r = lane; r < ck.n; r += 32) bulk_s2g(ck.dst + r * ck.dst_ss, src + r * ck.len, ck.len);
      }
      bulk_commit();                                         // every lane, also the ones without a store
      ++pp.stores;
    }
  }
  bulk_wait<0>();                                            // writes of this lane's groups have been performed
  __syncwarp();
}

// smem_base: 1024-byte aligned dynamic shared memory of at least kPushSmemBytes - 1024 bytes.
// wait_o: the legacy contract (push CTAs leave the kernel afterwards): block 0 also waits for my output buffer.
static __device__ __noinline__ void comm_cta_bulk(const CommParams& c, uint32_t smem_base, bool wait_o) {
  using namespace ptx;
  const int me = c.r * c.U + c.u;
  const int esz = 2;
  const int cta = static_cast<int>(blockIdx.x);
  if (threadIdx.x >= 32) return;                    // one warp drives the TMA unit
  const int lane = static_cast<int>(threadIdx.x);
  if (cta == 0 && lane < c.P)                       // tell every peer my staging is free for this call
    st_release_sys(c.peer_sig[lane] + kSigRTR + me, c.epoch);
  BulkPipe pp;
  pp.stage0 = smem_base;
  pp.bar0 = smem_base + kPushStages * kPushStageBytes;
  pp.loads = pp.stores = 0;
  // one barrier per lane: with a warp-uniform address under `if (lane == 0)` ptxas (12.9) if-converts the block, picks
  // the uniform-datapath form of SYNCS.EXCH and then drops it (`@!P1 NOP` in the SASS) -- the first arrive.expect_tx
  // on the never-initialised barrier then faults.  Per-lane addresses force the vector form.
  if (lane < kPushStages) mbar_init(pp.bar0 + 8 * lane, 1);
  fence_mbar_init();
  __syncwarp();
  const long long row_off_kv = (static_cast<long long>(c.r) * c.U + c.u) * c.rows;
  const long long row_off_q = static_cast<long long>(c.u) * c.rows;
  const bool have_q = c.n_q > 0 || c.n_stat > 0;
  const int n_sweeps = (c.q_to_all && have_q) ? 2 : 1;
  int rot = 0;                                      // rotates the chunk -> CTA assignment so short messages spread out
  for (int sweep = 0; sweep < n_sweeps; ++sweep) {
    for (int i = 0; i < c.P; ++i) {
      const int d = (me + i) % c.P;
      const int du = d % c.U, dr = d / c.U;
      const bool do_kv = sweep == 0;
      const bool do_q = have_q && (c.q_to_all ? sweep == 1 : dr == c.r);
      // destinations whose tiles can never see these rows (causal order, sliding window) get the arrival signal
      // but no data: the consumers' tile iterators skip exactly those segments (host: FusedUSPEngine._push_masks)
      const bool kv_data = do_kv && ((c.kv_dst_mask >> d) & 1u);
      const bool q_data = do_q && ((c.q_dst_mask >> d) & 1u);
      BulkMsg msgs[6];
      int n_msg = 0, total = 0;
      if (q_data) {                                   // forward: the consumer needs its Q tile before any K/V tile
        const long long q_rows = c.q_to_all ? c.stage_kv_rows : c.stage_q_rows;     // rows of the destination staging
        const long long q_off = c.q_to_all ? row_off_kv : row_off_q;
        for (int t = 0; t < c.n_q; ++t) {
          BulkMsg& m = msgs[n_msg++];
          m.nrows = c.rows;
          m.row_bytes = c.Hl * c.D * esz;
          m.src = static_cast<const unsigned char*>(c.qt[t].src) + static_cast<long long>(du) * c.Hl * c.D * esz;
          m.src_sb = c.qt[t].sb * esz; m.src_ss = c.qt[t].ss * esz;
          m.dst_ss = m.row_bytes;
          m.dst_sb = q_rows * m.dst_ss;
          m.dst = c.peer_slab[d] + c.qt[t].off + q_off * m.dst_ss;
          bulk_msg_finish(m, c.B);
          total += m.nchunks;
        }
        for (int t = 0; t < c.n_stat; ++t) {        // (B, H, rows) fp32 -> destination (B, Hl, q_rows) at column q_off
          BulkMsg& m = msgs[n_msg++];
          m.nrows = c.Hl;
          m.row_bytes = c.rows * 4;
          m.src = reinterpret_cast<const unsigned char*>(c.stat[t]) + static_cast<long long>(du) * c.Hl * c.rows * 4;
          m.src_sb = static_cast<long long>(c.H) * c.rows * 4;
          m.src_ss = static_cast<long long>(c.rows) * 4;
          m.dst = c.peer_slab[d] + c.stat_off[t] + q_off * 4;
          m.dst_sb = static_cast<long long>(c.Hl) * q_rows * 4;
          m.dst_ss = q_rows * 4;
          bulk_msg_finish(m, c.B);
          total += m.nchunks;
        }
      }
      if (kv_data) {
        const int h0 = (c.Hkv >= c.U) ? du * c.Hkvl : (du * c.Hkv) / c.U;      // kv head(s) of destination du
        for (int t = 0; t < c.n_kv; ++t) {
          BulkMsg& m = msgs[n_msg++];
          m.nrows = c.rows;
          m.row_bytes = c.Hkvl * c.D * esz;
          m.src = static_cast<const unsigned char*>(c.kvt[t].src) + static_cast<long long>(h0) * c.D * esz;
          m.src_sb = c.kvt[t].sb * esz; m.src_ss = c.kvt[t].ss * esz;
          m.dst_ss = m.row_bytes;
          m.dst_sb = c.stage_kv_rows * m.dst_ss;
          m.dst = c.peer_slab[d] + c.kvt[t].off + row_off_kv * m.dst_ss;
          bulk_msg_finish(m, c.B);
          total += m.nchunks;
        }
      }
      if (n_msg > 0) {
        if (lane == 0) spin_until_ge(c.my_sig + kSigRTR + d, c.epoch, 32, c.watchdog_ns);
        __syncwarp();
        const int first = ((cta - rot) % c.n_comm + c.n_comm) % c.n_comm;
        bulk_stream(pp, msgs, n_msg, total, first, c.n_comm, lane);
        rot = (rot + total) % c.n_comm;
        fence_proxy_async();
        __threadfence_system();
        __syncwarp();
      }
      if (lane == 0) {
        if (do_kv) red_add_release_sys(c.peer_sig[d] + kSigKV + me, 1u);
        if (c.q_to_all ? sweep == n_sweeps - 1 : true) {
          if (dr == c.r) red_add_release_sys(c.peer_sig[d] + kSigQ + c.u, 1u);
          red_add_release_sys(c.peer_sig[d] + kSigQA + me, 1u);
        }
      }
    }
  }
  if (lane < kPushStages) mbar_inval(pp.bar0 + 8 * lane);
  __syncwarp();
  // my output buffer is complete once every compute rank has scattered its tiles into it
  if (wait_o && cta == 0 && lane == 0 && c.o_target != 0) spin_until_ge(c.my_sig + kSigODone, c.o_target, 64, c.watchdog_ns);
}

// kernel-side dispatch of the communication role; returns after this CTA's transfers are out
__device__ __forceinline__ void comm_role(const CommParams& c, uint32_t smem_base, bool wait_o) {
  if (c.push_mode == 1) {
    comm_cta_bulk(c, smem_base, wait_o);
  } else {
    comm_cta(c, wait_o);
  }
}

__device__ __forceinline__ void wait_arrival(const uint32_t* flags, uint32_t epoch, int idx,
                                             unsigned long long watchdog_ns) {
  using namespace ptx;
  if (idx < 0) return;
  spin_until_ge(flags + idx, epoch, 64, watchdog_ns);
  fence_proxy_async();   // order the acquire before the async-proxy (TMA) reads that follow
}

}  // namespace lca
