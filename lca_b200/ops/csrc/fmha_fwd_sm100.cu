// Forward flash attention for sm_100a (B200): TMA -> smem -> tcgen05.mma -> TMEM.
//
// One persistent CTA per SM, 384 threads (3 warpgroups; warps 10-11 idle), warp-specialised:
//   warps 0-3  softmax warpgroup for Q tile 0 (one thread per query row; S/P/O live in TMEM)
//   warps 4-7  softmax warpgroup for Q tile 1
//   warp  8    MMA issuer (single elected thread issues every tcgen05.mma + tcgen05.commit)
//   warp  9    TMA producer (Q tiles + a ring of K/V tiles), also polls NVLink arrival flags
// A work item is a pair of 128-row Q tiles of one (batch, head) sharing one K/V stream:
//   tensor pipe order  QK0(j+1) | PV1(j) | QK1(j+1) | PV0(j+1) ...  so the softmax of one tile
//   overlaps the MMAs of the other (ping-pong), accumulators never leave TMEM, P overwrites S in
//   place as bf16/fp16 and feeds the PV MMA straight from TMEM (tcgen05.mma TS form).
// Masks are evaluated on GLOBAL token positions (segments, see fmha_params.h): causal, sliding
// window, ALiBi, softcap, GQA, ragged segment tails; fully masked K/V tiles are skipped by every
// role with the same deterministic iterator, which is what makes zigzag/stripe load-balanced.
//
// Capability parity: replaces flash_attn::_flash_attn_forward as used by
// yunchang/kernels/attention.py:165-203 (returns out + true LSE).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fmha_params.h"
#include "sm100_ptx.cuh"
#include "usp_comm.cuh"

namespace lca {
using namespace ptx;

namespace {

constexpr int BM = 128;        // query rows per tile (= TMEM lanes)
constexpr int BN = 128;        // key rows per tile
constexpr int kThreads = 384;      // 3 warpgroups: softmax0, softmax1, {MMA, TMA, 2 idle}
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr float kRescaleThreshold = 8.0f;   // lazy rescale: tolerate a stale max up to 2^8

template <int kD>
struct Cfg {
  static constexpr int DBLK = kD / 64;                 // 128-byte column blocks per row
  static constexpr int BLK_BYTES = 128 * 128;          // [128 rows][64 elem] sub-block
  static constexpr int TILE_BYTES = DBLK * BLK_BYTES;  // one Q / K / V tile
  static constexpr int STAGES = (kD == 128) ? 5 : 10;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = 2 * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_KV + STAGES * TILE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;  // + alignment slack
  static constexpr int TMEM_S = 0;                     // S_t / P_t at column t*128
  static constexpr int TMEM_O = 256;                   // O_t at column 256 + t*kD
};

#include "fmha_fwd_common.cuh"   // Work, decode_work, sched_work, TileIter, Ring

__device__ __forceinline__ void wait_flag(const FwdParams& p, int idx) {
  wait_arrival(p.flags, p.flag_epoch, idx, p.comm.watchdog_ns);
}

template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (kBf16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

struct Bars {
  uint32_t q_full[2], q_empty[2], s_full[2], p_full[2], o_full[2];
  uint32_t kv_full, kv_empty;   // base addresses of STAGES-long arrays
};

}  // namespace

// kPolyEvery: 1 of every kPolyEvery element pairs of an unmasked tile uses ex2_poly (0 = MUFU only)
// kDyn: work items are claimed from a global atomic counter by the producer warp and broadcast to the other roles
//       through a 2-deep smem ring (default of the fused multi-GPU launches: the push CTAs join the compute pool);
//       otherwise the static snake schedule is used.
// kDrop: attention dropout regenerated from global coordinates (sm100_ptx.cuh: dropout_*; scalar arithmetic).
// Every other instantiation runs the softmax on packed fp32x2 instructions (FFMA2 / FADD2): scale-and-subtract, the
// polynomial exp2 and the row sum issue once per element pair (validated in round 2: D=64 forward -18 % time,
// D=128 flat; the 64-row-K/V-tile variant this replaced measured 8 % slower and was deleted).
// Empty work items (no visible K/V tile) hand their Q tiles back through o_full, so q_full can never complete two
// phases under the MMA warp's parity wait (tests/test_fwd_pipeline_model_cpu.py).
template <int kD, bool kBf16, int kPolyEvery, bool kDyn, bool kDrop>
__global__ void __launch_bounds__(kThreads, 1) fmha_fwd_kernel(const __grid_constant__ FwdParams p) {
  using C = Cfg<kD>;
  constexpr bool kPk = !kDrop;        // packed fp32x2 softmax arithmetic
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  if (static_cast<int>(blockIdx.x) < p.comm.n_comm) {   // communication role (fused USP path)
    comm_role(p.comm, smem, !kDyn);
    if constexpr (!kDyn) return;
    // kDyn: work is claimed dynamically, so a push CTA joins the compute pool as soon as its transfers are out
    // instead of leaving its SM idle for the rest of the kernel (n_comm of 148 SMs = 5 % at the default of 8).
    // The warps that do not drive the TMA unit wait here: the compute prologue below writes the TMEM base address and
    // its mbarriers into shared memory that the bulk-copy stages are still using.
    __syncthreads();
  }
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;

  // ---- barrier carve-up
  const uint32_t bar0 = smem + C::OFF_BAR;
  Bars B;
  {
    uint32_t a = bar0;
    for (int t = 0; t < 2; ++t) { B.q_full[t] = a; a += 8; }
    for (int t = 0; t < 2; ++t) { B.q_empty[t] = a; a += 8; }
    for (int t = 0; t < 2; ++t) { B.s_full[t] = a; a += 8; }
    for (int t = 0; t < 2; ++t) { B.p_full[t] = a; a += 8; }
    for (int t = 0; t < 2; ++t) { B.o_full[t] = a; a += 8; }
    B.kv_full = a; a += 8 * C::STAGES;
    B.kv_empty = a; a += 8 * C::STAGES;
  }
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + C::OFF_BAR + 480);
  const uint32_t sc_full = bar0 + 400, sc_empty = bar0 + 416;              // dynamic scheduler ring (2 slots)
  volatile int* sched_idx = reinterpret_cast<volatile int*>(smem_gen + C::OFF_BAR + 432);
  // role-specific "next work item" (round = items fetched so far by this role)
  auto producer_next = [&](int round) -> int {
    if constexpr (kDyn) {
      const uint32_t slot = round & 1, par = (round >> 1) & 1;
      mbar_wait(sc_empty + 8 * slot, par ^ 1);
      const int w = static_cast<int>(atomicAdd(p.sched_counter, 1u) - p.sched_base);
      sched_idx[slot] = w;
      mbar_arrive(sc_full + 8 * slot);
      return w;
    } else {
      return sched_work(round, p.comm.n_comm);
    }
  };
  auto consumer_next = [&](int round) -> int {
    if constexpr (kDyn) {
      const uint32_t slot = round & 1, par = (round >> 1) & 1;
      mbar_wait(sc_full + 8 * slot, par);
      const int w = sched_idx[slot];
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(sc_empty + 8 * slot);
      return w;
    } else {
      return sched_work(round, p.comm.n_comm);
    }
  };

  if (threadIdx.x == 0) {
    if constexpr (kDyn) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(sc_full + 8 * s, 1);
        mbar_init(sc_empty + 8 * s, 9);    // MMA warp + 8 softmax warps
      }
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(B.q_full[t], 1);
      mbar_init(B.q_empty[t], 1);
      mbar_init(B.s_full[t], 1);
      mbar_init(B.p_full[t], 4);   // one arrive per softmax warp
      mbar_init(B.o_full[t], 1);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(B.kv_full + 8 * s, 1);
      mbar_init(B.kv_empty + 8 * s, 1);
    }
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&p.tm_q);
    prefetch_tmap(&p.tm_k);
    prefetch_tmap(&p.tm_v);
  }
  if (warp == kMmaWarp) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int hk_div = p.H / p.Hkv;

  if (warp >= kMmaWarp) {
   setmaxnreg_dec<96>();
   if (warp == kTmaWarp) {
    // =========================================================== TMA producer
    if (lane == 0) {
      uint32_t qc[2] = {0, 0};
      Ring<C::STAGES> kr;
      int q_flag_ok = -1, k_flag_ok = -1;
      for (int round = 0;; ++round) {
        Work wk;
        if (!decode_work(p, producer_next(round), wk)) break;
        const int qf = p.qseg[wk.qseg].flag;
        if (qf >= 0 && qf != q_flag_ok) { wait_flag(p, qf); q_flag_ok = qf; }
        for (int t = 0; t < wk.ntile; ++t) {
          mbar_wait(B.q_empty[t], (qc[t] & 1) ^ 1);
          mbar_arrive_expect_tx(B.q_full[t], C::TILE_BYTES);
#pragma unroll
          for (int db = 0; db < C::DBLK; ++db)
            tma_load_4d(smem + C::OFF_Q + t * C::TILE_BYTES + db * C::BLK_BYTES, &p.tm_q,
                        B.q_full[t], db * 64, wk.h, wk.row0 + t * BM, wk.b);
          ++qc[t];
        }
        const int hk = wk.h / hk_div;
        TileIter it;
        it.init(p, wk);
        while (it.next(p)) {
          if (it.flag >= 0 && it.flag != k_flag_ok) { wait_flag(p, it.flag); k_flag_ok = it.flag; }
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            const uint32_t slot = kr.idx;
            const uint32_t par = kr.phase;
            mbar_wait(B.kv_empty + 8 * slot, par ^ 1);
            mbar_arrive_expect_tx(B.kv_full + 8 * slot, C::TILE_BYTES);
#pragma unroll
            for (int db = 0; db < C::DBLK; ++db)
              tma_load_4d(smem + C::OFF_KV + slot * C::TILE_BYTES + db * C::BLK_BYTES,
                          kv == 0 ? &p.tm_k : &p.tm_v, B.kv_full + 8 * slot, db * 64, hk,
                          it.k_row0, wk.b);
            kr.advance();
          }
        }
      }
    }
   } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer (whole warp, elected lane issues)
    {
      // This warp's instruction stream is on the critical path of every K/V tile (round-2 ncu: the softmax warpgroups
      // waited on s_full 44 % of the time with the tensor pipe at 50 %): descriptors are built once and stepped with
      // one add on their low word, ring slots / phases are counters, the tile iterator is the range form.
      constexpr uint32_t idesc_qk = make_idesc_f16(kBf16 ? 1 : 0, BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(kBf16 ? 1 : 0, BM, kD, 0, 1);
      constexpr uint32_t kStage16 = C::TILE_BYTES >> 4;
      const uint64_t dq0 = make_sw128_desc(smem + C::OFF_Q, 16, 1024);                   // Q tile 0 (tile 1: + TILE_BYTES)
      const uint64_t dkk = make_sw128_desc(smem + C::OFF_KV, 16, 1024);                  // K/V slot 0 as the K-major B operand
      const uint64_t dvv = make_sw128_desc(smem + C::OFF_KV, C::BLK_BYTES, 1024);        // ... as the MN-major B operand
      uint32_t qc[2] = {0, 0}, pc[2] = {0, 0};
      Ring<C::STAGES> kr;
      auto issue_qk = [&](int t, uint32_t kslot) {
        const uint32_t ko = kslot * kStage16;
        const uint32_t qo = t * kStage16;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t off = ((kk >> 2) * C::BLK_BYTES + (kk & 3) * 32) >> 4;
          mma_ss(tmem + C::TMEM_S + t * 128, desc_step(dq0, qo + off), desc_step(dkk, ko + off), idesc_qk, kk > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, uint32_t vslot, bool acc) {
        const uint32_t vo = vslot * kStage16;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          mma_ts(tmem + C::TMEM_O + t * kD, tmem + C::TMEM_S + t * 128 + kk * 8, desc_step(dvv, vo + ((kk * 2048) >> 4)),
                 idesc_pv, (acc || kk > 0) ? 1u : 0u);
        }
      };
      for (int round = 0;; ++round) {
        Work wk;
        if (!decode_work(p, consumer_next(round), wk)) break;
        const int nt = wk.ntile;
        TileIter it;
        it.init(p, wk);
        bool have = it.next(p);
        // every loop over the two Q tiles is unrolled with a compile-time tile index: barrier addresses, TMEM columns and
        // descriptor offsets stay in uniform registers (a runtime index cost an R2UR + ELECT per tcgen05.mma: 106 R2UR
        // per K/V tile pair in the round-2 profile)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < nt) {
            mbar_wait(B.q_full[t], qc[t] & 1);
            ++qc[t];
          }
        }
        if (!have) {
#pragma unroll
          for (int t = 0; t < 2; ++t)
            if (t < nt) mma_commit(B.o_full[t]);      // empty item: hand the Q tiles back
          continue;
        }
        // first tile: S_t = Q_t K_0^T
        uint32_t kslot = kr.idx;
        mbar_wait(B.kv_full + 8 * kslot, kr.phase);
        kr.advance();
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < nt) {
            issue_qk(t, kslot);
            mma_commit(B.s_full[t]);
          }
        }
        mma_commit(B.kv_empty + 8 * kslot);
        for (int j = 0;; ++j) {
          const uint32_t vslot = kr.idx;
          const uint32_t vpar = kr.phase;
          kr.advance();
          const bool have_next = it.next(p);
          uint32_t kpar = 0;
          if (have_next) {
            kslot = kr.idx;
            kpar = kr.phase;
            kr.advance();
          }
          mbar_wait(B.kv_full + 8 * vslot, vpar);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t < nt) {
              mbar_wait(B.p_full[t], pc[t] & 1);
              ++pc[t];
              tc_fence_after();
              issue_pv(t, vslot, j > 0);
              if (t == nt - 1) mma_commit(B.kv_empty + 8 * vslot);
              if (have_next) {
                if (t == 0) {
                  mbar_wait(B.kv_full + 8 * kslot, kpar);
                  tc_fence_after();
                }
                issue_qk(t, kslot);
                mma_commit(B.s_full[t]);
                if (t == nt - 1) mma_commit(B.kv_empty + 8 * kslot);
              } else {
                mma_commit(B.o_full[t]);
              }
            }
          }
          if (!have_next) break;
        }
      }
    }
   }
  } else {
    setmaxnreg_inc<200>();
    // =========================================================== softmax / epilogue warpgroups
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + C::TMEM_S + t * 128;
    const uint32_t tO = tmem + lane_base + C::TMEM_O + t * kD;
    uint32_t sc = 0, oc = 0;
    const bool plain = (p.softcap == 0.f) && (p.alibi == nullptr);
    for (int round = 0;; ++round) {
      Work wk;
      if (!decode_work(p, consumer_next(round), wk)) break;
      if (t >= wk.ntile) continue;
      const int qpos = wk.pos0 + (t * BM + row) * p.q_pos_stride;
      const int qlo_t = wk.pos0 + t * BM * p.q_pos_stride;
      const int qhi_t = qlo_t + (BM - 1) * p.q_pos_stride;
      const float slope = p.alibi ? p.alibi[wk.b * p.alibi_bstride + wk.h] : 0.f;
      float m = -INFINITY;   // running max, log2 domain (already multiplied by scale*log2e)
      float l = 0.f;
      TileIter it;
      it.init(p, wk);
      int j = 0;
      while (it.next(p)) {
        mbar_wait(B.s_full[t], sc & 1);
        ++sc;
        tc_fence_after();
        const int kb = it.kpos0 + (it.nvalid - 1) * p.k_pos_stride;
        const bool need_mask = (it.nvalid < BN) || (p.wr >= 0 && kb - qlo_t > p.wr) ||
                               (p.wl >= 0 && qhi_t - it.kpos0 > p.wl);
        float mx = -INFINITY;
        float mul = p.scale_log2;   // multiplier applied to the TMEM value inside exp2
        const bool general = !plain || need_mask;
        if (general) {
          // general path: rewrite S in TMEM as log2-domain logits with scale / softcap / ALiBi /
          // position masks applied, 32 columns at a time (keeps the register footprint small)
          for (int c = 0; c < 4; ++c) {
            uint32_t u[32];
            tmem_ld32(tS + c * 32, u);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x = __uint_as_float(u[i]) * p.scale;
              if (p.softcap > 0.f) x = p.softcap * tanh_approx(x / p.softcap);
              const int col = c * 32 + i;
              const int rel = it.kpos0 + col * p.k_pos_stride - qpos;
              if (p.alibi) x -= slope * fabsf(static_cast<float>(rel));
              x *= 1.4426950408889634f;
              const bool masked = (col >= it.nvalid) || (p.wr >= 0 && rel > p.wr) || (p.wl >= 0 && -rel > p.wl);
              x = masked ? -INFINITY : x;
              u[i] = __float_as_uint(x);
              mx = fmaxf(mx, x);
            }
            tmem_st32(tS + c * 32, u);
          }
          tmem_wait_st();
          mul = 1.f;
        }
        uint32_t v[128];
        if constexpr (kPk) {
          if (!general) {
            // the row max of the first 64 columns overlaps the TMEM load of the last 64 (tcgen05.wait::ld is all-or-nothing)
            tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tmem_wait_ld();
            tmem_ld32(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
            tmem_ld32(tS + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
            // four independent chains: one serial FMNMX3 chain over 128 values was ~300 clk of pure dependency latency
            // on the critical path of every tile (two warps per scheduler cannot hide it)
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 64; c += 8) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                m4[i] = fmaxf(m4[i], fmaxf(__uint_as_float(v[c + 2 * i]), __uint_as_float(v[c + 2 * i + 1])));
            }
            tmem_wait_ld();
#pragma unroll
            for (int c = 64; c < 128; c += 8) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                m4[i] = fmaxf(m4[i], fmaxf(__uint_as_float(v[c + 2 * i]), __uint_as_float(v[c + 2 * i + 1])));
            }
            mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
            tmem_wait_ld();
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
          tmem_wait_ld();
          if (!general) {
#pragma unroll
            for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(v[c]));
            mx *= p.scale_log2;
          }
        }
        // ---- running max with lazy rescale
        const float m_new = fmaxf(m, mx);
        bool need = (m_new - m > kRescaleThreshold) || (m == -INFINITY && m_new > -INFINITY);
        if (j == 0) {
          m = m_new;
        } else if (__any_sync(0xffffffffu, need)) {
          float alpha = 1.f;
          if (need) {
            alpha = (m == -INFINITY) ? 0.f : ex2(m - m_new);   // old O rows are exact zeros if m == -inf
            m = m_new;
          }
          l *= alpha;
#pragma unroll
          for (int c = 0; c < kD / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + c * 32, o);
          }
        }
        const float sub = (m == -INFINITY) ? 0.f : m;
        // ---- P = exp2(x*mul - m), row sum, pack, store over S
        float rs = 0.f;
        if constexpr (kDrop) {
          // P = exp2(.), row sum over the UNdropped probabilities, then zero the dropped scores before they feed PV;
          // the 1/(1-p) rescale is folded into the final normalisation.  One hash word covers four key positions.
          const uint32_t rkey = ptx::dropout_row_key(static_cast<uint32_t>(qpos), p.drop_seed,
                                                     static_cast<uint32_t>(wk.b + p.qseg[wk.qseg].group),
                                                     static_cast<uint32_t>(wk.h + p.drop_head_off));
          const uint32_t p8 = static_cast<uint32_t>(p.drop_p8);
          if (p.k_pos_stride == 1 && (it.kpos0 & 3) == 0) {
#pragma unroll
            for (int c = 0; c < 128; c += 4) {
              const uint32_t w = ptx::dropout_word(rkey, static_cast<uint32_t>(it.kpos0 + c));
              float e[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) e[i] = ex2(fmaf(__uint_as_float(v[c + i]), mul, -sub));
              rs += (e[0] + e[1]) + (e[2] + e[3]);
#pragma unroll
              for (int i = 0; i < 4; ++i) e[i] = (((w >> (8 * i)) & 0xFFu) >= p8) ? e[i] : 0.f;
              v[c >> 1] = pack2<kBf16>(e[0], e[1]);
              v[(c >> 1) + 1] = pack2<kBf16>(e[2], e[3]);
            }
          } else {        // strided (stripe) or unaligned segments: one hash per score
#pragma unroll
            for (int c = 0; c < 128; c += 2) {
              const uint32_t k0 = static_cast<uint32_t>(it.kpos0 + c * p.k_pos_stride);
              const uint32_t k1 = static_cast<uint32_t>(it.kpos0 + (c + 1) * p.k_pos_stride);
              float p0 = ex2(fmaf(__uint_as_float(v[c]), mul, -sub));
              float p1 = ex2(fmaf(__uint_as_float(v[c + 1]), mul, -sub));
              rs += p0 + p1;
              p0 = ptx::dropout_keep(ptx::dropout_word(rkey, k0), k0, p8) ? p0 : 0.f;
              p1 = ptx::dropout_keep(ptx::dropout_word(rkey, k1), k1, p8) ? p1 : 0.f;
              v[c >> 1] = pack2<kBf16>(p0, p1);
            }
          }
        } else if (general) {
          if constexpr (kPk) {
            const uint64_t mul2 = ptx::pack_f32x2(mul, mul), nsub2 = ptx::pack_f32x2(-sub, -sub);
            uint64_t acc_a = ptx::pack_f32x2(0.f, 0.f), acc_b = acc_a;
#pragma unroll
            for (int c = 0; c < 128; c += 2) {
              float x0, x1;
              ptx::unpack_f32x2(
                  ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), mul2, nsub2), x0, x1);
              const float p0 = ex2(x0), p1 = ex2(x1);      // masked entries are -inf -> exactly 0
              if (c & 2) acc_b = ptx::add_f32x2(acc_b, ptx::pack_f32x2(p0, p1));
              else acc_a = ptx::add_f32x2(acc_a, ptx::pack_f32x2(p0, p1));
              v[c >> 1] = pack2<kBf16>(p0, p1);
            }
            float s0, s1;
            ptx::unpack_f32x2(ptx::add_f32x2(acc_a, acc_b), s0, s1);
            rs = s0 + s1;
          } else {
#pragma unroll
            for (int c = 0; c < 128; c += 2) {
              const float p0 = ex2(fmaf(__uint_as_float(v[c]), mul, -sub));
              const float p1 = ex2(fmaf(__uint_as_float(v[c + 1]), mul, -sub));
              rs += p0 + p1;
              v[c >> 1] = pack2<kBf16>(p0, p1);
            }
          }
        } else if constexpr (kPk) {
          const uint64_t mul2 = ptx::pack_f32x2(mul, mul), nsub2 = ptx::pack_f32x2(-sub, -sub);
          uint64_t acc_a = ptx::pack_f32x2(0.f, 0.f), acc_b = acc_a;     // two chains: the adds of a row are serial
#pragma unroll
          for (int c = 0; c < 128; c += 2) {
            const uint64_t x =
                ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), mul2, nsub2);
            float p0, p1;
            if (kPolyEvery > 0 && ((c >> 1) % (kPolyEvery > 0 ? kPolyEvery : 1)) == 0) {
              ptx::ex2_poly_x2(x, p0, p1);
            } else {
              float x0, x1;
              ptx::unpack_f32x2(x, x0, x1);
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            if (c & 2) acc_b = ptx::add_f32x2(acc_b, ptx::pack_f32x2(p0, p1));
            else acc_a = ptx::add_f32x2(acc_a, ptx::pack_f32x2(p0, p1));
            v[c >> 1] = pack2<kBf16>(p0, p1);
            // columns 0-63 are final: their 32 packed words go back to TMEM while the second half is computed
            if (c == 62) tmem_st32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          }
          float s0, s1;
          ptx::unpack_f32x2(ptx::add_f32x2(acc_a, acc_b), s0, s1);
          rs = s0 + s1;
        } else {
          // unmasked tiles (the bulk of the work): every kPolyEvery-th pair takes the FMA-pipe exp2
#pragma unroll
          for (int c = 0; c < 128; c += 2) {
            const float x0 = fmaf(__uint_as_float(v[c]), mul, -sub);
            const float x1 = fmaf(__uint_as_float(v[c + 1]), mul, -sub);
            float p0, p1;
            if (kPolyEvery > 0 && ((c >> 1) % (kPolyEvery > 0 ? kPolyEvery : 1)) == 0) {
              p0 = ex2_poly(x0);
              p1 = ex2_poly(x1);
            } else {
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            rs += p0 + p1;
            v[c >> 1] = pack2<kBf16>(p0, p1);
          }
        }
        l += rs;
        if constexpr (kPk && !kDrop) {
          if (general) tmem_st32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));   // unmasked tiles stored it mid-loop
          tmem_st32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        } else {
          tmem_st32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_st32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.p_full[t]);
        ++j;
      }
      // ---- epilogue: O / l -> 16-bit -> smem (XOR-swizzled 16B chunks) -> coalesced global stores
      uint8_t* stage = smem_gen + C::OFF_Q + t * C::TILE_BYTES;
      mbar_wait(B.o_full[t], oc & 1);     // committed by the MMA warp for empty items too: Q smem is reusable
      ++oc;
      tc_fence_after();
      float inv = (l > 0.f) ? 1.f / l : 0.f;
      if constexpr (kDrop) inv *= p.drop_rscale;
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t o[32];
        if (j > 0) {
          tmem_ld32(tO + c * 32, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // 4 chunks of 8 elements (16 bytes)
          uint4 w;
          w.x = pack2<kBf16>(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          w.y = pack2<kBf16>(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          w.z = pack2<kBf16>(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          w.w = pack2<kBf16>(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          const int chunk = c * 4 + g;
          *reinterpret_cast<uint4*>(stage + row * (kD * 2) + ((chunk ^ (row & 7)) << 4)) = w;
        }
      }
      tc_fence_before();
      const int rows_t = min(BM, wk.nrows - t * BM);
      if (row < rows_t) {
        const float lse = (l > 0.f) ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
        p.lse[wk.b * p.lse_sb + wk.h * p.lse_sh + wk.row0 + t * BM + row] = lse;
        float* lown = p.qseg[wk.qseg].lse_base;     // the token owner keeps the LSE too (used by the fused backward)
        if (lown != nullptr)
          lown[wk.b * p.lse_own_sb + static_cast<int64_t>(wk.h + p.o_head_off) * p.lse_own_sh + p.qseg[wk.qseg].o_row0 +
               wk.seg_row0 + t * BM + row] = lse;
      }
      named_bar_sync(1 + t, 128);
      {
        constexpr int LPR = kD / 8;          // lanes per row (16-byte chunks per row)
        constexpr int RPI = 32 / LPR;        // rows per warp instruction
        const QSegD qs = p.qseg[wk.qseg];
        uint8_t* obase = reinterpret_cast<uint8_t*>(qs.o_base) +
                         2 * (wk.b * p.o_sb + static_cast<int64_t>(wk.h + p.o_head_off) * p.o_sh);
        const int64_t orow0 = static_cast<int64_t>(qs.o_row0) + wk.seg_row0 + t * BM;
        const int chunk = lane % LPR;
#pragma unroll 4
        for (int i = 0; i < BM / (4 * RPI); ++i) {
          const int r = i * 4 * RPI + (warp & 3) * RPI + lane / LPR;
          if (r < rows_t) {
            const uint4 w = *reinterpret_cast<const uint4*>(stage + r * (kD * 2) + ((chunk ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(obase + 2 * (orow0 + r) * p.o_ss + chunk * 16) = w;
          }
        }
        if (qs.o_sig != nullptr) __threadfence_system();
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if ((warp & 3) == 0 && lane == 0) {
        mbar_arrive(B.q_empty[t]);
        uint32_t* sig = p.qseg[wk.qseg].o_sig;
        if (sig != nullptr) red_add_release_sys(sig, 1u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
  if constexpr (kDyn) {
    // the push CTAs joined the compute pool, so the "my output buffer is complete" wait moved to the end of the kernel
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.comm.n_comm > 0 && p.comm.o_target != 0)
      spin_until_ge(p.comm.my_sig + kSigODone, p.comm.o_target, 64, p.comm.watchdog_ns);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int kD, bool kBf16, int kPoly, bool kDyn, bool kDrop = false>
static cudaError_t launch_impl(const FwdParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<kD>;
  auto kern = fmha_fwd_kernel<kD, kBf16, kPoly, kDyn, kDrop>;
  // fused launches: the push CTAs stage their bulk copies in the same dynamic shared memory (usp_comm.cuh)
  constexpr int kSmem = C::SMEM_BYTES > kPushSmemBytes ? C::SMEM_BYTES : kPushSmemBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int avail = num_sms - p.comm.n_comm;
  int grid = p.total_work < avail ? p.total_work : avail;
  if (grid < 1) grid = 1;
  grid += p.comm.n_comm;       // comm CTAs first; all CTAs are co-resident (1 CTA/SM, grid <= #SMs)
  kern<<<grid, kThreads, p.comm.n_comm > 0 ? kSmem : C::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

template <int kD, bool kBf16, bool kDyn>
static cudaError_t launch_sched(const FwdParams& p, int num_sms, cudaStream_t stream) {
  switch (p.poly_every) {       // exp2 offload ratio: 0 = MUFU only; 3 (head_dim 64 default), 4 (head_dim 128 default), 6
    case 0: return launch_impl<kD, kBf16, 0, kDyn>(p, num_sms, stream);
    case 3: return launch_impl<kD, kBf16, 3, kDyn>(p, num_sms, stream);
    case 6: return launch_impl<kD, kBf16, 6, kDyn>(p, num_sms, stream);
    default: return launch_impl<kD, kBf16, 4, kDyn>(p, num_sms, stream);
  }
}

template <int kD, bool kBf16>
static cudaError_t launch_poly(const FwdParams& p, int num_sms, cudaStream_t stream) {
  if (p.drop_p8 > 0) return launch_impl<kD, kBf16, 0, false, true>(p, num_sms, stream);   // dropout: static schedule
  return p.dyn_sched ? launch_sched<kD, kBf16, true>(p, num_sms, stream) : launch_sched<kD, kBf16, false>(p, num_sms, stream);
}

cudaError_t launch_fmha_fwd(const FwdParams& p, int head_dim, bool bf16, int num_sms, cudaStream_t stream) {
  if (head_dim == 128) return bf16 ? launch_poly<128, true>(p, num_sms, stream) : launch_poly<128, false>(p, num_sms, stream);
  if (head_dim == 64) return bf16 ? launch_poly<64, true>(p, num_sms, stream) : launch_poly<64, false>(p, num_sms, stream);
  return cudaErrorInvalidValue;
}

}  // namespace lca
