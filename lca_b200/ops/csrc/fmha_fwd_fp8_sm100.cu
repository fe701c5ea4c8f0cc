// FP8 (e4m3) forward flash attention for sm_100a (validated on B200 in round 2 against the PyTorch emulation of its
// arithmetic, tests/test_fp8.py; reached through AttnType.SAGE_FP8* / SAGE_AUTO, ops/fp8.py).
//
// Same warp-specialised structure as fmha_fwd_sm100.cu (this file is derived from it), with 1-byte operands:
//   * Q/K/V tiles are [128 rows][128 B] (head_dim 128 only), TMA box (128, 1, 128, 1) over a uint8 tensor map;
//   * tcgen05.mma kind::f8f6f4, K = 32 per instruction: 4 MMAs per QK^T tile, 4 per PV tile;
//   * block scaling: Q and K carry one fp32 scale per (batch, head, 128-row block) -- they fold exactly into the
//     softmax argument (s * sq_i * sk_j * scale); V carries one scale per (batch, kv head), applied in the epilogue;
//   * P is written back to TMEM as e4m3 (32 columns) and feeds the PV MMA as the TMEM A operand.
// Role of the reference's "FA3 fp8" / SAGE_FP8 forward-only paths (kernels/attention.py:258-292,
// kernels/__init__.py:177-254), which cannot run on sm_100 at all.
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
  static_assert(kD == 128, "fp8 path: head_dim 128 only");
  static constexpr int DBLK = kD / 128;                // 128-byte column blocks per row (1-byte elements)
  static constexpr int BLK_BYTES = 128 * 128;          // [128 rows][64 elem] sub-block
  static constexpr int TILE_BYTES = DBLK * BLK_BYTES;  // one Q / K / V tile
  static constexpr int STAGES = 10;                    // 16 KB tiles
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

// four fp32 -> four e4m3 (element 0 in the lowest byte)
__device__ __forceinline__ uint32_t pack4_e4m3(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}

__device__ __forceinline__ void mma_ss_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (elect_one()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void mma_ts_f8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (elect_one()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// kind::f8f6f4 instruction descriptor: e4m3 x e4m3 -> fp32 (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t make_idesc_f8(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
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
template <int kD, int kPolyEvery>
__global__ void __launch_bounds__(kThreads, 1) fmha_fwd_fp8_kernel(const __grid_constant__ FwdParams p) {
  using C = Cfg<kD>;
  constexpr bool kBf16 = true;                          // output dtype
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
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

  if (threadIdx.x == 0) {
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
        if (!decode_work(p, sched_work(round, 0), wk)) break;
        const int qf = p.qseg[wk.qseg].flag;
        if (qf >= 0 && qf != q_flag_ok) { wait_flag(p, qf); q_flag_ok = qf; }
        for (int t = 0; t < wk.ntile; ++t) {
          mbar_wait(B.q_empty[t], (qc[t] & 1) ^ 1);
          mbar_arrive_expect_tx(B.q_full[t], C::TILE_BYTES);
#pragma unroll
          for (int db = 0; db < C::DBLK; ++db)
            tma_load_4d(smem + C::OFF_Q + t * C::TILE_BYTES + db * C::BLK_BYTES, &p.tm_q,
                        B.q_full[t], db * 128, wk.h, wk.row0 + t * BM, wk.b);
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
                          kv == 0 ? &p.tm_k : &p.tm_v, B.kv_full + 8 * slot, db * 128, hk,
                          it.k_row0, wk.b);
            kr.advance();
          }
        }
      }
    }
   } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer (whole warp, elected lane issues)
    {
      // (same issue-path diet as fmha_fwd_sm100.cu: descriptors built once and stepped by one add, ring counters,
      //  compile-time tile index so every tcgen05.mma operand stays in uniform registers)
      constexpr uint32_t idesc_qk = make_idesc_f8(BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f8(BM, kD, 0, 1);
      constexpr uint32_t kStage16 = C::TILE_BYTES >> 4;
      const uint64_t dq0 = make_sw128_desc(smem + C::OFF_Q, 16, 1024);
      const uint64_t dkk = make_sw128_desc(smem + C::OFF_KV, 16, 1024);
      const uint64_t dvv = make_sw128_desc(smem + C::OFF_KV, C::BLK_BYTES, 1024);
      uint32_t qc[2] = {0, 0}, pc[2] = {0, 0};
      Ring<C::STAGES> kr;
      auto issue_qk = [&](int t, uint32_t kslot) {
        const uint32_t ko = kslot * kStage16, qo = t * kStage16;
#pragma unroll
        for (int kk = 0; kk < kD / 32; ++kk) {             // K = 32 one-byte elements per MMA
          const uint32_t off = ((kk >> 2) * C::BLK_BYTES + (kk & 3) * 32) >> 4;
          mma_ss_f8(tmem + C::TMEM_S + t * 128, desc_step(dq0, qo + off), desc_step(dkk, ko + off), idesc_qk, kk > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, uint32_t vslot, bool acc) {
        const uint32_t vo = vslot * kStage16;
#pragma unroll
        for (int kk = 0; kk < BN / 32; ++kk) {             // 32 key rows (= 8 TMEM columns of e4m3 P) per MMA
          mma_ts_f8(tmem + C::TMEM_O + t * kD, tmem + C::TMEM_S + t * 128 + kk * 8, desc_step(dvv, vo + ((kk * 4096) >> 4)),
                    idesc_pv, (acc || kk > 0) ? 1u : 0u);
        }
      };
      for (int round = 0;; ++round) {
        Work wk;
        if (!decode_work(p, sched_work(round, 0), wk)) break;
        const int nt = wk.ntile;
        TileIter it;
        it.init(p, wk);
        bool have = it.next(p);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < nt) {
            mbar_wait(B.q_full[t], qc[t] & 1);
            ++qc[t];
          }
        }
        if (!have) continue;
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
    uint32_t sc = 0, oc = 0, qc = 0;
    const bool plain = (p.softcap == 0.f) && (p.alibi == nullptr);
    for (int round = 0;; ++round) {
      Work wk;
      if (!decode_work(p, sched_work(round, 0), wk)) break;
      if (t >= wk.ntile) continue;
      const int qpos = wk.pos0 + (t * BM + row) * p.q_pos_stride;
      const int qlo_t = wk.pos0 + t * BM * p.q_pos_stride;
      const int qhi_t = qlo_t + (BM - 1) * p.q_pos_stride;
      const float slope = p.alibi ? p.alibi[wk.b * p.alibi_bstride + wk.h] : 0.f;
      // block scales: one per (batch, head, 128-row block); segments are 128-row aligned on this path
      const float sq = p.q_scale[wk.b * p.q_scale_sb + wk.h * p.q_scale_sh + ((wk.row0 + t * BM) >> 7)];
      const float sv = p.v_scale[wk.b * p.Hkv + wk.h / hk_div];
      const float* ksc = p.k_scale + wk.b * p.k_scale_sb + (wk.h / hk_div) * p.k_scale_sh;
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
        const float qk = sq * ksc[it.k_row0 >> 7];      // de-quantisation factor of this (Q block, K block) tile
        float mul = p.scale_log2 * qk;                  // multiplier applied to the TMEM value inside exp2
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
              float x = __uint_as_float(u[i]) * (p.scale * qk);
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
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
        tmem_wait_ld();
        if (!general) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // four independent chains
#pragma unroll
          for (int c = 0; c < 128; c += 8) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              m4[i] = fmaxf(m4[i], fmaxf(__uint_as_float(v[c + 2 * i]), __uint_as_float(v[c + 2 * i + 1])));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * mul;
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
        if (general) {
#pragma unroll
          for (int c = 0; c < 128; c += 4) {
            float pe[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) pe[e] = ex2(fmaf(__uint_as_float(v[c + e]), mul, -sub));
            rs += (pe[0] + pe[1]) + (pe[2] + pe[3]);
            v[c >> 2] = pack4_e4m3(pe[0], pe[1], pe[2], pe[3]);
          }
        } else {
          // unmasked tiles (the bulk of the work): every kPolyEvery-th pair takes the FMA-pipe exp2
#pragma unroll
          for (int c = 0; c < 128; c += 4) {
            float pe[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x = fmaf(__uint_as_float(v[c + e]), mul, -sub);
              pe[e] = (kPolyEvery > 0 && e < 2 && ((c >> 2) % (kPolyEvery > 0 ? kPolyEvery : 1)) == 0) ? ex2_poly(x) : ex2(x);
            }
            rs += (pe[0] + pe[1]) + (pe[2] + pe[3]);
            v[c >> 2] = pack4_e4m3(pe[0], pe[1], pe[2], pe[3]);
          }
        }
        l += rs;
        tmem_st32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));     // 128 e4m3 = 32 TMEM columns
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.p_full[t]);
        ++j;
      }
      // ---- epilogue: O * (sv / l) -> bf16, stored straight from registers (each thread owns one 256-byte output row).
      // NOTE: unlike the 16-bit kernel there is no smem staging here: the Q tile buffers are only 16 KB each, a
      // 32 KB bf16 O tile would not fit in them.
      if (j > 0) {
        mbar_wait(B.o_full[t], oc & 1);
        ++oc;
        tc_fence_after();
      } else {
        mbar_wait(B.q_full[t], qc & 1);   // keep the Q barrier phases in step even when nothing was consumed
      }
      ++qc;
      const float inv = (l > 0.f) ? sv / l : 0.f;        // V's per-head scale folds into the normalisation
      const int rows_t = min(BM, wk.nrows - t * BM);
      const QSegD qs = p.qseg[wk.qseg];
      uint8_t* orow = reinterpret_cast<uint8_t*>(qs.o_base) +
                      2 * (wk.b * p.o_sb + static_cast<int64_t>(wk.h + p.o_head_off) * p.o_sh +
                           (static_cast<int64_t>(qs.o_row0) + wk.seg_row0 + t * BM + row) * p.o_ss);
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
        if (row < rows_t) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {   // 4 chunks of 8 elements (16 bytes)
            uint4 w;
            w.x = pack2<kBf16>(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
            w.y = pack2<kBf16>(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
            w.z = pack2<kBf16>(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
            w.w = pack2<kBf16>(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + (c * 4 + g) * 16) = w;
          }
        }
      }
      tc_fence_before();
      if (row < rows_t) {
        const float lse = (l > 0.f) ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
        p.lse[wk.b * p.lse_sb + wk.h * p.lse_sh + wk.row0 + t * BM + row] = lse;
      }
      named_bar_sync(1 + t, 128);            // every thread has finished reading O from TMEM
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
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int kPoly>
static cudaError_t launch_impl(const FwdParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<128>;
  auto kern = fmha_fwd_fp8_kernel<128, kPoly>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int grid = p.total_work < num_sms ? p.total_work : num_sms;
  if (grid < 1) grid = 1;
  kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_fmha_fwd_fp8(const FwdParams& p, int head_dim, int num_sms, cudaStream_t stream) {
  if (head_dim != 128) return cudaErrorInvalidValue;
  return p.poly_every == 0 ? launch_impl<0>(p, num_sms, stream) : launch_impl<6>(p, num_sms, stream);
}

}  // namespace lca
