// Forward flash attention for sm_100a, 64-row K/V tiles with DOUBLE-BUFFERED scores -- EXPERIMENTAL
// (compile-checked, opt-in: LCA_B200_FWD_BN64=1; hardware validation is the first GPU call of the next round).
//
// Same roles, data model, masks and epilogue as fmha_fwd_sm100.cu (this file is derived from it).  What changes is the
// pipeline between the tensor pipe and the softmax warpgroups:
//   * fmha_fwd_sm100.cu : one S buffer (128 columns) per Q tile; P overwrites S, so QK(j+1) of a tile can only be
//     issued after PV(j) of that tile -> per-tile chain  QK 512 clk -> softmax W -> PV 512 clk.  The softmax of one tile
//     has exactly the other tile's 1024 clk of MMA to hide behind; measured W ~ 1600-2200 clk => tensor pipe 0.5-0.63.
//   * here              : K/V tiles of 64 rows; the 128 score columns of a Q tile hold TWO stages of 64.  QK(j+1) and
//     QK(j+2) are in flight while the warpgroup works on S(j): the tensor pipe always has queued work and only stalls if
//     the softmax THROUGHPUT (not latency) falls behind: per 64 columns ~800-1000 clk of softmax against 1024 clk of MMA
//     for both tiles.
// Consequence: PV(j) of a tile may still be accumulating into O when its warpgroup looks at S(j+1).  The (rare, lazy)
// rescale of O therefore waits for a per-tile `pv_done` barrier that the MMA warp commits after every PV.
//
// Capability parity: same as fmha_fwd_sm100.cu (replaces flash_attn::_flash_attn_forward,
// yunchang/kernels/attention.py:165-203).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fmha_params.h"
#include "sm100_ptx.cuh"
#include "usp_comm.cuh"

namespace lca {
using namespace ptx;

namespace {

constexpr int BM = 128;        // query rows per tile (= TMEM lanes)
constexpr int BN = 64;         // key rows per tile (= columns of one score stage)
constexpr int kThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr float kRescaleThreshold = 8.0f;

template <int kD>
struct Cfg {
  static constexpr int DBLK = kD / 64;
  static constexpr int QBLK_BYTES = BM * 128;              // [128 rows][64 elem]
  static constexpr int QTILE_BYTES = DBLK * QBLK_BYTES;
  static constexpr int KVBLK_BYTES = BN * 128;             // [64 rows][64 elem]
  static constexpr int KVTILE_BYTES = DBLK * KVBLK_BYTES;
  static constexpr int STAGES = (kD == 128) ? 10 : 12;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = 2 * QTILE_BYTES;
  static constexpr int OFF_BAR = OFF_KV + STAGES * KVTILE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int TMEM_S = 0;                         // S_t stage s at column t*128 + s*64
  static constexpr int TMEM_O = 256;                       // O_t at column 256 + t*kD
};

struct Work {
  int qseg, seg_row0, row0, nrows, pos0, b, h, ntile;
};

__device__ __forceinline__ bool decode_work(const FwdParams& p, int w, Work& wk) {
  if (w >= p.total_work) return false;
  const int bh = p.B * p.H;
  int pr = w / bh;
  const int r = w - pr * bh;
  wk.b = r / p.H;
  wk.h = r - wk.b * p.H;
  for (int s = 0; s < p.n_qseg; ++s) {
    const int np = (p.qseg[s].nrows + 2 * BM - 1) / (2 * BM);
    if (pr < np) {
      const int pi = np - 1 - pr;  // heaviest (latest positions) first
      wk.qseg = s;
      wk.seg_row0 = pi * 2 * BM;
      wk.row0 = p.qseg[s].row0 + wk.seg_row0;
      wk.nrows = min(2 * BM, p.qseg[s].nrows - wk.seg_row0);
      wk.pos0 = p.qseg[s].pos0 + wk.seg_row0 * p.q_pos_stride;
      wk.ntile = wk.nrows > BM ? 2 : 1;
      return true;
    }
    pr -= np;
  }
  return false;
}

__device__ __forceinline__ int sched_work(int round, int n_comm) {
  const int G = static_cast<int>(gridDim.x) - n_comm;
  const int me = static_cast<int>(blockIdx.x) - n_comm;
  const int c = (round & 1) ? (G - 1 - me) : me;
  return round * G + c;
}

// K/V tiles (64 rows) a Q pair has to visit; identical in every role
struct TileIter {
  int seg, kt;
  int qmin, qmax, qgroup;
  int k_row0, nvalid, kpos0, flag;
  __device__ __forceinline__ void init(const FwdParams& p, const Work& wk) {
    seg = 0;
    kt = -1;
    qmin = wk.pos0;
    qmax = wk.pos0 + (wk.nrows - 1) * p.q_pos_stride;
    qgroup = p.qseg[wk.qseg].group;
  }
  __device__ __forceinline__ bool next(const FwdParams& p) {
    while (seg < p.n_kseg) {
      const KSegD s = p.kseg[seg];
      const int nt = (s.group == qgroup) ? (s.nrows + BN - 1) / BN : 0;
      while (++kt < nt) {
        const int r0 = kt * BN;
        const int nv = min(BN, s.nrows - r0);
        const int ka = s.pos0 + r0 * p.k_pos_stride;
        const int kb = ka + (nv - 1) * p.k_pos_stride;
        if (p.wr >= 0 && ka - qmax > p.wr) break;
        if (p.wl >= 0 && qmin - kb > p.wl) continue;
        k_row0 = s.row0 + r0;
        nvalid = nv;
        kpos0 = ka;
        flag = s.flag;
        return true;
      }
      ++seg;
      kt = -1;
    }
    return false;
  }
};

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

}  // namespace

// kPolyEvery: 1 of every kPolyEvery element pairs of an unmasked tile uses ex2_poly (0 = MUFU only)
// kPk: packed fp32x2 softmax arithmetic (FFMA2 / FADD2), same code as the kPk variant of fmha_fwd_sm100.cu on 64 columns
template <int kD, bool kBf16, int kPolyEvery, bool kPk>
__global__ void __launch_bounds__(kThreads, 1) fmha_fwd_bn64_kernel(const __grid_constant__ FwdParams p) {
  using C = Cfg<kD>;
  if (static_cast<int>(blockIdx.x) < p.comm.n_comm) {   // communication role (fused USP path)
    comm_cta<false>(p.comm);
    return;
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  // ---- barriers: q_full[2] q_empty[2] o_full[2] pv_done[2] s_full[2][2] p_full[2][2] kv_full[STAGES] kv_empty[STAGES]
  const uint32_t bar0 = smem + C::OFF_BAR;
  const uint32_t q_full = bar0, q_empty = bar0 + 16, o_full = bar0 + 32, pv_done = bar0 + 48;
  const uint32_t s_full = bar0 + 64;      // + 8 * (2*t + stage)
  const uint32_t p_full = bar0 + 96;      // + 8 * (2*t + stage)
  const uint32_t kv_full = bar0 + 128, kv_empty = kv_full + 8 * C::STAGES;
  static_assert(128 + 16 * C::STAGES <= 480, "barrier area");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + C::OFF_BAR + 480);

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full + 8 * t, 1);
      mbar_init(q_empty + 8 * t, 1);
      mbar_init(o_full + 8 * t, 1);
      mbar_init(pv_done + 8 * t, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(s_full + 8 * (2 * t + s), 1);
        mbar_init(p_full + 8 * (2 * t + s), 4);   // one arrive per softmax warp
      }
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(kv_full + 8 * s, 1);
      mbar_init(kv_empty + 8 * s, 1);
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
      uint32_t kvc = 0;                 // cumulative K/V tile loads: K(j), V(j) of a work item = base + 2j, base + 2j + 1
      int q_flag_ok = -1, k_flag_ok = -1;
      for (int round = 0;; ++round) {
        Work wk;
        if (!decode_work(p, sched_work(round, p.comm.n_comm), wk)) break;
        const int qf = p.qseg[wk.qseg].flag;
        if (qf >= 0 && qf != q_flag_ok) { wait_arrival(p.flags, p.flag_epoch, qf); q_flag_ok = qf; }
        for (int t = 0; t < wk.ntile; ++t) {
          mbar_wait(q_empty + 8 * t, (qc[t] & 1) ^ 1);
          mbar_arrive_expect_tx(q_full + 8 * t, C::QTILE_BYTES);
#pragma unroll
          for (int db = 0; db < C::DBLK; ++db)
            tma_load_4d(smem + C::OFF_Q + t * C::QTILE_BYTES + db * C::QBLK_BYTES, &p.tm_q, q_full + 8 * t, db * 64, wk.h,
                        wk.row0 + t * BM, wk.b);
          ++qc[t];
        }
        const int hk = wk.h / hk_div;
        TileIter it;
        it.init(p, wk);
        while (it.next(p)) {
          if (it.flag >= 0 && it.flag != k_flag_ok) { wait_arrival(p.flags, p.flag_epoch, it.flag); k_flag_ok = it.flag; }
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            const uint32_t slot = kvc % C::STAGES;
            const uint32_t par = (kvc / C::STAGES) & 1;
            mbar_wait(kv_empty + 8 * slot, par ^ 1);
            mbar_arrive_expect_tx(kv_full + 8 * slot, C::KVTILE_BYTES);
#pragma unroll
            for (int db = 0; db < C::DBLK; ++db)        // tm_k / tm_v carry 64-row boxes for this kernel
              tma_load_4d(smem + C::OFF_KV + slot * C::KVTILE_BYTES + db * C::KVBLK_BYTES, kv == 0 ? &p.tm_k : &p.tm_v,
                          kv_full + 8 * slot, db * 64, hk, it.k_row0, wk.b);
            ++kvc;
          }
        }
      }
    }
   } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer (whole warp, elected lane issues)
    {
      constexpr uint32_t idesc_qk = make_idesc_f16(kBf16 ? 1 : 0, BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(kBf16 ? 1 : 0, BM, kD, 0, 1);
      uint32_t qc[2] = {0, 0};
      uint32_t pcbits = 0;      // p_full phase parity, bit 2*t + stage (a dynamically indexed array would leave the
                                // uniform datapath and cost ~20 instructions per UTCHMMA instead of ~5)
      uint32_t kvc = 0;
      auto issue_qk = [&](int t, int s, uint32_t kslot) {
        const uint32_t qa = smem + C::OFF_Q + t * C::QTILE_BYTES;
        const uint32_t ka = smem + C::OFF_KV + kslot * C::KVTILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t qo = (kk >> 2) * C::QBLK_BYTES + (kk & 3) * 32;
          const uint32_t ko = (kk >> 2) * C::KVBLK_BYTES + (kk & 3) * 32;
          mma_ss(tmem + C::TMEM_S + t * 128 + s * BN, make_sw128_desc(qa + qo, 16, 1024), make_sw128_desc(ka + ko, 16, 1024),
                 idesc_qk, kk > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int s, uint32_t vslot, bool acc) {
        const uint32_t va = smem + C::OFF_KV + vslot * C::KVTILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          mma_ts(tmem + C::TMEM_O + t * kD, tmem + C::TMEM_S + t * 128 + s * BN + kk * 8,
                 make_sw128_desc(va + kk * 2048, C::KVBLK_BYTES, 1024), idesc_pv, (acc || kk > 0) ? 1u : 0u);
        }
      };
      for (int round = 0;; ++round) {
        Work wk;
        if (!decode_work(p, sched_work(round, p.comm.n_comm), wk)) break;
        const int nt = wk.ntile;
        TileIter it;
        it.init(p, wk);
        bool more = it.next(p);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(q_full + 8 * t, qc[t] & 1);
          ++qc[t];
        }
        if (!more) {
          // empty work item (no visible K/V tile): still hand the Q tiles back through o_full, so the warpgroups can
          // only release q_empty AFTER this warp has passed its q_full wait (a one-bit phase parity must never be
          // allowed to advance twice under a waiter)
          for (int t = 0; t < nt; ++t) mma_commit(o_full + 8 * t);
          continue;
        }
        const uint32_t base = kvc;
        int n_qk = 0;                               // score tiles issued so far
        // prologue: fill both score stages
        while (more && n_qk < 2) {
          const uint32_t idx = base + 2 * n_qk, slot = idx % C::STAGES;
          mbar_wait(kv_full + 8 * slot, (idx / C::STAGES) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) {
            issue_qk(t, n_qk, slot);
            mma_commit(s_full + 8 * (2 * t + n_qk));
          }
          mma_commit(kv_empty + 8 * slot);
          ++n_qk;
          more = it.next(p);
        }
        for (int j = 0; j < n_qk; ++j) {
          const int s = j & 1;
          const uint32_t vidx = base + 2 * j + 1, vslot = vidx % C::STAGES;
          const bool issue_next = more;             // score tile n_qk exists: it reuses stage s after PV(j)
          const uint32_t kidx = base + 2 * n_qk, kslot = kidx % C::STAGES;
          mbar_wait(kv_full + 8 * vslot, (vidx / C::STAGES) & 1);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(p_full + 8 * (2 * t + s), (pcbits >> (2 * t + s)) & 1u);
            pcbits ^= 1u << (2 * t + s);
            tc_fence_after();
            issue_pv(t, s, vslot, j > 0);
            mma_commit(pv_done + 8 * t);
            if (t == nt - 1) mma_commit(kv_empty + 8 * vslot);
            if (issue_next) {
              if (t == 0) {
                mbar_wait(kv_full + 8 * kslot, (kidx / C::STAGES) & 1);
                tc_fence_after();
              }
              issue_qk(t, s, kslot);
              mma_commit(s_full + 8 * (2 * t + s));
              if (t == nt - 1) mma_commit(kv_empty + 8 * kslot);
            } else if (j == n_qk - 1) {
              mma_commit(o_full + 8 * t);
            }
          }
          if (issue_next) {
            ++n_qk;
            more = it.next(p);
          }
        }
        kvc = base + 2 * n_qk;
      }
    }
   }
  } else {
    setmaxnreg_inc<200>();
    // =========================================================== softmax / epilogue warpgroups
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS0 = tmem + lane_base + C::TMEM_S + t * 128;
    const uint32_t tO = tmem + lane_base + C::TMEM_O + t * kD;
    uint32_t scbits = 0, oc = 0;             // s_full phase parity of this tile, bit = stage
    uint32_t pvc = 0;                    // PV MMAs of this Q tile committed before the current work item
    const bool plain = (p.softcap == 0.f) && (p.alibi == nullptr);
    for (int round = 0;; ++round) {
      Work wk;
      if (!decode_work(p, sched_work(round, p.comm.n_comm), wk)) break;
      if (t >= wk.ntile) continue;
      const int qpos = wk.pos0 + (t * BM + row) * p.q_pos_stride;
      const int qlo_t = wk.pos0 + t * BM * p.q_pos_stride;
      const int qhi_t = qlo_t + (BM - 1) * p.q_pos_stride;
      const float slope = p.alibi ? p.alibi[wk.b * p.alibi_bstride + wk.h] : 0.f;
      float m = -INFINITY;
      float l = 0.f;
      TileIter it;
      it.init(p, wk);
      int j = 0;
      while (it.next(p)) {
        const int s = j & 1;
        const uint32_t tS = tS0 + s * BN;
        mbar_wait(s_full + 8 * (2 * t + s), (scbits >> s) & 1u);
        scbits ^= 1u << s;
        tc_fence_after();
        const int kb = it.kpos0 + (it.nvalid - 1) * p.k_pos_stride;
        const bool need_mask = (it.nvalid < BN) || (p.wr >= 0 && kb - qlo_t > p.wr) ||
                               (p.wl >= 0 && qhi_t - it.kpos0 > p.wl);
        float mx = -INFINITY;
        float mul = p.scale_log2;
        const bool general = !plain || need_mask;
        if (general) {
          for (int c = 0; c < 2; ++c) {
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
        uint32_t v[64];
        tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_wait_ld();
        if (!general) {
#pragma unroll
          for (int c = 0; c < 64; ++c) mx = fmaxf(mx, __uint_as_float(v[c]));
          mx *= p.scale_log2;
        }
        // ---- running max with lazy rescale
        const float m_new = fmaxf(m, mx);
        bool need = (m_new - m > kRescaleThreshold) || (m == -INFINITY && m_new > -INFINITY);
        if (j == 0) {
          m = m_new;
        } else if (__any_sync(0xffffffffu, need)) {
          // PV(j-1) of this tile may still be accumulating into O (it was queued behind QK(j), QK(j+1)): wait for it.
          // Completed PV commits so far are pvc + j - 1 or pvc + j (PV(j) needs the p_full we have not given yet),
          // so the one-bit phase parity is unambiguous.
          mbar_wait(pv_done + 8 * t, (pvc + static_cast<uint32_t>(j) - 1u) & 1u);
          tc_fence_after();
          float alpha = 1.f;
          if (need) {
            alpha = (m == -INFINITY) ? 0.f : ex2(m - m_new);
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
        // ---- P = exp2(x*mul - m), row sum, pack, store over the first 32 columns of this stage
        float rs = 0.f;
        if constexpr (kPk) {
          const uint64_t mul2 = ptx::pack_f32x2(mul, mul), nsub2 = ptx::pack_f32x2(-sub, -sub);
          uint64_t acc_a = ptx::pack_f32x2(0.f, 0.f), acc_b = acc_a;
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            const uint64_t x =
                ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), mul2, nsub2);
            float p0, p1;
            // masked tiles carry -inf logits: the polynomial needs finite arguments, so it only runs on unmasked tiles
            if (kPolyEvery > 0 && ((c >> 1) % (kPolyEvery > 0 ? kPolyEvery : 1)) == 0 && !general) {
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
          }
          float s0, s1;
          ptx::unpack_f32x2(ptx::add_f32x2(acc_a, acc_b), s0, s1);
          rs = s0 + s1;
        } else if (general) {
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(v[c]), mul, -sub));
            const float p1 = ex2(fmaf(__uint_as_float(v[c + 1]), mul, -sub));
            rs += p0 + p1;
            v[c >> 1] = pack2<kBf16>(p0, p1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
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
        tmem_st32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + 8 * (2 * t + s));
        ++j;
      }
      pvc += static_cast<uint32_t>(j);
      // ---- epilogue: O / l -> 16-bit -> smem (XOR-swizzled 16B chunks) -> coalesced global stores
      uint8_t* stage = smem_gen + C::OFF_Q + t * C::QTILE_BYTES;
      mbar_wait(o_full + 8 * t, oc & 1);     // committed by the MMA warp for every work item, empty ones included
      ++oc;
      tc_fence_after();
      const float inv = (l > 0.f) ? 1.f / l : 0.f;
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
        for (int g = 0; g < 4; ++g) {
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
        float* lown = p.qseg[wk.qseg].lse_base;
        if (lown != nullptr)
          lown[wk.b * p.lse_own_sb + static_cast<int64_t>(wk.h + p.o_head_off) * p.lse_own_sh + p.qseg[wk.qseg].o_row0 +
               wk.seg_row0 + t * BM + row] = lse;
      }
      named_bar_sync(1 + t, 128);
      {
        constexpr int LPR = kD / 8;
        constexpr int RPI = 32 / LPR;
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
        mbar_arrive(q_empty + 8 * t);
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
template <int kD, bool kBf16, int kPoly, bool kPk = false>
static cudaError_t launch_impl(const FwdParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<kD>;
  auto kern = fmha_fwd_bn64_kernel<kD, kBf16, kPoly, kPk>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int avail = num_sms - p.comm.n_comm;
  int grid = p.total_work < avail ? p.total_work : avail;
  if (grid < 1) grid = 1;
  grid += p.comm.n_comm;
  kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

template <int kD, bool kBf16>
static cudaError_t launch_poly(const FwdParams& p, int num_sms, cudaStream_t stream) {
  if (p.f32x2) {
    switch (p.poly_every) {
      case 0: return launch_impl<kD, kBf16, 0, true>(p, num_sms, stream);
      case 3: return launch_impl<kD, kBf16, 3, true>(p, num_sms, stream);
      case 4: return launch_impl<kD, kBf16, 4, true>(p, num_sms, stream);
      default: return launch_impl<kD, kBf16, 6, true>(p, num_sms, stream);
    }
  }
  switch (p.poly_every) {
    case 0: return launch_impl<kD, kBf16, 0>(p, num_sms, stream);
    case 3: return launch_impl<kD, kBf16, 3>(p, num_sms, stream);
    case 4: return launch_impl<kD, kBf16, 4>(p, num_sms, stream);
    default: return launch_impl<kD, kBf16, 6>(p, num_sms, stream);
  }
}

// tm_k / tm_v of `p` must have been encoded with 64-row boxes (bindings.cpp does that when LCA_B200_FWD_BN64=1)
cudaError_t launch_fmha_fwd_bn64(const FwdParams& p, int head_dim, bool bf16, int num_sms, cudaStream_t stream) {
  if (head_dim == 128) return bf16 ? launch_poly<128, true>(p, num_sms, stream) : launch_poly<128, false>(p, num_sms, stream);
  if (head_dim == 64) return bf16 ? launch_poly<64, true>(p, num_sms, stream) : launch_poly<64, false>(p, num_sms, stream);
  return cudaErrorInvalidValue;
}

}  // namespace lca
