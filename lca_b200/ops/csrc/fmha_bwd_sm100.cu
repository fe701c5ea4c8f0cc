// Backward flash attention for sm_100a (B200): two tcgen05 passes sharing ONE kernel template.
//
//   pass dQ  (kIsDKV = false): stationary X = (Q_i, dO_i) tile of 128 query rows on the TMEM lanes,
//            streamed   Y = (K_j, V_j) tiles of 64 key rows.
//              T0 = Q K^T, T1 = dO V^T            (SS MMAs, fp32 in TMEM, NS = 2 or 3 stages)
//              P  = exp2(T0*scale*log2e - lse2_row), dS = P o (T1 - delta_row)   (softmax warpgroups)
//              dQ += dS K                          (TS MMA: A = dS from TMEM, B = K as MN-major smem)
//   pass dKV (kIsDKV = true):  stationary X = (K_j, V_j) tile of 128 key rows on the lanes,
//            streamed   Y = (Q_i, dO_i) tiles of 64 query rows (+ their lse2/delta columns),
//            looped over the G = H/Hkv query heads of the KV head (GQA reduces inside TMEM).
//              T0 = K Q^T (= S^T), T1 = V dO^T (= dP^T)
//              P^T, dS^T as above with per-COLUMN lse2/delta
//              dK += dS^T Q,  dV += P^T dO         (TS MMAs)
// Both passes are deterministic (no atomics: every output element is owned by one CTA) and use the
// same segment/global-position masking as the forward kernel; for the dKV pass the host swaps
// the window bounds because rows are keys and columns are queries.
// Seven GEMMs instead of the five of a fused dQ/dK/dV kernel, but nothing leaves TMEM between
// them, no fp32 dQ atomics cross L2, and TMEM (512 columns) is never oversubscribed:
//   [0, 64 NS) T0 stages | [64 NS, 128 NS) T1 stages | acc0 (D columns) | acc1 (dK/dV pass only); NS: see Cfg.
//
// Capability parity: flash_attn::_flash_attn_backward as called from
// yunchang/kernels/attention.py:205-250 (dq/dk/dv from dout,q,k,v,out,lse).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fmha_params.h"
#include "sm100_ptx.cuh"
#include "usp_comm.cuh"

namespace lca {
using namespace ptx;

namespace {

constexpr int BX = 128;   // stationary rows (TMEM lanes)
constexpr int kThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;

// TMEM budget (512 columns).  dK/dV pass: two accumulators (2 x D) leave room for NS = 2 stages of (T0, T1);
// dQ pass: one accumulator, so NS = 3 stages fit: the T GEMMs of tile j+2 are issued while tile j is still in the
// element-wise stage and tile j+1 waits for it -- the tensor pipe no longer idles for a whole
// "commit -> poll -> tcgen05.ld -> exp/dS -> tcgen05.st -> arrive -> poll" round trip per tile (measured r2 before the
// change: tensor pipe 48 % active, issue slots 32 %, i.e. latency-bound, not throughput-bound).
// dK/dV pass: two accumulators; at D = 128 only NS = 2 stages fit, at D = 64 three do.  (Measured and dropped in
// round 2: 32-row streamed tiles with NS = 4 -- 5.56 ms against 3.62 ms at S = 32K, the per-tile cost does not shrink
// with the tile, see profiles/README.md section 1.)
template <int kD, bool kIsDKV, int kBY>
struct Cfg {
  static constexpr int BY = kBY;                          // streamed rows per tile (TMEM columns of T0/T1)
  static constexpr int DBLK = kD / 64;
  static constexpr int XBLK_BYTES = BX * 128;             // [128 rows][64 elem]
  static constexpr int YBLK_BYTES = BY * 128;             // [BY rows][64 elem]
  static constexpr int XTILE_BYTES = DBLK * XBLK_BYTES;   // one stationary operand
  static constexpr int YTILE_BYTES = DBLK * YBLK_BYTES;   // one streamed operand
  static constexpr int NACC = kIsDKV ? 2 : 1;              // accumulators of kD columns
  static constexpr int NS_FIT = (512 - NACC * kD) / (2 * kBY);
  static constexpr int NS = NS_FIT < 3 ? NS_FIT : 3;      // (T0, T1) stages in TMEM
  static constexpr int STAGES = NS + 2;                   // streamed (Y0,Y1) tile pairs in flight (NS under MMA + loads)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_Y = 2 * XTILE_BYTES;
  static constexpr int OFF_STAT = OFF_Y + STAGES * 2 * YTILE_BYTES;   // dK/dV pass, per stage: lse2[64], delta[64]
  static constexpr int OFF_BAR = OFF_STAT + (kIsDKV ? STAGES * 2 * BY * 4 : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int TMEM_T0 = 0, TMEM_T1 = NS * BY, TMEM_ACC0 = 2 * NS * BY, TMEM_ACC1 = TMEM_ACC0 + kD;
  static_assert(TMEM_ACC0 + (kIsDKV ? 2 : 1) * kD <= 512, "TMEM columns");
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct Work {
  int xseg, seg_row0, row0, nrows, pos0, b, hx;
};

__device__ __forceinline__ bool decode_work(const BwdParams& p, int w, Work& wk) {
  if (w >= p.total_work) return false;
  const int bh = p.B * p.Hx;
  int t = w / bh;
  const int r = w - t * bh;
  wk.b = r / p.Hx;
  wk.hx = r - wk.b * p.Hx;
  for (int s = 0; s < p.n_xseg; ++s) {
    const int nt = (p.xseg[s].nrows + BX - 1) / BX;
    if (t < nt) {
      const int ti = p.x_heavy_last ? (nt - 1 - t) : t;
      wk.xseg = s;
      wk.seg_row0 = ti * BX;
      wk.row0 = p.xseg[s].row0 + wk.seg_row0;
      wk.nrows = min(BX, p.xseg[s].nrows - wk.seg_row0);
      wk.pos0 = p.xseg[s].pos0 + wk.seg_row0 * p.x_pos_stride;
      return true;
    }
    t -= nt;
  }
  return false;
}

__device__ __forceinline__ int sched_work(int round, int n_comm) {
  const int G = static_cast<int>(gridDim.x) - n_comm;
  const int me = static_cast<int>(blockIdx.x) - n_comm;
  const int c = (round & 1) ? (G - 1 - me) : me;
  return round * G + c;
}

// streamed-tile enumeration: (inner head gi) x (segment) x (BY-row tile), skipping tiles that are
// entirely masked for the stationary tile's position range.  Identical in every warp role.
// Positions grow with the tile index inside a segment, so the visible tiles of a segment are ONE contiguous range
// [lo, hi): it is computed when the iterator enters a segment (two integer divisions, rare); the per-tile step is a
// compare and three multiply-adds.  Round-2 profile: the single-thread roles (MMA issuer, TMA producer) -- not the
// tensor pipe, not the element-wise warpgroups -- bounded all three kernels, and the per-tile window tests + the
// constant-bank reload of the segment record were a third of their instruction stream
// (tests/test_properties_cpu.py checks the range form against the per-tile tests it replaces).
template <int BY>
struct TileIter {
  int gi, seg, yt, yt_end;
  int xmin, xmax, xgroup;
  int s_row0, s_nrows, s_pos0, s_flag;    // current segment (cached)
  int y_row0, nvalid, ypos0, flag;
  __device__ __forceinline__ void init(const BwdParams& p, const Work& wk) {
    gi = 0; seg = -1; yt = 0; yt_end = 0;
    xmin = wk.pos0;
    xmax = wk.pos0 + (wk.nrows - 1) * p.x_pos_stride;
    xgroup = p.xseg[wk.xseg].group;
  }
  __device__ __forceinline__ bool next(const BwdParams& p) {
    for (;;) {
      if (++yt < yt_end) {
        const int r0 = yt * BY;
        y_row0 = s_row0 + r0;
        nvalid = min(BY, s_nrows - r0);
        ypos0 = s_pos0 + r0 * p.y_pos_stride;
        flag = s_flag;
        return true;
      }
      if (++seg >= p.n_yseg) {
        seg = 0;
        if (p.n_yseg == 0 || ++gi >= p.n_inner) return false;
      }
      const KSegD s = p.yseg[seg];
      s_row0 = s.row0; s_nrows = s.nrows; s_pos0 = s.pos0; s_flag = s.flag;
      const int nt = (s.group == xgroup) ? (s.nrows + BY - 1) / BY : 0;
      int lo = 0, hi = nt;
      if (p.wr >= 0 && nt > 0) {            // a tile is right of the window iff ypos0 - xmax > wr
        const int lim = xmax + p.wr - s.pos0;
        hi = lim < 0 ? 0 : min(nt, lim / (BY * p.y_pos_stride) + 1);
      }
      if (p.wl >= 0 && hi > 0) {            // ... left of it iff xmin - (position of its last valid row) > wl
        const int need = xmin - p.wl - s.pos0;
        if (need > 0) {
          const int e_min = (need + p.y_pos_stride - 1) / p.y_pos_stride + 1;   // rows the tile prefix must span
          lo = e_min > s.nrows ? hi : (e_min + BY - 1) / BY - 1;
        }
      }
      yt = lo - 1;
      yt_end = hi;
    }
  }
};

// stage index + phase bit of a ring of N mbarrier-guarded buffers (no division on the single-thread roles' paths)
template <int N>
struct Ring {
  uint32_t idx = 0, phase = 0;
  __device__ __forceinline__ void advance() {
    if (++idx == N) { idx = 0; phase ^= 1u; }
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

// write one row-slice of an accumulator: 32 fp32 values -> out (mode 0: 16-bit, 1: fp32 store, 2: fp32 +=)
template <bool kBf16>
__device__ __forceinline__ void store_slice(void* base, int col0, const uint32_t (&o)[32], float mul, int mode) {
  if (mode == 0) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(base) + col0 * 2);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 w;
      w.x = pack2<kBf16>(__uint_as_float(o[g * 8 + 0]) * mul, __uint_as_float(o[g * 8 + 1]) * mul);
      w.y = pack2<kBf16>(__uint_as_float(o[g * 8 + 2]) * mul, __uint_as_float(o[g * 8 + 3]) * mul);
      w.z = pack2<kBf16>(__uint_as_float(o[g * 8 + 4]) * mul, __uint_as_float(o[g * 8 + 5]) * mul);
      w.w = pack2<kBf16>(__uint_as_float(o[g * 8 + 6]) * mul, __uint_as_float(o[g * 8 + 7]) * mul);
      dst[g] = w;
    }
  } else if (mode == 3) {
    // cross-rank reduction: fp32 vector reductions straight into the owner's accumulator (NVLink)
    float* dst = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(base) + col0 * 4);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4),
                   "f"(__uint_as_float(o[g * 4 + 0]) * mul), "f"(__uint_as_float(o[g * 4 + 1]) * mul),
                   "f"(__uint_as_float(o[g * 4 + 2]) * mul), "f"(__uint_as_float(o[g * 4 + 3]) * mul)
                   : "memory");
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(base) + col0 * 4);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float4 w = make_float4(__uint_as_float(o[g * 4 + 0]) * mul, __uint_as_float(o[g * 4 + 1]) * mul,
                             __uint_as_float(o[g * 4 + 2]) * mul, __uint_as_float(o[g * 4 + 3]) * mul);
      if (mode == 2) {
        const float4 old = dst[g];
        w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
      }
      dst[g] = w;
    }
  }
}

}  // namespace

// kDyn: dynamic tile scheduler (global atomic counter + 2-deep smem ring): the default of the fused multi-GPU
//   launches, where the push CTAs join the compute pool once their transfers are out; else the static snake schedule.
// kDrop: attention dropout regenerated from global coordinates (scalar arithmetic); every other instantiation runs the
//   element-wise stage on packed fp32x2 instructions (FFMA2 / FADD2 / FMUL2; statistics are staged NEGATED because the
//   packed forms take no negate modifier).  Round-2 hardware validation: packed = -4 % time in both passes; the
//   "both warpgroups on every tile" split (kSplit) and the 64-row forward tiles it was modelled after measured SLOWER
//   than this pipeline and were deleted.
// x_empty protocol (dQ pass): the element-wise warps wait x_full once per work item (their lse2/delta reads must be
//   ordered after a peer's push).  If only the MMA warp released x_empty, a work item WITHOUT any visible streamed tile
//   would be released at once, the producer would reload X, and x_full could complete two phases while a warpgroup is
//   still in the previous item's epilogue: its one-bit parity wait would block forever (found by the protocol model
//   tests/test_bwd_pipeline_model_cpu.py; reproduced on hardware in round 2: tools/gpu_repro_xfix.py hangs with the old
//   rule, finishes with this one).  So every element-wise warp also arrives on x_empty (count 9) right after its
//   x_full wait: X cannot be reloaded under a waiter.
template <int kD, bool kBf16, bool kIsDKV, bool kDyn, bool kDrop, int kBY>
__global__ void __launch_bounds__(kThreads, 1) fmha_bwd_kernel(const __grid_constant__ BwdParams p) {
  using C = Cfg<kD, kIsDKV, kBY>;
  constexpr int NS = C::NS;
  constexpr int BY = C::BY;
  constexpr bool kXf = !kIsDKV;       // see "x_empty protocol" above
  constexpr bool kPk = !kDrop;        // packed fp32x2 element-wise stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
  if (static_cast<int>(blockIdx.x) < p.comm.n_comm) {   // communication role (fused USP backward)
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

  // ---- barriers
  uint32_t a = smem + C::OFF_BAR;
  const uint32_t x_full = a; a += 8;
  const uint32_t x_empty = a; a += 8;
  const uint32_t acc_full = a; a += 8;
  const uint32_t acc_empty = a; a += 8;
  const uint32_t t_full = a; a += 8 * NS;
  const uint32_t p_full = a; a += 8 * NS;
  const uint32_t y_full = a; a += 8 * C::STAGES;
  const uint32_t y_empty = a; a += 8 * C::STAGES;
  const uint32_t st_full = a; a += 8 * C::STAGES;
  const uint32_t st_empty = a; a += 8 * C::STAGES;
  static_assert(32 + 16 * NS + 32 * C::STAGES <= 400, "barrier block overlaps the scheduler ring");
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + C::OFF_BAR + 480);
  float* stat = reinterpret_cast<float*>(smem_gen + C::OFF_STAT);
  const uint32_t sc_full = smem + C::OFF_BAR + 400, sc_empty = smem + C::OFF_BAR + 416;   // dynamic scheduler ring
  volatile int* sched_idx = reinterpret_cast<volatile int*>(smem_gen + C::OFF_BAR + 432);
  auto producer_next = [&](int round) -> int {          // called by ALL lanes of the producer warp
    if constexpr (kDyn) {
      const uint32_t slot = round & 1, par = (round >> 1) & 1;
      if (lane == 0) {
        mbar_wait(sc_empty + 8 * slot, par ^ 1);
        sched_idx[slot] = static_cast<int>(atomicAdd(p.sched_counter, 1u) - p.sched_base);
        mbar_arrive(sc_full + 8 * slot);
      }
      __syncwarp();
      return sched_idx[slot];
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
      if (lane == 0) mbar_arrive(sc_empty + 8 * slot);
      return w;
    } else {
      return sched_work(round, p.comm.n_comm);
    }
  };

  if (threadIdx.x == 0) {
    if constexpr (kDyn) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(sc_full + 8 * s, 1);
        mbar_init(sc_empty + 8 * s, 9);    // MMA warp + 8 element-wise warps
      }
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, kXf ? 9 : 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    for (int s = 0; s < NS; ++s) {
      mbar_init(t_full + 8 * s, 1);
      mbar_init(p_full + 8 * s, 4);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(y_full + 8 * s, 1);
      mbar_init(y_empty + 8 * s, 1);
      mbar_init(st_full + 8 * s, 32);
      mbar_init(st_empty + 8 * s, 4);
    }
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&p.tm_x0); prefetch_tmap(&p.tm_x1); prefetch_tmap(&p.tm_y0); prefetch_tmap(&p.tm_y1);
  }
  if (warp == kMmaWarp) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp >= kMmaWarp) {
    setmaxnreg_dec<96>();
    if (warp == kTmaWarp) {
      // =========================================================== producer (whole warp: lane 0 drives TMA,
      // all lanes fetch the per-column lse2/delta of streamed query tiles in the dKV pass)
      uint32_t xc = 0;
      Ring<C::STAGES> yr;           // streamed-tile ring (stage index + phase), advanced once per tile
      int x_ok = -1, y_ok = -1;     // last arrival flags already acquired (a flag only ever needs one acquire per launch)
      for (int round = 0;; ++round) {
        Work wk;
        if (!decode_work(p, producer_next(round), wk)) break;
        if (lane == 0) {
          const int xf = p.xseg[wk.xseg].flag;
          if (xf >= 0 && xf != x_ok) { wait_arrival(p.flags, p.flag_epoch, xf, p.comm.watchdog_ns); x_ok = xf; }
          mbar_wait(x_empty, (xc & 1) ^ 1);
          mbar_arrive_expect_tx(x_full, 2 * C::XTILE_BYTES);
#pragma unroll
          for (int db = 0; db < C::DBLK; ++db) {
            tma_load_4d(smem + C::OFF_X + db * C::XBLK_BYTES, &p.tm_x0, x_full, db * 64, wk.hx, wk.row0, wk.b);
            tma_load_4d(smem + C::OFF_X + C::XTILE_BYTES + db * C::XBLK_BYTES, &p.tm_x1, x_full, db * 64, wk.hx, wk.row0, wk.b);
          }
        }
        ++xc;
        TileIter<BY> it;
        it.init(p, wk);
        const int hy0 = kIsDKV ? wk.hx * p.n_inner : wk.hx / p.hx_per_hy;      // (one division per work item, not per tile)
        const float* l2_b = p.lse2 + wk.b * p.stat_sb;
        const float* dl_b = p.delta + wk.b * p.stat_sb;
        while (it.next(p)) {
          const int hy = kIsDKV ? hy0 + it.gi : hy0;
          const uint32_t st = yr.idx;
          const uint32_t par = yr.phase;
          if (lane == 0) {
            if (it.flag >= 0 && it.flag != y_ok) { wait_arrival(p.flags, p.flag_epoch, it.flag, p.comm.watchdog_ns); y_ok = it.flag; }
            mbar_wait(y_empty + 8 * st, par ^ 1);
            mbar_arrive_expect_tx(y_full + 8 * st, 2 * C::YTILE_BYTES);
            const uint32_t ydst = smem + C::OFF_Y + st * 2 * C::YTILE_BYTES;
#pragma unroll
            for (int db = 0; db < C::DBLK; ++db) {
              tma_load_4d(ydst + db * C::YBLK_BYTES, &p.tm_y0, y_full + 8 * st, db * 64, hy, it.y_row0, wk.b);
              tma_load_4d(ydst + C::YTILE_BYTES + db * C::YBLK_BYTES, &p.tm_y1, y_full + 8 * st, db * 64, hy, it.y_row0, wk.b);
            }
          }
          if constexpr (kIsDKV) {
            // per-column statistics of the streamed query rows: asynchronous 4-byte copies global -> smem that arrive
            // on st_full when they land (cp.async.mbarrier.arrive.noinc), so STAGES tiles of statistics are in flight.
            // Round 2 finding: the synchronous ld.global -> st.shared -> arrive this replaces cost one global-load
            // latency PER TILE on the producer's critical path and bounded the whole dK/dV pass.  Invalid columns are
            // zero-filled (their scores are masked anyway).
            __syncwarp();                      // lane 0 has acquired the arrival flag of this tile's rows
            mbar_wait(st_empty + 8 * st, par ^ 1);
            const float* l2 = l2_b + hy * p.stat_sh;
            const float* dl = dl_b + hy * p.stat_sh;
            const uint32_t sdst = smem + C::OFF_STAT + st * 2 * BY * 4;
#pragma unroll
            for (int i = 0; i < BY / 32; ++i) {
              const int c = lane + 32 * i;
              const bool ok = c < it.nvalid;
              cp_async_f32_zfill(sdst + c * 4, ok ? l2 + it.y_row0 + c : l2, ok);
              cp_async_f32_zfill(sdst + (BY + c) * 4, ok ? dl + it.y_row0 + c : dl, ok);
            }
            cp_async_mbar_arrive_noinc(st_full + 8 * st);
          }
          yr.advance();
        }
      }
    } else if (warp == kMmaWarp) {
      // =========================================================== MMA issuer (whole warp, elected lane issues)
      // This warp's instruction stream is on the critical path of every streamed tile (round-2 ncu: the element-wise
      // warpgroups waited on t_full 60-70 % of the time while the tensor pipe sat at 55-60 %: the issuer executed
      // ~240 mostly dependent uniform-datapath instructions per tile).  Hence: shared-memory descriptors are built
      // ONCE and stepped by adding (byte offset >> 4) to their low word, ring indices / phases are counters (no
      // division), and the tile iterator is the range form above.
      {
        constexpr uint32_t idesc_t = make_idesc_f16(kBf16 ? 1 : 0, BX, BY, 0, 0);
        constexpr uint32_t idesc_acc = make_idesc_f16(kBf16 ? 1 : 0, BX, kD, 0, 1);
        constexpr uint32_t kYStage16 = (2 * C::YTILE_BYTES) >> 4;        // descriptor step per smem stage
        const uint64_t dx0 = make_sw128_desc(smem + C::OFF_X, 16, 1024);                       // X0, K-major
        const uint64_t dx1 = make_sw128_desc(smem + C::OFF_X + C::XTILE_BYTES, 16, 1024);      // X1
        const uint64_t dyt = make_sw128_desc(smem + C::OFF_Y, 16, 1024);                       // Y0 of stage 0 as the K-major B operand
        const uint64_t dya = make_sw128_desc(smem + C::OFF_Y, C::YBLK_BYTES, 1024);            // ... as the MN-major B operand
        uint32_t xc = 0, ac = 0;
        Ring<C::STAGES> yi;        // smem stage of the next tile whose T GEMMs are issued
        Ring<C::STAGES> ya;        // smem stage of the next tile whose accumulate GEMMs are issued
        Ring<NS> ti;               // TMEM stage (+ t_full phase) of the next T issue
        Ring<NS> ta;               // TMEM stage (+ p_full phase) of the next accumulate issue
        auto issue_t = [&](uint32_t stage, uint32_t s) {
          const uint32_t so = stage * kYStage16;
          const uint32_t d0 = tmem + C::TMEM_T0 + s * BY, d1 = tmem + C::TMEM_T1 + s * BY;
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t xo = (kk >> 2) * C::XBLK_BYTES + (kk & 3) * 32;
            const uint32_t yo = (kk >> 2) * C::YBLK_BYTES + (kk & 3) * 32;
            mma_ss(d0, desc_step(dx0, xo >> 4), desc_step(dyt, so + (yo >> 4)), idesc_t, kk > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t xo = (kk >> 2) * C::XBLK_BYTES + (kk & 3) * 32;
            const uint32_t yo = C::YTILE_BYTES + (kk >> 2) * C::YBLK_BYTES + (kk & 3) * 32;
            mma_ss(d1, desc_step(dx1, xo >> 4), desc_step(dyt, so + (yo >> 4)), idesc_t, kk > 0 ? 1u : 0u);
          }
        };
        auto issue_acc = [&](uint32_t stage, uint32_t s, bool acc) {
          const uint32_t so = stage * kYStage16;
          const uint32_t a1 = tmem + C::TMEM_T1 + s * BY, a0 = tmem + C::TMEM_T0 + s * BY;
#pragma unroll
          for (int kk = 0; kk < BY / 16; ++kk)   // acc0 += dS * Y0
            mma_ts(tmem + C::TMEM_ACC0, a1 + kk * 8, desc_step(dya, so + ((kk * 2048) >> 4)), idesc_acc, (acc || kk > 0) ? 1u : 0u);
          if constexpr (kIsDKV) {
#pragma unroll
            for (int kk = 0; kk < BY / 16; ++kk)   // acc1 += P * Y1
              mma_ts(tmem + C::TMEM_ACC1, a0 + kk * 8, desc_step(dya, so + ((C::YTILE_BYTES + kk * 2048) >> 4)), idesc_acc,
                     (acc || kk > 0) ? 1u : 0u);
          }
        };
        auto t_step = [&]() {      // T GEMMs of the next streamed tile into the next TMEM stage
          mbar_wait(y_full + 8 * yi.idx, yi.phase);
          tc_fence_after();
          issue_t(yi.idx, ti.idx);
          mma_commit(t_full + 8 * ti.idx);
          yi.advance();
          ti.advance();
        };
        for (int round = 0;; ++round) {
          Work wk;
          if (!decode_work(p, consumer_next(round), wk)) break;
          TileIter<BY> it;
          it.init(p, wk);
          mbar_wait(x_full, xc & 1);
          ++xc;
          bool more = it.next(p);
          if (!more) {
            mma_commit(x_empty);
            continue;
          }
          tc_fence_after();
          // prologue: the T GEMMs of the first NS tiles (as far as present)
          int pending = 0;                         // tiles whose T GEMMs are issued and whose accumulate GEMMs are not
          for (int i = 0; i < NS && more; ++i) {
            t_step();
            ++pending;
            more = it.next(p);
          }
          if (!more) mma_commit(x_empty);          // all T GEMMs (the only readers of X) are issued
          for (bool first = true; pending > 0; first = false) {
            mbar_wait(p_full + 8 * ta.idx, ta.phase);
            if (first) {
              mbar_wait(acc_empty, (ac & 1) ^ 1);
              ++ac;
            }
            tc_fence_after();
            issue_acc(ya.idx, ta.idx, !first);
            mma_commit(y_empty + 8 * ya.idx);
            ya.advance();
            ta.advance();
            --pending;
            if (more) {                              // the next tile takes over the TMEM stage that was just consumed
              t_step();
              ++pending;
              more = it.next(p);
              if (!more) mma_commit(x_empty);
            }
          }
          mma_commit(acc_full);
        }
      }
    }
  } else {
    setmaxnreg_inc<200>();
    // =========================================================== elementwise warpgroups
    // streamed tile g (running count over the CTA's work items) sits in TMEM stage g % NS and belongs to warpgroup g % 2
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t g = 0, afc = 0, xcw = 0;
    Ring<C::STAGES> yr;      // smem stage (+ phase) of the current streamed tile
    Ring<NS> tr;             // TMEM stage (+ phase) of the current streamed tile
    const bool plain = (p.softcap == 0.f) && (p.alibi == nullptr);
    for (int round = 0;; ++round) {
      Work wk;
      if (!decode_work(p, consumer_next(round), wk)) break;
      const int xpos = wk.pos0 + row * p.x_pos_stride;
      const int xhi = wk.pos0 + (BX - 1) * p.x_pos_stride;
      const bool row_ok = row < wk.nrows;
      float lse2_r = INFINITY, delta_r = 0.f;
      if constexpr (!kIsDKV) {
        // the per-row statistics may have been pushed by a peer together with this Q/dO tile: the producer
        // acquired the arrival flag before issuing the tile's TMA, so x_full orders our reads after the push
        mbar_wait(x_full, xcw & 1);
        ++xcw;
        if constexpr (kXf) {
          if (lane == 0) mbar_arrive(x_empty);     // this warp has consumed the phase: X may be reloaded
        }
        if (row_ok) {
          lse2_r = p.lse2[wk.b * p.stat_sb + wk.hx * p.stat_sh + wk.row0 + row];
          delta_r = p.delta[wk.b * p.stat_sb + wk.hx * p.stat_sh + wk.row0 + row];
        }
      }
      TileIter<BY> it;
      it.init(p, wk);
      int j = 0;
      while (it.next(p)) {
        const uint32_t st = yr.idx;
        const uint32_t ypar = yr.phase;
        const int sT = static_cast<int>(tr.idx);                  // TMEM stage of this tile
        const uint32_t tpar = tr.phase;
        yr.advance();
        tr.advance();
        const uint32_t gt = g++;
        if (static_cast<int>(gt & 1u) != wg) { ++j; continue; }   // warpgroup wg owns the tiles with g % 2 == wg
        const uint32_t tT0 = tmem + lane_base + C::TMEM_T0 + sT * BY, tT1 = tmem + lane_base + C::TMEM_T1 + sT * BY;
        constexpr int h_begin = 0, h_end = BY / 32;               // 32-column halves per tile
        const int hq = kIsDKV ? wk.hx * p.n_inner + it.gi : wk.hx;     // query head (ALiBi slope index)
        const float slope = p.alibi ? p.alibi[wk.b * p.alibi_bstride + hq] : 0.f;
        mbar_wait(t_full + 8 * sT, tpar);
        if constexpr (kIsDKV) mbar_wait(st_full + 8 * st, ypar);
        tc_fence_after();
        const int yb = it.ypos0 + (it.nvalid - 1) * p.y_pos_stride;
        const bool need_mask = (it.nvalid < BY) || (p.wr >= 0 && yb - wk.pos0 > p.wr) ||
                               (p.wl >= 0 && xhi - it.ypos0 > p.wl) || (wk.nrows < BX);
        const float* st_l2 = stat + st * 2 * BY;
        const float* st_dl = st_l2 + BY;
        float mul = p.scale_log2;
        if (!plain || need_mask) {
          // general pre-pass (rolled, keeps the hot loop small): rewrite T0 as masked log2-domain logits
          // and T1 as dl + (T1 - dl) * softcap'(s), so the common loop below needs no special cases.
#pragma unroll 1
          for (int half = h_begin; half < h_end; ++half) {
            uint32_t t0[32], t1[32];
            tmem_ld32(tT0 + half * 32, t0);
            tmem_ld32(tT1 + half * 32, t1);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = half * 32 + i;
              const float dl = kIsDKV ? st_dl[col] : delta_r;
              float x = __uint_as_float(t0[i]) * p.scale;
              float extra = 1.f;
              if (p.softcap > 0.f) {
                const float th = tanh_approx(x / p.softcap);
                x = p.softcap * th;
                extra = 1.f - th * th;
              }
              const int rel = it.ypos0 + col * p.y_pos_stride - xpos;
              if (p.alibi) x -= slope * fabsf(static_cast<float>(rel));
              const bool masked = (col >= it.nvalid) || !row_ok || (p.wr >= 0 && rel > p.wr) || (p.wl >= 0 && -rel > p.wl);
              t0[i] = __float_as_uint(masked ? -INFINITY : x * 1.4426950408889634f);
              t1[i] = __float_as_uint(masked ? dl : fmaf(__uint_as_float(t1[i]) - dl, extra, dl));
            }
            tmem_st32(tT0 + half * 32, t0);
            tmem_st32(tT1 + half * 32, t1);
          }
          tmem_wait_st();
          mul = 1.f;
        }
        // 16-column chunks, double buffered: the tcgen05.ld of chunk q+1 is in flight while chunk q is computed
        // (before: load 2 x 32 columns -> wait -> compute -> store, twice per tile, i.e. two exposed TMEM round trips;
        // the element-wise warpgroups, not the tensor pipe, bounded both passes: ~3000 clk per 64-column tile).
        auto chunk = [&](int q, const uint32_t (&t0)[16], const uint32_t (&t1)[16], uint32_t* pp, uint32_t* ds) {
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            float l2v[4], dlv[4];
            if constexpr (kIsDKV) {
              const float4 a4 = *reinterpret_cast<const float4*>(st_l2 + q * 16 + c);
              const float4 b4 = *reinterpret_cast<const float4*>(st_dl + q * 16 + c);
              l2v[0] = a4.x; l2v[1] = a4.y; l2v[2] = a4.z; l2v[3] = a4.w;
              dlv[0] = b4.x; dlv[1] = b4.y; dlv[2] = b4.z; dlv[3] = b4.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) { l2v[e] = kPk ? -lse2_r : lse2_r; dlv[e] = kPk ? -delta_r : delta_r; }
            }
            float pv[4], dv[4];
            if constexpr (kDrop) {
              // dS = P o (keep * dP / (1-p) - delta); the dV GEMM consumes P_drop = keep * P / (1-p).  The keep bits
              // are regenerated from the global coordinates (ops/dropout.py).  Rows are queries in the dQ pass
              // (columns = keys: four consecutive keys share a hash word) and keys in the dK/dV pass (columns =
              // queries: one hash per score).
              float pd[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = q * 16 + c + e;
                const uint32_t ypos = static_cast<uint32_t>(it.ypos0 + col * p.y_pos_stride);
                const uint32_t qp = kIsDKV ? ypos : static_cast<uint32_t>(xpos);
                const uint32_t kp = kIsDKV ? static_cast<uint32_t>(xpos) : ypos;
                const uint32_t w = ptx::dropout_word(
                    ptx::dropout_row_key(qp, p.drop_seed, static_cast<uint32_t>(wk.b + p.xseg[wk.xseg].group),
                                         static_cast<uint32_t>(hq + p.drop_head_off)), kp);
                const bool keep = ptx::dropout_keep(w, kp, static_cast<uint32_t>(p.drop_p8));
                pv[e] = ex2(fmaf(__uint_as_float(t0[c + e]), mul, -l2v[e]));
                dv[e] = pv[e] * ((keep ? __uint_as_float(t1[c + e]) * p.drop_rscale : 0.f) - dlv[e]);
                pd[e] = keep ? pv[e] * p.drop_rscale : 0.f;
              }
              if constexpr (kIsDKV) {
                pv[0] = pd[0]; pv[1] = pd[1]; pv[2] = pd[2]; pv[3] = pd[3];
              }
            } else {
              // packed fp32x2 arithmetic: one FFMA2 / FADD2 / FMUL2 per element pair (the packed forms take no negate
              // modifier).  dQ pass: l2v / dlv are the row statistics negated once per work item:
              //   x = T0*mul + (-lse2), d = T1 + (-delta), dS = P*d.
              // dK/dV pass: l2v / dlv are the TRUE per-column statistics as the asynchronous copies delivered them:
              //   y = T0*(-mul) + lse2 = -x (the negation folds into the MUFU operand), d' = T1*(-1) + delta = -d,
              //   and P*d' = -dS goes to the dK GEMM; the epilogue multiplies dK by -scale.
              const uint64_t mul2 = kIsDKV ? ptx::pack_f32x2(-mul, -mul) : ptx::pack_f32x2(mul, mul);
#pragma unroll
              for (int e = 0; e < 4; e += 2) {
                float x0, x1;
                ptx::unpack_f32x2(ptx::fma_f32x2(
                    ptx::pack_f32x2(__uint_as_float(t0[c + e]), __uint_as_float(t0[c + e + 1])), mul2,
                    ptx::pack_f32x2(l2v[e], l2v[e + 1])), x0, x1);
                pv[e] = ex2(kIsDKV ? -x0 : x0);
                pv[e + 1] = ex2(kIsDKV ? -x1 : x1);
                const uint64_t t1p = ptx::pack_f32x2(__uint_as_float(t1[c + e]), __uint_as_float(t1[c + e + 1]));
                const uint64_t d = kIsDKV ? ptx::fma_f32x2(t1p, ptx::pack_f32x2(-1.f, -1.f), ptx::pack_f32x2(dlv[e], dlv[e + 1]))
                                          : ptx::add_f32x2(t1p, ptx::pack_f32x2(dlv[e], dlv[e + 1]));
                ptx::unpack_f32x2(ptx::mul_f32x2(ptx::pack_f32x2(pv[e], pv[e + 1]), d), dv[e], dv[e + 1]);
              }
            }
            if constexpr (kIsDKV) {
              pp[(c >> 1)] = pack2<kBf16>(pv[0], pv[1]);
              pp[(c >> 1) + 1] = pack2<kBf16>(pv[2], pv[3]);
            }
            ds[(c >> 1)] = pack2<kBf16>(dv[0], dv[1]);
            ds[(c >> 1) + 1] = pack2<kBf16>(dv[2], dv[3]);
          }
        };
        uint32_t ta0[16], ta1[16], tb0[16], tb1[16];
        tmem_ld16(tT0, ta0);
        tmem_ld16(tT1, ta1);
        tmem_wait_ld();
#pragma unroll
        for (int half = h_begin; half < h_end; ++half) {
          uint32_t pp[16], ds[16];
          tmem_ld16(tT0 + half * 32 + 16, tb0);
          tmem_ld16(tT1 + half * 32 + 16, tb1);
          chunk(2 * half, ta0, ta1, pp, ds);
          tmem_wait_ld();
          if (half + 1 < h_end) {
            tmem_ld16(tT0 + (half + 1) * 32, ta0);
            tmem_ld16(tT1 + (half + 1) * 32, ta1);
          }
          chunk(2 * half + 1, tb0, tb1, pp + 8, ds + 8);
          if constexpr (kIsDKV) tmem_st16(tT0 + half * 16, pp);
          tmem_st16(tT1 + half * 16, ds);
          if (half + 1 < h_end) tmem_wait_ld();
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(p_full + 8 * sT);
          if constexpr (kIsDKV) mbar_arrive(st_empty + 8 * st);
        }
        ++j;
      }
      // ---- epilogue: both warpgroups drain the accumulators (dQ: column halves; dKV: wg0 -> dK, wg1 -> dV)
      if (j > 0) {
        mbar_wait(acc_full, afc & 1);
        ++afc;
        tc_fence_after();
      }
      {
        const XSegD xs = p.xseg[wk.xseg];
        const int64_t orow = static_cast<int64_t>(xs.o_row0) + wk.seg_row0 + row;
        int ncols, col_begin;
        uint32_t tacc;
        uint8_t* obase;
        float mul;
        int esz = (p.out_mode == 0) ? 2 : 4;
        if constexpr (kIsDKV) {
          ncols = kD; col_begin = 0;
          tacc = tmem + lane_base + (wg == 0 ? C::TMEM_ACC0 : C::TMEM_ACC1);
          void* ob = wg == 0 ? (xs.o_base0 ? xs.o_base0 : p.out0) : (xs.o_base1 ? xs.o_base1 : p.out1);
          obase = reinterpret_cast<uint8_t*>(ob);
          mul = wg == 0 ? (kPk ? -p.scale : p.scale) : 1.f;     // packed dK/dV pass accumulates -dS (see above)
        } else {
          ncols = kD / 2; col_begin = wg * (kD / 2);
          tacc = tmem + lane_base + C::TMEM_ACC0;
          obase = reinterpret_cast<uint8_t*>(xs.o_base0 ? xs.o_base0 : p.out0);
          mul = p.scale;
        }
        uint8_t* orow_ptr = obase + esz * (wk.b * p.o_sb + orow * p.o_ss + static_cast<int64_t>(wk.hx + p.o_head_off) * p.o_sh);
        const bool skip_store = (p.out_mode == 3) && (j == 0);      // nothing to reduce: no visible tile
        for (int c = 0; c < ncols; c += 32) {
          uint32_t o[32];
          if (j > 0) {
            tmem_ld32(tacc + col_begin + c, o);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0u;
          }
          if (row_ok && !skip_store) store_slice<kBf16>(orow_ptr, col_begin + c, o, mul, p.out_mode);
        }
        if (xs.o_sig != nullptr) {          // publish this warpgroup's part of the tile to the owner rank
          __threadfence_system();
          named_bar_sync(1 + wg, 128);
          if ((warp & 3) == 0 && lane == 0) red_add_release_sys(xs.o_sig, 1u);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && j > 0) mbar_arrive(acc_empty);
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
template <int kD, bool kBf16, bool kIsDKV, bool kDyn, bool kDrop = false, int kBY = 64>
static cudaError_t launch_impl(const BwdParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<kD, kIsDKV, kBY>;
  auto kern = fmha_bwd_kernel<kD, kBf16, kIsDKV, kDyn, kDrop, kBY>;
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

template <int kD, bool kBf16>
static cudaError_t launch_pass(const BwdParams& p, bool is_dkv, int num_sms, cudaStream_t stream) {
  if (p.drop_p8 > 0)                  // dropout instantiations (static schedule, scalar arithmetic)
    return is_dkv ? launch_impl<kD, kBf16, true, false, true>(p, num_sms, stream)
                  : launch_impl<kD, kBf16, false, false, true>(p, num_sms, stream);
  if (p.dyn_sched)
    return is_dkv ? launch_impl<kD, kBf16, true, true>(p, num_sms, stream) : launch_impl<kD, kBf16, false, true>(p, num_sms, stream);
  return is_dkv ? launch_impl<kD, kBf16, true, false>(p, num_sms, stream) : launch_impl<kD, kBf16, false, false>(p, num_sms, stream);
}

cudaError_t launch_fmha_bwd(const BwdParams& p, int head_dim, bool bf16, bool is_dkv, int num_sms, cudaStream_t stream) {
  if (head_dim == 128) return bf16 ? launch_pass<128, true>(p, is_dkv, num_sms, stream) : launch_pass<128, false>(p, is_dkv, num_sms, stream);
  if (head_dim == 64) return bf16 ? launch_pass<64, true>(p, is_dkv, num_sms, stream) : launch_pass<64, false>(p, is_dkv, num_sms, stream);
  return cudaErrorInvalidValue;
}

}  // namespace lca
