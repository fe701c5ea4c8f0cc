// Pieces shared by the forward translation units (16-bit: fmha_fwd_sm100.cu, e4m3: fmha_fwd_fp8_sm100.cu): work-item
// decoding, the static schedule, the K/V tile iterator and the ring counter.  Included inside each unit's anonymous
// namespace AFTER it has defined BM / BN (query / key rows per tile).
#pragma once

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

// static "snake" schedule: round k visits work k*G + c on even rounds and k*G + (G-1-c) on odd
// rounds, which cancels the cost gradient of the heaviest-first ordering across CTAs.
__device__ __forceinline__ int sched_work(int round, int n_comm) {
  const int G = static_cast<int>(gridDim.x) - n_comm;          // compute CTAs
  const int me = static_cast<int>(blockIdx.x) - n_comm;
  const int c = (round & 1) ? (G - 1 - me) : me;
  return round * G + c;
}

// Deterministic enumeration of the K/V tiles a Q pair has to visit (identical in every role).
// Positions grow with the tile index inside a segment, so the visible tiles of a segment are ONE contiguous range
// [lo, hi) computed when the iterator enters the segment; the per-tile step is a compare and three multiply-adds
// (the single-thread roles -- MMA issuer, TMA producer -- are the critical path of this kernel, see the round-2
// profile in profiles/r2/; tests/test_properties_cpu.py checks the range form against the per-tile window tests).
struct TileIter {
  int seg, kt, kt_end;
  int qmin, qmax, qgroup;
  int s_row0, s_nrows, s_pos0, s_flag;    // current segment (cached)
  // current tile
  int k_row0, nvalid, kpos0, flag;
  __device__ __forceinline__ void init(const FwdParams& p, const Work& wk) {
    seg = -1;
    kt = 0;
    kt_end = 0;
    qmin = wk.pos0;
    qmax = wk.pos0 + (wk.nrows - 1) * p.q_pos_stride;
    qgroup = p.qseg[wk.qseg].group;
  }
  __device__ __forceinline__ bool next(const FwdParams& p) {
    for (;;) {
      if (++kt < kt_end) {
        const int r0 = kt * BN;
        k_row0 = s_row0 + r0;
        nvalid = min(BN, s_nrows - r0);
        kpos0 = s_pos0 + r0 * p.k_pos_stride;
        flag = s_flag;
        return true;
      }
      if (++seg >= p.n_kseg) return false;
      const KSegD s = p.kseg[seg];
      s_row0 = s.row0; s_nrows = s.nrows; s_pos0 = s.pos0; s_flag = s.flag;
      const int nt = (s.group == qgroup) ? (s.nrows + BN - 1) / BN : 0;
      int lo = 0, hi = nt;
      if (p.wr >= 0 && nt > 0) {            // a tile is right of the window iff kpos0 - qmax > wr
        const int lim = qmax + p.wr - s.pos0;
        hi = lim < 0 ? 0 : min(nt, lim / (BN * p.k_pos_stride) + 1);
      }
      if (p.wl >= 0 && hi > 0) {            // ... left of it iff qmin - (position of its last valid row) > wl
        const int need = qmin - p.wl - s.pos0;
        if (need > 0) {
          const int e_min = (need + p.k_pos_stride - 1) / p.k_pos_stride + 1;   // rows the tile prefix must span
          lo = e_min > s.nrows ? hi : (e_min + BN - 1) / BN - 1;
        }
      }
      kt = lo - 1;
      kt_end = hi;
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

