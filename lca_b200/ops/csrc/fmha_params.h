// Host/device parameter blocks of the sm_100a attention kernels.
//
// Everything the kernels know about sequence parallelism arrives as *segments*: contiguous row
// ranges of the Q / K / V tensors whose tokens sit at global positions pos0 + i*stride.  The
// ring variants (basic / zigzag / stripe), the Ulysses gather order, sliding windows and ALiBi
// all reduce to this description, so one kernel serves every parallel layout.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace lca {

constexpr int kMaxSeg = 128;      // kernel parameters are ~12 KiB with 128 segments per side (limit 32 KiB since CUDA 12.1)

struct QSegD {
  int row0;        // first row of the segment in the Q tensor (dim S)
  int nrows;       // rows in the segment
  int pos0;        // global position of row0
  int flag;        // index into FwdParams::flags that must reach flag_epoch before Q rows are read (-1: none)
  int o_row0;      // first destination row in the output tensor addressed by o_base
  int group;       // attention group id: a Q segment only visits K segments of the same group (varlen)
  void* o_base;    // output tensor base for this segment (may be a peer-mapped pointer)
  uint32_t* o_sig; // optional: system-scope counter incremented once per finished 128-row tile
  float* lse_base; // optional: token owner's (B, H_total, rows) fp32 LSE buffer (fused path; may be peer-mapped)
};

struct KSegD {
  int row0;
  int nrows;
  int pos0;
  int flag;        // index into flags (-1: none)
  int group;
};

constexpr int kMaxPeers = 16;

// Signal-pad layout (uint32 slots) of every rank's symmetric slab.  All counters are monotonic
// across calls ("epochs"), so nothing is ever reset while peers may still be polling.
constexpr int kSigKV = 0;        // [kSigKV + src]   : K/V shard of sp-rank `src` has landed   (+n_comm per call)
constexpr int kSigQ = 16;        // [kSigQ + src_u]  : Q shard of Ulysses-rank `src_u` landed  (+n_comm per call)
constexpr int kSigRTR = 32;      // [kSigRTR + dst]  : `dst` entered call `epoch` (its staging may be overwritten)
constexpr int kSigODone = 48;    // output tiles written into my out buffer by all compute ranks (cumulative)
constexpr int kSigDKV = 49;      // dK/dV tiles (or reductions) delivered into my buffers (cumulative)
constexpr int kSigQA = 64;       // [kSigQA + src]   : Q-like shard of sp-rank `src` landed in the ALL-ranks staging (+n_comm per call)
constexpr int kSigSlots = 96;
// Every call bumps every arrival class (KV, Q, QA) by n_comm, whether or not data of that class was sent, so all
// classes stay in lock-step with the call epoch and one expected value serves every flag.

// The communication half of the fused USP kernels: CTAs [0, n_comm) push this rank's shards into the
// peers' staging buffers with plain st.global over NVLink (see usp_comm.cuh).
//   "Q-like" tensors (q, and dO in the backward) go to the Ulysses peers of my ring index,
//   "KV-like" tensors (k, v) go to every sp-rank, optionally a per-row fp32 statistic (delta) rides with Q.
struct PushTensor {
  const void* src;                   // (B, rows, heads, D) shard, (head, dim) dense
  long long sb, ss;                  // element strides (batch, row)
  long long off;                     // byte offset of the destination staging tensor inside a slab
};

struct CommParams {
  int n_comm;                        // 0: no comm role (plain single-device attention)
  int P, U, R, u, r;                 // mesh: sp size, degrees, my coordinates; sp-rank = r*U + u
  int rows;                          // local tokens S/P
  int B, H, Hkv, D;
  int Hl, Hkvl;                      // heads per destination (Hkvl = max(1, Hkv/U))
  int n_q, n_kv;                     // number of Q-like (0 when U == 1) and KV-like tensors to push
  int q_to_all;                      // 1: Q-like tensors (and statistics) go to EVERY sp-rank (backward dK/dV pass)
  int n_stat;                        // per-row fp32 statistics (B, H, rows) pushed with the Q-like tensors (delta, lse2)
  PushTensor qt[2], kvt[2];
  const float* stat[2];
  long long stat_off[2];             // byte offsets of the destination (B, Hl, stage rows) fp32 tensors
  unsigned char* peer_slab[kMaxPeers];   // mapped base of every sp-rank's slab (index = sp-rank)
  unsigned int* peer_sig[kMaxPeers];     // mapped signal pad of every sp-rank
  unsigned int* my_sig;
  long long stage_q_rows, stage_kv_rows;   // rows of the staging tensors (S/R and S)
  unsigned int epoch;                // 1-based call counter
  unsigned int o_target;             // value kSigODone must reach before this rank's kernel may exit (0: skip)
  unsigned int kv_dst_mask, q_dst_mask;   // bit d: sp-rank d needs the DATA of my KV-like / Q-like tensors (its arrival
                                     // counters are bumped either way); causal / windowed layouts leave most peers out
  int push_mode;                     // 1: TMA bulk-copy push engine (default), 0: scalar st.global loop (LCA_B200_PUSH=scalar)
  unsigned long long watchdog_ns;    // spin-wait budget before the kernel traps (LCA_B200_WATCHDOG_S, 0 = wait forever)
};

struct FwdParams {
  CUtensorMap tm_q, tm_k, tm_v;       // 4-D (D, H, S, B) bf16/fp16, box (64, 1, 128, 1), SWIZZLE_128B
  int n_qseg, n_kseg;
  QSegD qseg[kMaxSeg];
  KSegD kseg[kMaxSeg];
  int q_pos_stride, k_pos_stride;     // position step between consecutive rows (stripe: R)
  int B, H, Hkv;
  int total_work;                     // sum over q segments of ceil(nrows/256) * B * H
  int wl, wr;                         // visible iff -wl <= kpos - qpos <= wr; -1 = unbounded (causal => wr = 0)
  float scale;                        // softmax scale
  float scale_log2;                   // softmax scale * log2(e)
  float softcap;                      // 0 = off
  const float* alibi;                 // per-head slopes or nullptr
  int alibi_bstride;                  // 0 for (H,), H for (B,H)
  int64_t o_sb, o_ss, o_sh;           // output strides in elements (batch, row, head)
  int o_head_off;                     // destination head index = h + o_head_off
  float* lse;                         // (B, H, lse_rows) fp32, row index = row in the Q tensor
  int64_t lse_sb, lse_sh;
  int64_t lse_own_sb, lse_own_sh;     // strides of the owners' LSE buffers (QSegD::lse_base), indexed like the output
  const uint32_t* flags;              // arrival flags written by peers (fused paths)
  uint32_t flag_epoch;
  int poly_every;                     // exp2 offload ratio: 1 of every N element pairs on the FMA pipe (0, 3, 4, 6)
  CommParams comm;
  // fp8 path only (fmha_fwd_fp8_sm100.cu): block scales of the e4m3 operands
  const float* q_scale;               // (B, H, ceil(Sq/128))
  const float* k_scale;               // (B, Hkv, ceil(Sk/128))
  const float* v_scale;               // (B, Hkv)
  int64_t q_scale_sb, q_scale_sh, k_scale_sb, k_scale_sh;
  // dynamic tile scheduler (kDyn instantiations: default of the fused launches, LCA_B200_DYN_SCHED)
  uint32_t* sched_counter;            // monotonic device counter shared by the compute CTAs of a launch
  uint32_t sched_base;                // counter value at launch start (host-tracked: += total_work + compute CTAs)
  int dyn_sched;
  // attention dropout (kDrop instantiations; default for dropout_p > 0 since round 2): keep decisions are a pure
  // function of (seed, batch + group, global query head, global q position, global k position) -- ops/dropout.py
  int drop_p8;                        // 0 = off; a score is dropped when its hash byte < drop_p8
  uint32_t drop_seed;
  float drop_rscale;                  // 256 / (256 - drop_p8)
  int drop_head_off;                  // global index of local query head 0
};

// ---- backward -----------------------------------------------------------------------------------
struct XSegD {
  int row0;      // first row of the segment in the stationary tensors
  int nrows;
  int pos0;
  int group;
  int o_row0;    // first destination row in out0/out1
  int flag;      // arrival flag of the stationary rows (-1: none)
  void* o_base0; // per-segment destination (nullptr: BwdParams::out0 / out1); may be peer-mapped
  void* o_base1;
  uint32_t* o_sig;   // optional system-scope completion counter (+1 per finished 128-row tile and warpgroup)
};

// One parameter block serves both passes (see fmha_bwd_sm100.cu):
//   dQ  pass: X = (Q, dO)  [Hx = H],   Y = (K, V)  [Hy = Hkv], n_inner = 1,  hx_per_hy = H/Hkv
//   dKV pass: X = (K, V)   [Hx = Hkv], Y = (Q, dO) [Hy = H],   n_inner = H/Hkv (query heads per KV head)
struct BwdParams {
  CUtensorMap tm_x0, tm_x1;           // box (64,1,128,1)
  CUtensorMap tm_y0, tm_y1;           // box (64,1,y_rows,1)
  int y_rows;                         // streamed rows per tile (64)
  int n_xseg, n_yseg;
  XSegD xseg[kMaxSeg];
  KSegD yseg[kMaxSeg];
  int x_pos_stride, y_pos_stride;
  int x_heavy_last;                   // 1: later stationary tiles cost more (walk them first)
  int B, Hx;
  int n_inner, hx_per_hy;
  int total_work;
  int wl, wr;                         // bounds on (ypos - xpos): visible iff -wl <= ypos - xpos <= wr
  float scale, scale_log2, softcap;
  const float* alibi;                 // indexed by QUERY head
  int alibi_bstride;
  const float* lse2;                  // (B, H, Sq) log2-domain LSE (+inf for empty rows), indexed by query row
  const float* delta;                 // (B, H, Sq) rowsum(dO o O)
  int64_t stat_sb, stat_sh;
  void* out0;                         // dQ (dQ pass) / dK (dKV pass)
  void* out1;                         // dV (dKV pass)
  int64_t o_sb, o_ss, o_sh;           // output strides in elements
  int out_mode;                       // 0: 16-bit store, 1: fp32 store, 2: fp32 accumulate, 3: fp32 red.add (peer reduction)
  int o_head_off;                     // destination head index = hx + o_head_off
  const uint32_t* flags;              // arrival flags (fused path)
  uint32_t flag_epoch;
  CommParams comm;
  uint32_t* sched_counter;            // dynamic tile scheduler (see FwdParams)
  uint32_t sched_base;
  int dyn_sched;
  // attention dropout (kDrop instantiations; default for dropout_p > 0 since round 2): keep decisions are a pure
  // function of (seed, batch + group, global query head, global q position, global k position) -- ops/dropout.py
  int drop_p8;                        // 0 = off; a score is dropped when its hash byte < drop_p8
  uint32_t drop_seed;
  float drop_rscale;                  // 256 / (256 - drop_p8)
  int drop_head_off;                  // global index of local query head 0
};

}  // namespace lca
