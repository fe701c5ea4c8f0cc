// Python bindings (pybind11 via torch/extension.h) for the sm_100a kernels.
// Host-side work done here: argument validation, CUtensorMap encoding through the driver entry
// point (no link-time libcuda dependency, so the module imports on a CPU-only box), launching on
// the current torch CUDA stream.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <torch/extension.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "launchers.h"
#include "symm.h"

namespace lca {

// ------------------------------------------------------------------------------------ tensor maps
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool encode_tmap_4d(CUtensorMap* out, const void* base, int64_t D, int64_t H, int64_t S, int64_t B, int64_t stride_h,
                    int64_t stride_s, int64_t stride_b, int box_rows, const char** err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  if (reinterpret_cast<uintptr_t>(base) % 16) { *err = "tensor base not 16-byte aligned"; return false; }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(S),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(stride_h * 2), static_cast<cuuint64_t>(stride_s * 2),
                           static_cast<cuuint64_t>(stride_b * 2)};
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16) { *err = "tensor strides must be multiples of 8 elements"; return false; }
  cuuint32_t box[4] = {64, 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 /* 16-bit payload; fp16 uses the same map */, 4,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed"; return false; }
  return true;
}

// 1-byte (e4m3) tensors: box (128 bytes, 1, 128 rows, 1)
static void make_tmap_u8(CUtensorMap* m, const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.dim() == 4 && t.element_size() == 1 && t.stride(3) == 1, name, " must be a (B, S, H, D) 1-byte tensor");
  const int64_t B = t.size(0), S = t.size(1), H = t.size(2), D = t.size(3);
  int64_t sb = t.stride(0), ss = t.stride(1), sh = t.stride(2);
  if (B == 1) sb = S * ss;
  if (H == 1) sh = D;
  EncodeTiledFn fn = get_encode_fn();
  TORCH_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0 && sh % 16 == 0 && ss % 16 == 0 && sb % 16 == 0, name,
              ": 16-byte alignment");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(sh), static_cast<cuuint64_t>(ss), static_cast<cuuint64_t>(sb)};
  cuuint32_t box[4] = {128, 1, 128, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, t.data_ptr(), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, name, ": cuTensorMapEncodeTiled failed");
}

static void make_tmap(CUtensorMap* m, const at::Tensor& t, const char* name, int box_rows = 128) {
  TORCH_CHECK(t.dim() == 4, name, " must be (B, S, H, D)");
  TORCH_CHECK(t.stride(3) == 1, name, " last dim must be contiguous");
  const int64_t B = t.size(0), S = t.size(1), H = t.size(2), D = t.size(3);
  int64_t sb = t.stride(0), ss = t.stride(1), sh = t.stride(2);
  if (B == 1) sb = S * ss;           // stride of a size-1 dim is arbitrary; keep the map valid
  if (H == 1) sh = D;
  const char* err = nullptr;
  TORCH_CHECK(encode_tmap_4d(m, t.data_ptr(), D, H, S, B, sh, ss, sb, box_rows, &err), name, ": ", err ? err : "?");
}

static int dtype_code(const at::Tensor& t) {
  if (t.scalar_type() == at::kFloat) return 0;
  if (t.scalar_type() == at::kBFloat16) return 1;
  if (t.scalar_type() == at::kHalf) return 2;
  TORCH_CHECK(false, "unsupported dtype ", t.scalar_type());
}

static int num_sms() {
  static int n = 0;
  if (!n) n = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  return n;
}

// ---- dynamic tile scheduler state (EXPERIMENTAL, LCA_B200_DYN_SCHED=1): one monotonic counter per device; the host
// tracks its value at the start of each launch (every launch claims total_work + #CTAs indices: each CTA, push CTAs
// included, makes exactly one failing claim).  Assumes the
// launches of one device are issued in stream order (single compute stream).
struct SchedState {
  at::Tensor counter;
  uint32_t base = 0;
};
// ---- build-time defaults of the opt-in kernel variants: ONE place to flip once a variant has been validated on
// hardware (tools/validate_experimental.sh).  An environment variable, when set, always overrides the default.
namespace defaults {
constexpr int kDynSched = -1;    // LCA_B200_DYN_SCHED  dynamic tile scheduler: -1 = fused launches only (push CTAs join the compute pool)
constexpr int kPolyEvery = 4;    // LCA_B200_POLY_EVERY exp2 offload ratio of the forward (0, 2*, 3, 4, 6; *packed variant only), head_dim 128
constexpr int kPolyEveryD64 = 3; //                     the same for head_dim 64
}  // namespace defaults
static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// Dynamic tile scheduler.  Measured in round 2: on ONE GPU the static snake schedule is 1-4 % faster (CTAs advance in
// lock-step over neighbouring Q tiles, which keeps K/V tiles hot in L2), on the fused multi-GPU launches dynamic claiming
// wins (N=2, S=128K forward: 2098 vs 1856 TFLOPS; Ulysses S=32K forward 5.1 vs 7.3 ms) because the push CTAs rejoin
// the compute pool and late-arriving shards no longer stall a fixed share of the work.  Hence: default on for
// launches with a communication role, off otherwise; LCA_B200_DYN_SCHED=0/1 forces either.
static bool dyn_sched_enabled(int n_comm) {
  static int e = env_int("LCA_B200_DYN_SCHED", defaults::kDynSched);
  return e < 0 ? n_comm > 0 : e == 1;
}
template <typename P>
static void attach_sched(P& p, const at::Tensor& like, int sms, int n_comm) {
  if (!dyn_sched_enabled(n_comm)) return;
  static std::mutex mu;
  static std::map<int, SchedState> states;
  std::lock_guard<std::mutex> lock(mu);
  SchedState& st = states[like.get_device()];
  if (!st.counter.defined()) st.counter = at::zeros({1}, like.options().dtype(at::kInt));
  p.sched_counter = reinterpret_cast<uint32_t*>(st.counter.data_ptr<int>());
  p.sched_base = st.base;
  p.dyn_sched = 1;
  const int avail = sms - n_comm;
  int gc = p.total_work < avail ? p.total_work : avail;
  if (gc < 1) gc = 1;
  // every claiming CTA makes exactly one failing claim; the push CTAs claim too once their transfers are out
  st.base += static_cast<uint32_t>(p.total_work) + static_cast<uint32_t>(gc) + static_cast<uint32_t>(n_comm);
}

#define LCA_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    TORCH_CHECK(_e == cudaSuccess, #expr, " failed: ", cudaGetErrorString(_e));             \
  } while (0)

// ------------------------------------------------------------------------------------ fmha fwd
// qsegs[i] = {row0, nrows, pos0, flag, o_row0, o_base_ptr (0 -> `out`), o_sig_ptr, group [, lse_owner_ptr]}
// ksegs[i] = {row0, nrows, pos0, flag, group}
static void fill_fwd_params(FwdParams& p, const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
              const std::vector<std::vector<int64_t>>& qsegs, const std::vector<std::vector<int64_t>>& ksegs,
              int64_t q_pos_stride, int64_t k_pos_stride, at::Tensor& out, int64_t o_head_off, at::Tensor& lse,
              double scale, int64_t wl, int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi,
              int64_t flags_ptr, int64_t flag_epoch) {
  TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(), "q/k/v must be CUDA tensors");
  TORCH_CHECK(q.scalar_type() == at::kBFloat16 || q.scalar_type() == at::kHalf, "q must be bf16 or fp16");
  TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type(), "dtype mismatch");
  TORCH_CHECK(out.scalar_type() == q.scalar_type() && out.stride(3) == 1, "bad out");
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.dim() == 3 && lse.stride(2) == 1, "lse must be (B,H,S) fp32");
  const int64_t B = q.size(0), H = q.size(2), D = q.size(3), Hkv = k.size(2);
  TORCH_CHECK(D == 64 || D == 128, "head_dim must be 64 or 128 (got ", D, ")");
  TORCH_CHECK(k.size(3) == D && v.size(3) == D && v.size(2) == Hkv && H % Hkv == 0, "bad k/v shape");
  TORCH_CHECK(k.size(0) == B && v.size(0) == B && v.size(1) == k.size(1), "bad k/v batch/seq");
  TORCH_CHECK(!qsegs.empty() && qsegs.size() <= kMaxSeg && !ksegs.empty() && ksegs.size() <= kMaxSeg, "segment count");
  TORCH_CHECK(scale > 0, "softmax_scale must be positive");
  std::memset(&p, 0, sizeof(p));
  make_tmap(&p.tm_q, q, "q");
  make_tmap(&p.tm_k, k, "k");
  make_tmap(&p.tm_v, v, "v");
  p.n_qseg = static_cast<int>(qsegs.size());
  p.n_kseg = static_cast<int>(ksegs.size());
  int64_t pairs = 0;
  for (int i = 0; i < p.n_qseg; ++i) {
    const auto& s = qsegs[i];
    TORCH_CHECK(s.size() == 8 || s.size() == 9, "qseg needs 8 or 9 fields");
    TORCH_CHECK(s[0] >= 0 && s[1] > 0 && s[0] + s[1] <= q.size(1), "qseg rows out of range");
    p.qseg[i].row0 = static_cast<int>(s[0]);
    p.qseg[i].nrows = static_cast<int>(s[1]);
    p.qseg[i].pos0 = static_cast<int>(s[2]);
    p.qseg[i].flag = static_cast<int>(s[3]);
    p.qseg[i].o_row0 = static_cast<int>(s[4]);
    p.qseg[i].o_base = s[5] ? reinterpret_cast<void*>(s[5]) : out.data_ptr();
    p.qseg[i].o_sig = reinterpret_cast<uint32_t*>(s[6]);
    p.qseg[i].group = static_cast<int>(s[7]);
    p.qseg[i].lse_base = s.size() == 9 ? reinterpret_cast<float*>(s[8]) : nullptr;
    if (!s[5]) TORCH_CHECK(s[4] >= 0 && s[4] + s[1] <= out.size(1), "qseg output rows out of range");
    pairs += (s[1] + 255) / 256;
  }
  for (int i = 0; i < p.n_kseg; ++i) {
    const auto& s = ksegs[i];
    TORCH_CHECK(s.size() == 5, "kseg needs 5 fields");
    TORCH_CHECK(s[0] >= 0 && s[1] > 0 && s[0] + s[1] <= k.size(1), "kseg rows out of range");
    p.kseg[i].row0 = static_cast<int>(s[0]);
    p.kseg[i].nrows = static_cast<int>(s[1]);
    p.kseg[i].pos0 = static_cast<int>(s[2]);
    p.kseg[i].flag = static_cast<int>(s[3]);
    p.kseg[i].group = static_cast<int>(s[4]);
  }
  p.q_pos_stride = static_cast<int>(q_pos_stride);
  p.k_pos_stride = static_cast<int>(k_pos_stride);
  TORCH_CHECK(p.q_pos_stride > 0 && p.k_pos_stride > 0, "position strides must be positive");
  p.B = static_cast<int>(B);
  p.H = static_cast<int>(H);
  p.Hkv = static_cast<int>(Hkv);
  p.total_work = static_cast<int>(pairs * B * H);
  p.wl = static_cast<int>(wl);
  p.wr = static_cast<int>(wr);
  p.scale = static_cast<float>(scale);
  p.scale_log2 = static_cast<float>(scale * 1.4426950408889634);
  p.softcap = static_cast<float>(softcap);
  if (alibi.has_value() && alibi->defined()) {
    const at::Tensor& a = *alibi;
    TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kFloat && a.is_contiguous(), "alibi_slopes must be fp32 CUDA contiguous");
    TORCH_CHECK((a.dim() == 1 && a.size(0) == H) || (a.dim() == 2 && a.size(0) == B && a.size(1) == H), "alibi shape");
    p.alibi = a.data_ptr<float>();
    p.alibi_bstride = a.dim() == 2 ? static_cast<int>(H) : 0;
  }
  p.o_sb = out.stride(0);
  p.o_ss = out.stride(1);
  p.o_sh = out.stride(2);
  TORCH_CHECK(p.o_ss % 8 == 0 && p.o_sh % 8 == 0 && (B == 1 || p.o_sb % 8 == 0), "out strides must be multiples of 8");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(out.data_ptr()) % 16 == 0, "out not 16B aligned");
  p.o_head_off = static_cast<int>(o_head_off);
  p.lse = lse.data_ptr<float>();
  p.lse_sb = lse.stride(0);
  p.lse_sh = lse.stride(1);
  TORCH_CHECK(lse.size(0) == B && lse.size(1) == H && lse.size(2) >= q.size(1), "lse shape");
  p.flags = reinterpret_cast<const uint32_t*>(flags_ptr);
  p.flag_epoch = static_cast<uint32_t>(flag_epoch);
  {
    static int poly = env_int("LCA_B200_POLY_EVERY", -1);
    // measured r2 (S=32K, packed arithmetic): D=128 is flat in the offload ratio (4: 1.971 ms, 6: 1.978 ms, 3: 2.03 ms);
    // D=64 has twice the exponentials per tensor FLOP and wants the heavier offload (3: 3.01 ms, 6: 3.24 ms, 2: 3.83 ms)
    p.poly_every = poly >= 0 ? poly : (D == 64 ? defaults::kPolyEveryD64 : defaults::kPolyEvery);
  }
  p.lse_own_sb = out.size(2) * out.size(1);      // owners keep (B, H_total, rows) next to their (B, rows, H_total, D) output
  p.lse_own_sh = out.size(1);
}

// drop = {} (no dropout) or {p8, seed, global index of local query head 0}: EXPERIMENTAL dropout instantiations
template <typename P>
static void set_dropout(P& p, const std::vector<int64_t>& drop, double softcap) {
  if (drop.empty() || drop[0] <= 0) return;
  TORCH_CHECK(drop.size() == 3 && drop[0] < 256, "drop = {p8, seed, head_offset}");
  TORCH_CHECK(softcap == 0.0, "native dropout does not combine with softcap (use the PyTorch engine)");
  p.drop_p8 = static_cast<int>(drop[0]);
  p.drop_seed = static_cast<uint32_t>(drop[1] & 0xFFFFFFFFll);
  p.drop_rscale = 256.0f / static_cast<float>(256 - p.drop_p8);
  p.drop_head_off = static_cast<int>(drop[2]);
}

// The fused launches (usp_fwd / usp_bwd_pass) have long positional signatures; dropout for the NEXT fused launch of
// this thread is handed over separately (EXPERIMENTAL, consumed by exactly one launch).
static thread_local std::vector<int64_t> g_next_drop;
void set_next_dropout(const std::vector<int64_t>& drop) { g_next_drop = drop; }
template <typename P>
static void take_next_dropout(P& p, double softcap) {
  if (g_next_drop.empty()) return;
  std::vector<int64_t> d;
  d.swap(g_next_drop);
  set_dropout(p, d, softcap);
}

static void fmha_fwd_impl(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                          const std::vector<std::vector<int64_t>>& qsegs, const std::vector<std::vector<int64_t>>& ksegs,
                          int64_t q_pos_stride, int64_t k_pos_stride, at::Tensor& out, int64_t o_head_off, at::Tensor& lse,
                          double scale, int64_t wl, int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi,
                          int64_t flags_ptr, int64_t flag_epoch, int64_t sm_limit, const std::vector<int64_t>& drop) {
  c10::cuda::CUDAGuard guard(q.device());
  FwdParams p;
  fill_fwd_params(p, q, k, v, qsegs, ksegs, q_pos_stride, k_pos_stride, out, o_head_off, lse, scale, wl, wr, softcap,
                  alibi, flags_ptr, flag_epoch);
  set_dropout(p, drop, softcap);
  int sms = num_sms();
  if (sm_limit > 0 && sm_limit < sms) sms = static_cast<int>(sm_limit);
  if (p.drop_p8 == 0) attach_sched(p, q, sms, 0);     // the dropout instantiations use the static schedule
  LCA_CUDA_OK(launch_fmha_fwd(p, static_cast<int>(q.size(3)), q.scalar_type() == at::kBFloat16, sms,
                              at::cuda::getCurrentCUDAStream()));
}

void fmha_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
              const std::vector<std::vector<int64_t>>& qsegs, const std::vector<std::vector<int64_t>>& ksegs,
              int64_t q_pos_stride, int64_t k_pos_stride, at::Tensor& out, int64_t o_head_off, at::Tensor& lse,
              double scale, int64_t wl, int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi,
              int64_t flags_ptr, int64_t flag_epoch, int64_t sm_limit) {
  fmha_fwd_impl(q, k, v, qsegs, ksegs, q_pos_stride, k_pos_stride, out, o_head_off, lse, scale, wl, wr, softcap, alibi,
                flags_ptr, flag_epoch, sm_limit, {});
}

void fmha_fwd_drop(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v,
                   const std::vector<std::vector<int64_t>>& qsegs, const std::vector<std::vector<int64_t>>& ksegs,
                   int64_t q_pos_stride, int64_t k_pos_stride, at::Tensor& out, int64_t o_head_off, at::Tensor& lse,
                   double scale, int64_t wl, int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi,
                   int64_t flags_ptr, int64_t flag_epoch, int64_t sm_limit, const std::vector<int64_t>& drop) {
  fmha_fwd_impl(q, k, v, qsegs, ksegs, q_pos_stride, k_pos_stride, out, o_head_off, lse, scale, wl, wr, softcap, alibi,
                flags_ptr, flag_epoch, sm_limit, drop);
}

// mesh = {P, U, R, u, r, rows, n_comm};  qlike/kvlike: user shards to push with their destination byte offsets.
static void fill_comm(CommParams& c, const std::vector<int64_t>& mesh, const std::vector<at::Tensor>& qlike,
                      const std::vector<int64_t>& q_offs, const std::vector<at::Tensor>& kvlike,
                      const std::vector<int64_t>& kv_offs, const std::vector<at::Tensor>& stats,
                      const std::vector<int64_t>& stat_offs, bool q_to_all, int64_t stage_q_rows, int64_t stage_kv_rows, const std::vector<int64_t>& peer_slabs,
                      const std::vector<int64_t>& peer_sigs, int64_t my_sig, int64_t epoch, int64_t o_target, int64_t H,
                      int64_t Hkv) {
  TORCH_CHECK(mesh.size() >= 7 && mesh.size() <= 9, "mesh arity");   // [P, U, R, u, r, rows, n_comm(, kv_dst_mask, q_dst_mask)]
  const int P = static_cast<int>(mesh[0]), U = static_cast<int>(mesh[1]), R = static_cast<int>(mesh[2]);
  TORCH_CHECK(P == U * R && P <= kMaxPeers && static_cast<int>(peer_slabs.size()) == P &&
              static_cast<int>(peer_sigs.size()) == P, "peer tables");
  c.n_comm = static_cast<int>(mesh[6]);
  TORCH_CHECK(c.n_comm >= 1 && c.n_comm <= 64, "n_comm");
  c.P = P; c.U = U; c.R = R;
  c.u = static_cast<int>(mesh[3]); c.r = static_cast<int>(mesh[4]);
  c.rows = static_cast<int>(mesh[5]);
  c.kv_dst_mask = mesh.size() > 7 ? static_cast<unsigned int>(mesh[7]) : 0xFFFFFFFFu;   // destinations that need my K/V data
  c.q_dst_mask = mesh.size() > 8 ? static_cast<unsigned int>(mesh[8]) : 0xFFFFFFFFu;    // ... my Q-like data
  TORCH_CHECK(qlike.size() <= 2 && kvlike.size() <= 2 && qlike.size() == q_offs.size() && kvlike.size() == kv_offs.size() &&
              !kvlike.empty(), "push tensor lists");
  const at::Tensor& k0 = kvlike[0];
  c.B = static_cast<int>(k0.size(0));
  c.H = static_cast<int>(H); c.Hkv = static_cast<int>(Hkv); c.D = static_cast<int>(k0.size(3));
  TORCH_CHECK(c.H % U == 0, "query heads must be divisible by the Ulysses degree");
  TORCH_CHECK(c.Hkv % U == 0 || U % c.Hkv == 0, "kv heads must divide or be divisible by the Ulysses degree");
  c.Hl = c.H / U;
  c.Hkvl = c.Hkv >= U ? c.Hkv / U : 1;
  auto fill = [&](PushTensor& pt, const at::Tensor& t, int64_t off, int64_t heads) {
    TORCH_CHECK(t.is_cuda() && t.dim() == 4 && t.size(1) == c.rows && t.size(2) == heads && t.size(3) == c.D &&
                t.size(0) == c.B, "push shard shape");
    TORCH_CHECK(t.stride(3) == 1 && t.stride(2) == t.size(3), "push shards need dense (head, dim) axes");
    TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0 && t.stride(1) % 8 == 0 && t.stride(0) % 8 == 0,
                "push shards must be 16-byte aligned");
    pt.src = t.data_ptr(); pt.sb = t.stride(0); pt.ss = t.stride(1); pt.off = off;
  };
  c.n_q = static_cast<int>(qlike.size());
  c.n_kv = static_cast<int>(kvlike.size());
  for (int i = 0; i < c.n_q; ++i) fill(c.qt[i], qlike[i], q_offs[i], c.H);
  for (int i = 0; i < c.n_kv; ++i) fill(c.kvt[i], kvlike[i], kv_offs[i], c.Hkv);
  TORCH_CHECK(stats.size() <= 2 && stats.size() == stat_offs.size(), "stat lists");
  c.n_stat = static_cast<int>(stats.size());
  c.q_to_all = q_to_all ? 1 : 0;
  for (int i = 0; i < c.n_stat; ++i) {
    const at::Tensor& st = stats[i];
    TORCH_CHECK(st.scalar_type() == at::kFloat && st.is_contiguous() && st.dim() == 3 && st.size(0) == c.B &&
                st.size(1) == c.H && st.size(2) == c.rows && c.rows % 4 == 0, "stat must be contiguous (B,H,rows) fp32");
    c.stat[i] = st.data_ptr<float>();
    c.stat_off[i] = stat_offs[i];
  }
  for (int i = 0; i < P; ++i) {
    c.peer_slab[i] = reinterpret_cast<unsigned char*>(peer_slabs[i]);
    c.peer_sig[i] = reinterpret_cast<unsigned int*>(peer_sigs[i]);
  }
  c.my_sig = reinterpret_cast<unsigned int*>(my_sig);
  c.stage_q_rows = stage_q_rows; c.stage_kv_rows = stage_kv_rows;
  c.epoch = static_cast<unsigned int>(epoch);
  c.o_target = static_cast<unsigned int>(o_target);
  static const int push_mode = [] {
    const char* v = std::getenv("LCA_B200_PUSH");
    if (!v || !*v || std::string(v) == "bulk") return 1;
    TORCH_CHECK(std::string(v) == "scalar", "LCA_B200_PUSH must be 'bulk' or 'scalar'");
    return 0;
  }();
  c.push_mode = push_mode;
  // rows wider than one 32 KiB stage are split by the bulk engine; every size it moves is a multiple of 16 bytes
  // because head slices are >= 128 bytes and `rows` is a multiple of 8 (FusedUSPEngine.supports_shapes)
  static const long long wd_s = env_int("LCA_B200_WATCHDOG_S", 600);
  c.watchdog_ns = wd_s > 0 ? static_cast<unsigned long long>(wd_s) * 1000000000ull : 0ull;
}

// EXPERIMENTAL fp8 (e4m3) forward.  q8/k8/v8: (B,S,H|Hkv,128) uint8/float8 views, scales from quantize_e4m3.
void fmha_fwd_fp8(const at::Tensor& q8, const at::Tensor& k8, const at::Tensor& v8, const at::Tensor& q_scale,
                  const at::Tensor& k_scale, const at::Tensor& v_scale, const std::vector<std::vector<int64_t>>& qsegs,
                  const std::vector<std::vector<int64_t>>& ksegs, int64_t q_pos_stride, int64_t k_pos_stride, at::Tensor& out,
                  at::Tensor& lse, double scale, int64_t wl, int64_t wr, double softcap,
                  const c10::optional<at::Tensor>& alibi) {
  TORCH_CHECK(q8.is_cuda() && q8.element_size() == 1 && k8.element_size() == 1 && v8.element_size() == 1, "fp8 operands");
  const int64_t B = q8.size(0), H = q8.size(2), D = q8.size(3), Hkv = k8.size(2);
  TORCH_CHECK(D == 128 && k8.size(3) == 128 && v8.size(3) == 128, "fp8 path: head_dim 128 only");
  TORCH_CHECK(H % Hkv == 0 && v8.size(2) == Hkv && k8.size(0) == B && v8.size(1) == k8.size(1), "bad k/v shape");
  TORCH_CHECK(out.scalar_type() == at::kBFloat16 && out.stride(3) == 1 && lse.scalar_type() == at::kFloat, "out bf16 / lse fp32");
  for (const at::Tensor* t : {&q_scale, &k_scale, &v_scale}) TORCH_CHECK(t->scalar_type() == at::kFloat && t->is_contiguous(), "scales");
  TORCH_CHECK(q_scale.dim() == 3 && q_scale.size(0) == B && q_scale.size(1) == H && q_scale.size(2) == (q8.size(1) + 127) / 128, "q_scale");
  TORCH_CHECK(k_scale.dim() == 3 && k_scale.size(1) == Hkv && k_scale.size(2) == (k8.size(1) + 127) / 128, "k_scale");
  TORCH_CHECK(v_scale.dim() == 2 && v_scale.size(0) == B && v_scale.size(1) == Hkv, "v_scale");
  TORCH_CHECK(!qsegs.empty() && qsegs.size() <= kMaxSeg && !ksegs.empty() && ksegs.size() <= kMaxSeg, "segment count");
  c10::cuda::CUDAGuard guard(q8.device());
  FwdParams p;
  std::memset(&p, 0, sizeof(p));
  make_tmap_u8(&p.tm_q, q8, "q8");
  make_tmap_u8(&p.tm_k, k8, "k8");
  make_tmap_u8(&p.tm_v, v8, "v8");
  p.n_qseg = static_cast<int>(qsegs.size());
  p.n_kseg = static_cast<int>(ksegs.size());
  int64_t pairs = 0;
  for (int i = 0; i < p.n_qseg; ++i) {
    const auto& s = qsegs[i];
    TORCH_CHECK(s.size() >= 8 && s[0] % 128 == 0 && s[0] + s[1] <= q8.size(1), "fp8 q segments must start on 128-row blocks");
    p.qseg[i].row0 = static_cast<int>(s[0]); p.qseg[i].nrows = static_cast<int>(s[1]); p.qseg[i].pos0 = static_cast<int>(s[2]);
    p.qseg[i].flag = -1; p.qseg[i].o_row0 = static_cast<int>(s[4]); p.qseg[i].o_base = out.data_ptr();
    p.qseg[i].o_sig = nullptr; p.qseg[i].group = static_cast<int>(s[7]); p.qseg[i].lse_base = nullptr;
    pairs += (s[1] + 255) / 256;
  }
  for (int i = 0; i < p.n_kseg; ++i) {
    const auto& s = ksegs[i];
    TORCH_CHECK(s.size() == 5 && s[0] % 128 == 0 && s[0] + s[1] <= k8.size(1), "fp8 k segments must start on 128-row blocks");
    p.kseg[i].row0 = static_cast<int>(s[0]); p.kseg[i].nrows = static_cast<int>(s[1]); p.kseg[i].pos0 = static_cast<int>(s[2]);
    p.kseg[i].flag = -1; p.kseg[i].group = static_cast<int>(s[4]);
  }
  p.q_pos_stride = static_cast<int>(q_pos_stride); p.k_pos_stride = static_cast<int>(k_pos_stride);
  p.B = static_cast<int>(B); p.H = static_cast<int>(H); p.Hkv = static_cast<int>(Hkv);
  p.total_work = static_cast<int>(pairs * B * H);
  p.wl = static_cast<int>(wl); p.wr = static_cast<int>(wr);
  p.scale = static_cast<float>(scale); p.scale_log2 = static_cast<float>(scale * 1.4426950408889634);
  p.softcap = static_cast<float>(softcap);
  if (alibi.has_value() && alibi->defined()) {
    TORCH_CHECK(alibi->is_cuda() && alibi->scalar_type() == at::kFloat && alibi->is_contiguous(), "alibi");
    p.alibi = alibi->data_ptr<float>();
    p.alibi_bstride = alibi->dim() == 2 ? static_cast<int>(H) : 0;
  }
  p.o_sb = out.stride(0); p.o_ss = out.stride(1); p.o_sh = out.stride(2);
  p.lse = lse.data_ptr<float>(); p.lse_sb = lse.stride(0); p.lse_sh = lse.stride(1);
  p.poly_every = 6;
  p.q_scale = q_scale.data_ptr<float>(); p.k_scale = k_scale.data_ptr<float>(); p.v_scale = v_scale.data_ptr<float>();
  p.q_scale_sb = q_scale.stride(0); p.q_scale_sh = q_scale.stride(1);
  p.k_scale_sb = k_scale.stride(0); p.k_scale_sh = k_scale.stride(1);
  LCA_CUDA_OK(launch_fmha_fwd_fp8(p, 128, num_sms(), at::cuda::getCurrentCUDAStream()));
}

// bf16/fp16 (B,S,H,D) -> (e4m3 bytes (B,S,H,D), scale).  per_head: one scale per (b,h) [V]; else per 128-row block [Q, K].
std::vector<at::Tensor> quantize_e4m3(const at::Tensor& x, bool per_head) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.stride(3) == 1 && (x.scalar_type() == at::kBFloat16 || x.scalar_type() == at::kHalf));
  const int B = x.size(0), S = x.size(1), H = x.size(2), D = x.size(3);
  TORCH_CHECK(D % 8 == 0 && x.stride(2) % 8 == 0 && x.stride(1) % 8 == 0, "alignment");
  c10::cuda::CUDAGuard guard(x.device());
  at::Tensor y = at::empty({B, S, H, D}, x.options().dtype(at::kByte));
  at::Tensor scale, ext;
  const float* ext_p = nullptr;
  const int nblk = (S + 127) / 128;
  if (per_head) {
    ext = (x.abs().amax({1, 3}).to(at::kFloat) / 448.0).clamp_min(1e-12).contiguous();     // (B, H)
    ext_p = ext.data_ptr<float>();
    scale = ext;
  } else {
    scale = at::empty({B, H, nblk}, x.options().dtype(at::kFloat));
  }
  LCA_CUDA_OK(launch_quant_e4m3(x.data_ptr(), dtype_code(x), y.data_ptr<uint8_t>(), per_head ? nullptr : scale.data_ptr<float>(),
                                ext_p, B, S, H, D, x.stride(0), x.stride(1), x.stride(2), at::cuda::getCurrentCUDAStream()));
  return {y, scale};
}

// Fused USP forward: the same kernel with `n_comm` communication CTAs that push this rank's q/k/v shards
// (`uq`, `uk`, `uv`: user tensors (B, S/P, H|Hkv, D)) into the peers' staging buffers while the compute CTAs
// consume `q`, `k`, `v` (= views of MY staging buffers, or the user q when U == 1) segment by segment.
// mesh = {P, U, R, u, r, rows, n_comm};  offs = {off_q, off_k, off_v, stage_q_rows, stage_kv_rows}
void usp_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& uq, const at::Tensor& uk,
             const at::Tensor& uv, const std::vector<std::vector<int64_t>>& qsegs,
             const std::vector<std::vector<int64_t>>& ksegs, int64_t q_pos_stride, int64_t k_pos_stride, at::Tensor& out,
             int64_t o_head_off, at::Tensor& lse, double scale, int64_t wl, int64_t wr, double softcap,
             const c10::optional<at::Tensor>& alibi, const std::vector<int64_t>& mesh, const std::vector<int64_t>& offs,
             const std::vector<int64_t>& peer_slabs, const std::vector<int64_t>& peer_sigs, int64_t my_sig, int64_t epoch,
             int64_t o_target) {
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(offs.size() == 5, "offs arity");
  FwdParams p;
  const int64_t n_comm = mesh.at(6);
  fill_fwd_params(p, q, k, v, qsegs, ksegs, q_pos_stride, k_pos_stride, out, o_head_off, lse, scale, wl, wr, softcap,
                  alibi, my_sig, epoch * n_comm);
  std::vector<at::Tensor> ql, kvl = {uk, uv};
  std::vector<int64_t> qo, kvo = {offs[1], offs[2]};
  if (mesh.at(1) > 1) { ql.push_back(uq); qo.push_back(offs[0]); }
  fill_comm(p.comm, mesh, ql, qo, kvl, kvo, {}, {}, false, offs[3], offs[4], peer_slabs, peer_sigs, my_sig, epoch,
            o_target, uq.size(2), uk.size(2));
  take_next_dropout(p, softcap);
  if (p.drop_p8 == 0) attach_sched(p, q, num_sms(), p.comm.n_comm);
  LCA_CUDA_OK(launch_fmha_fwd(p, static_cast<int>(q.size(3)), q.scalar_type() == at::kBFloat16, num_sms(),
                              at::cuda::getCurrentCUDAStream()));
}

// ------------------------------------------------------------------------------------ fmha bwd
// One pass of the backward (see fmha_bwd_sm100.cu).  x0/x1 stationary, y0/y1 streamed.
// xsegs[i] = {row0, nrows, pos0, group, o_row0 [, flag, o_base0, o_base1, o_sig]};  ysegs[i] = {row0, nrows, pos0, flag, group}
static void fill_bwd_params(BwdParams& p, bool is_dkv, const at::Tensor& x0, const at::Tensor& x1, const at::Tensor& y0,
                   const at::Tensor& y1, const std::vector<std::vector<int64_t>>& xsegs,
                   const std::vector<std::vector<int64_t>>& ysegs, int64_t x_pos_stride, int64_t y_pos_stride,
                   const at::Tensor& lse2, const at::Tensor& delta, at::Tensor& out0, const c10::optional<at::Tensor>& out1,
                   int64_t out_mode, double scale, int64_t wl, int64_t wr, double softcap,
                   const c10::optional<at::Tensor>& alibi) {
  TORCH_CHECK(x0.is_cuda() && (x0.scalar_type() == at::kBFloat16 || x0.scalar_type() == at::kHalf), "x0 must be CUDA bf16/fp16");
  for (const at::Tensor* t : {&x1, &y0, &y1}) TORCH_CHECK(t->scalar_type() == x0.scalar_type() && t->is_cuda(), "dtype mismatch");
  const int64_t B = x0.size(0), Hx = x0.size(2), Hy = y0.size(2), D = x0.size(3);
  TORCH_CHECK(D == 64 || D == 128, "head_dim must be 64 or 128");
  TORCH_CHECK(x1.sizes() == x0.sizes() && y1.sizes() == y0.sizes() && y0.size(0) == B && y0.size(3) == D, "shape mismatch");
  TORCH_CHECK(!xsegs.empty() && xsegs.size() <= kMaxSeg && !ysegs.empty() && ysegs.size() <= kMaxSeg, "segment count");
  const int64_t Hq = is_dkv ? Hy : Hx;
  TORCH_CHECK(is_dkv ? (Hy % Hx == 0) : (Hx % Hy == 0), "head counts");
  TORCH_CHECK(lse2.scalar_type() == at::kFloat && delta.scalar_type() == at::kFloat && lse2.dim() == 3 &&
              lse2.sizes() == delta.sizes() && lse2.stride(2) == 1 && delta.strides() == lse2.strides(), "lse2/delta");
  TORCH_CHECK(lse2.size(0) == B && lse2.size(1) == Hq, "lse2 shape");
  std::memset(&p, 0, sizeof(p));
  make_tmap(&p.tm_x0, x0, "x0", 128);
  make_tmap(&p.tm_x1, x1, "x1", 128);
  p.y_rows = 64;                 // streamed rows per tile (both passes)
  make_tmap(&p.tm_y0, y0, "y0", p.y_rows);
  make_tmap(&p.tm_y1, y1, "y1", p.y_rows);
  p.n_xseg = static_cast<int>(xsegs.size());
  p.n_yseg = static_cast<int>(ysegs.size());
  int64_t tiles = 0;
  for (int i = 0; i < p.n_xseg; ++i) {
    const auto& s = xsegs[i];
    TORCH_CHECK((s.size() == 5 || s.size() == 9) && s[0] >= 0 && s[1] > 0 && s[0] + s[1] <= x0.size(1), "bad xseg");
    const bool ext = s.size() == 9;
    if (!ext || s[6] == 0) TORCH_CHECK(s[4] >= 0 && s[4] + s[1] <= out0.size(1), "xseg output rows out of range");
    p.xseg[i].row0 = static_cast<int>(s[0]);
    p.xseg[i].nrows = static_cast<int>(s[1]);
    p.xseg[i].pos0 = static_cast<int>(s[2]);
    p.xseg[i].group = static_cast<int>(s[3]);
    p.xseg[i].o_row0 = static_cast<int>(s[4]);
    p.xseg[i].flag = ext ? static_cast<int>(s[5]) : -1;
    p.xseg[i].o_base0 = ext ? reinterpret_cast<void*>(s[6]) : nullptr;
    p.xseg[i].o_base1 = ext ? reinterpret_cast<void*>(s[7]) : nullptr;
    p.xseg[i].o_sig = ext ? reinterpret_cast<uint32_t*>(s[8]) : nullptr;
    tiles += (s[1] + 127) / 128;
  }
  for (int i = 0; i < p.n_yseg; ++i) {
    const auto& s = ysegs[i];
    TORCH_CHECK(s.size() == 5 && s[0] >= 0 && s[1] > 0 && s[0] + s[1] <= y0.size(1), "bad yseg");
    p.yseg[i] = {static_cast<int>(s[0]), static_cast<int>(s[1]), static_cast<int>(s[2]), static_cast<int>(s[3]),
                 static_cast<int>(s[4])};
  }
  p.x_pos_stride = static_cast<int>(x_pos_stride);
  p.y_pos_stride = static_cast<int>(y_pos_stride);
  p.x_heavy_last = is_dkv ? 0 : 1;
  p.B = static_cast<int>(B);
  p.Hx = static_cast<int>(Hx);
  p.n_inner = is_dkv ? static_cast<int>(Hy / Hx) : 1;
  p.hx_per_hy = is_dkv ? 1 : static_cast<int>(Hx / Hy);
  p.total_work = static_cast<int>(tiles * B * Hx);
  p.wl = static_cast<int>(wl);
  p.wr = static_cast<int>(wr);
  p.scale = static_cast<float>(scale);
  p.scale_log2 = static_cast<float>(scale * 1.4426950408889634);
  p.softcap = static_cast<float>(softcap);
  if (alibi.has_value() && alibi->defined()) {
    const at::Tensor& a = *alibi;
    TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kFloat && a.is_contiguous(), "alibi_slopes must be fp32 CUDA contiguous");
    TORCH_CHECK((a.dim() == 1 && a.size(0) == Hq) || (a.dim() == 2 && a.size(0) == B && a.size(1) == Hq), "alibi shape");
    p.alibi = a.data_ptr<float>();
    p.alibi_bstride = a.dim() == 2 ? static_cast<int>(Hq) : 0;
  }
  p.lse2 = lse2.data_ptr<float>();
  p.delta = delta.data_ptr<float>();
  p.stat_sb = lse2.stride(0);
  p.stat_sh = lse2.stride(1);
  TORCH_CHECK(out0.dim() == 4 && out0.size(0) == B && out0.size(3) == D && out0.stride(3) == 1, "out0 shape");
  const bool f32 = out0.scalar_type() == at::kFloat;
  TORCH_CHECK(f32 || out0.scalar_type() == x0.scalar_type(), "out dtype");
  TORCH_CHECK(out_mode == 0 ? !f32 : f32, "out_mode 1/2/3 need fp32 outputs, mode 0 a 16-bit output");
  p.out0 = out0.data_ptr();
  p.o_sb = out0.stride(0);
  p.o_ss = out0.stride(1);
  p.o_sh = out0.stride(2);
  TORCH_CHECK(p.o_ss % 8 == 0 && p.o_sh % 8 == 0 && reinterpret_cast<uintptr_t>(p.out0) % 16 == 0, "out0 alignment");
  if (is_dkv) {
    TORCH_CHECK(out1.has_value() && out1->sizes() == out0.sizes() && out1->strides() == out0.strides() &&
                out1->scalar_type() == out0.scalar_type(), "out1 must match out0");
    p.out1 = out1->data_ptr();
    TORCH_CHECK(reinterpret_cast<uintptr_t>(p.out1) % 16 == 0, "out1 alignment");
  }
  p.out_mode = static_cast<int>(out_mode);
}

static void fmha_bwd_pass_impl(bool is_dkv, const at::Tensor& x0, const at::Tensor& x1, const at::Tensor& y0,
                   const at::Tensor& y1,
                   const std::vector<std::vector<int64_t>>& xsegs, const std::vector<std::vector<int64_t>>& ysegs,
                   int64_t x_pos_stride, int64_t y_pos_stride, const at::Tensor& lse2, const at::Tensor& delta,
                   at::Tensor& out0, const c10::optional<at::Tensor>& out1, bool accumulate, double scale, int64_t wl,
                   int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi, int64_t sm_limit,
                   const std::vector<int64_t>& drop) {
  c10::cuda::CUDAGuard guard(x0.device());
  BwdParams p;
  const bool f32 = out0.scalar_type() == at::kFloat;
  TORCH_CHECK(!accumulate || f32, "accumulate needs fp32 outputs");
  TORCH_CHECK(out0.size(2) == x0.size(2), "out heads");
  fill_bwd_params(p, is_dkv, x0, x1, y0, y1, xsegs, ysegs, x_pos_stride, y_pos_stride, lse2, delta, out0, out1,
                  f32 ? (accumulate ? 2 : 1) : 0, scale, wl, wr, softcap, alibi);
  set_dropout(p, drop, softcap);
  int sms = num_sms();
  if (sm_limit > 0 && sm_limit < sms) sms = static_cast<int>(sm_limit);
  if (p.drop_p8 == 0) attach_sched(p, x0, sms, 0);
  LCA_CUDA_OK(launch_fmha_bwd(p, static_cast<int>(x0.size(3)), x0.scalar_type() == at::kBFloat16, is_dkv, sms,
                              at::cuda::getCurrentCUDAStream()));
}

void fmha_bwd_pass(bool is_dkv, const at::Tensor& x0, const at::Tensor& x1, const at::Tensor& y0, const at::Tensor& y1,
                   const std::vector<std::vector<int64_t>>& xsegs, const std::vector<std::vector<int64_t>>& ysegs,
                   int64_t x_pos_stride, int64_t y_pos_stride, const at::Tensor& lse2, const at::Tensor& delta,
                   at::Tensor& out0, const c10::optional<at::Tensor>& out1, bool accumulate, double scale, int64_t wl,
                   int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi, int64_t sm_limit) {
  fmha_bwd_pass_impl(is_dkv, x0, x1, y0, y1, xsegs, ysegs, x_pos_stride, y_pos_stride, lse2, delta, out0, out1, accumulate,
                     scale, wl, wr, softcap, alibi, sm_limit, {});
}

void fmha_bwd_pass_drop(bool is_dkv, const at::Tensor& x0, const at::Tensor& x1, const at::Tensor& y0,
                        const at::Tensor& y1, const std::vector<std::vector<int64_t>>& xsegs,
                        const std::vector<std::vector<int64_t>>& ysegs, int64_t x_pos_stride, int64_t y_pos_stride,
                        const at::Tensor& lse2, const at::Tensor& delta, at::Tensor& out0,
                        const c10::optional<at::Tensor>& out1, bool accumulate, double scale, int64_t wl, int64_t wr,
                        double softcap, const c10::optional<at::Tensor>& alibi, int64_t sm_limit,
                        const std::vector<int64_t>& drop) {
  fmha_bwd_pass_impl(is_dkv, x0, x1, y0, y1, xsegs, ysegs, x_pos_stride, y_pos_stride, lse2, delta, out0, out1, accumulate,
                     scale, wl, wr, softcap, alibi, sm_limit, drop);
}

// Fused USP backward pass.  dQ pass (is_dkv = false) carries the push CTAs (q, dO | k, v | delta) and scatters dQ tiles
// to the token owners; the dK/dV pass reduces its partial tiles into the owners' fp32 accumulators with red.add over
// NVLink (out_mode 3).  `out0`/`out1` give the destination layout (strides, dtype); per-segment bases live in xsegs.
// comm_lists = {} for a pass without push CTAs.
void usp_bwd_pass(bool is_dkv, const at::Tensor& x0, const at::Tensor& x1, const at::Tensor& y0, const at::Tensor& y1,
                  const std::vector<std::vector<int64_t>>& xsegs, const std::vector<std::vector<int64_t>>& ysegs,
                  int64_t x_pos_stride, int64_t y_pos_stride, const at::Tensor& lse2, const at::Tensor& delta,
                  at::Tensor& out0, const c10::optional<at::Tensor>& out1, int64_t out_mode, int64_t o_head_off, double scale,
                  int64_t wl, int64_t wr, double softcap, const c10::optional<at::Tensor>& alibi, int64_t flags_ptr,
                  int64_t flag_epoch, const std::vector<int64_t>& mesh, const std::vector<at::Tensor>& qlike,
                  const std::vector<int64_t>& q_offs, const std::vector<at::Tensor>& kvlike,
                  const std::vector<int64_t>& kv_offs, const std::vector<at::Tensor>& stats,
                  const std::vector<int64_t>& stat_offs, bool q_to_all, int64_t stage_q_rows, int64_t stage_kv_rows,
                  const std::vector<int64_t>& peer_slabs,
                  const std::vector<int64_t>& peer_sigs, int64_t my_sig, int64_t epoch, int64_t o_target, int64_t H,
                  int64_t Hkv) {
  c10::cuda::CUDAGuard guard(x0.device());
  BwdParams p;
  fill_bwd_params(p, is_dkv, x0, x1, y0, y1, xsegs, ysegs, x_pos_stride, y_pos_stride, lse2, delta, out0, out1, out_mode,
                  scale, wl, wr, softcap, alibi);
  p.o_head_off = static_cast<int>(o_head_off);
  p.flags = reinterpret_cast<const uint32_t*>(flags_ptr);
  p.flag_epoch = static_cast<uint32_t>(flag_epoch);
  if (!mesh.empty())
    fill_comm(p.comm, mesh, qlike, q_offs, kvlike, kv_offs, stats, stat_offs, q_to_all, stage_q_rows, stage_kv_rows,
              peer_slabs, peer_sigs, my_sig, epoch, o_target, H, Hkv);
  take_next_dropout(p, softcap);
  if (p.drop_p8 == 0) attach_sched(p, x0, num_sms(), p.comm.n_comm);
  LCA_CUDA_OK(launch_fmha_bwd(p, static_cast<int>(x0.size(3)), x0.scalar_type() == at::kBFloat16, is_dkv, num_sms(),
                              at::cuda::getCurrentCUDAStream()));
}

void symm_wait(int64_t sig_ptr, int64_t target) {
  LCA_CUDA_OK(launch_wait_counter(reinterpret_cast<const uint32_t*>(sig_ptr), static_cast<uint32_t>(target),
                                  at::cuda::getCurrentCUDAStream()));
}

// ------------------------------------------------------------------------------------ utilities
void merge_out_lse(at::Tensor& out_acc, at::Tensor& lse_acc, const at::Tensor& block_out, const at::Tensor& block_lse) {
  TORCH_CHECK(out_acc.is_cuda() && out_acc.scalar_type() == at::kFloat && out_acc.is_contiguous() && out_acc.dim() == 4);
  TORCH_CHECK(lse_acc.scalar_type() == at::kFloat && lse_acc.is_contiguous() && block_lse.scalar_type() == at::kFloat &&
              block_lse.is_contiguous());
  TORCH_CHECK(block_out.is_contiguous() && block_out.sizes() == out_acc.sizes());
  const int B = out_acc.size(0), S = out_acc.size(1), H = out_acc.size(2), D = out_acc.size(3);
  TORCH_CHECK(lse_acc.size(0) == B && lse_acc.size(1) == H && lse_acc.size(2) == S && block_lse.sizes() == lse_acc.sizes());
  c10::cuda::CUDAGuard guard(out_acc.device());
  LCA_CUDA_OK(launch_merge_out_lse(out_acc.data_ptr<float>(), lse_acc.data_ptr<float>(), block_out.data_ptr(),
                                   dtype_code(block_out), block_lse.data_ptr<float>(), B, S, H, D,
                                   at::cuda::getCurrentCUDAStream()));
}

at::Tensor finalize_out(const at::Tensor& out_acc, at::ScalarType dtype) {
  TORCH_CHECK(out_acc.is_cuda() && out_acc.scalar_type() == at::kFloat && out_acc.is_contiguous());
  TORCH_CHECK(out_acc.numel() % 4 == 0);
  at::Tensor out = at::empty(out_acc.sizes(), out_acc.options().dtype(dtype));
  c10::cuda::CUDAGuard guard(out_acc.device());
  LCA_CUDA_OK(launch_finalize_out(out_acc.data_ptr<float>(), out.data_ptr(), dtype_code(out), out_acc.numel(),
                                  at::cuda::getCurrentCUDAStream()));
  return out;
}

at::Tensor flatten_varlen_lse(const at::Tensor& lse, const at::Tensor& cu_seqlens, int64_t total) {
  TORCH_CHECK(lse.is_cuda() && lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.dim() == 3);
  TORCH_CHECK(cu_seqlens.is_cuda() && cu_seqlens.scalar_type() == at::kInt && cu_seqlens.is_contiguous());
  const int B = lse.size(0), H = lse.size(1), M = lse.size(2);
  TORCH_CHECK(cu_seqlens.numel() == B + 1);
  at::Tensor flat = at::empty({H, total}, lse.options());
  c10::cuda::CUDAGuard guard(lse.device());
  LCA_CUDA_OK(launch_flatten_lse(lse.data_ptr<float>(), flat.data_ptr<float>(), cu_seqlens.data_ptr<int>(), B, H, M,
                                 static_cast<int>(total), at::cuda::getCurrentCUDAStream()));
  return flat;
}

at::Tensor unflatten_varlen_lse(const at::Tensor& flat, const at::Tensor& cu_seqlens, int64_t max_seqlen) {
  TORCH_CHECK(flat.is_cuda() && flat.scalar_type() == at::kFloat && flat.is_contiguous() && flat.dim() == 2);
  TORCH_CHECK(cu_seqlens.is_cuda() && cu_seqlens.scalar_type() == at::kInt && cu_seqlens.is_contiguous());
  const int H = flat.size(0), total = flat.size(1), B = cu_seqlens.numel() - 1;
  at::Tensor pad = at::empty({B, H, max_seqlen}, flat.options());
  c10::cuda::CUDAGuard guard(flat.device());
  LCA_CUDA_OK(launch_unflatten_lse(flat.data_ptr<float>(), pad.data_ptr<float>(), cu_seqlens.data_ptr<int>(), B, H,
                                   static_cast<int>(max_seqlen), total, at::cuda::getCurrentCUDAStream()));
  return pad;
}

// (B,S,G,X...) -> (G,B,S,X...) when to_group_major, else the inverse
at::Tensor permute_group(const at::Tensor& src, int64_t G, bool to_group_major) {
  TORCH_CHECK(src.is_cuda() && src.is_contiguous());
  c10::cuda::CUDAGuard guard(src.device());
  const int64_t esz = src.element_size();
  if (to_group_major) {
    TORCH_CHECK(src.dim() >= 4 && src.size(2) == G, "expected (B,S,G,...)");
    const int64_t B = src.size(0), S = src.size(1);
    const int64_t x = src.numel() / (B * S * G) * esz;
    std::vector<int64_t> shape = {G, B, S};
    for (int i = 3; i < src.dim(); ++i) shape.push_back(src.size(i));
    at::Tensor dst = at::empty(shape, src.options());
    LCA_CUDA_OK(launch_permute_heads_out(src.data_ptr(), dst.data_ptr(), B, S, G, static_cast<int>(x),
                                         at::cuda::getCurrentCUDAStream()));
    return dst;
  }
  TORCH_CHECK(src.dim() >= 4 && src.size(0) == G, "expected (G,B,S,...)");
  const int64_t B = src.size(1), S = src.size(2);
  const int64_t x = src.numel() / (B * S * G) * esz;
  std::vector<int64_t> shape = {B, S, G};
  for (int i = 3; i < src.dim(); ++i) shape.push_back(src.size(i));
  at::Tensor dst = at::empty(shape, src.options());
  LCA_CUDA_OK(launch_permute_heads_in(src.data_ptr(), dst.data_ptr(), B, S, G, static_cast<int>(x),
                                      at::cuda::getCurrentCUDAStream()));
  return dst;
}

// delta = rowsum(out o dout) (B,H,S); with `lse` also returns the log2-domain LSE the backward kernels read
std::vector<at::Tensor> attn_delta(const at::Tensor& out, const at::Tensor& dout, const c10::optional<at::Tensor>& lse) {
  TORCH_CHECK(out.is_cuda() && out.dim() == 4 && out.sizes() == dout.sizes() && out.scalar_type() == dout.scalar_type());
  TORCH_CHECK(out.stride(3) == 1 && dout.stride(3) == 1);
  const int B = out.size(0), S = out.size(1), H = out.size(2), D = out.size(3);
  at::Tensor delta = at::empty({B, H, S}, out.options().dtype(at::kFloat));
  at::Tensor lse2;
  const float* lse_p = nullptr;
  float* lse2_p = nullptr;
  if (lse.has_value() && lse->defined()) {
    TORCH_CHECK(lse->scalar_type() == at::kFloat && lse->is_contiguous() && lse->dim() == 3 && lse->size(0) == B &&
                lse->size(1) == H && lse->size(2) == S, "lse must be contiguous (B,H,S) fp32");
    lse2 = at::empty({B, H, S}, delta.options());
    lse_p = lse->data_ptr<float>();
    lse2_p = lse2.data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(out.device());
  LCA_CUDA_OK(launch_delta(out.data_ptr(), dout.data_ptr(), dtype_code(out), delta.data_ptr<float>(), lse_p, lse2_p, B, S, H,
                           D, out.stride(0), out.stride(1), out.stride(2), dout.stride(0), dout.stride(1), dout.stride(2),
                           at::cuda::getCurrentCUDAStream()));
  if (lse2.defined()) return {delta, lse2};
  return {delta};
}

}  // namespace lca

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "lca_b200 sm_100a kernels";
  m.def("fmha_fwd", &lca::fmha_fwd, "tcgen05 flash-attention forward (segments + global positions)");
  m.def("fmha_fwd_fp8", &lca::fmha_fwd_fp8, "EXPERIMENTAL: e4m3 block-scaled forward (tcgen05 kind::f8f6f4)");
  m.def("quantize_e4m3", &lca::quantize_e4m3, "bf16/fp16 -> e4m3 with per-128-row-block (or per-head) fp32 scales");
  m.def("usp_fwd", &lca::usp_fwd, "fused USP forward: NVLink push CTAs + tcgen05 attention CTAs in one kernel");
  m.def("usp_bwd_pass", &lca::usp_bwd_pass, "fused USP backward pass (push CTAs / peer scatter / NVLink red.add)");
  m.def("symm_wait", &lca::symm_wait, "device-side wait until a system-scope counter reaches a target");
  m.def("fmha_bwd_pass", &lca::fmha_bwd_pass, "tcgen05 flash-attention backward pass (dQ or dK/dV)");
  m.def("set_next_dropout", &lca::set_next_dropout, "EXPERIMENTAL: {p8, seed, head_offset} for the next fused launch");
  m.def("fmha_fwd_drop", &lca::fmha_fwd_drop, "EXPERIMENTAL: forward with coordinate-keyed dropout");
  m.def("fmha_bwd_pass_drop", &lca::fmha_bwd_pass_drop, "EXPERIMENTAL: backward pass with coordinate-keyed dropout");
  m.def("merge_out_lse", &lca::merge_out_lse, "in-place online-softmax merge");
  m.def("finalize_out", &lca::finalize_out, "fp32 accumulator -> 16-bit output");
  m.def("flatten_varlen_lse", &lca::flatten_varlen_lse);
  m.def("unflatten_varlen_lse", &lca::unflatten_varlen_lse);
  m.def("permute_group", &lca::permute_group);
  m.def("attn_delta", &lca::attn_delta);
  lca::bind_symm(m);
  m.attr("arch") = "sm_100a";
}
