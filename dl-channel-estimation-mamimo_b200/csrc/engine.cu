// C-ABI implementation (include/mamimo.h) of the channel-estimation engine.
// Host side only orchestrates: tables, BN folding + operand splitting of the weights,
// workspace, tensor maps, launches and the chunked host<->device pipeline.  All math on the
// hot path runs in the kernels of ls.cuh / fc.cuh.
#include <cuda_runtime.h>
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/mamimo.h"
#include "fc.cuh"
#include "lmmse.cuh"
#include "ls.cuh"
#include "ofdm.cuh"
#include "svd.cuh"
#include "omp.cuh"
#include "tables.h"

using namespace mm;

namespace {

struct HostLayer {
  int in = 0, out = 0;
  bool loaded = false, has_bn = false;
  std::vector<float> W, b, g, be, mu, var;   // W: Keras [in][out]
};

struct DevLayer {
  Operand w;               // [planes][Npad][Kpad]
  float* bias = nullptr;   // [N]
  int N = 0, K = 0;
  float w_scale = 1.f;
  float rowsum = 0.f, bmax = 0.f;   // max_n sum_k |W[k][n]| and max_n |b[n]| of the folded layer (FP16X3 range bounds)
  CUtensorMap tmap_b;       // box of 256 weight rows (1-CTA kernel)
  CUtensorMap tmap_b_half;  // box of 128 weight rows (CTA-pair kernel: each CTA stages half the tile)
  CUtensorMap tmap_b_q;     // box of 64 weight rows (few-row calls: 128 x 64 tiles)
  CUtensorMap tmap_a;      // A operand of this layer (net-specific buffer for layer 0)
};

thread_local std::string g_create_err;

int round_up(int x, int m) { return (x + m - 1) / m * m; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace

constexpr int kDynSlotsMax = 8;

struct mamimo_engine {
  mamimo_config cfg;
  std::string err;
  int num_sms = 0;
  int fc_sms = 0;               // SMs the persistent FC kernels may occupy
  int rows_per_pkt = 0;
  int max_pkts = 0;
  int host_chunk = 0;           // units per chunk of the host-buffer pipeline
  int kb_per_chunk = 4;
  bool fc_pair = true;          // CTA-pair (cta_group::2) FC kernel
  bool fc_small = true;         // few-row calls: 128 x 128 tiles on single CTAs instead (MAMIMO_FC_SMALL=0 disables)
  bool fc_tiny = true;          // and 128 x 64 tiles when even those leave most SMs idle (MAMIMO_FC_TINY=0 disables)
  bool ofdm_tma = true;         // persistent bulk-copy-fed OFDM kernel (MAMIMO_OFDM_TMA=0: plain three-pass kernel)
  bool ls_tma = true;           // TMA-fed persistent LS kernel (MAMIMO_LS_TMA=0: plain split kernel)
  int ls_tma_ctas = 4;          // resident CTAs per SM of that kernel (MAMIMO_LS_TMA_CTAS)
  int ls_tile = 64;             // tones per CTA of the split LS kernel (MAMIMO_LS_TILE=128: experiment)
  int ls_tile64 = 64;           // tones per tile of the TMA-fed LS kernel at 64 antennas.  MAMIMO_LS_TILE64=32 (16 KB stages,
                                // 6 CTAs/SM instead of 3) measured SLOWER: 4.72 vs 4.93 TB/s at 64x8x2048 (r2G) -- the 256-byte
                                // row segments cost more than the occupancy gives; kept as an experiment knob
  bool ls_split = true;         // LS: FWHT split over threads for 32/64 antennas (MAMIMO_LS_SPLIT=0 disables)
  unsigned long long* d_dbg = nullptr;   // MAMIMO_FC_DEBUG=1: role wait-cycle counters of the pair kernel
  int l2_prefetch = 0;             // measured slower (426 vs 442 TFLOP/s): kept as an experiment knob (MAMIMO_L2_PREFETCH)
  int rows_alloc = 0;           // plane stride (rows) of every activation operand
  int n_pil = 0;
  int n_layers = 0;             // n_hidden + 1 when an MLP is configured, else 0
  int elem_bytes = 4, planes = 1, block_k = 32;
  float act_scale = 1.f;
  // FP16X3 range management (schemes.cuh): device-resolved per-level scales; dyn_fixed = act_scale_log2 pinned
  DynState* d_dyn = nullptr;    // [kDynSlots]; slot = sub-batch in flight
  bool dyn_fixed = false;
  bool ls_verify = true;        // automatic scale: provisional + verify passes of the LS kernel (MAMIMO_LS_VERIFY=0: one pass
                                // after an exact amax read of Y)
  int ls_prof_cls = 0;          // profile class the LS launches are booked under (the verify pass counts as bookkeeping)
  int dyn_slot = 0;
  float ls_gain = 1.f;          // bound on |H_ls component| / amax |Y component|
  float tmax[2] = {0.f, 0.f};   // mode A: max |T| of the de-duplicated first layer
  bool hadamard = false;
  bool finalized = false;
  std::vector<float> hP;        // [n_tx][n_ltf] complex interleaved
  float2* dP = nullptr;
  float2* d_inv_den = nullptr;
  std::vector<double> hPd;      // the same tables in double (mamimo_set_pilots_f64; FP64 LS of the MATLAB-facing surface)
  double2* dPd = nullptr;
  double2* d_inv_den_d = nullptr;
  // OMP hybrid-precoder consumer (omp.cuh): steering dictionary and per-chunk workspace
  double2* d_omp_At = nullptr;      // [n_rays][n_tx]
  double2* d_omp_AtcT = nullptr;    // conj, transposed, zero padded: [n_tx][n_rays_pad]
  int omp_rays = 0, omp_rays_pad = 0;
  double2* d_omp_wres = nullptr;
  uint8_t* d_omp_active = nullptr;
  size_t omp_wres_bytes = 0, omp_active_bytes = 0;
  bool omp_generic = false;
  HostLayer hl[2][MAMIMO_MAX_HIDDEN + 1];
  DevLayer dl[2][MAMIMO_MAX_HIDDEN + 1];
  Operand act_in[2];            // layer-0 A operand per net
  Operand act_h[2][2];          // [net][ping-pong] hidden activations (per net, so the nets can overlap)
  // mode A with the de-duplicated first layer (see expand_pairs_kernel)
  bool dedup_a = false;
  float* d_z[2] = {nullptr, nullptr};        // [n_prx][h0] first-layer LTF term
  float* d_T[2] = {nullptr, nullptr};        // [n_tx][h0]  W1_p^T p_j + b1
  float* d_zero_bias = nullptr;              // [h0] zeros (the bias rides in T)
  std::vector<double> hW0p[2], hb0[2];       // host copies to rebuild T when P changes
  // fused all-gather (optional): gathered planes [world * gather_rows][d_out] float32 per rank
  int gather_world = 0, gather_rank = 0;
  int64_t gather_rows = 0;                 // rows per rank slot
  float* gather_local[2] = {nullptr, nullptr};                       // this rank's gathered planes
  bool gather_owned = true;                                          // false: attached (caller's memory, never freed here)
  float* gather_mc[2] = {nullptr, nullptr};                          // NVSwitch multicast addresses of the planes (or null)
  float* gather_peer[2][kMaxGatherRanks] = {};                       // every rank's planes as mapped here
  // OFDM front-end (optional)
  int fft_len = 0, cp_len = 0, sym_offset = 0, n_twiddle = 0;
  float2* d_twiddle = nullptr;
  float2* d_tw256 = nullptr;    // register-FFT kernels (256..4096): pass-2 twiddles [15][16]
  float2* d_tw3 = nullptr;      // 512..4096: pass-3 twiddles [fft_len/256 - 1][256]
  int* d_kmap = nullptr;        // [fft_len] FFT bin -> output column (-1 = dropped)
  int* d_bins = nullptr;
  float2* d_ydemod = nullptr;   // [max_pkts][n_rx][n_ltf][n_sc] scratch between demod and LS
  uint32_t* d_flags = nullptr;
  uint32_t* h_flags = nullptr;  // pinned
  // host-memory pipeline
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr, s_side = nullptr;
  cudaEvent_t ev_side[2] = {nullptr, nullptr};
  int gather_sms = 0;           // > 0: SMs given to the gathering final layers (side stream) while other work computes
  double gather_growth = 1.2;   // size ratio of consecutive sub-batches (see mamimo_estimate_stages)
  bool gather_ce = false;       // MAMIMO_GATHER_MODE=ce: the sub-batches' planes travel by copy engine (peer memcpy on per-peer
                                // streams) instead of TMA stores from the final-layer kernels; SMs only compute
  bool gather_push = false;     // MAMIMO_GATHER_MODE=push: like ce, but peer_push_kernel (push_ctas CTAs on the side stream)
  int push_ctas = 24;           //   moves the rows with bulk copies instead of the copy engines
  cudaStream_t s_copy[kMaxGatherRanks] = {};
  cudaEvent_t ev_copy[kMaxGatherRanks] = {};
  int gather_sub = 1;           // > 1: pipelined step -- the batch is cut into this many sub-batches and the gathering
                                // layers of sub-batch i run under LS + hidden layers of sub-batch i+1
  int cur_row_off = 0;          // first pair row of the sub-batch being enqueued (operand buffers + gather slot)
  int ls_sm_limit = 0;          // > 0: SMs the LS kernel may fill (the rest hold gathering CTAs)
  cudaEvent_t ev_sub[kDynSlotsMax] = {};
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  void* st_in[2] = {nullptr, nullptr};
  size_t st_in_bytes = 0;
  void* st_in2[2] = {nullptr, nullptr};   // second input plane (modes A/B)
  size_t st_in2_bytes = 0;
  float* st_hr[2] = {nullptr, nullptr};
  float* st_hi[2] = {nullptr, nullptr};
  size_t st_h_bytes = 0;
  void* st_hls[2] = {nullptr, nullptr};
  size_t st_hls_bytes = 0;
  // LMMSE smoother workspace (lazy)
  double2* lm_M = nullptr;       // [lm_slabs][R][n_pad]
  double2* lm_Dinv = nullptr;    // [lm_slabs][nb][32][32]
  double2* lm_par = nullptr;     // [lm_slabs] (c, 1/snr)
  void* lm_in = nullptr;         // host-buffer staging of H_ls / H_mmse chunks
  void* lm_out = nullptr;
  int lm_slabs = 0;
  bool lm_schur = true;          // Toeplitz (Schur) factorisation; MAMIMO_LMMSE_SCHUR=0: blocked dense Cholesky
  int lm_groups = 4;             // slab groups run on separate streams (MAMIMO_LMMSE_STREAMS; measured 1: 6.81, 2: 6.41, 4: 6.39 ms)
  cudaEvent_t lm_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t lm_io_bytes = 0;
  // CUDA-graph cache of the device-resident full path: one graph launch per batch (MAMIMO_GRAPH=0 disables)
  struct GraphEntry {
    const void* y; void* hls; float* hr; float* hi; int64_t n_pkt; int y_type; cudaStream_t st;
    uint32_t mode;                      // stage mask incl. MAMIMO_STAGE_GATHER: a gathering graph is not a plain one
    cudaGraphExec_t exec; uint64_t launches; uint64_t stamp;
  };
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;
  bool small_batch_overlap = true;   // MAMIMO_SMALL_OVERLAP=0: never run the two nets on two streams
  bool always_overlap = true;        // two streams at every batch size (MAMIMO_SMALL_OVERLAP=1: small batches only)
  uint64_t graph_stamp = 0, graph_replays = 0;
  mamimo_stats stats;
  // optional per-kernel-class device timing (mamimo_profile_begin/end)
  struct ProfRec { cudaEvent_t a, b; int cls; cudaStream_t st; };
  std::vector<ProfRec> prof;
  bool profiling = false;
};

namespace {

mamimo_status fail(mamimo_engine* e, mamimo_status s, const std::string& msg) {
  if (e) e->err = msg; else g_create_err = msg;
  return s;
}
mamimo_status fail_cuda(mamimo_engine* e, cudaError_t ce, const char* what) {
  return fail(e, MAMIMO_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
}
#define CK(e, call)                                         \
  do {                                                      \
    cudaError_t ce_ = (call);                               \
    if (ce_ != cudaSuccess) return fail_cuda(e, ce_, #call); \
  } while (0)

void invalidate_graphs(mamimo_engine* e) {      // any change of tables / weights / workspace pointers
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

enum { kClsLs = 0, kClsFc = 1, kClsStage = 2, kClsLmmse = 3 };   // the OFDM demod kernel is booked under kClsStage
// RAII event bracket around one launch (no-op unless profiling)
struct ProfScope {
  mamimo_engine* e; cudaStream_t st; int idx = -1;
  ProfScope(mamimo_engine* e_, cudaStream_t st_, int cls) : e(e_), st(st_) {
    if (!e->profiling) return;
    mamimo_engine::ProfRec r; r.cls = cls; r.st = st_;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    e->prof.push_back(r);
    idx = static_cast<int>(e->prof.size()) - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(e->prof[idx].b, st); }
};

bool is_sylvester(const std::vector<float>& P, int n_tx, int n_ltf) {
  if (n_tx != n_ltf || n_tx < 1 || n_tx > 64 || (n_tx & (n_tx - 1))) return false;
  for (int j = 0; j < n_tx; ++j)
    for (int n = 0; n < n_ltf; ++n) {
      const float want = (__builtin_popcount(j & n) & 1) ? -1.f : 1.f;
      if (P[2 * (j * n_ltf + n)] != want || P[2 * (j * n_ltf + n) + 1] != 0.f) return false;
    }
  return true;
}

mamimo_status alloc_operand(mamimo_engine* e, Operand& op, int planes, int rows_alloc, int kpad, int elem_bytes) {
  op.planes = planes; op.rows_alloc = rows_alloc; op.kpad = kpad; op.elem_bytes = elem_bytes;
  CK(e, cudaMalloc(&op.ptr, op.bytes()));
  CK(e, cudaMemset(op.ptr, 0, op.bytes()));
  return MAMIMO_OK;
}

mamimo_status make_map(mamimo_engine* e, CUtensorMap* map, const Operand& op, int kpad, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(e, MAMIMO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMapDataType dt;
  switch (e->cfg.precision) {
    case MAMIMO_PREC_TF32X3: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case MAMIMO_PREC_FP16X3: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; break;
    default: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
  }
  // the buffer may be used with a smaller row pitch than it was allocated with (hidden sizes differ)
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kpad), static_cast<cuuint64_t>(op.planes) * op.rows_alloc};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kpad) * op.elem_bytes};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(e->block_k), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, op.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(e, MAMIMO_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
  return MAMIMO_OK;
}

constexpr int kTcBN = 256;
constexpr int kTcSmallBN = 128;   // few-row calls (a handful of packets): 128 x 128 tiles on single CTAs
constexpr int kTcTinyBN = 64;     // one to four packets: 128 x 64 tiles
constexpr int kDynSlots = kDynSlotsMax;
static_assert(kMaxLevels >= MAMIMO_MAX_HIDDEN + 1, "DynState levels");

DynState* dyn_of(mamimo_engine* e) {
  return (e->cfg.precision == MAMIMO_PREC_FP16X3 && e->d_dyn) ? e->d_dyn + e->dyn_slot : nullptr;
}

// start of a call's range bookkeeping: zero the slot, then (auto mode) the exact amax of the input planes
mamimo_status dyn_begin(mamimo_engine* e, const void* in0, size_t n0, const void* in1, size_t n1, bool is_double,
                        cudaStream_t st, bool may_sample = false) {
  DynState* d = dyn_of(e);
  if (!d) return MAMIMO_OK;
  CK(e, cudaMemsetAsync(d, 0, sizeof(DynState), st));
  if (e->dyn_fixed) return MAMIMO_OK;
  const void* ptr[2] = {in0, in1};
  const size_t cnt[2] = {n0, n1};
  for (int i = 0; i < 2; ++i) {
    if (!ptr[i] || !cnt[i]) continue;
    // LS path (may_sample): large inputs are sampled 1 cache line in 8 -- the LS kernel's verify pass makes the result
    // exact again (ls_resolve_scale); small ones and the staging paths read every element
    const bool sample = may_sample && e->ls_verify && cnt[i] >= (static_cast<size_t>(1) << 22);
    const size_t vec = cnt[i] / (is_double ? 2 : 4) / (sample ? 8 : 1);
    const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>((vec + 255) / 256, static_cast<size_t>(e->num_sms) * 8)));
    ProfScope ps(e, st, kClsStage);
    if (is_double) {
      if (sample) amax_kernel<double, true><<<grid, 256, 0, st>>>(static_cast<const double*>(ptr[i]), cnt[i], &d->in_amax[i]);
      else amax_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(ptr[i]), cnt[i], &d->in_amax[i]);
    } else {
      if (sample) amax_kernel<float, true><<<grid, 256, 0, st>>>(static_cast<const float*>(ptr[i]), cnt[i], &d->in_amax[i]);
      else amax_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(ptr[i]), cnt[i], &d->in_amax[i]);
    }
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
  }
  return MAMIMO_OK;
}

template <int S>
mamimo_status set_tc_attr(mamimo_engine* e) {
  using Cfg = FcTcCfg<S, kTcBN>;
  const int smem = Cfg::kStages * Cfg::kStageBytes + Cfg::kAuxBytes + 1024;
  CK(e, cudaFuncSetAttribute(fc_tc_kernel<S, kTcBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  using CfgS = FcTcCfg<S, kTcSmallBN>;
  CK(e, cudaFuncSetAttribute(fc_tc_kernel<S, kTcSmallBN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             CfgS::kStages * CfgS::kStageBytes + CfgS::kAuxBytes + 1024));
  using CfgT = FcTcCfg<S, kTcTinyBN>;
  CK(e, cudaFuncSetAttribute(fc_tc_kernel<S, kTcTinyBN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             CfgT::kStages * CfgT::kStageBytes + CfgT::kAuxBytes + 1024));
  using Cfg2 = FcTc2Cfg<S>;
  const int smem2 = Cfg2::kStages * Cfg2::kStageBytes + Cfg2::kAuxBytes + 1024;
  CK(e, cudaFuncSetAttribute(fc_tc2_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  CK(e, cudaFuncSetAttribute(fc_tc2_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             FcTc2Cfg<S, true>::kSmemBytes));
  return MAMIMO_OK;
}

// ------------------------------------------------------------------ weight split + upload
template <int S>
mamimo_status upload_layer(mamimo_engine* e, DevLayer& d, const std::vector<double>& Wf /*[in][out]*/,
                           const std::vector<double>& bf, int in, int out) {
  using Sch = Scheme<S>;
  using E = typename Sch::elem;
  const int kpad = round_up(in, Sch::kBlockK);
  const int npad = round_up(out, 256);
  d.N = out; d.K = kpad;
  double wmax = 0;
  for (double v : Wf) wmax = std::max(wmax, std::fabs(v));
  {
    std::vector<double> colsum(out, 0.0);
    for (int k = 0; k < in; ++k)
      for (int n = 0; n < out; ++n) colsum[n] += std::fabs(Wf[static_cast<size_t>(k) * out + n]);
    double rs = 0, bm = 0;
    for (int n = 0; n < out; ++n) { rs = std::max(rs, colsum[n]); bm = std::max(bm, std::fabs(bf[n])); }
    d.rowsum = static_cast<float>(rs * (1.0 + 1e-6));
    d.bmax = static_cast<float>(bm * (1.0 + 1e-6));
  }
  d.w_scale = 1.f;
  if (S == kFp16x3 && wmax > 0) {
    int ex;
    std::frexp(wmax, &ex);                 // wmax = f * 2^ex, f in [0.5,1)
    d.w_scale = std::ldexp(1.0f, 13 - ex); // max |w| * scale in [2^12, 2^13)
  }
  std::vector<E> host(static_cast<size_t>(Sch::kPlanes) * npad * kpad);
  memset(host.data(), 0, host.size() * sizeof(E));
  bool ovf = false;
  for (int n = 0; n < out; ++n)
    for (int k = 0; k < in; ++k) {
      E p[Sch::kPlanes];
      Sch::split(static_cast<float>(Wf[static_cast<size_t>(k) * out + n]), d.w_scale, p, &ovf);
      for (int q = 0; q < Sch::kPlanes; ++q) host[(static_cast<size_t>(q) * npad + n) * kpad + k] = p[q];
    }
  if (ovf) return fail(e, MAMIMO_ERR_RANGE, "weight split overflow");
  d.w.planes = Sch::kPlanes; d.w.rows_alloc = npad; d.w.kpad = kpad; d.w.elem_bytes = sizeof(E);
  CK(e, cudaMalloc(&d.w.ptr, d.w.bytes()));
  CK(e, cudaMemcpy(d.w.ptr, host.data(), d.w.bytes(), cudaMemcpyHostToDevice));
  std::vector<float> b32(out);
  for (int n = 0; n < out; ++n) b32[n] = static_cast<float>(bf[n]);
  CK(e, cudaMalloc(&d.bias, out * sizeof(float)));
  CK(e, cudaMemcpy(d.bias, b32.data(), out * sizeof(float), cudaMemcpyHostToDevice));
  return MAMIMO_OK;
}

// ------------------------------------------------------------------ launches
template <int S, int NLTF, bool HAD>
mamimo_status launch_ls_t(mamimo_engine* e, const LsArgs& a, cudaStream_t st) {
  const int n_tiles = (a.n_pil + a.pil_per_tile - 1) / a.pil_per_tile;
  const long long grid = static_cast<long long>(a.n_pkt) * a.n_rx * n_tiles;
  const size_t smem = (static_cast<size_t>(a.n_tx) * (a.pil_per_tile + 4) + (HAD ? 0 : a.n_tx * a.n_ltf)) * sizeof(float2);
  if (smem > 48 * 1024)
    CK(e, cudaFuncSetAttribute(ls_kernel<S, NLTF, HAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)));
  {
    ProfScope ps(e, st, e->ls_prof_cls);
    ls_kernel<S, NLTF, HAD><<<static_cast<unsigned>(grid), 128, smem, st>>>(a);
  }
  CK(e, cudaGetLastError());
  e->stats.kernel_launches++;
  return MAMIMO_OK;
}

template <int S, int NLTF, int T>
mamimo_status launch_ls_split(mamimo_engine* e, const LsArgs& a, cudaStream_t st) {
  const int n_tiles = (a.n_pil + T - 1) / T;
  const long long grid = static_cast<long long>(a.n_pkt) * a.n_rx * n_tiles;
  const size_t smem = static_cast<size_t>(NLTF) * (T + 4) * sizeof(float2);
  {
    ProfScope ps(e, st, e->ls_prof_cls);
    ls_had_split_kernel<S, NLTF, T><<<static_cast<unsigned>(grid), T * (NLTF / 16), smem, st>>>(a);
  }
  CK(e, cudaGetLastError());
  e->stats.kernel_launches++;
  return MAMIMO_OK;
}

// TMA-fed persistent LS kernel: per-call tensor map over Y viewed as float32 [n_pkt*n_rx*n_ltf][2*n_sc]
template <int S, int NLTF, int STAGES, int NPS, int T = 64>
mamimo_status launch_ls_tma(mamimo_engine* e, const LsArgs& a, cudaStream_t st) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(e, MAMIMO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(2) * a.n_sc, static_cast<cuuint64_t>(a.n_pkt) * a.n_rx * a.n_ltf};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.n_sc) * 8};
  const cuuint32_t box[2] = {2u * ls_tma_row_tones<NPS, T>(), static_cast<cuuint32_t>(NLTF)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(a.Y), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(e, MAMIMO_ERR_CUDA, "cuTensorMapEncodeTiled (LS) failed: " + std::to_string(r));
  constexpr int smem = ls_tma_smem_bytes<NLTF, STAGES, NPS, T>();
  // per launch, not cached in a static: the attribute is per device and a process may hold engines on several
  CK(e, cudaFuncSetAttribute(ls_tma_kernel<S, NLTF, STAGES, NPS, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (a.n_sc + T - 1) / T;
  const long long total = static_cast<long long>(a.n_pkt) * a.n_rx * n_tiles;
  const int want_ctas = T == 32 ? std::max(e->ls_tma_ctas, 6) : e->ls_tma_ctas;
  const int per_sm = std::max(1, std::min(want_ctas, (227 * 1024) / (smem + 1024)));
  const int sms = e->ls_sm_limit > 0 ? std::min(e->ls_sm_limit, e->num_sms) : e->num_sms;
  const int grid = static_cast<int>(std::min<long long>(total, static_cast<long long>(sms) * per_sm));
  {
    ProfScope ps(e, st, e->ls_prof_cls);
    ls_tma_kernel<S, NLTF, STAGES, NPS, T><<<grid, T * (NLTF / 16), smem, st>>>(map, a);
  }
  CK(e, cudaGetLastError());
  e->stats.kernel_launches++;
  return MAMIMO_OK;
}

template <int S>
mamimo_status launch_ls(mamimo_engine* e, const LsArgs& a, cudaStream_t st) {
  if (e->hadamard && e->ls_split && e->ls_tma && !a.y_double && (a.n_sc % 2) == 0 && (a.n_ltf == 32 || a.n_ltf == 64)) {
    // persistent TMA-fed kernel; comb pilots (n_ps = 2, 4, 8) interpolate in its emit phase
    const bool comb_ok = (a.n_sc % 64) == 0 && (a.kpad & 3) == 0 && a.n_pil >= 2;
#define LS_TMA_CASE(NPS)                                                                        \
  return a.n_ltf == 32 ? launch_ls_tma<S, 32, 2, NPS>(e, a, st) : launch_ls_tma<S, 64, 2, NPS>(e, a, st);
    if (a.n_ps == 1 && a.n_ltf == 64 && e->ls_tile64 == 32) return launch_ls_tma<S, 64, 2, 1, 32>(e, a, st);
    if (a.n_ps == 1) { LS_TMA_CASE(1) }
    if (comb_ok && a.n_ps == 2) { LS_TMA_CASE(2) }
    if (comb_ok && a.n_ps == 4) { LS_TMA_CASE(4) }
    if (comb_ok && a.n_ps == 8) { LS_TMA_CASE(8) }
#undef LS_TMA_CASE
  }
  // every reference call site (n_ps = 1, Hadamard P, 32 or 64 antennas): transform split over threads
  if (e->hadamard && a.n_ps == 1 && e->ls_split) {
    if (a.n_ltf == 32) return e->ls_tile == 128 ? launch_ls_split<S, 32, 128>(e, a, st) : launch_ls_split<S, 32, 64>(e, a, st);
    if (a.n_ltf == 64) return launch_ls_split<S, 64, 64>(e, a, st);
  }
#define LS_CASE(n)                                                    \
  case n:                                                             \
    return e->hadamard ? launch_ls_t<S, n, true>(e, a, st) : launch_ls_t<S, n, false>(e, a, st);
  switch (a.n_ltf) {
    LS_CASE(1) LS_CASE(2) LS_CASE(4) LS_CASE(8) LS_CASE(16) LS_CASE(32) LS_CASE(64)
    default: return launch_ls_t<S, 0, false>(e, a, st);
  }
#undef LS_CASE
}

template <int S>
mamimo_status launch_fc(mamimo_engine* e, const DevLayer& d, FcArgs a, cudaStream_t st, const GatherMaps* gm = nullptr) {
  ProfScope ps(e, st, kClsFc);
  if constexpr (S == kFp32Simt) {
    const int grid = ((a.M + 127) / 128) * ((a.N + 127) / 128);
    fc_simt_kernel<S><<<grid, 256, 0, st>>>(a);
  } else {
    // Latency regime: a call of a few packets gives the CTA-pair kernel only ceil(N/256) tiles per layer (one packet:
    // 4 pairs, half of each pair on rows that do not exist) and every tile walks the whole K serially.  128 x 128
    // tiles on single CTAs double the number of SMs at work and halve the MMA time per tile; the accumulation order per
    // output element is the same, so the results are bit-identical to the pair kernel's.
    const int small_tiles = ((a.M + kFcBlockM - 1) / kFcBlockM) * ((a.N + kTcSmallBN - 1) / kTcSmallBN);
    const int tiny_tiles = ((a.M + kFcBlockM - 1) / kFcBlockM) * ((a.N + kTcTinyBN - 1) / kTcTinyBN);
    if (e->fc_pair && e->fc_small && e->fc_tiny && !gm && a.row_off == 0 && tiny_tiles * 2 <= e->fc_sms) {
      using CfgT = FcTcCfg<S, kTcTinyBN>;       // one to four packets: 128 x 64 tiles, twice the SMs again
      const int smem = CfgT::kStages * CfgT::kStageBytes + CfgT::kAuxBytes + 1024;
      fc_tc_kernel<S, kTcTinyBN><<<tiny_tiles, kFcThreads, smem, st>>>(d.tmap_a, d.tmap_b_q, a);
      CK(e, cudaGetLastError());
      e->stats.kernel_launches++;
      return MAMIMO_OK;
    }
    if (e->fc_pair && e->fc_small && !gm && a.row_off == 0 && small_tiles * 2 <= e->fc_sms) {
      using CfgS = FcTcCfg<S, kTcSmallBN>;
      const int smem = CfgS::kStages * CfgS::kStageBytes + CfgS::kAuxBytes + 1024;
      fc_tc_kernel<S, kTcSmallBN><<<small_tiles, kFcThreads, smem, st>>>(d.tmap_a, d.tmap_b_half, a);
      CK(e, cudaGetLastError());
      e->stats.kernel_launches++;
      return MAMIMO_OK;
    }
    if (e->fc_pair) {
      using Cfg2 = FcTc2Cfg<S>;
      const int pair_tiles = ((a.M + 2 * kFcBlockM - 1) / (2 * kFcBlockM)) * ((a.N + kTcBN - 1) / kTcBN);
      const int grid2 = 2 * std::min(pair_tiles, e->fc_sms / 2);
      const int smem2 = Cfg2::kStages * Cfg2::kStageBytes + Cfg2::kAuxBytes + 1024;
      if (gm) {
        fc_tc2_kernel<S, true><<<grid2, kFcThreads, FcTc2Cfg<S, true>::kSmemBytes, st>>>(d.tmap_a, d.tmap_b_half, *gm, a);
      } else {
        static const GatherMaps no_maps = {};
        fc_tc2_kernel<S, false><<<grid2, kFcThreads, smem2, st>>>(d.tmap_a, d.tmap_b_half, no_maps, a);
      }
      CK(e, cudaGetLastError());
      e->stats.kernel_launches++;
      return MAMIMO_OK;
    }
    using Cfg = FcTcCfg<S, kTcBN>;
    const int tiles = ((a.M + kFcBlockM - 1) / kFcBlockM) * ((a.N + kTcBN - 1) / kTcBN);
    const int grid = std::min(tiles, e->fc_sms);
    const int smem = Cfg::kStages * Cfg::kStageBytes + Cfg::kAuxBytes + 1024;
    fc_tc_kernel<S, kTcBN><<<grid, kFcThreads, smem, st>>>(d.tmap_a, d.tmap_b, a);
  }
  CK(e, cudaGetLastError());
  e->stats.kernel_launches++;
  return MAMIMO_OK;
}

// all FC layers of both nets for n_rows rows already staged in act_in[0/1]
mamimo_status make_gather_maps(mamimo_engine* e, int net, int n_rows, GatherMaps* gm);

template <int S>
mamimo_status run_mlp(mamimo_engine* e, int n_rows, float* out_r, float* out_i, cudaStream_t st,
                      unsigned net_mask = 3u, bool gather = false, int l_begin = 0, int l_end = -1) {
  if (l_end < 0) l_end = e->n_layers;
  for (int net = 0; net < 2; ++net) {
    if (!(net_mask & (1u << net))) continue;
    for (int l = l_begin; l < l_end; ++l) {
      const DevLayer& d = e->dl[net][l];
      const Operand& A = (l == 0) ? e->act_in[net] : e->act_h[net][(l - 1) & 1];
      const bool last = (l == e->n_layers - 1);
      FcArgs a;
      memset(&a, 0, sizeof(a));
      a.M = n_rows; a.N = d.N; a.num_k_blocks = d.K / e->block_k; a.kb_per_chunk = e->kb_per_chunk;
      a.a_plane_rows = A.rows_alloc; a.b_plane_rows = d.w.rows_alloc;
      a.bias = d.bias; a.alpha = 1.0f / (e->act_scale * d.w_scale); a.relu = last ? 0 : 1;
      a.flags = e->d_flags;
      a.l2_prefetch = e->l2_prefetch;
      a.dbg = e->d_dbg;
      a.dyn = dyn_of(e); a.net = net; a.level = l; a.fixed_scale = e->dyn_fixed ? 1 : 0;
      a.w_inv_scale = 1.0f / d.w_scale; a.rowsum = d.rowsum; a.bmax = d.bmax;
      a.row_off = e->cur_row_off;
      if (a.row_off && (S == kFp32Simt || !e->fc_pair || (a.row_off % (2 * kFcBlockM))))
        return fail(e, MAMIMO_ERR_UNSUPPORTED, "sub-batch row offsets need the CTA-pair kernel and 256-row alignment");
      a.A = reinterpret_cast<const float*>(A.ptr); a.W = reinterpret_cast<const float*>(d.w.ptr); a.kpad = d.K;
      if (last) {
        a.out_f32 = net == 0 ? out_r : out_i;
        a.out_ld = d.N;
      } else {
        const Operand& O = e->act_h[net][l & 1];
        a.out_planes = O.ptr; a.out_kpad = e->dl[net][l + 1].K; a.out_plane_rows = O.rows_alloc;
        a.out_scale = e->act_scale;
      }
      GatherMaps gm;
      const bool fused = gather && last;
      if (fused) {
        if (S == kFp32Simt || !e->fc_pair) return fail(e, MAMIMO_ERR_UNSUPPORTED, "fused all-gather needs the CTA-pair tensor-core kernel");
        mamimo_status gs = make_gather_maps(e, net, n_rows, &gm);
        if (gs != MAMIMO_OK) return gs;
        a.gather_world = e->gather_world;
      }
      mamimo_status s = launch_fc<S>(e, d, a, st, fused ? &gm : nullptr);
      if (s != MAMIMO_OK) return s;
    }
  }
  return MAMIMO_OK;
}

// Both nets of a batch: on two streams / graph branches when allowed (see the comment inside), else back to back.
template <int S>
mamimo_status run_mlp_both(mamimo_engine* e, int rows, float* hr, float* hi, cudaStream_t st, int l_begin = 0) {
  if (e->fc_pair && S != kFp32Simt && e->small_batch_overlap && !e->profiling) {
    // The two (independent) nets run on two streams / graph branches instead of back to back.  Small batches are
    // latency-bound (both nets' tiles fit the machine at once): one 32x4x1024 packet 152 -> 79 us device-resident,
    // 239 -> 164 us from host buffers.  Large batches: the other net's CTAs fill the half-empty last round of
    // every persistent layer kernel (13.5 rounds of tiles per layer at 500 packets): 251.3 -> 256.6 k packets/s.
    // Not while profiling: per-kernel event brackets on two interleaved streams would overlap in time.
    int max_n = e->cfg.d_out;
    for (int i = 0; i < e->cfg.n_hidden; ++i) max_n = std::max(max_n, e->cfg.hidden[i]);
    const int tiles = ((rows + 2 * kFcBlockM - 1) / (2 * kFcBlockM)) * ((max_n + kTcBN - 1) / kTcBN);
    if (tiles * 2 <= e->fc_sms / 2 || e->always_overlap) {
      CK(e, cudaEventRecord(e->ev_side[0], st));
      CK(e, cudaStreamWaitEvent(e->s_side, e->ev_side[0], 0));
      mamimo_status s = run_mlp<S>(e, rows, hr, hi, e->s_side, 2u, false, l_begin, e->n_layers);
      if (s == MAMIMO_OK) s = run_mlp<S>(e, rows, hr, hi, st, 1u, false, l_begin, e->n_layers);
      CK(e, cudaEventRecord(e->ev_side[1], e->s_side));
      CK(e, cudaStreamWaitEvent(st, e->ev_side[1], 0));
      return s;
    }
  }
  return run_mlp<S>(e, rows, hr, hi, st, 3u, false, l_begin, e->n_layers);
}

template <int S>
mamimo_status run_ls(mamimo_engine* e, const void* dY, int y_double, int n_pkt, void* dHls, int h_double,
                     bool want_planes, cudaStream_t st) {
  LsArgs a;
  memset(&a, 0, sizeof(a));
  a.Y = dY; a.P = e->dP; a.inv_den = e->d_inv_den; a.H_ls = dHls;
  if (want_planes) {
    a.planes[0] = e->act_in[0].ptr; a.planes[1] = e->act_in[1].ptr;
    a.plane_rows = e->act_in[0].rows_alloc; a.kpad = e->act_in[0].kpad;
  }
  a.scale = e->act_scale;
  a.n_pkt = n_pkt; a.n_rx = e->cfg.n_rx; a.n_tx = e->cfg.n_tx; a.n_ltf = e->cfg.n_ltf;
  a.n_sc = e->cfg.n_sc; a.n_ps = e->cfg.n_ps; a.n_pil = e->n_pil;
  a.pil_per_tile = std::max(1, 128 / e->cfg.n_ps);
  a.y_double = y_double; a.h_double = h_double; a.flags = e->d_flags;
  a.dyn = want_planes ? dyn_of(e) : nullptr; a.in_gain = e->ls_gain; a.fixed_scale = e->dyn_fixed ? 1 : 0;
  a.row_off = want_planes ? e->cur_row_off : 0;
  a.pass = 0;
  e->ls_prof_cls = kClsLs;
  mamimo_status s = launch_ls<S>(e, a, st);
  if (s != MAMIMO_OK || S != kFp16x3 || !a.dyn || e->dyn_fixed || !e->ls_verify) return s;
  a.pass = 1;                      // verify the provisional scale against the exact amax; recomputes only if it must
  a.H_ls = nullptr;                // H_ls itself never depends on the operand scale
  e->ls_prof_cls = kClsStage;
  s = launch_ls<S>(e, a, st);
  e->ls_prof_cls = kClsLs;
  return s;
}

template <int S>
mamimo_status run_stage_planes(mamimo_engine* e, const float* dXr, const float* dXi, int64_t rows, cudaStream_t st) {
  const int threads = 256;
  const int64_t total = rows * e->cfg.d_in;
  const int grid = static_cast<int>(std::min<int64_t>((total + threads - 1) / threads, e->num_sms * 16));
  for (int net = 0; net < 2; ++net) {
    const Operand& A = e->act_in[net];
    ProfScope ps(e, st, kClsStage);
    stage_planes_kernel<S><<<grid, threads, 0, st>>>(net == 0 ? dXr : dXi, A.ptr, rows, e->cfg.d_in, A.rows_alloc,
                                                     A.kpad, e->act_scale, e->d_flags, dyn_of(e), net,
                                                     e->dyn_fixed ? 1 : 0);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
  }
  return MAMIMO_OK;
}

template <int S>
mamimo_status run_stage_time(mamimo_engine* e, const float* dSr, const float* dSi, int64_t n_pkt, cudaStream_t st) {
  const int threads = 256;
  const int64_t n_prx = n_pkt * e->cfg.n_rx;
  const int64_t total = n_prx * e->cfg.n_tx * e->cfg.d_in;
  const int grid = static_cast<int>(std::min<int64_t>((total + threads - 1) / threads, e->num_sms * 16));
  for (int net = 0; net < 2; ++net) {
    const Operand& A = e->act_in[net];
    ProfScope ps(e, st, kClsStage);
    stage_time_p_kernel<S><<<grid, threads, 0, st>>>(net == 0 ? dSr : dSi, e->dP, A.ptr, n_prx, e->cfg.n_tx,
                                                     e->cfg.n_ltf, e->cfg.len_ltf, A.rows_alloc, A.kpad,
                                                     e->act_scale, e->d_flags, dyn_of(e), net, e->dyn_fixed ? 1 : 0);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
  }
  return MAMIMO_OK;
}

// Mode A with the de-duplicated first layer: T[j][n] = b1[n] + sum_m Re P[j][m] * W1[len_ltf + m][n]
// (the generator feeds P(:, iTx) of the pickled P = row j of the MATLAB-oriented P held here).
mamimo_status rebuild_mode_a_table(mamimo_engine* e) {
  if (!e->dedup_a || e->hW0p[0].empty()) return MAMIMO_OK;
  const int nt = e->cfg.n_tx, nl = e->cfg.n_ltf, h0 = e->cfg.hidden[0];
  if (e->hP.empty()) return fail(e, MAMIMO_ERR_STATE, "P not set (mamimo_set_pilots)");
  std::vector<float> T(static_cast<size_t>(nt) * h0);
  for (int net = 0; net < 2; ++net) {
    double tm = 0;
    for (int j = 0; j < nt; ++j)
      for (int n = 0; n < h0; ++n) {
        double acc = e->hb0[net][n];
        for (int m = 0; m < nt; ++m)
          acc += static_cast<double>(e->hP[2 * (j * nl + m)]) * e->hW0p[net][static_cast<size_t>(m) * h0 + n];
        T[static_cast<size_t>(j) * h0 + n] = static_cast<float>(acc);
        tm = std::max(tm, std::fabs(acc));
      }
    e->tmax[net] = static_cast<float>(tm * (1.0 + 1e-6));
    CK(e, cudaMemcpy(e->d_T[net], T.data(), T.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  return MAMIMO_OK;
}

// rows of this chunk: LTF planes per (pkt, rx) -> Z = X_ltf W1_ltf^T (one GEMM per net, n_pkt*n_rx rows) ->
// expand to pair rows with relu(Z + T) -> remaining layers on n_pkt*n_rx*n_tx rows
template <int S>
mamimo_status run_mode_a_dedup(mamimo_engine* e, const float* dSr, const float* dSi, int64_t n_pkt, float* hr,
                               float* hi, cudaStream_t st) {
  const int64_t n_prx = n_pkt * e->cfg.n_rx;
  const int h0 = e->cfg.hidden[0];
  const int threads = 256;
  for (int net = 0; net < 2; ++net) {
    const Operand& A = e->act_in[net];
    {
      const int64_t total = n_prx * e->cfg.len_ltf;
      const int grid = static_cast<int>(std::min<int64_t>((total + threads - 1) / threads, e->num_sms * 16));
      ProfScope ps(e, st, kClsStage);
      stage_planes_kernel<S><<<grid, threads, 0, st>>>(net == 0 ? dSr : dSi, A.ptr, n_prx, e->cfg.len_ltf, A.rows_alloc,
                                                       A.kpad, e->act_scale, e->d_flags, dyn_of(e), net,
                                                       e->dyn_fixed ? 1 : 0);
    }
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    const DevLayer& d = e->dl[net][0];
    FcArgs a;
    memset(&a, 0, sizeof(a));
    a.M = static_cast<int>(n_prx); a.N = d.N; a.num_k_blocks = d.K / e->block_k; a.kb_per_chunk = e->kb_per_chunk;
    a.a_plane_rows = A.rows_alloc; a.b_plane_rows = d.w.rows_alloc;
    a.bias = d.bias; a.alpha = 1.0f / (e->act_scale * d.w_scale); a.relu = 0;
    a.flags = e->d_flags; a.dbg = e->d_dbg;
    a.dyn = dyn_of(e); a.net = net; a.level = 0; a.fixed_scale = e->dyn_fixed ? 1 : 0;
    a.w_inv_scale = 1.0f / d.w_scale; a.rowsum = d.rowsum; a.bmax = d.bmax;
    a.A = reinterpret_cast<const float*>(A.ptr); a.W = reinterpret_cast<const float*>(d.w.ptr); a.kpad = d.K;
    a.out_f32 = e->d_z[net]; a.out_ld = h0;
    mamimo_status s = launch_fc<S>(e, d, a, st);
    if (s != MAMIMO_OK) return s;
    const Operand& O = e->act_h[net][0];
    {
      const int64_t total = n_prx * e->cfg.n_tx * (e->dl[net][1].K / 4);
      const int grid = static_cast<int>(std::min<int64_t>((total + threads - 1) / threads, e->num_sms * 16));
      ProfScope ps(e, st, kClsStage);
      expand_pairs_kernel<S><<<grid, threads, 0, st>>>(e->d_z[net], e->d_T[net], O.ptr, n_prx, e->cfg.n_tx, h0,
                                                       O.rows_alloc, e->dl[net][1].K, e->act_scale, e->d_flags,
                                                       dyn_of(e), net, e->dyn_fixed ? 1 : 0, d.rowsum, e->tmax[net]);
    }
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
  }
  return run_mlp_both<S>(e, static_cast<int>(n_prx) * e->cfg.n_tx, hr, hi, st, 1);
}

// One map per rank: this rank's row slot inside that rank's gathered plane, clipped to the rows of this call.
mamimo_status make_gather_maps(mamimo_engine* e, int net, int n_rows, GatherMaps* gm) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(e, MAMIMO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (e->gather_world < 1) return fail(e, MAMIMO_ERR_STATE, "fused all-gather not connected (mamimo_gather_connect)");
  if (e->cur_row_off + n_rows > e->gather_rows) return fail(e, MAMIMO_ERR_INVALID, "more rows than the gather slot holds");
  const int d_out = e->cfg.d_out;
  if (d_out % 4) return fail(e, MAMIMO_ERR_UNSUPPORTED, "fused all-gather needs d_out % 4 == 0 (TMA row pitch)");
  memset(gm, 0, sizeof(*gm));
  for (int p = 0; p < e->gather_world; ++p) {
    float* base = e->gather_peer[net][p] + (static_cast<size_t>(e->gather_rank) * e->gather_rows + e->cur_row_off) * d_out;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d_out), static_cast<cuuint64_t>(n_rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d_out) * sizeof(float)};
    const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(kFcBlockM)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&gm->m[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, MAMIMO_ERR_CUDA, "cuTensorMapEncodeTiled (gather) failed: " + std::to_string(r));
  }
  return MAMIMO_OK;
}

mamimo_status run_ofdm(mamimo_engine* e, const void* dx, int x_double, int64_t n_pkt, float2* dY, cudaStream_t st) {
  if (e->fft_len == 256 && e->d_tw256 && !x_double && e->ofdm_tma && (e->cp_len % 2) == 0 && (e->sym_offset % 2) == 0 &&
      getenv("MAMIMO_OFDM_GENERIC") == nullptr) {
    OfdmR16Args b;
    memset(&b, 0, sizeof(b));
    b.x = dx; b.Y = dY; b.tw2 = e->d_tw256; b.tw3 = nullptr; b.kmap = e->d_kmap;
    b.cp_len = e->cp_len; b.sym_offset = e->sym_offset; b.n_sc = e->cfg.n_sc;
    b.total_syms = n_pkt * e->cfg.n_rx * e->cfg.n_ltf; b.x_double = 0; b.flags = e->d_flags;
    const unsigned grid = static_cast<unsigned>(std::min<long long>((b.total_syms + 15) / 16, static_cast<long long>(e->num_sms) * 3));
    CK(e, cudaFuncSetAttribute(ofdm_r16_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ofdm_r16_tma_smem<8>()));
    {
      ProfScope ps(e, st, kClsStage);
      ofdm_r16_tma_kernel<8><<<grid, 256, ofdm_r16_tma_smem<8>(), st>>>(b);
    }
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return MAMIMO_OK;
  }
  if (e->fft_len == 256 && e->d_tw256 && getenv("MAMIMO_OFDM_GENERIC") == nullptr) {
    Ofdm256Args b;
    memset(&b, 0, sizeof(b));
    b.x = dx; b.Y = dY; b.tw2 = e->d_tw256; b.kmap = e->d_kmap;
    b.cp_len = e->cp_len; b.sym_offset = e->sym_offset; b.n_sc = e->cfg.n_sc;
    b.total_syms = n_pkt * e->cfg.n_rx * e->cfg.n_ltf; b.x_double = x_double;
    const long long grid = (b.total_syms + 7) / 8;
    {
      ProfScope ps(e, st, kClsStage);
      ofdm256_kernel<<<static_cast<unsigned>(grid), 128, 0, st>>>(b);
    }
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return MAMIMO_OK;
  }
  if (e->fft_len >= 512 && e->fft_len <= 4096 && e->d_tw256 && e->d_tw3 && getenv("MAMIMO_OFDM_GENERIC") == nullptr) {
    OfdmR16Args b;
    memset(&b, 0, sizeof(b));
    b.x = dx; b.Y = dY; b.tw2 = e->d_tw256; b.tw3 = e->d_tw3; b.kmap = e->d_kmap;
    b.cp_len = e->cp_len; b.sym_offset = e->sym_offset; b.n_sc = e->cfg.n_sc;
    b.total_syms = n_pkt * e->cfg.n_rx * e->cfg.n_ltf; b.x_double = x_double;
    const int syms = 4096 / e->fft_len;
    const unsigned grid = static_cast<unsigned>((b.total_syms + syms - 1) / syms);
    b.flags = e->d_flags;
    if (!x_double && e->ofdm_tma && (e->cp_len % 2) == 0 && (e->sym_offset % 2) == 0) {
      // persistent kernel, next tile's FFT windows prefetched by cp.async.bulk
      const unsigned pgrid = std::min<unsigned>(grid, static_cast<unsigned>(e->num_sms) * 3u);
      ProfScope ps(e, st, kClsStage);
#define OFDM_TMA_CASE(L)                                                                                              \
  {                                                                                                                   \
    CK(e, cudaFuncSetAttribute(ofdm_r16_tma_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, ofdm_r16_tma_smem<L>())); \
    ofdm_r16_tma_kernel<L><<<pgrid, 256, ofdm_r16_tma_smem<L>(), st>>>(b);                                            \
  }
      switch (e->fft_len) {
        case 512: OFDM_TMA_CASE(9) break;
        case 1024: OFDM_TMA_CASE(10) break;
        case 2048: OFDM_TMA_CASE(11) break;
        default: OFDM_TMA_CASE(12) break;
      }
#undef OFDM_TMA_CASE
      CK(e, cudaGetLastError());
      e->stats.kernel_launches++;
      return MAMIMO_OK;
    }
    {
      ProfScope ps(e, st, kClsStage);
      switch (e->fft_len) {
        case 512: ofdm_r16_kernel<9><<<grid, 256, 0, st>>>(b); break;
        case 1024: ofdm_r16_kernel<10><<<grid, 256, 0, st>>>(b); break;
        case 2048: ofdm_r16_kernel<11><<<grid, 256, 0, st>>>(b); break;
        default: ofdm_r16_kernel<12><<<grid, 256, 0, st>>>(b); break;
      }
    }
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return MAMIMO_OK;
  }
  OfdmArgs a;
  memset(&a, 0, sizeof(a));
  a.x = dx; a.Y = dY; a.twiddle = e->d_twiddle; a.n_twiddle = e->n_twiddle; a.bins = e->d_bins;
  a.fft_len = e->fft_len; a.cp_len = e->cp_len; a.sym_offset = e->sym_offset;
  a.n_sym = e->cfg.n_ltf; a.n_sc = e->cfg.n_sc; a.x_double = x_double;
  a.total_syms = n_pkt * e->cfg.n_rx * e->cfg.n_ltf;
  a.syms_per_cta = std::max(1, std::min(16, 1024 / e->fft_len));   // >= one radix-4 butterfly per thread per stage
  const long long grid = (a.total_syms + a.syms_per_cta - 1) / a.syms_per_cta;
  const int threads = 256;
  const size_t smem = (static_cast<size_t>(a.syms_per_cta) * 2 * e->fft_len + e->n_twiddle) * sizeof(float2);
  if (smem > 48 * 1024)
    CK(e, cudaFuncSetAttribute(ofdm_demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    ProfScope ps(e, st, kClsStage);
    ofdm_demod_kernel<<<static_cast<unsigned>(grid), threads, smem, st>>>(a);
  }
  CK(e, cudaGetLastError());
  e->stats.kernel_launches++;
  return MAMIMO_OK;
}

#define DISPATCH_S(e, expr)                                               \
  [&]() -> mamimo_status {                                                \
    switch ((e)->cfg.precision) {                                         \
      case MAMIMO_PREC_FP32_SIMT: { constexpr int S = kFp32Simt; return expr; } \
      case MAMIMO_PREC_TF32X3: { constexpr int S = kTf32x3; return expr; }      \
      case MAMIMO_PREC_FP16X3: { constexpr int S = kFp16x3; return expr; }      \
      case MAMIMO_PREC_BF16X1: { constexpr int S = kBf16x1; return expr; }      \
      default: return fail(e, MAMIMO_ERR_INVALID, "bad precision");      \
    }                                                                     \
  }()

mamimo_status check_flags(mamimo_engine* e, cudaStream_t st) {
  CK(e, cudaMemcpyAsync(e->h_flags, e->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CK(e, cudaStreamSynchronize(st));
  const uint32_t f = *e->h_flags;
  e->stats.last_device_flags = f;
  if (f) {
    CK(e, cudaMemsetAsync(e->d_flags, 0, sizeof(uint32_t), st));
    if (f & kFlagTimeout) return fail(e, MAMIMO_ERR_TIMEOUT, "device pipeline wait timed out (kernel aborted)");
    if (f & kFlagRange) return fail(e, MAMIMO_ERR_RANGE, "fp16 split operand overflow (non-finite input, or a pinned act_scale_log2 too high): use act_scale_log2 = 0 (auto) or TF32X3");
    if (f & kFlagUnderflow) return fail(e, MAMIMO_ERR_RANGE, "fp16 split operand below the accuracy window of the pinned act_scale_log2: use act_scale_log2 = 0 (auto) or TF32X3");
    if (f & kFlagNotPd) return fail(e, MAMIMO_ERR_RANGE, "LMMSE: Rpp is not positive definite in FP64 (SNR too high for this tau_rms)");
  }
  return MAMIMO_OK;
}

mamimo_status ensure_staging(mamimo_engine* e, size_t in_bytes, size_t in2_bytes, size_t h_bytes, size_t hls_bytes) {
  for (int i = 0; i < 2; ++i) {
    if (in_bytes > e->st_in_bytes) { if (e->st_in[i]) cudaFree(e->st_in[i]); CK(e, cudaMalloc(&e->st_in[i], in_bytes)); }
    if (in2_bytes > e->st_in2_bytes) { if (e->st_in2[i]) cudaFree(e->st_in2[i]); CK(e, cudaMalloc(&e->st_in2[i], in2_bytes)); }
    if (h_bytes > e->st_h_bytes) {
      if (e->st_hr[i]) cudaFree(e->st_hr[i]);
      if (e->st_hi[i]) cudaFree(e->st_hi[i]);
      CK(e, cudaMalloc(&e->st_hr[i], h_bytes));
      CK(e, cudaMalloc(&e->st_hi[i], h_bytes));
    }
    if (hls_bytes > e->st_hls_bytes) { if (e->st_hls[i]) cudaFree(e->st_hls[i]); CK(e, cudaMalloc(&e->st_hls[i], hls_bytes)); }
  }
  e->st_in_bytes = std::max(e->st_in_bytes, in_bytes);
  e->st_in2_bytes = std::max(e->st_in2_bytes, in2_bytes);
  e->st_h_bytes = std::max(e->st_h_bytes, h_bytes);
  e->st_hls_bytes = std::max(e->st_hls_bytes, hls_bytes);
  return MAMIMO_OK;
}

// Generic chunked runner.  `units` are packets (modes A/C) or rows (mode B).
//   stage(chunk_units, in0, in1, hls, hr, hi, stream): enqueue kernels for one device-resident chunk
template <class StageFn>
mamimo_status run_chunked(mamimo_engine* e, int64_t n_units, int64_t units_per_chunk, const void* in0,
                          size_t in0_unit_bytes, const void* in1, size_t in1_unit_bytes, void* hls,
                          size_t hls_unit_bytes, float* hr, float* hi, size_t h_unit_bytes, mamimo_mem mem,
                          cudaStream_t user_stream, StageFn stage) {
  if (n_units == 0) return MAMIMO_OK;
  if (mem == MAMIMO_MEM_DEVICE) {
    for (int64_t u0 = 0; u0 < n_units; u0 += units_per_chunk) {
      const int64_t n = std::min(units_per_chunk, n_units - u0);
      mamimo_status s = stage(n, static_cast<const char*>(in0) + u0 * in0_unit_bytes,
                              in1 ? static_cast<const char*>(in1) + u0 * in1_unit_bytes : nullptr,
                              hls ? static_cast<char*>(hls) + u0 * hls_unit_bytes : nullptr,
                              hr ? reinterpret_cast<float*>(reinterpret_cast<char*>(hr) + u0 * h_unit_bytes) : nullptr,
                              hi ? reinterpret_cast<float*>(reinterpret_cast<char*>(hi) + u0 * h_unit_bytes) : nullptr,
                              user_stream);
      if (s != MAMIMO_OK) return s;
    }
    return MAMIMO_OK;
  }
  // host buffers: double-buffered H2D -> compute -> D2H on three streams, in small chunks so the
  // three stages of neighbouring chunks overlap (PCIe is full duplex)
  units_per_chunk = std::min<int64_t>(units_per_chunk, e->host_chunk);
  const int64_t cu = std::min(units_per_chunk, n_units);
  mamimo_status s = ensure_staging(e, cu * in0_unit_bytes, in1 ? cu * in1_unit_bytes : 0, hr ? cu * h_unit_bytes : 0,
                                   hls ? cu * hls_unit_bytes : 0);
  if (s != MAMIMO_OK) return s;
  if (n_units <= units_per_chunk) {
    // one chunk (latency regime): nothing to overlap, so copy in, compute and copy out on ONE stream -- no event
    // hand-offs between streams
    cudaStream_t st = e->s_comp;
    CK(e, cudaMemcpyAsync(e->st_in[0], in0, n_units * in0_unit_bytes, cudaMemcpyHostToDevice, st));
    e->stats.h2d_bytes += n_units * in0_unit_bytes;
    if (in1) {
      CK(e, cudaMemcpyAsync(e->st_in2[0], in1, n_units * in1_unit_bytes, cudaMemcpyHostToDevice, st));
      e->stats.h2d_bytes += n_units * in1_unit_bytes;
    }
    s = stage(n_units, e->st_in[0], in1 ? e->st_in2[0] : nullptr, hls ? e->st_hls[0] : nullptr, hr ? e->st_hr[0] : nullptr,
              hi ? e->st_hi[0] : nullptr, st);
    if (s != MAMIMO_OK) return s;
    if (hls) {
      CK(e, cudaMemcpyAsync(hls, e->st_hls[0], n_units * hls_unit_bytes, cudaMemcpyDeviceToHost, st));
      e->stats.d2h_bytes += n_units * hls_unit_bytes;
    }
    if (hr) {
      CK(e, cudaMemcpyAsync(hr, e->st_hr[0], n_units * h_unit_bytes, cudaMemcpyDeviceToHost, st));
      e->stats.d2h_bytes += n_units * h_unit_bytes;
      if (hi) {
        CK(e, cudaMemcpyAsync(hi, e->st_hi[0], n_units * h_unit_bytes, cudaMemcpyDeviceToHost, st));
        e->stats.d2h_bytes += n_units * h_unit_bytes;
      }
    }
    return check_flags(e, st);          // copies the flag word behind the outputs and synchronises the stream
  }
  int64_t c = 0;
  for (int64_t u0 = 0; u0 < n_units; u0 += units_per_chunk, ++c) {
    const int b = static_cast<int>(c & 1);
    const int64_t n = std::min(units_per_chunk, n_units - u0);
    if (c >= 2) {
      CK(e, cudaStreamWaitEvent(e->s_h2d, e->ev_comp[b], 0));   // previous user of st_in[b] has consumed it
      CK(e, cudaStreamWaitEvent(e->s_comp, e->ev_out[b], 0));   // previous outputs in st_h*[b] are on the host
    }
    CK(e, cudaMemcpyAsync(e->st_in[b], static_cast<const char*>(in0) + u0 * in0_unit_bytes, n * in0_unit_bytes,
                          cudaMemcpyHostToDevice, e->s_h2d));
    e->stats.h2d_bytes += n * in0_unit_bytes;
    if (in1) {
      CK(e, cudaMemcpyAsync(e->st_in2[b], static_cast<const char*>(in1) + u0 * in1_unit_bytes, n * in1_unit_bytes,
                            cudaMemcpyHostToDevice, e->s_h2d));
      e->stats.h2d_bytes += n * in1_unit_bytes;
    }
    CK(e, cudaEventRecord(e->ev_in[b], e->s_h2d));
    CK(e, cudaStreamWaitEvent(e->s_comp, e->ev_in[b], 0));
    s = stage(n, e->st_in[b], in1 ? e->st_in2[b] : nullptr, hls ? e->st_hls[b] : nullptr, hr ? e->st_hr[b] : nullptr,
              hi ? e->st_hi[b] : nullptr, e->s_comp);
    if (s != MAMIMO_OK) return s;
    CK(e, cudaEventRecord(e->ev_comp[b], e->s_comp));
    CK(e, cudaStreamWaitEvent(e->s_d2h, e->ev_comp[b], 0));
    if (hls) {
      CK(e, cudaMemcpyAsync(static_cast<char*>(hls) + u0 * hls_unit_bytes, e->st_hls[b], n * hls_unit_bytes,
                            cudaMemcpyDeviceToHost, e->s_d2h));
      e->stats.d2h_bytes += n * hls_unit_bytes;
    }
    if (hr) {
      CK(e, cudaMemcpyAsync(reinterpret_cast<char*>(hr) + u0 * h_unit_bytes, e->st_hr[b], n * h_unit_bytes,
                            cudaMemcpyDeviceToHost, e->s_d2h));
      e->stats.d2h_bytes += n * h_unit_bytes;
      if (hi) {
        CK(e, cudaMemcpyAsync(reinterpret_cast<char*>(hi) + u0 * h_unit_bytes, e->st_hi[b], n * h_unit_bytes,
                              cudaMemcpyDeviceToHost, e->s_d2h));
        e->stats.d2h_bytes += n * h_unit_bytes;
      }
    }
    CK(e, cudaEventRecord(e->ev_out[b], e->s_d2h));
  }
  CK(e, cudaStreamSynchronize(e->s_d2h));
  return check_flags(e, e->s_comp);
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int32_t mamimo_abi_version(void) { return MAMIMO_ABI_VERSION; }

const char* mamimo_status_string(mamimo_status s) {
  switch (s) {
    case MAMIMO_OK: return "ok";
    case MAMIMO_ERR_INVALID: return "invalid argument";
    case MAMIMO_ERR_CUDA: return "CUDA error";
    case MAMIMO_ERR_NOMEM: return "out of memory";
    case MAMIMO_ERR_STATE: return "invalid engine state";
    case MAMIMO_ERR_UNSUPPORTED: return "unsupported device or configuration";
    case MAMIMO_ERR_RANGE: return "operand range overflow";
    case MAMIMO_ERR_TIMEOUT: return "device pipeline timeout";
  }
  return "unknown";
}

const char* mamimo_last_error(const mamimo_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

void mamimo_config_init(mamimo_config* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = MAMIMO_ABI_VERSION;
  c->n_ps = 1;
  c->precision = MAMIMO_PREC_FP16X3;   /* range-managed on the device (act_scale_log2 = 0): same 2e-6 as TF32X3 at twice the rate */
  c->input_mode = MAMIMO_INPUT_LS;
  c->act_scale_log2 = 0;   /* auto */
}

void mamimo_vht_ltf256(int8_t out[256]) {
  for (int i = 0; i < 256; ++i) out[i] = kVhtLtf256[i] == '+' ? 1 : (kVhtLtf256[i] == '-' ? -1 : 0);
}

int32_t mamimo_carriers_locations(int32_t* out, int32_t capacity) {
  int32_t n = 0;
  for (int i = 1; i <= kFftLen; ++i) {
    if (is_null_carrier(i) || is_pilot_carrier(i)) continue;
    if (out && n < capacity) out[n] = i;
    ++n;
  }
  return n;
}

mamimo_status mamimo_default_p(int32_t n, float* out) {
  if (n < 1 || (n & (n - 1)) || !out) return MAMIMO_ERR_INVALID;
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) out[j * n + k] = (__builtin_popcount(j & k) & 1) ? -1.f : 1.f;
  return MAMIMO_OK;
}

int64_t mamimo_pair_row(int64_t p, int32_t i_rx, int32_t i_tx, int32_t n_rx, int32_t n_tx) {
  return p * (static_cast<int64_t>(n_rx) * n_tx) + static_cast<int64_t>(i_rx) * n_tx + i_tx;
}

void* mamimo_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void mamimo_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

mamimo_status mamimo_create(const mamimo_config* cfg, mamimo_engine** out) {
  if (!cfg || !out) return fail(nullptr, MAMIMO_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != MAMIMO_ABI_VERSION) return fail(nullptr, MAMIMO_ERR_INVALID, "abi_version mismatch");
  if (cfg->n_tx < 1 || cfg->n_rx < 1 || cfg->n_ltf < 1 || cfg->n_sc < 1 || cfg->n_ps < 1)
    return fail(nullptr, MAMIMO_ERR_INVALID, "n_tx, n_rx, n_ltf, n_sc, n_ps must be >= 1");
  if (cfg->n_ltf > 64 || cfg->n_tx > 64) return fail(nullptr, MAMIMO_ERR_INVALID, "n_tx and n_ltf are limited to 64");
  if (cfg->n_hidden < 0 || cfg->n_hidden > MAMIMO_MAX_HIDDEN) return fail(nullptr, MAMIMO_ERR_INVALID, "n_hidden out of range");
  if (cfg->precision < 0 || cfg->precision > MAMIMO_PREC_BF16X1) return fail(nullptr, MAMIMO_ERR_INVALID, "bad precision");
  if (cfg->input_mode < 0 || cfg->input_mode > MAMIMO_INPUT_TIME_P) return fail(nullptr, MAMIMO_ERR_INVALID, "bad input_mode");
  const bool has_mlp = cfg->d_out > 0;
  if (has_mlp) {
    if (cfg->d_in < 1) return fail(nullptr, MAMIMO_ERR_INVALID, "d_in must be >= 1");
    for (int i = 0; i < cfg->n_hidden; ++i)
      if (cfg->hidden[i] < 1) return fail(nullptr, MAMIMO_ERR_INVALID, "hidden sizes must be >= 1");
    if (cfg->input_mode == MAMIMO_INPUT_LS && cfg->d_in != cfg->n_sc)
      return fail(nullptr, MAMIMO_ERR_INVALID, "mode C needs d_in == n_sc");
    if (cfg->input_mode == MAMIMO_INPUT_TIME_P && (cfg->len_ltf < 1 || cfg->d_in != cfg->len_ltf + cfg->n_tx || cfg->n_ltf != cfg->n_tx))
      return fail(nullptr, MAMIMO_ERR_INVALID, "mode A needs d_in == len_ltf + n_tx and n_ltf == n_tx");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, MAMIMO_ERR_UNSUPPORTED, "no CUDA device: this engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MAMIMO_ERR_INVALID, "bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, MAMIMO_ERR_CUDA, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, MAMIMO_ERR_UNSUPPORTED, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");

  mamimo_engine* e = new mamimo_engine();
  e->cfg = *cfg;
  memset(&e->stats, 0, sizeof(e->stats));
  e->num_sms = prop.multiProcessorCount;
  e->fc_sms = std::max(2, e->num_sms - std::max(0, cfg->fc_sm_reserve));
  e->rows_per_pkt = cfg->n_tx * cfg->n_rx;
  e->n_pil = (cfg->n_sc + cfg->n_ps - 1) / cfg->n_ps;
  e->n_layers = has_mlp ? cfg->n_hidden + 1 : 0;
  mamimo_status s = MAMIMO_OK;
  auto bail = [&](mamimo_status st) { g_create_err = e->err; mamimo_destroy(e); return st; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { e->err = "cudaSetDevice failed"; return bail(MAMIMO_ERR_CUDA); }

  switch (cfg->precision) {
    case MAMIMO_PREC_FP32_SIMT: e->elem_bytes = 4; e->planes = 1; e->block_k = Scheme<kFp32Simt>::kBlockK; break;
    case MAMIMO_PREC_TF32X3: e->elem_bytes = 4; e->planes = 2; e->block_k = Scheme<kTf32x3>::kBlockK; break;
    case MAMIMO_PREC_FP16X3: e->elem_bytes = 2; e->planes = 2; e->block_k = Scheme<kFp16x3>::kBlockK; break;
    default: e->elem_bytes = 2; e->planes = 1; e->block_k = Scheme<kBf16x1>::kBlockK; break;
  }
  e->dyn_fixed = cfg->precision == MAMIMO_PREC_FP16X3 && cfg->act_scale_log2 != 0;
  if (cfg->act_scale_log2 < -60 || cfg->act_scale_log2 > 60) { e->err = "act_scale_log2 out of range [-60, 60]"; return bail(MAMIMO_ERR_INVALID); }
  e->act_scale = e->dyn_fixed ? std::ldexp(1.0f, cfg->act_scale_log2) : 1.0f;

  // chunking: ~64K rows per chunk by default
  int rows_per_unit = cfg->input_mode == MAMIMO_INPUT_PLANES ? 1 : e->rows_per_pkt;
  e->max_pkts = cfg->max_pkts > 0 ? cfg->max_pkts : std::max(1, 65536 / rows_per_unit);
  // host pipeline granularity: ~2K rows per chunk so H2D, compute and D2H of neighbouring chunks overlap.  The D2H
  // stream trails the H2D stream by one chunk, so a call of n chunks runs at n / (n + 1.2) of the duplex PCIe rate:
  // measured at 500 packets of 32x4x1024 (r2s): 64-packet chunks 41.3 k pkt/s, 32: 42.1 k, 16: 43.4 k, 8: 41.1 k
  // (below ~1K rows the per-chunk launches stop hiding under the copies; shorter first / last chunks: no gain, 43.2 k)
  e->host_chunk = cfg->host_chunk_pkts > 0 ? std::min(cfg->host_chunk_pkts, e->max_pkts)
                                           : std::min(e->max_pkts, std::max(1, 2048 / rows_per_unit));
  if (const char* env = getenv("MAMIMO_HOST_CHUNK")) { if (atoi(env) > 0) e->host_chunk = std::min(atoi(env), e->max_pkts); }
  e->kb_per_chunk = cfg->kb_per_chunk > 0 ? cfg->kb_per_chunk : 4;
  if (const char* env = getenv("MAMIMO_KB_PER_CHUNK")) { if (atoi(env) > 0) e->kb_per_chunk = atoi(env); }
  e->fc_pair = cfg->fc_single_cta == 0;
  if (const char* env = getenv("MAMIMO_FC_PAIR")) e->fc_pair = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_FC_SMALL")) e->fc_small = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_FC_TINY")) e->fc_tiny = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_L2_PREFETCH")) e->l2_prefetch = atoi(env);
  if (const char* env = getenv("MAMIMO_LS_SPLIT")) e->ls_split = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_GRAPH")) e->use_graphs = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_SMALL_OVERLAP")) { e->small_batch_overlap = atoi(env) != 0; e->always_overlap = atoi(env) != 1; }
  if (const char* env = getenv("MAMIMO_LS_TMA")) e->ls_tma = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_LS_VERIFY")) e->ls_verify = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_LS_TILE64")) e->ls_tile64 = atoi(env) == 32 ? 32 : 64;
  if (const char* env = getenv("MAMIMO_OFDM_TMA")) e->ofdm_tma = atoi(env) != 0;
  if (const char* env = getenv("MAMIMO_LS_TMA_CTAS")) if (atoi(env) > 0) e->ls_tma_ctas = atoi(env);
  if (const char* env = getenv("MAMIMO_LS_TILE")) e->ls_tile = atoi(env) == 128 ? 128 : 64;
  if (const char* env = getenv("MAMIMO_FC_DEBUG")) {
    if (atoi(env) && cudaMalloc(&e->d_dbg, 8 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(e->d_dbg, 0, 8 * sizeof(unsigned long long));
  }
  const long long rows = static_cast<long long>(e->max_pkts) * rows_per_unit;
  if (rows > (1ll << 30)) { e->err = "max_pkts too large"; return bail(MAMIMO_ERR_INVALID); }
  e->rows_alloc = round_up(static_cast<int>(rows), 256);   // whole CTA-pair row tiles

  auto ck = [&](cudaError_t ce, const char* what) { if (ce != cudaSuccess && s == MAMIMO_OK) s = fail_cuda(e, ce, what); };
  ck(cudaMalloc(&e->d_flags, sizeof(uint32_t)), "cudaMalloc flags");
  if (s == MAMIMO_OK) ck(cudaMemset(e->d_flags, 0, sizeof(uint32_t)), "memset flags");
  if (cfg->precision == MAMIMO_PREC_FP16X3) {
    ck(cudaMalloc(&e->d_dyn, kDynSlots * sizeof(DynState)), "cudaMalloc dyn");
    if (s == MAMIMO_OK) ck(cudaMemset(e->d_dyn, 0, kDynSlots * sizeof(DynState)), "memset dyn");
  }
  ck(cudaMallocHost(&e->h_flags, sizeof(uint32_t)), "cudaMallocHost flags");
  ck(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking), "stream");
  ck(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking), "stream");
  ck(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking), "stream");
  ck(cudaStreamCreateWithFlags(&e->s_side, cudaStreamNonBlocking), "stream");
  for (int i = 0; i < 2; ++i) ck(cudaEventCreateWithFlags(&e->ev_side[i], cudaEventDisableTiming), "event");
  for (int i = 0; i < kDynSlotsMax; ++i) ck(cudaEventCreateWithFlags(&e->ev_sub[i], cudaEventDisableTiming), "event");
  for (int i = 0; i < 2; ++i) {
    ck(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming), "event");
    ck(cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming), "event");
    ck(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming), "event");
  }
  if (s != MAMIMO_OK) return bail(s);

  if (has_mlp) {
    const int kin = round_up(cfg->d_in, e->block_k);
    int kh = e->block_k;
    for (int i = 0; i < cfg->n_hidden; ++i) kh = std::max(kh, round_up(cfg->hidden[i], e->block_k));
    // mode A: the first layer runs once per (pkt, rx) on the LTF part only when there is a hidden layer to expand into
    e->dedup_a = cfg->input_mode == MAMIMO_INPUT_TIME_P && cfg->n_hidden >= 1 && (cfg->hidden[0] % 4) == 0 &&
                 getenv("MAMIMO_NO_DEDUP") == nullptr;
    const int in_rows = e->dedup_a ? round_up(e->max_pkts * cfg->n_rx, 256) : e->rows_alloc;
    const int in_k = e->dedup_a ? round_up(cfg->len_ltf, e->block_k) : kin;
    for (int net = 0; net < 2 && s == MAMIMO_OK; ++net) s = alloc_operand(e, e->act_in[net], e->planes, in_rows, in_k, e->elem_bytes);
    if (e->dedup_a && s == MAMIMO_OK) {
      const size_t h0 = cfg->hidden[0];
      for (int net = 0; net < 2; ++net) {
        ck(cudaMalloc(&e->d_z[net], static_cast<size_t>(in_rows) * h0 * sizeof(float)), "cudaMalloc z");
        ck(cudaMalloc(&e->d_T[net], static_cast<size_t>(cfg->n_tx) * h0 * sizeof(float)), "cudaMalloc T");
      }
      ck(cudaMalloc(&e->d_zero_bias, h0 * sizeof(float)), "cudaMalloc zero bias");
      if (s == MAMIMO_OK) ck(cudaMemset(e->d_zero_bias, 0, h0 * sizeof(float)), "memset zero bias");
    }
    if (cfg->n_hidden > 0)
      for (int i = 0; i < (cfg->n_hidden > 1 ? 2 : 1) && s == MAMIMO_OK; ++i)
        for (int net = 0; net < 2 && s == MAMIMO_OK; ++net)
          s = alloc_operand(e, e->act_h[net][i], e->planes, e->rows_alloc, kh, e->elem_bytes);
    if (s != MAMIMO_OK) return bail(s);
    int in = cfg->d_in;
    for (int l = 0; l <= cfg->n_hidden; ++l) {
      const int outw = l < cfg->n_hidden ? cfg->hidden[l] : cfg->d_out;
      for (int net = 0; net < 2; ++net) { e->hl[net][l].in = in; e->hl[net][l].out = outw; }
      in = outw;
    }
    if (cfg->precision == MAMIMO_PREC_TF32X3) s = set_tc_attr<kTf32x3>(e);
    else if (cfg->precision == MAMIMO_PREC_FP16X3) s = set_tc_attr<kFp16x3>(e);
    else if (cfg->precision == MAMIMO_PREC_BF16X1) s = set_tc_attr<kBf16x1>(e);
    if (s != MAMIMO_OK) return bail(s);
  }
  // default tables (all-ones pilots, Sylvester-Hadamard P) when a default P exists for this shape
  if ((cfg->n_tx == cfg->n_ltf && (cfg->n_tx & (cfg->n_tx - 1)) == 0) || cfg->n_ltf == 1) {
    s = mamimo_set_pilots(e, nullptr, nullptr);
    if (s != MAMIMO_OK) return bail(s);
  }
  *out = e;
  return MAMIMO_OK;
}

void mamimo_destroy(mamimo_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  invalidate_graphs(e);
  if (e->d_dbg) {          // diagnostic dump: average cycles per cluster over the engine's lifetime
    unsigned long long h[8] = {0};
    cudaMemcpy(h, e->d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = h[5] ? static_cast<double>(h[5]) : 1.0;
    fprintf(stderr, "[mamimo fc debug] per cluster-launch: producer wait-empty %.0f of %.0f cyc | MMA wait-tmem %.0f, wait-operands %.0f of %.0f cyc (%llu cluster-launches)\n",
            h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5]);
    cudaFree(e->d_dbg);
  }
  auto fr = [](void* p) { if (p) cudaFree(p); };
  if (e->gather_owned) { fr(e->gather_local[0]); fr(e->gather_local[1]); }
  fr(e->d_z[0]); fr(e->d_z[1]); fr(e->d_T[0]); fr(e->d_T[1]); fr(e->d_zero_bias);
  fr(e->dP); fr(e->d_inv_den); fr(e->dPd); fr(e->d_inv_den_d); fr(e->d_flags); fr(e->d_dyn); fr(e->d_twiddle); fr(e->d_tw256); fr(e->d_tw3); fr(e->d_kmap); fr(e->lm_M); fr(e->lm_Dinv); fr(e->lm_par); fr(e->lm_in); fr(e->lm_out); fr(e->d_bins); fr(e->d_ydemod);
  fr(e->d_omp_At); fr(e->d_omp_AtcT); fr(e->d_omp_wres); fr(e->d_omp_active);
  if (e->h_flags) cudaFreeHost(e->h_flags);
  for (int net = 0; net < 2; ++net) {
    fr(e->act_in[net].ptr);
    for (int l = 0; l <= MAMIMO_MAX_HIDDEN; ++l) { fr(e->dl[net][l].w.ptr); fr(e->dl[net][l].bias); }
  }
  for (int i = 0; i < 2; ++i) {
    fr(e->act_h[0][i].ptr); fr(e->act_h[1][i].ptr); fr(e->st_in[i]); fr(e->st_in2[i]); fr(e->st_hr[i]); fr(e->st_hi[i]); fr(e->st_hls[i]);
    if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
    if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
    if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
  }
  if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
  if (e->s_side) cudaStreamDestroy(e->s_side);
  for (int i = 0; i < 2; ++i) if (e->ev_side[i]) cudaEventDestroy(e->ev_side[i]);
  for (int i = 0; i < kDynSlotsMax; ++i) if (e->ev_sub[i]) cudaEventDestroy(e->ev_sub[i]);
  for (int p = 0; p < kMaxGatherRanks; ++p) {
    if (e->s_copy[p]) cudaStreamDestroy(e->s_copy[p]);
    if (e->ev_copy[p]) cudaEventDestroy(e->ev_copy[p]);
  }
  for (int i = 0; i < 4; ++i) if (e->lm_ev[i]) cudaEventDestroy(e->lm_ev[i]);
  delete e;
}

namespace {
// tables in double: x_pilot [n_pil] and P [n_tx][n_ltf], complex interleaved (either may be NULL = defaults)
mamimo_status set_pilots_impl(mamimo_engine* e, const double* x_pilot, const double* P) {
  invalidate_graphs(e);
  CK(e, cudaSetDevice(e->cfg.device));
  const int nt = e->cfg.n_tx, nl = e->cfg.n_ltf;
  e->hPd.assign(static_cast<size_t>(2) * nt * nl, 0.0);
  if (P) {
    memcpy(e->hPd.data(), P, e->hPd.size() * sizeof(double));
  } else {
    if (nt == nl && (nt & (nt - 1)) == 0) {
      for (int j = 0; j < nt; ++j)
        for (int n = 0; n < nl; ++n) e->hPd[2 * (j * nl + n)] = (__builtin_popcount(j & n) & 1) ? -1.0 : 1.0;
    } else if (nl == 1) {
      for (int j = 0; j < nt; ++j) e->hPd[2 * j] = 1.0;
    } else {
      return fail(e, MAMIMO_ERR_INVALID, "no default P for this (n_tx, n_ltf): pass P explicitly");
    }
  }
  e->hP.resize(e->hPd.size());
  for (size_t i = 0; i < e->hPd.size(); ++i) e->hP[i] = static_cast<float>(e->hPd[i]);
  e->hadamard = is_sylvester(e->hP, nt, nl);
  std::vector<float> inv(static_cast<size_t>(2) * e->n_pil);
  std::vector<double> invd(static_cast<size_t>(2) * e->n_pil);
  double inv_max = 0;
  for (int i = 0; i < e->n_pil; ++i) {
    const double xr = x_pilot ? x_pilot[2 * i] : 1.0, xi = x_pilot ? x_pilot[2 * i + 1] : 0.0;
    const double den = (xr * xr + xi * xi) * nl;
    if (den == 0.0) return fail(e, MAMIMO_ERR_INVALID, "pilot tone " + std::to_string(i) + " is zero");
    invd[2 * i] = xr / den;                              // 1/(nl*x) = conj(x)/(nl*|x|^2)
    invd[2 * i + 1] = -xi / den;
    inv[2 * i] = static_cast<float>(invd[2 * i]);
    inv[2 * i + 1] = static_cast<float>(invd[2 * i + 1]);
    inv_max = std::max(inv_max, std::sqrt(xr * xr + xi * xi) / den);
  }
  {
    // |H component| <= |H| <= sum_n |Y_n| |P[j][n]| |inv| <= sqrt(2) amax|Y comp| * max_j sum_n |P[j][n]| * max |inv|;
    // comb pilots: the last segment extrapolates by at most one pilot spacing (|h0 + w (h1 - h0)|, w < 2)
    double prow = 0;
    for (int j = 0; j < nt; ++j) {
      double r = 0;
      for (int n = 0; n < nl; ++n) r += std::hypot(e->hPd[2 * (j * nl + n)], e->hPd[2 * (j * nl + n) + 1]);
      prow = std::max(prow, r);
    }
    e->ls_gain = static_cast<float>(1.41421357 * prow * inv_max * (e->cfg.n_ps > 1 ? 5.0 : 1.0) * (1.0 + 1e-5));
  }
  if (!e->dP) CK(e, cudaMalloc(&e->dP, e->hP.size() * sizeof(float)));
  if (!e->d_inv_den) CK(e, cudaMalloc(&e->d_inv_den, inv.size() * sizeof(float)));
  if (!e->dPd) CK(e, cudaMalloc(&e->dPd, e->hPd.size() * sizeof(double)));
  if (!e->d_inv_den_d) CK(e, cudaMalloc(&e->d_inv_den_d, invd.size() * sizeof(double)));
  CK(e, cudaMemcpy(e->dP, e->hP.data(), e->hP.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(e, cudaMemcpy(e->d_inv_den, inv.data(), inv.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(e, cudaMemcpy(e->dPd, e->hPd.data(), e->hPd.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(e, cudaMemcpy(e->d_inv_den_d, invd.data(), invd.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (e->finalized) return rebuild_mode_a_table(e);       // mode A: the P-row term of the first layer depends on P
  return MAMIMO_OK;
}
}  // namespace

mamimo_status mamimo_set_pilots(mamimo_engine* e, const float* x_pilot, const float* P) {
  if (!e) return MAMIMO_ERR_INVALID;
  std::vector<double> xd, pd;
  if (x_pilot) xd.assign(x_pilot, x_pilot + static_cast<size_t>(2) * e->n_pil);
  if (P) pd.assign(P, P + static_cast<size_t>(2) * e->cfg.n_tx * e->cfg.n_ltf);
  return set_pilots_impl(e, x_pilot ? xd.data() : nullptr, P ? pd.data() : nullptr);
}

mamimo_status mamimo_set_pilots_f64(mamimo_engine* e, const double* x_pilot, const double* P) {
  if (!e) return MAMIMO_ERR_INVALID;
  return set_pilots_impl(e, x_pilot, P);
}

mamimo_status mamimo_load_layer(mamimo_engine* e, int32_t net, int32_t layer, const float* W, const float* b,
                                const float* g, const float* be, const float* mu, const float* var) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (e->n_layers == 0) return fail(e, MAMIMO_ERR_STATE, "engine was created without an MLP (d_out == 0)");
  if (net < 0 || net > 1 || layer < 0 || layer >= e->n_layers || !W || !b) return fail(e, MAMIMO_ERR_INVALID, "bad net/layer/pointer");
  const bool bn = g || be || mu || var;
  if (bn && !(g && be && mu && var)) return fail(e, MAMIMO_ERR_INVALID, "BN needs gamma, beta, mean and var");
  if (bn && layer == e->n_layers - 1) return fail(e, MAMIMO_ERR_INVALID, "the final linear layer has no BatchNormalization");
  HostLayer& L = e->hl[net][layer];
  L.W.assign(W, W + static_cast<size_t>(L.in) * L.out);
  L.b.assign(b, b + L.out);
  L.has_bn = bn;
  if (bn) { L.g.assign(g, g + L.out); L.be.assign(be, be + L.out); L.mu.assign(mu, mu + L.out); L.var.assign(var, var + L.out); }
  L.loaded = true;
  e->finalized = false;
  return MAMIMO_OK;
}

mamimo_status mamimo_finalize_weights(mamimo_engine* e) {
  if (!e) return MAMIMO_ERR_INVALID;
  invalidate_graphs(e);
  if (e->n_layers == 0) return fail(e, MAMIMO_ERR_STATE, "no MLP configured");
  CK(e, cudaSetDevice(e->cfg.device));
  for (int net = 0; net < 2; ++net)
    for (int l = 0; l < e->n_layers; ++l)
      if (!e->hl[net][l].loaded) return fail(e, MAMIMO_ERR_STATE, "layer " + std::to_string(l) + " of net " + std::to_string(net) + " not loaded");
  for (int net = 0; net < 2; ++net) {
    std::vector<double> scale, shift;   // BN of the previous layer (ReLU precedes BN: folds forward)
    for (int l = 0; l < e->n_layers; ++l) {
      const HostLayer& L = e->hl[net][l];
      std::vector<double> Wf(static_cast<size_t>(L.in) * L.out), bf(L.out);
      for (int n = 0; n < L.out; ++n) bf[n] = L.b[n];
      for (int k = 0; k < L.in; ++k)
        for (int n = 0; n < L.out; ++n) {
          const double w = L.W[static_cast<size_t>(k) * L.out + n];
          if (!scale.empty()) {
            bf[n] += shift[k] * w;
            Wf[static_cast<size_t>(k) * L.out + n] = scale[k] * w;
          } else {
            Wf[static_cast<size_t>(k) * L.out + n] = w;
          }
        }
      scale.clear(); shift.clear();
      if (L.has_bn) {
        scale.resize(L.out); shift.resize(L.out);
        for (int n = 0; n < L.out; ++n) {
          scale[n] = L.g[n] / std::sqrt(static_cast<double>(L.var[n]) + 1e-3);   // Keras BN epsilon
          shift[n] = L.be[n] - L.mu[n] * scale[n];
        }
      }
      DevLayer& d = e->dl[net][l];
      if (d.w.ptr) { cudaFree(d.w.ptr); d.w.ptr = nullptr; }
      if (d.bias) { cudaFree(d.bias); d.bias = nullptr; }
      int in_used = L.in;
      if (l == 0 && e->dedup_a) {
        // keep only the LTF rows of the first kernel in the GEMM; the P-row rows and the bias go into the table T
        in_used = e->cfg.len_ltf;
        e->hW0p[net].assign(Wf.begin() + static_cast<size_t>(in_used) * L.out, Wf.end());   // [n_tx][h0]
        e->hb0[net] = bf;
        Wf.resize(static_cast<size_t>(in_used) * L.out);
        std::fill(bf.begin(), bf.end(), 0.0);
      }
      mamimo_status s = DISPATCH_S(e, (upload_layer<S>(e, d, Wf, bf, in_used, L.out)));
      if (s != MAMIMO_OK) return s;
      if (e->cfg.precision != MAMIMO_PREC_FP32_SIMT) {
        s = make_map(e, &d.tmap_b, d.w, d.w.kpad, kTcBN);
        if (s != MAMIMO_OK) return s;
        s = make_map(e, &d.tmap_b_half, d.w, d.w.kpad, kTcBN / 2);
        if (s != MAMIMO_OK) return s;
        s = make_map(e, &d.tmap_b_q, d.w, d.w.kpad, kTcTinyBN);
        if (s != MAMIMO_OK) return s;
        const Operand& A = (l == 0) ? e->act_in[net] : e->act_h[net][(l - 1) & 1];
        s = make_map(e, &d.tmap_a, A, d.K, kFcBlockM);
        if (s != MAMIMO_OK) return s;
      }
    }
  }
  {
    mamimo_status ts = rebuild_mode_a_table(e);
    if (ts != MAMIMO_OK) return ts;
  }
  e->finalized = true;
  return MAMIMO_OK;
}

mamimo_status mamimo_ls_estimate(mamimo_engine* e, const void* Y, mamimo_ctype y_type, mamimo_mem y_mem,
                                 int64_t n_pkt, void* H_ls, mamimo_ctype h_type, mamimo_mem h_mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (n_pkt < 0 || (n_pkt > 0 && (!Y || !H_ls))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if (!e->dP) return fail(e, MAMIMO_ERR_STATE, "pilots / P not set (mamimo_set_pilots)");
  if (y_mem != h_mem) return fail(e, MAMIMO_ERR_INVALID, "Y and H_ls must live in the same memory kind");
  if ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(H_ls)) & 15)
    return fail(e, MAMIMO_ERR_INVALID, "Y and H_ls must be 16-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t yb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.n_ltf * e->cfg.n_sc * (y_type == MAMIMO_C128 ? 16 : 8);
  const size_t hb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.n_tx * e->cfg.n_sc * (h_type == MAMIMO_C128 ? 16 : 8);
  const bool f64 = y_type == MAMIMO_C128 && h_type == MAMIMO_C128 && e->dPd && getenv("MAMIMO_LS_F64_OFF") == nullptr;
  auto stage = [&](int64_t n, const void* in0, const void*, void* hls, float*, float*, cudaStream_t st) {
    if (f64) {      // complex double in and out: double arithmetic, as helperMIMOChannelEstimate.m:33-36 computes it
      LsF64Args a;
      a.Y = static_cast<const double2*>(in0); a.P = e->dPd; a.inv_den = e->d_inv_den_d; a.H = static_cast<double2*>(hls);
      a.n_rx = e->cfg.n_rx; a.n_tx = e->cfg.n_tx; a.n_ltf = e->cfg.n_ltf; a.n_sc = e->cfg.n_sc; a.n_ps = e->cfg.n_ps;
      a.n_pil = e->n_pil;
      const long long gy = static_cast<long long>(n) * a.n_rx * a.n_tx;
      for (long long y0 = 0; y0 < gy; y0 += 65535 / a.n_tx * a.n_tx) {      // grid.y limit; whole (pkt, rx) groups per launch
        const long long ny = std::min<long long>(65535 / a.n_tx * a.n_tx, gy - y0);
        LsF64Args b = a;
        b.Y = a.Y + (y0 / a.n_tx) * a.n_ltf * static_cast<size_t>(a.n_sc);
        b.H = a.H + y0 * static_cast<size_t>(a.n_sc);
        ProfScope ps(e, st, kClsLs);
        ls_f64_kernel<<<dim3((a.n_sc + 127) / 128, static_cast<unsigned>(ny)), 128, 0, st>>>(b);
        CK(e, cudaGetLastError());
        e->stats.kernel_launches++;
      }
      return MAMIMO_OK;
    }
    return DISPATCH_S(e, (run_ls<S>(e, in0, y_type == MAMIMO_C128, static_cast<int>(n), hls, h_type == MAMIMO_C128, false, st)));
  };
  return run_chunked(e, n_pkt, e->max_pkts, Y, yb, nullptr, 0, H_ls, hb, nullptr, nullptr, 0, y_mem,
                     static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_estimate_stages(mamimo_engine* e, const void* Y, mamimo_ctype y_type, int64_t n_pkt, void* H_ls,
                                     float* H_real, float* H_imag, mamimo_mem mem, void* stream, uint32_t stages) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (e->cfg.input_mode != MAMIMO_INPUT_LS) return fail(e, MAMIMO_ERR_STATE, "engine not configured for mode C (INPUT_LS)");
  if (!e->finalized) return fail(e, MAMIMO_ERR_STATE, "weights not finalised");
  if (!e->dP) return fail(e, MAMIMO_ERR_STATE, "pilots / P not set (mamimo_set_pilots)");
  const uint32_t all = MAMIMO_STAGE_LS | MAMIMO_STAGE_NET_REAL | MAMIMO_STAGE_NET_IMAG;
  const bool gather = (stages & MAMIMO_STAGE_GATHER) != 0;
  stages &= ~MAMIMO_STAGE_GATHER;
  if (stages == 0 || (stages & ~all)) return fail(e, MAMIMO_ERR_INVALID, "bad stage mask");
  if (gather && (mem != MAMIMO_MEM_DEVICE || n_pkt > e->max_pkts || e->gather_world < 1))
    return fail(e, MAMIMO_ERR_INVALID, "MAMIMO_STAGE_GATHER needs a connected gather, device buffers and one chunk");
  if (stages != all && (mem != MAMIMO_MEM_DEVICE || n_pkt > e->max_pkts))
    return fail(e, MAMIMO_ERR_INVALID, "partial stage masks need device buffers and n_pkt <= max_pkts (one chunk)");
  if (n_pkt < 0 || (n_pkt > 0 && (((stages & MAMIMO_STAGE_LS) && !Y) ||
                                  (!gather && (stages & MAMIMO_STAGE_NET_REAL) && !H_real) ||
                                  (!gather && (stages & MAMIMO_STAGE_NET_IMAG) && !H_imag))))
    return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(H_ls) | reinterpret_cast<uintptr_t>(H_real) |
       reinterpret_cast<uintptr_t>(H_imag)) & 15)
    return fail(e, MAMIMO_ERR_INVALID, "Y, H_ls, H_real and H_imag must be 16-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t yb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.n_ltf * e->cfg.n_sc * (y_type == MAMIMO_C128 ? 16 : 8);
  const size_t hlsb = static_cast<size_t>(e->rows_per_pkt) * e->cfg.n_sc * 8;
  const size_t hb = static_cast<size_t>(e->rows_per_pkt) * e->cfg.d_out * sizeof(float);
  // packets per 256-row pair tile: sub-batch boundaries must fall on whole tiles
  const int pkt_align = (2 * kFcBlockM) / std::gcd(2 * kFcBlockM, e->rows_per_pkt);
  const bool tc_pair = e->fc_pair && e->cfg.precision != MAMIMO_PREC_FP32_SIMT;
  const bool ce_mode = gather && stages == all && (e->gather_ce || e->gather_push) && tc_pair && e->n_layers >= 1;
  const bool push_mode = ce_mode && e->gather_push;
  const bool pipelined = ce_mode || (gather && stages == all && e->gather_sub > 1 && e->gather_sms > 0 && e->n_layers >= 2 &&
                                     tc_pair && n_pkt >= 2LL * pkt_align);
  auto stage = [&](int64_t n, const void* in0, const void*, void* hls, float* hr, float* hi, cudaStream_t st) {
    mamimo_status s = MAMIMO_OK;
    if (pipelined) {
      // Pipelined fused all-gather.  The batch is cut into sub-batches that live side by side in the operand buffers
      // (row offsets); LS + hidden layers of sub-batch i+1 run on `st` while the gathering final layers of sub-batch
      // i -- the kernels whose epilogue TMA-stores every tile into every rank's plane over NVLink -- run on the side
      // stream with gather_sms SMs, so the links are busy from the end of the first sub-batch to the end of the step.
      const int L = e->n_layers, full = e->fc_sms, g = e->gather_sms;
      const int64_t units = (n + pkt_align - 1) / pkt_align;
      const int S = static_cast<int>(std::min<int64_t>(std::min(e->gather_sub, kDynSlotsMax), units));
      // Sub-batch sizes grow geometrically (ratio gather_growth): the side stream moves a packet's planes at the NVLink
      // rate, the main stream computes the next sub-batch's LS + hidden layers faster than that, so sub-batch i+1 may
      // be larger than i by the ratio of the two rates without ever making the links wait -- and the first one, during
      // which the links idle, is the smallest.
      int64_t sizes[kDynSlotsMax];
      {
        const double r = e->gather_growth;
        double w = 1.0, wsum = 0.0;
        for (int i = 0; i < S; ++i) { wsum += w; w *= r; }
        int64_t left = units;
        w = 1.0;
        for (int i = 0; i < S; ++i) {
          int64_t u = i == S - 1 ? left : std::max<int64_t>(1, static_cast<int64_t>(units * w / wsum + 0.5));
          u = std::min(u, left - (S - 1 - i));            // keep one unit for every later sub-batch
          sizes[i] = std::max<int64_t>(u, 1);
          left -= sizes[i];
          w *= r;
        }
      }
      int64_t pkt0 = 0;
      for (int i = 0; i < S && s == MAMIMO_OK; ++i) {
        const int64_t np_i = std::min<int64_t>(sizes[i] * pkt_align, n - pkt0);
        if (np_i <= 0) break;
        const int rows_i = static_cast<int>(np_i) * e->rows_per_pkt;
        e->dyn_slot = i;
        e->cur_row_off = static_cast<int>(pkt0) * e->rows_per_pkt;
        const char* y_i = static_cast<const char*>(in0) + pkt0 * yb;
        void* hls_i = hls ? static_cast<char*>(hls) + pkt0 * hlsb : nullptr;
        float* hr_i = hr ? hr + static_cast<size_t>(e->cur_row_off) * e->cfg.d_out : nullptr;
        float* hi_i = hi ? hi + static_cast<size_t>(e->cur_row_off) * e->cfg.d_out : nullptr;
        if (ce_mode) {
          // copy-engine variant: every layer of both nets on `st` with all SMs, the final layers writing straight into
          // this rank's own slot of its gathered planes; then one peer memcpy per (plane, remote rank) on that rank's
          // copy stream carries the sub-batch's rows over NVLink while the SMs start the next sub-batch
          const size_t slot_off = (static_cast<size_t>(e->gather_rank) * e->gather_rows + e->cur_row_off) * e->cfg.d_out;
          float* own_r = e->gather_local[0] + slot_off;
          float* own_i = e->gather_local[1] + slot_off;
          if (push_mode) {                                    // the push kernel of sub-batch i-1 holds push_ctas SMs
            e->fc_sms = i == 0 ? full : std::max(2, (full - e->push_ctas) & ~1);
            e->ls_sm_limit = i == 0 ? 0 : e->num_sms - e->push_ctas;
          }
          s = dyn_begin(e, y_i, static_cast<size_t>(np_i) * e->cfg.n_rx * e->cfg.n_ltf * e->cfg.n_sc * 2, nullptr, 0,
                        y_type == MAMIMO_C128, st, true);
          if (s == MAMIMO_OK) s = DISPATCH_S(e, (run_ls<S>(e, y_i, y_type == MAMIMO_C128, static_cast<int>(np_i), hls_i, 0, true, st)));
          if (s == MAMIMO_OK) s = DISPATCH_S(e, (run_mlp<S>(e, rows_i, own_r, own_i, st, 3u, false, 0, L)));
          if (s != MAMIMO_OK) break;
          CK(e, cudaEventRecord(e->ev_sub[i], st));
          const size_t bytes = static_cast<size_t>(rows_i) * e->cfg.d_out * sizeof(float);
          if (push_mode) {
            CK(e, cudaStreamWaitEvent(e->s_side, e->ev_sub[i], 0));
            for (int d = 0; d < 2 && e->gather_mc[0]; ++d) {       // multicast: one store stream, the switch replicates
              PushMcArgs ma{};
              ma.src = reinterpret_cast<const float4*>(d ? own_i : own_r);
              ma.mc = reinterpret_cast<float4*>(e->gather_mc[d] + slot_off);
              ma.n_vec = bytes / 16;
              ProfScope ps(e, e->s_side, kClsStage);
              peer_push_mc_kernel<<<e->push_ctas, kPushMcThreads, 0, e->s_side>>>(ma);
              CK(e, cudaGetLastError());
            }
            for (int d = 0; d < 2 && !e->gather_mc[0]; ++d) {
              PushArgs pa{};
              pa.src = reinterpret_cast<const uint8_t*>(d ? own_i : own_r);
              for (int p = 0; p < e->gather_world; ++p)
                if (p != e->gather_rank) pa.dst[pa.n_dst++] = reinterpret_cast<uint8_t*>(e->gather_peer[d][p] + slot_off);
              pa.bytes = bytes;
              pa.flags = e->d_flags;
              if (pa.n_dst) {
                ProfScope ps(e, e->s_side, kClsStage);
                peer_push_kernel<<<e->push_ctas, 32, kPushSmem, e->s_side>>>(pa);
                CK(e, cudaGetLastError());
              }
            }
          } else
          for (int p = 0; p < e->gather_world; ++p) {
            if (p == e->gather_rank) continue;
            CK(e, cudaStreamWaitEvent(e->s_copy[p], e->ev_sub[i], 0));
            CK(e, cudaMemcpyAsync(e->gather_peer[0][p] + slot_off, own_r, bytes, cudaMemcpyDeviceToDevice, e->s_copy[p]));
            CK(e, cudaMemcpyAsync(e->gather_peer[1][p] + slot_off, own_i, bytes, cudaMemcpyDeviceToDevice, e->s_copy[p]));
          }
          if (hr_i) CK(e, cudaMemcpyAsync(hr_i, own_r, bytes, cudaMemcpyDeviceToDevice, st));
          if (hi_i) CK(e, cudaMemcpyAsync(hi_i, own_i, bytes, cudaMemcpyDeviceToDevice, st));
          pkt0 += np_i;
          continue;
        }
        e->fc_sms = i == 0 ? full : full - g;                 // nothing runs on the side stream during the first one
        e->ls_sm_limit = i == 0 ? 0 : e->num_sms - g;
        s = dyn_begin(e, y_i, static_cast<size_t>(np_i) * e->cfg.n_rx * e->cfg.n_ltf * e->cfg.n_sc * 2, nullptr, 0,
                      y_type == MAMIMO_C128, st, true);
        if (s == MAMIMO_OK) s = DISPATCH_S(e, (run_ls<S>(e, y_i, y_type == MAMIMO_C128, static_cast<int>(np_i), hls_i, 0, true, st)));
        if (s == MAMIMO_OK) s = DISPATCH_S(e, (run_mlp<S>(e, rows_i, hr_i, hi_i, st, 3u, false, 0, L - 1)));
        if (s != MAMIMO_OK) break;
        CK(e, cudaEventRecord(e->ev_sub[i], st));
        CK(e, cudaStreamWaitEvent(e->s_side, e->ev_sub[i], 0));
        e->fc_sms = (i == S - 1 || pkt0 + np_i >= n) ? full : g;   // the last gather has the machine to itself
        s = DISPATCH_S(e, (run_mlp<S>(e, rows_i, hr_i, hi_i, e->s_side, 3u, true, L - 1, L)));
        pkt0 += np_i;
      }
      e->fc_sms = full; e->ls_sm_limit = 0; e->dyn_slot = 0; e->cur_row_off = 0;
      if (s != MAMIMO_OK) return s;
      if (push_mode) {
        CK(e, cudaEventRecord(e->ev_side[1], e->s_side));
        CK(e, cudaStreamWaitEvent(st, e->ev_side[1], 0));
        return MAMIMO_OK;
      }
      if (ce_mode) {
        for (int p = 0; p < e->gather_world; ++p) {
          if (p == e->gather_rank) continue;
          CK(e, cudaEventRecord(e->ev_copy[p], e->s_copy[p]));
          CK(e, cudaStreamWaitEvent(st, e->ev_copy[p], 0));
        }
        return MAMIMO_OK;
      }
      CK(e, cudaEventRecord(e->ev_side[1], e->s_side));
      CK(e, cudaStreamWaitEvent(st, e->ev_side[1], 0));
      return MAMIMO_OK;
    }
    if (stages & MAMIMO_STAGE_LS) {
      s = dyn_begin(e, in0, static_cast<size_t>(n) * e->cfg.n_rx * e->cfg.n_ltf * e->cfg.n_sc * 2, nullptr, 0,
                    y_type == MAMIMO_C128, st, true);
      if (s != MAMIMO_OK) return s;
      s = DISPATCH_S(e, (run_ls<S>(e, in0, y_type == MAMIMO_C128, static_cast<int>(n), hls, 0, true, st)));
    }
    if (s != MAMIMO_OK) return s;
    const unsigned nets = (stages >> 1) & 3u;
    if (!nets) return s;
    const int rows = static_cast<int>(n) * e->rows_per_pkt;
    const int L = e->n_layers;
    if (gather && nets == 3u && e->gather_sms > 0 && L >= 2) {
      // NVLink-bound regime (world >= 4): the real net's gathering final layer runs on a side stream with a
      // few SMs (it only has to keep NVLink busy) while the imaginary net's hidden layers use the rest.
      const int full = e->fc_sms;
      s = DISPATCH_S(e, (run_mlp<S>(e, rows, hr, hi, st, 1u, false, 0, L - 1)));
      if (s != MAMIMO_OK) return s;
      CK(e, cudaEventRecord(e->ev_side[0], st));
      CK(e, cudaStreamWaitEvent(e->s_side, e->ev_side[0], 0));
      e->fc_sms = e->gather_sms;
      s = DISPATCH_S(e, (run_mlp<S>(e, rows, hr, hi, e->s_side, 1u, true, L - 1, L)));
      e->fc_sms = full - e->gather_sms;
      if (s == MAMIMO_OK) s = DISPATCH_S(e, (run_mlp<S>(e, rows, hr, hi, st, 2u, false, 0, L - 1)));
      e->fc_sms = full;
      if (s != MAMIMO_OK) return s;
      CK(e, cudaEventRecord(e->ev_side[1], e->s_side));
      CK(e, cudaStreamWaitEvent(st, e->ev_side[1], 0));
      return DISPATCH_S(e, (run_mlp<S>(e, rows, hr, hi, st, 2u, true, L - 1, L)));
    }
    if (!gather && nets == 3u) return DISPATCH_S(e, (run_mlp_both<S>(e, rows, hr, hi, st)));
    return DISPATCH_S(e, (run_mlp<S>(e, rows, hr, hi, st, nets, gather)));
  };
  // Device-resident full path on a capturable stream: the 7 launches of a batch are captured once per
  // (buffers, batch size, stream) and replayed as ONE graph launch afterwards.
  cudaStream_t ust = static_cast<cudaStream_t>(stream);
  const uint32_t graph_mode = stages | (gather ? MAMIMO_STAGE_GATHER : 0u);
  if (mem == MAMIMO_MEM_DEVICE && stages == all && (!gather || pipelined) && e->use_graphs && !e->profiling && ust != nullptr &&
      ust != cudaStreamLegacy && ust != cudaStreamPerThread && n_pkt > 0) {
    cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ust, &cst) == cudaSuccess && cst == cudaStreamCaptureStatusNone) {
      for (auto& g : e->graphs)
        if (g.y == Y && g.hls == H_ls && g.hr == H_real && g.hi == H_imag && g.n_pkt == n_pkt &&
            g.y_type == static_cast<int>(y_type) && g.st == ust && g.mode == graph_mode) {
          g.stamp = ++e->graph_stamp;
          CK(e, cudaGraphLaunch(g.exec, ust));
          e->stats.kernel_launches += g.launches;
          e->graph_replays++;
          e->stats.graph_launches++;
          return MAMIMO_OK;
        }
      const uint64_t l0 = e->stats.kernel_launches;
      CK(e, cudaStreamBeginCapture(ust, cudaStreamCaptureModeRelaxed));
      const mamimo_status cs = run_chunked(e, n_pkt, e->max_pkts, Y, yb, nullptr, 0, H_ls, hlsb, H_real, H_imag, hb, mem, ust, stage);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(ust, &graph);
      if (cs != MAMIMO_OK) { if (graph) cudaGraphDestroy(graph); return cs; }
      if (ce != cudaSuccess) return fail_cuda(e, ce, "cudaStreamEndCapture");
      cudaGraphExec_t exec = nullptr;
      const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ci != cudaSuccess) return fail_cuda(e, ci, "cudaGraphInstantiate");
      if (e->graphs.size() >= 8) {                       // evict the least recently used entry
        size_t lru = 0;
        for (size_t i = 1; i < e->graphs.size(); ++i) if (e->graphs[i].stamp < e->graphs[lru].stamp) lru = i;
        cudaGraphExecDestroy(e->graphs[lru].exec);
        e->graphs.erase(e->graphs.begin() + lru);
      }
      e->graphs.push_back({Y, H_ls, H_real, H_imag, n_pkt, static_cast<int>(y_type), ust, graph_mode, exec,
                           e->stats.kernel_launches - l0, ++e->graph_stamp});
      CK(e, cudaGraphLaunch(exec, ust));
      e->stats.graph_launches++;
      return MAMIMO_OK;
    }
  }
  return run_chunked(e, n_pkt, e->max_pkts, Y, yb, nullptr, 0, H_ls, hlsb, H_real, H_imag, hb, mem,
                     static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_estimate(mamimo_engine* e, const void* Y, mamimo_ctype y_type, int64_t n_pkt, void* H_ls,
                              float* H_real, float* H_imag, mamimo_mem mem, void* stream) {
  if (n_pkt > 0 && (!Y || !H_real || !H_imag)) return e ? fail(e, MAMIMO_ERR_INVALID, "null buffer") : MAMIMO_ERR_INVALID;
  return mamimo_estimate_stages(e, Y, y_type, n_pkt, H_ls, H_real, H_imag, mem, stream,
                                MAMIMO_STAGE_LS | MAMIMO_STAGE_NET_REAL | MAMIMO_STAGE_NET_IMAG);
}

mamimo_status mamimo_predict_planes(mamimo_engine* e, const float* X_real, const float* X_imag, int64_t n_rows,
                                    float* Y_real, float* Y_imag, mamimo_mem mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (e->cfg.input_mode != MAMIMO_INPUT_PLANES) return fail(e, MAMIMO_ERR_STATE, "engine not configured for mode B (INPUT_PLANES)");
  if (!e->finalized) return fail(e, MAMIMO_ERR_STATE, "weights not finalised");
  if (n_rows < 0 || (n_rows > 0 && (!X_real || !X_imag || !Y_real || !Y_imag))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t xb = static_cast<size_t>(e->cfg.d_in) * sizeof(float);
  const size_t hb = static_cast<size_t>(e->cfg.d_out) * sizeof(float);
  auto stage = [&](int64_t n, const void* in0, const void* in1, void*, float* hr, float* hi, cudaStream_t st) {
    const size_t cnt = static_cast<size_t>(n) * e->cfg.d_in;
    mamimo_status s = dyn_begin(e, in0, cnt, in1, cnt, false, st);
    if (s != MAMIMO_OK) return s;
    s = DISPATCH_S(e, (run_stage_planes<S>(e, static_cast<const float*>(in0), static_cast<const float*>(in1), n, st)));
    if (s != MAMIMO_OK) return s;
    return DISPATCH_S(e, (run_mlp_both<S>(e, static_cast<int>(n), hr, hi, st)));
  };
  return run_chunked(e, n_rows, e->max_pkts, X_real, xb, X_imag, xb, nullptr, 0, Y_real, Y_imag, hb, mem,
                     static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_predict_time(mamimo_engine* e, const float* sig_real, const float* sig_imag, int64_t n_pkt,
                                  float* Y_real, float* Y_imag, mamimo_mem mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (e->cfg.input_mode != MAMIMO_INPUT_TIME_P) return fail(e, MAMIMO_ERR_STATE, "engine not configured for mode A (INPUT_TIME_P)");
  if (!e->finalized) return fail(e, MAMIMO_ERR_STATE, "weights not finalised");
  if (n_pkt < 0 || (n_pkt > 0 && (!sig_real || !sig_imag || !Y_real || !Y_imag))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t xb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.len_ltf * sizeof(float);
  const size_t hb = static_cast<size_t>(e->rows_per_pkt) * e->cfg.d_out * sizeof(float);
  auto stage = [&](int64_t n, const void* in0, const void* in1, void*, float* hr, float* hi, cudaStream_t st) {
    const size_t cnt = static_cast<size_t>(n) * e->cfg.n_rx * e->cfg.len_ltf;
    mamimo_status s = dyn_begin(e, in0, cnt, in1, cnt, false, st);
    if (s != MAMIMO_OK) return s;
    if (e->dedup_a)
      return DISPATCH_S(e, (run_mode_a_dedup<S>(e, static_cast<const float*>(in0), static_cast<const float*>(in1), n, hr, hi, st)));
    if (DynState* d = dyn_of(e); d && !e->dyn_fixed) {      // the P-row columns of the input belong to its range too
      for (int net = 0; net < 2; ++net)
        amax_kernel<float><<<1, 256, 0, st>>>(reinterpret_cast<const float*>(e->dP),
                                              static_cast<size_t>(2) * e->cfg.n_tx * e->cfg.n_ltf, &d->in_amax[net]);
      CK(e, cudaGetLastError());
      e->stats.kernel_launches += 2;
    }
    s = DISPATCH_S(e, (run_stage_time<S>(e, static_cast<const float*>(in0), static_cast<const float*>(in1), n, st)));
    if (s != MAMIMO_OK) return s;
    return DISPATCH_S(e, (run_mlp_both<S>(e, static_cast<int>(n) * e->rows_per_pkt, hr, hi, st)));
  };
  return run_chunked(e, n_pkt, e->max_pkts, sig_real, xb, sig_imag, xb, nullptr, 0, Y_real, Y_imag, hb, mem,
                     static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_ipc_export(const void* dev_ptr, uint8_t handle[MAMIMO_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == MAMIMO_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (!dev_ptr || !handle || cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)) != cudaSuccess) return MAMIMO_ERR_CUDA;
  memcpy(handle, &h, sizeof(h));
  return MAMIMO_OK;
}

mamimo_status mamimo_ipc_open(const uint8_t handle[MAMIMO_IPC_HANDLE_BYTES], void** dev_ptr) {
  cudaIpcMemHandle_t h;
  if (!handle || !dev_ptr) return MAMIMO_ERR_INVALID;
  memcpy(&h, handle, sizeof(h));
  return cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? MAMIMO_OK : MAMIMO_ERR_CUDA;
}

mamimo_status mamimo_ipc_close(void* dev_ptr) {
  return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? MAMIMO_OK : MAMIMO_ERR_CUDA;
}

mamimo_status mamimo_gather_create(mamimo_engine* e, int32_t world, int32_t rank, int64_t pkts_per_rank,
                                   float** real_plane, float** imag_plane) {
  if (!e || !real_plane || !imag_plane) return MAMIMO_ERR_INVALID;
  if (e->n_layers == 0) return fail(e, MAMIMO_ERR_STATE, "no MLP configured");
  if (world < 1 || world > kMaxGatherRanks || rank < 0 || rank >= world || pkts_per_rank < 1 || pkts_per_rank > e->max_pkts)
    return fail(e, MAMIMO_ERR_INVALID, "need 1 <= world <= 8, 0 <= rank < world, 1 <= pkts_per_rank <= max_pkts");
  CK(e, cudaSetDevice(e->cfg.device));
  e->gather_world = 0;
  for (int n = 0; n < 2; ++n) {
    if (e->gather_local[n] && e->gather_owned) cudaFree(e->gather_local[n]);
    e->gather_local[n] = nullptr; e->gather_mc[n] = nullptr;
  }
  e->gather_owned = true;
  e->gather_rank = rank;
  e->gather_rows = pkts_per_rank * e->rows_per_pkt;
  const size_t bytes = static_cast<size_t>(world) * e->gather_rows * e->cfg.d_out * sizeof(float);
  for (int n = 0; n < 2; ++n) {
    CK(e, cudaMalloc(&e->gather_local[n], bytes));
    CK(e, cudaMemset(e->gather_local[n], 0, bytes));
  }
  e->gather_world = -world;                 // allocated, not yet connected
  *real_plane = e->gather_local[0];
  *imag_plane = e->gather_local[1];
  return MAMIMO_OK;
}

mamimo_status mamimo_gather_attach(mamimo_engine* e, int32_t world, int32_t rank, int64_t pkts_per_rank,
                                   void* const* real_planes, void* const* imag_planes, void* mc_real, void* mc_imag) {
  if (!e || !real_planes || !imag_planes) return MAMIMO_ERR_INVALID;
  if (e->n_layers == 0) return fail(e, MAMIMO_ERR_STATE, "no MLP configured");
  if (world < 1 || world > kMaxGatherRanks || rank < 0 || rank >= world || pkts_per_rank < 1 || pkts_per_rank > e->max_pkts)
    return fail(e, MAMIMO_ERR_INVALID, "need 1 <= world <= 8, 0 <= rank < world, 1 <= pkts_per_rank <= max_pkts");
  if ((mc_real == nullptr) != (mc_imag == nullptr)) return fail(e, MAMIMO_ERR_INVALID, "give both multicast addresses or neither");
  for (int p = 0; p < world; ++p)
    if (!real_planes[p] || !imag_planes[p] || ((reinterpret_cast<uintptr_t>(real_planes[p]) | reinterpret_cast<uintptr_t>(imag_planes[p])) & 127))
      return fail(e, MAMIMO_ERR_INVALID, "peer planes must be non-null and 128-byte aligned");
  if ((reinterpret_cast<uintptr_t>(mc_real) | reinterpret_cast<uintptr_t>(mc_imag)) & 127)
    return fail(e, MAMIMO_ERR_INVALID, "multicast addresses must be 128-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  for (int n = 0; n < 2; ++n) { if (e->gather_local[n] && e->gather_owned) cudaFree(e->gather_local[n]); }
  e->gather_owned = false;
  e->gather_local[0] = static_cast<float*>(real_planes[rank]);
  e->gather_local[1] = static_cast<float*>(imag_planes[rank]);
  e->gather_mc[0] = static_cast<float*>(mc_real);
  e->gather_mc[1] = static_cast<float*>(mc_imag);
  e->gather_rank = rank;
  e->gather_rows = pkts_per_rank * e->rows_per_pkt;
  e->gather_world = -world;
  invalidate_graphs(e);
  mamimo_status s = mamimo_gather_connect(e, real_planes, imag_planes);
  // with multicast addresses the push schedule is the default: the links carry every row once
  if (s == MAMIMO_OK && mc_real && getenv("MAMIMO_GATHER_MODE") == nullptr) {
    e->gather_push = true;
    if (getenv("MAMIMO_GATHER_SUB") == nullptr) e->gather_sub = 6;
  }
  return s;
}

mamimo_status mamimo_gather_connect(mamimo_engine* e, void* const* real_planes, void* const* imag_planes) {
  if (!e || !real_planes || !imag_planes) return MAMIMO_ERR_INVALID;
  if (e->gather_world >= 0 && !e->gather_local[0]) return fail(e, MAMIMO_ERR_STATE, "call mamimo_gather_create first");
  const int world = e->gather_world < 0 ? -e->gather_world : e->gather_world;
  for (int p = 0; p < world; ++p) {
    if (!real_planes[p] || !imag_planes[p]) return fail(e, MAMIMO_ERR_INVALID, "null peer plane");
    e->gather_peer[0][p] = static_cast<float*>(real_planes[p]);
    e->gather_peer[1][p] = static_cast<float*>(imag_planes[p]);
  }
  if (e->gather_peer[0][e->gather_rank] != e->gather_local[0] || e->gather_peer[1][e->gather_rank] != e->gather_local[1])
    return fail(e, MAMIMO_ERR_INVALID, "entry [rank] must be this engine's own planes");
  e->gather_world = world;
  // world >= 4: a plane's NVLink time exceeds a layer's compute time -> overlap it with the other net (see
  // mamimo_estimate_stages).  The gathering layer only needs enough SMs to keep NVLink busy.
  e->gather_sms = world >= 3 ? 56 : 0;     // measured: 4 GPUs 590 k pkt/s (512 k without), 8 GPUs 623 k (585 k with 36 SMs)
  if (const char* env = getenv("MAMIMO_GATHER_SMS")) e->gather_sms = atoi(env) & ~1;
  if (e->gather_sms >= e->fc_sms - 2) e->gather_sms = 0;
  // pipelined step (sub-batches): NVLink busy under the next sub-batch's LS + hidden layers
  if (const char* env = getenv("MAMIMO_GATHER_MODE")) {
    e->gather_ce = strcmp(env, "ce") == 0;
    e->gather_push = strcmp(env, "push") == 0;
  }
  if (const char* env = getenv("MAMIMO_PUSH_CTAS")) e->push_ctas = std::max(1, std::min(atoi(env), e->num_sms - 4));
  if (e->gather_push)
    CK(e, cudaFuncSetAttribute(peer_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushSmem));
  if (e->gather_ce)
    for (int p = 0; p < world; ++p) {
      if (p == e->gather_rank || e->s_copy[p]) continue;
      CK(e, cudaStreamCreateWithFlags(&e->s_copy[p], cudaStreamNonBlocking));
      CK(e, cudaEventCreateWithFlags(&e->ev_copy[p], cudaEventDisableTiming));
    }
  e->gather_sub = world >= 3 ? 6 : 1;
  e->gather_growth = world >= 5 ? 1.8 : 1.35;     // r2w/r2x sweeps (4 GPUs): 6 sub-batches x 1.35 = 3.17 ms, one shot 3.27 ms
  if (const char* env = getenv("MAMIMO_GATHER_SUB")) e->gather_sub = std::max(1, std::min(atoi(env), kDynSlotsMax));
  if (const char* env = getenv("MAMIMO_GATHER_GROWTH")) e->gather_growth = std::max(1.0, std::min(atof(env), 4.0));
  return MAMIMO_OK;
}

mamimo_status mamimo_set_ofdm(mamimo_engine* e, int32_t fft_len, int32_t cp_len, int32_t sym_offset,
                              const int32_t* carriers) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (fft_len < 2 || fft_len > 4096 || (fft_len & (fft_len - 1))) return fail(e, MAMIMO_ERR_INVALID, "fft_len must be a power of two in [2, 4096]");
  if (cp_len < 0 || sym_offset < 0 || sym_offset > cp_len) return fail(e, MAMIMO_ERR_INVALID, "need 0 <= sym_offset <= cp_len");
  if (cp_len > fft_len) return fail(e, MAMIMO_ERR_INVALID, "cyclic prefix longer than the FFT (the window start would fall before the symbol)");
  if (!carriers) return fail(e, MAMIMO_ERR_INVALID, "carriers is NULL");
  if (e->cfg.n_sc > fft_len) return fail(e, MAMIMO_ERR_INVALID, "n_sc exceeds fft_len");
  CK(e, cudaSetDevice(e->cfg.device));
  std::vector<int> bins(e->cfg.n_sc);
  for (int k = 0; k < e->cfg.n_sc; ++k) {
    if (carriers[k] < 1 || carriers[k] > fft_len) return fail(e, MAMIMO_ERR_INVALID, "carrier index out of range");
    bins[k] = (carriers[k] - 1 + fft_len / 2) % fft_len;      // undo fftshift: shifted index -> natural FFT bin
  }
  // per-stage compact twiddle tables (FP64 -> FP32): radix-4 stage ns: w_r[k] = exp(-2 pi i r k / (4 ns)), r = 1..3,
  // k < ns; then the radix-2 tail stage (log2 fft odd): w[k] = exp(-2 pi i k / (2 ns))
  std::vector<float> tw;
  const double two_pi = 6.283185307179586476925286766559;
  int ns = 1;
  for (; ns * 4 <= fft_len; ns *= 4)
    for (int r = 1; r <= 3; ++r)
      for (int k = 0; k < ns; ++k) {
        const double ang = -two_pi * r * k / (4.0 * ns);
        tw.push_back(static_cast<float>(std::cos(ang)));
        tw.push_back(static_cast<float>(std::sin(ang)));
      }
  if (ns < fft_len)
    for (int k = 0; k < ns; ++k) {
      const double ang = -two_pi * k / (2.0 * ns);
      tw.push_back(static_cast<float>(std::cos(ang)));
      tw.push_back(static_cast<float>(std::sin(ang)));
    }
  e->n_twiddle = static_cast<int>(tw.size() / 2);
  if (e->d_twiddle) { cudaFree(e->d_twiddle); e->d_twiddle = nullptr; }
  if (e->d_bins) { cudaFree(e->d_bins); e->d_bins = nullptr; }
  CK(e, cudaMalloc(&e->d_twiddle, std::max<size_t>(8, tw.size() * sizeof(float))));
  CK(e, cudaMalloc(&e->d_bins, bins.size() * sizeof(int)));
  CK(e, cudaMemcpy(e->d_twiddle, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(e, cudaMemcpy(e->d_bins, bins.data(), bins.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (!e->d_ydemod)
    CK(e, cudaMalloc(&e->d_ydemod, static_cast<size_t>(e->max_pkts) * e->cfg.n_rx * e->cfg.n_ltf * e->cfg.n_sc * sizeof(float2)));
  {
    std::vector<int> kmap(fft_len, -1);
    for (int k = 0; k < e->cfg.n_sc; ++k) kmap[bins[k]] = k;
    if (e->d_kmap) { cudaFree(e->d_kmap); e->d_kmap = nullptr; }
    CK(e, cudaMalloc(&e->d_kmap, kmap.size() * sizeof(int)));
    CK(e, cudaMemcpy(e->d_kmap, kmap.data(), kmap.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (e->d_tw256) { cudaFree(e->d_tw256); e->d_tw256 = nullptr; }
    if (e->d_tw3) { cudaFree(e->d_tw3); e->d_tw3 = nullptr; }
    if (fft_len >= 512) {
      const int r3 = fft_len / 256;
      std::vector<float> t3(static_cast<size_t>(2) * (r3 - 1) * 256);
      for (int r = 1; r < r3; ++r)
        for (int k = 0; k < 256; ++k) {
          const double ang = -6.283185307179586476925286766559 * r * k / fft_len;
          t3[2 * ((r - 1) * 256 + k)] = static_cast<float>(std::cos(ang));
          t3[2 * ((r - 1) * 256 + k) + 1] = static_cast<float>(std::sin(ang));
        }
      CK(e, cudaMalloc(&e->d_tw3, t3.size() * sizeof(float)));
      CK(e, cudaMemcpy(e->d_tw3, t3.data(), t3.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (fft_len >= 256) {
      std::vector<float> t2(2 * 15 * 16);
      for (int r = 1; r < 16; ++r)
        for (int k = 0; k < 16; ++k) {
          const double ang = -6.283185307179586476925286766559 * r * k / 256.0;
          t2[2 * ((r - 1) * 16 + k)] = static_cast<float>(std::cos(ang));
          t2[2 * ((r - 1) * 16 + k) + 1] = static_cast<float>(std::sin(ang));
        }
      CK(e, cudaMalloc(&e->d_tw256, t2.size() * sizeof(float)));
      CK(e, cudaMemcpy(e->d_tw256, t2.data(), t2.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  e->fft_len = fft_len; e->cp_len = cp_len; e->sym_offset = sym_offset;
  return MAMIMO_OK;
}

mamimo_status mamimo_ofdm_demod(mamimo_engine* e, const void* x, mamimo_ctype x_type, int64_t n_pkt, void* Y,
                                mamimo_mem mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (!e->fft_len) return fail(e, MAMIMO_ERR_STATE, "OFDM front-end not configured (mamimo_set_ofdm)");
  if (n_pkt < 0 || (n_pkt > 0 && (!x || !Y))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(Y)) & 15) return fail(e, MAMIMO_ERR_INVALID, "x and Y must be 16-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t xb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.n_ltf * (e->fft_len + e->cp_len) * (x_type == MAMIMO_C128 ? 16 : 8);
  const size_t yb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.n_ltf * e->cfg.n_sc * 8;
  auto stage = [&](int64_t n, const void* in0, const void*, void* out, float*, float*, cudaStream_t st) {
    return run_ofdm(e, in0, x_type == MAMIMO_C128, n, static_cast<float2*>(out), st);
  };
  return run_chunked(e, n_pkt, e->max_pkts, x, xb, nullptr, 0, Y, yb, nullptr, nullptr, 0, mem,
                     static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_estimate_time(mamimo_engine* e, const void* x, mamimo_ctype x_type, int64_t n_pkt, void* H_ls,
                                   float* H_real, float* H_imag, mamimo_mem mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (!e->fft_len) return fail(e, MAMIMO_ERR_STATE, "OFDM front-end not configured (mamimo_set_ofdm)");
  if (!e->dP) return fail(e, MAMIMO_ERR_STATE, "pilots / P not set (mamimo_set_pilots)");
  const bool mlp = e->n_layers > 0;
  if (mlp && e->cfg.input_mode != MAMIMO_INPUT_LS) return fail(e, MAMIMO_ERR_STATE, "engine not configured for mode C (INPUT_LS)");
  if (mlp && !e->finalized) return fail(e, MAMIMO_ERR_STATE, "weights not finalised");
  if (n_pkt < 0 || (n_pkt > 0 && (!x || (mlp && (!H_real || !H_imag)) || (!mlp && !H_ls))))
    return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(H_ls) | reinterpret_cast<uintptr_t>(H_real) |
       reinterpret_cast<uintptr_t>(H_imag)) & 15)
    return fail(e, MAMIMO_ERR_INVALID, "buffers must be 16-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t xb = static_cast<size_t>(e->cfg.n_rx) * e->cfg.n_ltf * (e->fft_len + e->cp_len) * (x_type == MAMIMO_C128 ? 16 : 8);
  const size_t hlsb = static_cast<size_t>(e->rows_per_pkt) * e->cfg.n_sc * 8;
  const size_t hb = static_cast<size_t>(e->rows_per_pkt) * e->cfg.d_out * sizeof(float);
  auto stage = [&](int64_t n, const void* in0, const void*, void* hls, float* hr, float* hi, cudaStream_t st) {
    mamimo_status s = run_ofdm(e, in0, x_type == MAMIMO_C128, n, e->d_ydemod, st);
    if (s != MAMIMO_OK) return s;
    if (mlp) {
      s = dyn_begin(e, e->d_ydemod, static_cast<size_t>(n) * e->cfg.n_rx * e->cfg.n_ltf * e->cfg.n_sc * 2, nullptr, 0,
                    false, st, true);
      if (s != MAMIMO_OK) return s;
    }
    s = DISPATCH_S(e, (run_ls<S>(e, e->d_ydemod, 0, static_cast<int>(n), hls, 0, mlp, st)));
    if (s != MAMIMO_OK || !mlp) return s;
    return DISPATCH_S(e, (run_mlp_both<S>(e, static_cast<int>(n) * e->rows_per_pkt, hr, hi, st)));
  };
  return run_chunked(e, n_pkt, e->max_pkts, x, xb, nullptr, 0, H_ls, hlsb, mlp ? H_real : nullptr,
                     mlp ? H_imag : nullptr, hb, mem, static_cast<cudaStream_t>(stream), stage);
}

// ---- next row: LMMSE smoother, LMMSE_ce.m:23-39 via helperMIMOChannelEstimate.m:37-39 ------------
double mamimo_tau_rms(const double* h, int32_t n, int32_t is_complex) {
  // LMMSE_ce.m:27-30: k = 0:n-1; hh = h*h'; tmp = |h|.^2 .* k; r = sum(tmp)/hh; r2 = tmp*k.'/hh; sqrt(r2 - r^2)
  if (!h || n < 1) return 0.0;
  double hh = 0, s1 = 0, s2 = 0;
  for (int k = 0; k < n; ++k) {
    const double p = is_complex ? h[2 * k] * h[2 * k] + h[2 * k + 1] * h[2 * k + 1] : h[k] * h[k];
    hh += p; s1 += p * k; s2 += p * k * static_cast<double>(k);
  }
  const double r = s1 / hh, r2 = s2 / hh;
  return std::sqrt(r2 - r * r);
}

mamimo_status mamimo_lmmse(mamimo_engine* e, const void* H_ls, mamimo_ctype h_type, int64_t n_pkt,
                           const double* tau_rms, const double* snr_db, void* H_mmse, mamimo_ctype out_type,
                           mamimo_mem mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (n_pkt < 0 || (n_pkt > 0 && (!H_ls || !H_mmse || !tau_rms || !snr_db))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if ((reinterpret_cast<uintptr_t>(H_ls) | reinterpret_cast<uintptr_t>(H_mmse)) & 15) return fail(e, MAMIMO_ERR_INVALID, "H_ls and H_mmse must be 16-byte aligned");
  if (n_pkt == 0) return MAMIMO_OK;
  CK(e, cudaSetDevice(e->cfg.device));
  const int n = e->cfg.n_sc, nt = e->cfg.n_tx, nrx = e->cfg.n_rx;
  const int nb = (n + kLmNB - 1) / kLmNB, n_pad = nb * kLmNB, nt_pad = round_up(nt, 8), R = n_pad + nt_pad;
  const int64_t n_slab = n_pkt * nrx;
  const size_t per_slab = (static_cast<size_t>(R) * n_pad + static_cast<size_t>(nb) * kLmNB * kLmNB) * sizeof(double2);
  {
    size_t budget = static_cast<size_t>(8) << 30;                       // workspace budget (MAMIMO_LMMSE_WS_MB)
    if (const char* env = getenv("MAMIMO_LMMSE_WS_MB")) if (atoll(env) > 0) budget = static_cast<size_t>(atoll(env)) << 20;
    const int64_t cap = std::max<int64_t>(1, std::min<int64_t>(static_cast<int64_t>(budget / per_slab), 32768));
    const int want = static_cast<int>(std::min<int64_t>(cap, std::max<int64_t>(n_slab, 4LL * nrx)));
    if (want > e->lm_slabs) {                                            // first call, or a larger batch than before
      CK(e, cudaDeviceSynchronize());
      if (e->lm_M) cudaFree(e->lm_M);
      if (e->lm_Dinv) cudaFree(e->lm_Dinv);
      if (e->lm_par) cudaFree(e->lm_par);
      e->lm_M = e->lm_Dinv = e->lm_par = nullptr;
      const bool first = e->lm_slabs == 0;
      e->lm_slabs = 0;
      size_t free_b = 0, total_b = 0;
      CK(e, cudaMemGetInfo(&free_b, &total_b));
      const int fit = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, (free_b / 2) / per_slab)));
      CK(e, cudaMalloc(&e->lm_M, static_cast<size_t>(fit) * R * n_pad * sizeof(double2)));
      CK(e, cudaMalloc(&e->lm_Dinv, static_cast<size_t>(fit) * nb * kLmNB * kLmNB * sizeof(double2)));
      CK(e, cudaMalloc(&e->lm_par, static_cast<size_t>(fit) * sizeof(double2)));
      e->lm_slabs = fit;
      if (first) {
        CK(e, cudaFuncSetAttribute(lmmse_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLmPanelSmem));
        CK(e, cudaFuncSetAttribute(lmmse_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLmBsSmem));
        CK(e, cudaFuncSetAttribute(lmmse_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLmBsSmem));
        if (2 * n_pad * sizeof(double2) > 48 * 1024)
          CK(e, cudaFuncSetAttribute(lmmse_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(2 * n_pad * sizeof(double2))));
        if (const char* env = getenv("MAMIMO_LMMSE_SCHUR")) e->lm_schur = atoi(env) != 0;
        for (int i = 0; i < 4; ++i) CK(e, cudaEventCreateWithFlags(&e->lm_ev[i], cudaEventDisableTiming));
        if (const char* env = getenv("MAMIMO_LMMSE_STREAMS")) if (atoi(env) > 0) e->lm_groups = atoi(env);
      }
    }
  }
  const size_t in_el = h_type == MAMIMO_C128 ? 16 : 8, out_el = out_type == MAMIMO_C128 ? 16 : 8;
  const size_t slab_el = static_cast<size_t>(nt) * n;
  if (mem == MAMIMO_MEM_HOST && e->lm_io_bytes < static_cast<size_t>(e->lm_slabs) * slab_el * 16) {
    if (e->lm_in) cudaFree(e->lm_in);
    if (e->lm_out) cudaFree(e->lm_out);
    e->lm_in = e->lm_out = nullptr;
    e->lm_io_bytes = static_cast<size_t>(e->lm_slabs) * slab_el * 16;
    CK(e, cudaMalloc(&e->lm_in, e->lm_io_bytes));
    CK(e, cudaMalloc(&e->lm_out, e->lm_io_bytes));
  }
  cudaStream_t st = mem == MAMIMO_MEM_HOST ? e->s_comp : static_cast<cudaStream_t>(stream);
  const double two_pi = 6.283185307179586476925286766559;
  std::vector<double2> par;
  for (int64_t s0 = 0; s0 < n_slab; s0 += e->lm_slabs) {
    const int ns = static_cast<int>(std::min<int64_t>(e->lm_slabs, n_slab - s0));
    par.resize(ns);
    for (int i = 0; i < ns; ++i) {
      const int64_t slab = s0 + i;
      const double snr = std::pow(10.0, snr_db[slab] * 0.1);           // LMMSE_ce.m:23
      par[i] = make_double2(two_pi * tau_rms[slab / nrx] / n, 1.0 / snr);   // j2pi_tau_df, LMMSE_ce.m:31-32
    }
    CK(e, cudaStreamSynchronize(st));                                  // previous chunk done with lm_par / staging
    CK(e, cudaMemcpyAsync(e->lm_par, par.data(), ns * sizeof(double2), cudaMemcpyHostToDevice, st));
    const char* in_p = static_cast<const char*>(H_ls) + s0 * slab_el * in_el;
    char* out_p = static_cast<char*>(H_mmse) + s0 * slab_el * out_el;
    LmArgs a;
    memset(&a, 0, sizeof(a));
    if (mem == MAMIMO_MEM_HOST) {
      CK(e, cudaMemcpyAsync(e->lm_in, in_p, ns * slab_el * in_el, cudaMemcpyHostToDevice, st));
      e->stats.h2d_bytes += ns * slab_el * in_el;
      a.B = e->lm_in; a.out = e->lm_out;
    } else {
      a.B = in_p; a.out = out_p;
    }
    a.M = e->lm_M; a.Dinv = e->lm_Dinv; a.par = e->lm_par;
    a.b_double = h_type == MAMIMO_C128; a.out_double = out_type == MAMIMO_C128;
    a.n = n; a.n_pad = n_pad; a.nb = nb; a.n_tx = nt; a.nt_pad = nt_pad; a.R = R; a.n_ps = e->cfg.n_ps;
    a.flags = e->d_flags;
    // The chunk is split into groups of slabs that run on separate streams: the diag kernel is latency-bound
    // (64 dependent pivot steps per block), the panel kernel throughput-bound, so the groups' kernels fill each
    // other's gaps (measured at 32x4x234, 500 packets: 6.8 ms on one stream).
    int n_grp = std::max(1, std::min({e->lm_groups, 4, ns}));
    cudaStream_t gst[4] = {st, e->s_side, e->s_h2d, e->s_d2h};
    if (mem == MAMIMO_MEM_HOST) gst[2] = gst[3] = nullptr;             // keep the copy streams out of it
    if (mem == MAMIMO_MEM_HOST) n_grp = std::min(n_grp, 2);
    if (n_grp > 1) {
      CK(e, cudaEventRecord(e->lm_ev[0], st));
      for (int g = 1; g < n_grp; ++g) CK(e, cudaStreamWaitEvent(gst[g], e->lm_ev[0], 0));
    }
    const int per_grp = (ns + n_grp - 1) / n_grp;
    for (int J = 0; J <= nb; ++J) {
      for (int g = 0; g < n_grp; ++g) {
        const int s_lo = g * per_grp, s_n = std::min(per_grp, ns - s_lo);
        if (s_n <= 0) continue;
        LmArgs b = a;
        b.M = a.M + static_cast<size_t>(s_lo) * R * n_pad;
        b.Dinv = a.Dinv + static_cast<size_t>(s_lo) * nb * kLmNB * kLmNB;
        b.par = a.par + s_lo;
        b.B = static_cast<const char*>(a.B) + s_lo * slab_el * in_el;
        b.out = static_cast<char*>(a.out) + s_lo * slab_el * out_el;
        b.J = J;
        if (e->lm_schur) {                     // Toeplitz route: O(n^2) Schur factorisation + blocked triangular solves
          if (J != nb) continue;
          {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_schur_kernel<<<s_n, 128, 2 * n_pad * sizeof(double2), gst[g]>>>(b);
          }
          {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_linv_kernel<<<dim3(nb, s_n), 256, 0, gst[g]>>>(b);
          }
          {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_solve_kernel<<<dim3((nt_pad + 31) / 32, s_n), 128, kLmBsSmem, gst[g]>>>(b);
          }
          e->stats.kernel_launches += 3;
          if (e->cfg.n_ps != 1) {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_rhp_kernel<<<dim3((n + 127) / 128, nt, s_n), 128, 0, gst[g]>>>(b);
            e->stats.kernel_launches++;
          }
          continue;
        }
        if (J < nb) {
          {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_diag_kernel<<<s_n, 256, 0, gst[g]>>>(b);
          }
          const int rows_below = R - (J + 1) * kLmNB;
          {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_panel_kernel<<<dim3((rows_below + kLmPanelRows - 1) / kLmPanelRows, s_n), 256, kLmPanelSmem, gst[g]>>>(b);
          }
          e->stats.kernel_launches += 2;
        } else {
          {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_backsub_kernel<<<dim3((nt_pad + 31) / 32, s_n), 128, kLmBsSmem, gst[g]>>>(b);
          }
          e->stats.kernel_launches++;
          if (e->cfg.n_ps != 1) {
            ProfScope ps(e, gst[g], kClsLmmse);
            lmmse_rhp_kernel<<<dim3((n + 127) / 128, nt, s_n), 128, 0, gst[g]>>>(b);
            e->stats.kernel_launches++;
          }
        }
      }
    }
    for (int g = 1; g < n_grp; ++g) {
      CK(e, cudaEventRecord(e->lm_ev[g], gst[g]));
      CK(e, cudaStreamWaitEvent(st, e->lm_ev[g], 0));
    }
    CK(e, cudaGetLastError());
    if (mem == MAMIMO_MEM_HOST) {
      CK(e, cudaMemcpyAsync(out_p, e->lm_out, ns * slab_el * out_el, cudaMemcpyDeviceToHost, st));
      e->stats.d2h_bytes += ns * slab_el * out_el;
    }
  }
  if (mem == MAMIMO_MEM_HOST) {
    return check_flags(e, st);
  }
  return MAMIMO_OK;
}

// ---- next row: per-subcarrier SVD invariants of H-hat, omphybweights.m:169-176 -------------------
mamimo_status mamimo_svd(mamimo_engine* e, const void* H, mamimo_ctype h_type, int64_t n_pkt, void* sigma, void* V1,
                         mamimo_ctype out_type, mamimo_mem mem, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (n_pkt < 0 || (n_pkt > 0 && (!H || !sigma))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if (e->cfg.n_rx > 8) return fail(e, MAMIMO_ERR_UNSUPPORTED, "mamimo_svd supports n_rx <= 8");
  if ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(sigma) | reinterpret_cast<uintptr_t>(V1)) & 15)
    return fail(e, MAMIMO_ERR_INVALID, "H, sigma and V1 must be 16-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  const int nr = e->cfg.n_rx, nt = e->cfg.n_tx, nsc = e->cfg.n_sc;
  const size_t hb = static_cast<size_t>(nr) * nt * nsc * (h_type == MAMIMO_C128 ? 16 : 8);
  const size_t vb = static_cast<size_t>(nr) * nt * nsc * (out_type == MAMIMO_C128 ? 16 : 8);
  const size_t sb = static_cast<size_t>(nr) * nsc * (out_type == MAMIMO_C128 ? 8 : 4);
  auto stage = [&](int64_t n, const void* in0, const void*, void* v1, float* sg, float*, cudaStream_t st) {
    SvdArgs a;
    memset(&a, 0, sizeof(a));
    a.H = in0; a.sigma = sg; a.V1 = v1;
    a.h_double = h_type == MAMIMO_C128; a.out_double = out_type == MAMIMO_C128;
    a.n_tx = nt; a.n_sc = nsc;
    const dim3 grid((nsc + 127) / 128, static_cast<unsigned>(n));
    const dim3 grid_s((nsc + kSvdSmemThreads - 1) / kSvdSmemThreads, static_cast<unsigned>(n));
    ProfScope ps(e, st, kClsStage);
#define SVD_SMEM_CASE(NR)                                                                                              \
  CK(e, cudaFuncSetAttribute(svd_gram_smem_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, svd_smem_bytes<NR>())); \
  svd_gram_smem_kernel<NR><<<grid_s, kSvdSmemThreads, svd_smem_bytes<NR>(), st>>>(a);
    switch (nr) {
      case 1: if (a.h_double) svd_gram_kernel<1, true><<<grid, 128, 0, st>>>(a); else svd_gram_kernel<1, false><<<grid, 128, 0, st>>>(a); break;
      case 2: if (a.h_double) svd_gram_kernel<2, true><<<grid, 128, 0, st>>>(a); else svd_gram_kernel<2, false><<<grid, 128, 0, st>>>(a); break;
      case 3: if (a.h_double) svd_gram_kernel<3, true><<<grid, 128, 0, st>>>(a); else svd_gram_kernel<3, false><<<grid, 128, 0, st>>>(a); break;
      case 4: if (a.h_double) svd_gram_kernel<4, true><<<grid, 128, 0, st>>>(a); else svd_gram_kernel<4, false><<<grid, 128, 0, st>>>(a); break;
      case 5: { SVD_SMEM_CASE(5) } break;
      case 6: { SVD_SMEM_CASE(6) } break;
      case 7: { SVD_SMEM_CASE(7) } break;
      default: { SVD_SMEM_CASE(8) } break;
    }
#undef SVD_SMEM_CASE
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return MAMIMO_OK;
  };
  // chunks of at most 65535 packets (grid.y); the host path additionally streams in host_chunk units
  return run_chunked(e, n_pkt, std::min<int64_t>(e->max_pkts, 65535), H, hb, nullptr, 0, V1, vb, static_cast<float*>(sigma),
                     nullptr, sb, mem, static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_set_steering_dictionary(mamimo_engine* e, const double* At, int32_t n_rays) {
  if (!e || !At) return MAMIMO_ERR_INVALID;
  if (n_rays < 1 || n_rays > (1 << 20)) return fail(e, MAMIMO_ERR_INVALID, "need 1 <= n_rays <= 2^20");
  const int nt = e->cfg.n_tx;
  for (size_t i = 0; i < static_cast<size_t>(2) * n_rays * nt; ++i)
    if (!std::isfinite(At[i])) return fail(e, MAMIMO_ERR_INVALID, "dictionary holds a non-finite value");
  CK(e, cudaSetDevice(e->cfg.device));
  CK(e, cudaDeviceSynchronize());
  const int pad = (n_rays + kOmpTile - 1) / kOmpTile * kOmpTile;
  std::vector<double> t(static_cast<size_t>(2) * nt * pad, 0.0);
  for (int r = 0; r < n_rays; ++r)
    for (int j = 0; j < nt; ++j) {
      t[2 * (static_cast<size_t>(j) * pad + r)] = At[2 * (static_cast<size_t>(r) * nt + j)];
      t[2 * (static_cast<size_t>(j) * pad + r) + 1] = -At[2 * (static_cast<size_t>(r) * nt + j) + 1];
    }
  if (e->d_omp_At) { cudaFree(e->d_omp_At); e->d_omp_At = nullptr; }
  if (e->d_omp_AtcT) { cudaFree(e->d_omp_AtcT); e->d_omp_AtcT = nullptr; }
  e->omp_rays = 0;
  CK(e, cudaMalloc(&e->d_omp_At, sizeof(double2) * n_rays * nt));
  CK(e, cudaMalloc(&e->d_omp_AtcT, sizeof(double2) * nt * pad));
  CK(e, cudaMemcpy(e->d_omp_At, At, sizeof(double2) * n_rays * nt, cudaMemcpyHostToDevice));
  CK(e, cudaMemcpy(e->d_omp_AtcT, t.data(), sizeof(double2) * nt * pad, cudaMemcpyHostToDevice));
  e->omp_rays = n_rays; e->omp_rays_pad = pad;
  return MAMIMO_OK;
}

mamimo_status mamimo_omp(mamimo_engine* e, const void* F, mamimo_ctype f_type, int32_t f_rows, int64_t n_pkt, int32_t ns,
                         int32_t n_rf, int32_t* idx, float* err, void* Fbb, mamimo_ctype fbb_type, mamimo_mem mem,
                         void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  if (n_pkt < 0 || (n_pkt > 0 && (!F || !idx || !err || !Fbb))) return fail(e, MAMIMO_ERR_INVALID, "null buffer");
  if (e->omp_rays < 1) return fail(e, MAMIMO_ERR_STATE, "no steering dictionary (mamimo_set_steering_dictionary)");
  const int nt = e->cfg.n_tx, nsc = e->cfg.n_sc;
  if (ns < 1 || ns > kOmpMaxNs || ns > f_rows || ns > nt) return fail(e, MAMIMO_ERR_INVALID, "need 1 <= ns <= min(8, f_rows, n_tx)");
  if (n_rf < 1 || n_rf > kOmpMaxRf || n_rf > e->omp_rays) return fail(e, MAMIMO_ERR_INVALID, "need 1 <= n_rf <= min(8, n_rays)");
  if (nt > kOmpMaxTx) return fail(e, MAMIMO_ERR_UNSUPPORTED, "mamimo_omp supports n_tx <= 100");
  if ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(err) | reinterpret_cast<uintptr_t>(Fbb)) & 15)
    return fail(e, MAMIMO_ERR_INVALID, "F, idx, err and Fbb must be 16-byte aligned");
  CK(e, cudaSetDevice(e->cfg.device));
  const size_t fb = static_cast<size_t>(f_rows) * nt * nsc * (f_type == MAMIMO_C128 ? 16 : 8);
  const size_t bb = static_cast<size_t>(ns) * n_rf * nsc * (fbb_type == MAMIMO_C128 ? 16 : 8);
  const size_t ib = static_cast<size_t>(n_rf) * nsc * 4;
  // workspace: the residual of a chunk of packets (256 MB at most) and its active flags
  const size_t wres_pkt = static_cast<size_t>(ns) * nt * nsc * sizeof(double2);
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(e->max_pkts, 65535), (256u << 20) / wres_pkt));
  const int64_t cmax = std::min<int64_t>(chunk, std::max<int64_t>(n_pkt, 1));
  if (n_rf > 1 && e->omp_wres_bytes < cmax * wres_pkt) {
    CK(e, cudaDeviceSynchronize());
    if (e->d_omp_wres) cudaFree(e->d_omp_wres);
    e->d_omp_wres = nullptr; e->omp_wres_bytes = 0;
    CK(e, cudaMalloc(&e->d_omp_wres, cmax * wres_pkt));
    e->omp_wres_bytes = cmax * wres_pkt;
  }
  if (e->omp_active_bytes < static_cast<size_t>(cmax) * nsc) {
    CK(e, cudaDeviceSynchronize());
    if (e->d_omp_active) cudaFree(e->d_omp_active);
    e->d_omp_active = nullptr; e->omp_active_bytes = 0;
    CK(e, cudaMalloc(&e->d_omp_active, static_cast<size_t>(cmax) * nsc));
    e->omp_active_bytes = static_cast<size_t>(cmax) * nsc;
  }
  const size_t corr_smem = omp_corr_smem(nt, ns);
  const int w_res = omp_w_resident(nt, ns) ? 1 : 0;
  const int pf = (nt * kOmpTile + kOmpThreads - 1) / kOmpThreads;          // = ceil(n_tx / 4)
  auto corr = pf <= 8 ? omp_corr_kernel<8> : (pf <= 16 ? omp_corr_kernel<16> : omp_corr_kernel<25>);
  CK(e, cudaFuncSetAttribute(corr, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(corr_smem)));
  e->omp_generic = getenv("MAMIMO_OMP_GENERIC") != nullptr;     // diagnostics: the 4x4 kernel for Ns = 1 too
  const bool tall = ns == 1 && !e->omp_generic && omp_corr1_smem(nt) <= kOmpSmemBudget;
  if (tall)
    CK(e, cudaFuncSetAttribute(omp_corr1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(omp_corr1_smem(nt))));
  auto stage = [&](int64_t n, const void* in0, const void*, void* fbb, float* ix, float* er, cudaStream_t st) {
    OmpArgs a;
    memset(&a, 0, sizeof(a));
    a.F = in0; a.f_double = f_type == MAMIMO_C128; a.f_rows = f_rows;
    a.Wres = e->d_omp_wres; a.AtcT = e->d_omp_AtcT; a.At = e->d_omp_At;
    a.idx = reinterpret_cast<int32_t*>(ix); a.err = er; a.Fbb = fbb; a.fbb_double = fbb_type == MAMIMO_C128;
    a.active = e->d_omp_active;
    a.n_tx = nt; a.n_sc = nsc; a.n_rays = e->omp_rays; a.n_rays_pad = e->omp_rays_pad; a.ns = ns; a.n_rf = n_rf;
    CK(e, cudaMemsetAsync(e->d_omp_active, 1, static_cast<size_t>(n) * nsc, st));
    const dim3 grid_c((nsc + kOmpTile - 1) / kOmpTile, static_cast<unsigned>(n));
    const dim3 grid_r((nsc + 127) / 128, static_cast<unsigned>(n));
    const int need = std::max(ns, n_rf);
    for (int r = 0; r < n_rf; ++r) {
      a.round = r;
      {
        ProfScope ps(e, st, kClsStage);
        if (tall) omp_corr1_kernel<<<grid_c, kOmp1Threads, omp_corr1_smem(nt), st>>>(a);
        else corr<<<grid_c, kOmpThreads, corr_smem, st>>>(a, w_res);
      }
      {
        ProfScope ps(e, st, kClsStage);
        if (need <= 1) omp_refit_kernel<1, 1><<<grid_r, 128, 0, st>>>(a);
        else if (need <= 2) omp_refit_kernel<2, 2><<<grid_r, 128, 0, st>>>(a);
        else if (need <= 4) omp_refit_kernel<4, 4><<<grid_r, 128, 0, st>>>(a);
        else omp_refit_kernel<8, 8><<<grid_r, 128, 0, st>>>(a);
      }
      CK(e, cudaGetLastError());
      e->stats.kernel_launches += 2;
    }
    return MAMIMO_OK;
  };
  return run_chunked(e, n_pkt, chunk, F, fb, nullptr, 0, Fbb, bb, reinterpret_cast<float*>(idx), err, ib, mem,
                     static_cast<cudaStream_t>(stream), stage);
}

mamimo_status mamimo_synchronize(mamimo_engine* e) {
  if (!e) return MAMIMO_ERR_INVALID;
  CK(e, cudaSetDevice(e->cfg.device));
  CK(e, cudaDeviceSynchronize());
  return check_flags(e, e->s_comp);
}

mamimo_status mamimo_poll_flags(mamimo_engine* e, void* stream) {
  if (!e) return MAMIMO_ERR_INVALID;
  CK(e, cudaSetDevice(e->cfg.device));
  return check_flags(e, static_cast<cudaStream_t>(stream));
}

mamimo_status mamimo_get_stats(const mamimo_engine* e, mamimo_stats* out) {
  if (!e || !out) return MAMIMO_ERR_INVALID;
  *out = e->stats;
  return MAMIMO_OK;
}

mamimo_status mamimo_get_debug_counters(mamimo_engine* e, uint64_t out[8], int32_t reset) {
  if (!e || !out) return MAMIMO_ERR_INVALID;
  memset(out, 0, 8 * sizeof(uint64_t));
  if (!e->d_dbg) return fail(e, MAMIMO_ERR_STATE, "role counters are off: set MAMIMO_FC_DEBUG=1 before mamimo_create on a library built with -DMAMIMO_FC_DEBUG_COUNTERS");
  CK(e, cudaSetDevice(e->cfg.device));
  CK(e, cudaDeviceSynchronize());
  CK(e, cudaMemcpy(out, e->d_dbg, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (reset) CK(e, cudaMemset(e->d_dbg, 0, 8 * sizeof(uint64_t)));
  return MAMIMO_OK;
}

mamimo_status mamimo_profile_begin(mamimo_engine* e) {
  if (!e) return MAMIMO_ERR_INVALID;
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof.clear();
  e->profiling = true;
  return MAMIMO_OK;
}

mamimo_status mamimo_profile_end(mamimo_engine* e, mamimo_profile* out) {
  if (!e || !out) return MAMIMO_ERR_INVALID;
  e->profiling = false;
  CK(e, cudaSetDevice(e->cfg.device));
  CK(e, cudaDeviceSynchronize());
  memset(out, 0, sizeof(*out));
  if (const char* path = getenv("MAMIMO_TIMELINE")) {        // diagnostics: per-launch start/end (ms since the first launch)
    if (FILE* f = e->prof.empty() ? nullptr : fopen(path, "w")) {
      fprintf(f, "idx,class,stream,start_ms,end_ms\n");
      int i = 0;
      for (auto& r : e->prof) {
        float t0 = 0.f, t1 = 0.f;
        if (cudaEventElapsedTime(&t0, e->prof[0].a, r.a) == cudaSuccess && cudaEventElapsedTime(&t1, e->prof[0].a, r.b) == cudaSuccess)
          fprintf(f, "%d,%s,%s,%.4f,%.4f\n", i, r.cls == kClsLs ? "ls" : r.cls == kClsFc ? "fc" : r.cls == kClsLmmse ? "lmmse" : "stage",
                  r.st == e->s_side ? "side" : "main", t0, t1);
        ++i;
      }
      fclose(f);
    }
  }
  for (auto& r : e->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      if (r.cls == kClsLs) { out->ls_ms += ms; out->ls_launches++; }
      else if (r.cls == kClsFc) { out->fc_ms += ms; out->fc_launches++; }
      else if (r.cls == kClsLmmse) { out->lmmse_ms += ms; out->lmmse_launches++; }
      else { out->stage_ms += ms; out->stage_launches++; }
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  e->prof.clear();
  return MAMIMO_OK;
}

}  // extern "C"
