/*
 * mamimo.h -- C ABI of the B200-native massive-MIMO OFDM channel-estimation engine.
 *
 * This is the drop-in boundary for ONE hot path of
 * mauro-belgiovine/DL-channel-estimation-MaMIMO (citations relative to the
 * reference root; pg/ = packet_generation/phased_arr/):
 *
 *   LS estimate (P-matrix despread + divide by LTF tone)     pg/helperMIMOChannelEstimate.m:24-36
 *   [optional comb-pilot linear interpolation, north_star]   (no reference counterpart; Nps=1 everywhere)
 *   real/imag FC denoiser per (tx,rx) pair                   massiveMIMO_CSI_prediction_DNN.py:173-234,330-346
 *   CSIPredictor.inference planes -> two predicts            inference.py:24-32
 *
 * Plain C: opaque handle, integer status codes, raw pointers + sizes, no
 * exceptions, no torch / CUDA types in any signature (streams travel as void*).
 * The caller owns every buffer it passes; the engine owns weights, tables and
 * workspace.  One engine per device; calls on one engine are serialised by the
 * caller (MATLAB's interpreter thread / the Python GIL do that already).
 *
 * Layouts (all C order, last index contiguous):
 *   Y      complex [n_pkt][n_rx][n_ltf][n_sc]   == MATLAB rxData [Nsc x nltf x Nr] per packet (column-major)
 *   H_ls   complex [n_pkt][n_rx][n_tx][n_sc]    == MATLAB hD     [Nsc x numSTS x Nr] per packet
 *   H_real / H_imag  float32 [n_pkt*n_rx*n_tx][d_out], row = p*(n_rx*n_tx) + i_rx*n_tx + i_tx
 *          (create_massiveMIMO_CSIest_dnn_dataset.py:62; inverse pg/BER_test_maMIMO_LTF.m:213-218)
 */
#ifndef MAMIMO_H_
#define MAMIMO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MAMIMO_API __declspec(dllexport)
#else
#define MAMIMO_API __attribute__((visibility("default")))
#endif

#define MAMIMO_ABI_VERSION 2
#define MAMIMO_MAX_HIDDEN 8

typedef struct mamimo_engine mamimo_engine;

typedef enum {
  MAMIMO_OK = 0,
  MAMIMO_ERR_INVALID = 1,      /* bad argument / shape (reference: validateattributes / error(message(...))) */
  MAMIMO_ERR_CUDA = 2,         /* CUDA runtime / driver failure; see mamimo_last_error */
  MAMIMO_ERR_NOMEM = 3,
  MAMIMO_ERR_STATE = 4,        /* e.g. estimate before weights are finalised */
  MAMIMO_ERR_UNSUPPORTED = 5,  /* e.g. no sm_100 device present */
  MAMIMO_ERR_RANGE = 6,        /* split-fp16 operand outside its range window (overflow or underflow), detected on device */
  MAMIMO_ERR_TIMEOUT = 7       /* a device-side pipeline wait timed out (kernel aborted itself) */
} mamimo_status;

/* element types of caller buffers */
typedef enum { MAMIMO_C64 = 0, MAMIMO_C128 = 1 } mamimo_ctype;   /* interleaved re,im (float / double) */
typedef enum { MAMIMO_MEM_HOST = 0, MAMIMO_MEM_DEVICE = 1 } mamimo_mem;

/* arithmetic of the FC layers (accumulation is always FP32) */
typedef enum {
  MAMIMO_PREC_FP32_SIMT = 0,   /* CUDA-core FFMA, exact FP32: on-device accuracy anchor */
  MAMIMO_PREC_TF32X3 = 1,      /* tcgen05 kind::tf32, error-compensated hi/lo split, 3 MMA passes; bitwise independent of how a
                                  batch is chunked */
  MAMIMO_PREC_FP16X3 = 2,      /* DEFAULT.  tcgen05 kind::f16 (fp16), hi/lo split scaled by device-chosen powers of two
                                  (act_scale_log2), 3 MMA passes: same accuracy as TF32X3 at twice the MMA rate */
  MAMIMO_PREC_BF16X1 = 3       /* tcgen05 kind::f16 (bf16), single pass -- NOT within 1e-5; diagnostics only */
} mamimo_precision;

/* which stage feeds the first FC layer */
typedef enum {
  MAMIMO_INPUT_LS = 0,      /* mode C (north_star): real/imag planes of the in-engine LS(+interp) estimate */
  MAMIMO_INPUT_PLANES = 1,  /* mode B (inference.py:29-30): caller vector, .real -> real net, .imag -> imag net */
  MAMIMO_INPUT_TIME_P = 2   /* mode A (massiveMIMO_dataGenerator.py:303-316): [time-domain LTF || P(:,iTx)] */
} mamimo_input_mode;

typedef struct {
  int32_t abi_version;      /* MAMIMO_ABI_VERSION */
  int32_t device;           /* CUDA device ordinal */
  int32_t n_tx;             /* numSTS (helperMIMOChannelEstimate.m:9) */
  int32_t n_rx;             /* numRx  (:8) */
  int32_t n_ltf;            /* nltf   (:8); == n_tx during sounding, 1 in the data phase */
  int32_t n_sc;             /* estimated tones = numel(prm.CarriersLocations) (:26) */
  int32_t n_ps;             /* pilot spacing; 1 = every tone is a pilot (all reference call sites) */
  int32_t input_mode;       /* mamimo_input_mode */
  int32_t precision;        /* mamimo_precision */
  int32_t d_in;             /* FC input width: mode C n_sc; mode A len_ltf + n_tx; mode B anything */
  int32_t d_out;            /* FC output width = simParams['nSubCarr'] (..._DNN.py:227) */
  int32_t n_hidden;         /* number of Dense(relu) layers (--nn), 0 = LS only */
  int32_t hidden[MAMIMO_MAX_HIDDEN];
  int32_t len_ltf;          /* mode A only: time-domain samples per (pkt,rx) fed to the net */
  int32_t max_pkts;         /* packets per internal chunk (workspace sizing); 0 = default */
  int32_t act_scale_log2;   /* FP16X3 only.  0 (default) = auto: every activation level gets a power-of-two scale chosen
                               on the device per call from the measured input amax and weight-norm bounds, so the result
                               does not depend on the amplitude of Y.  != 0 pins 2^act_scale_log2 for all levels (no
                               amax pre-pass); overflow AND underflow of the fp16 window then raise MAMIMO_ERR_RANGE */
  int32_t kb_per_chunk;     /* FC: k-blocks accumulated in the tensor core between FP32 register drains; 0 = default (4) */
  int32_t host_chunk_pkts;  /* packets (rows in mode B) per H2D/compute/D2H pipeline chunk for HOST buffers; 0 = default (~2K rows) */
  int32_t fc_single_cta;    /* 1 = use the 1-CTA FC kernel instead of the CTA-pair (cta_group::2) kernel */
  int32_t fc_sm_reserve;    /* SMs the persistent FC kernels leave free (room for a concurrent NCCL all-gather) */
  int32_t reserved[3];
} mamimo_config;

typedef struct {
  uint64_t kernel_launches;   /* engine kernels launched since create */
  uint64_t h2d_bytes;         /* bytes copied host->device by the engine */
  uint64_t d2h_bytes;
  uint32_t last_device_flags; /* bit0 range overflow, bit1 pipeline timeout, bit2 LMMSE matrix not positive definite,
                                 bit3 range underflow (pinned act_scale_log2 only) */
  uint32_t graph_launches;    /* device-resident full-path calls issued as ONE CUDA-graph launch (captured once per
                                 (buffers, batch size, stream), replayed afterwards; MAMIMO_GRAPH=0 disables) */
} mamimo_stats;

/* device time per kernel class, measured with CUDA events recorded on the launching stream */
typedef struct {
  double ls_ms, fc_ms, stage_ms;                 /* summed event-to-event durations */
  uint64_t ls_launches, fc_launches, stage_launches;
  double lmmse_ms;                               /* LMMSE smoother kernels (fill, diag, panel, backsub) */
  uint64_t lmmse_launches;
} mamimo_profile;

/* ---- library-level ------------------------------------------------------- */
MAMIMO_API int32_t mamimo_abi_version(void);
MAMIMO_API const char* mamimo_status_string(mamimo_status s);
/* last error text for this engine (or the last create failure when e == NULL) */
MAMIMO_API const char* mamimo_last_error(const mamimo_engine* e);
MAMIMO_API void mamimo_config_init(mamimo_config* cfg);   /* zero + defaults */

/* ---- integer tables of the reference (bit-exact parity surface) ----------
 * replaces the literals in pg/helperMIMOChannelEstimate.m:16-23 and
 * pg/generate_maMIMO_LTF.m:99-102 */
MAMIMO_API void mamimo_vht_ltf256(int8_t out[256]);
MAMIMO_API int32_t mamimo_carriers_locations(int32_t* out, int32_t capacity);  /* 1-based; returns count (234) */
/* default P = Sylvester-Hadamard(n) stand-in for helperGetP (pg/helperMIMOChannelEstimate.m:13) */
MAMIMO_API mamimo_status mamimo_default_p(int32_t n, float* out /* [n][n] */);
/* row of pair (p, i_rx, i_tx): create_massiveMIMO_CSIest_dnn_dataset.py:62 */
MAMIMO_API int64_t mamimo_pair_row(int64_t p, int32_t i_rx, int32_t i_tx, int32_t n_rx, int32_t n_tx);

/* ---- engine lifetime ----------------------------------------------------- */
MAMIMO_API mamimo_status mamimo_create(const mamimo_config* cfg, mamimo_engine** out);
MAMIMO_API void mamimo_destroy(mamimo_engine* e);

/* pilot tones X_pilot [n_pilots] (complex64 interleaved, n_pilots = ceil(n_sc/n_ps); NULL = all +1) and
 * mapping matrix P [n_tx][n_ltf] (complex64 interleaved; NULL = default_p).  Replaces ltf(ind) and
 * helperGetP in pg/helperMIMOChannelEstimate.m:13,27. */
MAMIMO_API mamimo_status mamimo_set_pilots(mamimo_engine* e, const float* x_pilot, const float* P);
/* The same tables in double precision (complex128 interleaved).  mamimo_ls_estimate with complex128 Y AND complex128
 * H_ls then runs the despread in FP64 end to end, as MATLAB does (hD matches helperMIMOChannelEstimate.m to ~1e-15);
 * every other combination computes in FP32 (the hot path ends in FP32 operand planes). */
MAMIMO_API mamimo_status mamimo_set_pilots_f64(mamimo_engine* e, const double* x_pilot, const double* P);

/* One Dense layer of net (0 = 'real', 1 = 'imag'); layer in [0, n_hidden].  W is the Keras kernel
 * [in][out] row-major, b [out]; BN vectors [out] (all NULL when the layer has no BatchNormalization;
 * the final linear layer never has one).  Replaces Model.load_weights (..._DNN.py:334). */
MAMIMO_API mamimo_status mamimo_load_layer(mamimo_engine* e, int32_t net, int32_t layer,
                                           const float* W, const float* b,
                                           const float* bn_gamma, const float* bn_beta,
                                           const float* bn_mean, const float* bn_var);
/* fold BN into the following Dense, split/convert for the selected precision, upload */
MAMIMO_API mamimo_status mamimo_finalize_weights(mamimo_engine* e);

/* ---- the hot path -------------------------------------------------------- */
/* LS (+interp) only: drop-in for pg/helperMIMOChannelEstimate.m:33-36 over a batch of packets. */
MAMIMO_API mamimo_status mamimo_ls_estimate(mamimo_engine* e, const void* Y, mamimo_ctype y_type,
                                            mamimo_mem y_mem, int64_t n_pkt,
                                            void* H_ls, mamimo_ctype h_type, mamimo_mem h_mem,
                                            void* stream);

/* Full path, mode C: Y -> LS -> interp -> two FC nets.  H_ls may be NULL.  mem applies to Y, H_ls,
 * H_real, H_imag alike.  Host buffers are streamed through the device in chunks of max_pkts with
 * copies overlapped with compute.  Synchronous w.r.t. the host when mem == HOST; asynchronous on
 * `stream` when mem == DEVICE.  On a non-default stream the device-resident call is captured into a CUDA graph the
 * first time it is seen with given buffers / batch size and replayed as ONE graph launch afterwards. */
MAMIMO_API mamimo_status mamimo_estimate(mamimo_engine* e, const void* Y, mamimo_ctype y_type,
                                         int64_t n_pkt, void* H_ls, float* H_real, float* H_imag,
                                         mamimo_mem mem, void* stream);

/* The same path run piecewise so a caller can overlap its own work (e.g. the all-gather of the real plane)
 * with the remaining stages.  stages is a mask of MAMIMO_STAGE_*; LS leaves the operand planes of both
 * nets in the engine's workspace, the NET stages consume them.  Partial masks: device buffers, one chunk. */
#define MAMIMO_STAGE_LS 1u
#define MAMIMO_STAGE_NET_REAL 2u
#define MAMIMO_STAGE_NET_IMAG 4u
#define MAMIMO_STAGE_GATHER 8u   /* final layers also TMA-store every tile into every rank's gathered plane */
MAMIMO_API mamimo_status mamimo_estimate_stages(mamimo_engine* e, const void* Y, mamimo_ctype y_type,
                                                int64_t n_pkt, void* H_ls, float* H_real, float* H_imag,
                                                mamimo_mem mem, void* stream, uint32_t stages);

/* ---- fused all-gather of H-hat (multi-GPU, one process per GPU) ------------
 * Each rank owns two gathered planes float32 [world * pkts_per_rank * n_rx*n_tx][d_out] (rank r's rows start at
 * r * pkts_per_rank * n_rx*n_tx).  gather_create allocates them (cudaMalloc, IPC-exportable) and returns their
 * device pointers; the host exchanges them (same process: raw pointers; other processes: mamimo_ipc_export /
 * mamimo_ipc_open) and passes every rank's pair to gather_connect (entry [rank] = this engine's own planes).
 * Then mamimo_estimate_stages(..., stages | MAMIMO_STAGE_GATHER) makes the final FC layer of each net store
 * its output tiles into all `world` planes from inside the kernel (TMA stores to peer memory over NVLink);
 * H_real / H_imag may be NULL.  Peers may read their plane after a cross-rank barrier following the call. */
#define MAMIMO_IPC_HANDLE_BYTES 64
MAMIMO_API mamimo_status mamimo_gather_create(mamimo_engine* e, int32_t world, int32_t rank, int64_t pkts_per_rank,
                                              float** real_plane, float** imag_plane);
MAMIMO_API mamimo_status mamimo_gather_connect(mamimo_engine* e, void* const* real_planes, void* const* imag_planes);
/* Same, over planes the CALLER allocated and mapped (e.g. a symmetric-memory allocator: cuMemCreate + peer mappings):
 * real_planes[r] / imag_planes[r] = rank r's plane as mapped in this process, entry [rank] this rank's own; the
 * engine never frees them.  mc_real / mc_imag: NVSwitch multicast addresses bound to all `world` planes
 * (cuMulticastCreate / cuMulticastBindMem), or NULL.  With multicast addresses the gather is a multimem.st stream on
 * a side stream -- every row leaves this GPU once and the switch replicates it -- unless MAMIMO_GATHER_MODE says
 * otherwise. */
MAMIMO_API mamimo_status mamimo_gather_attach(mamimo_engine* e, int32_t world, int32_t rank, int64_t pkts_per_rank,
                                              void* const* real_planes, void* const* imag_planes, void* mc_real,
                                              void* mc_imag);
MAMIMO_API mamimo_status mamimo_ipc_export(const void* dev_ptr, uint8_t handle[MAMIMO_IPC_HANDLE_BYTES]);
MAMIMO_API mamimo_status mamimo_ipc_open(const uint8_t handle[MAMIMO_IPC_HANDLE_BYTES], void** dev_ptr);
MAMIMO_API mamimo_status mamimo_ipc_close(void* dev_ptr);

/* Mode B: rows of caller-supplied planes (float32 [n_rows][d_in]) -> (float32 [n_rows][d_out]) x2.
 * This is what CSIPredictor.inference does with X.real / X.imag (inference.py:29-30). */
MAMIMO_API mamimo_status mamimo_predict_planes(mamimo_engine* e, const float* X_real, const float* X_imag,
                                               int64_t n_rows, float* Y_real, float* Y_imag,
                                               mamimo_mem mem, void* stream);

/* Mode A: time-domain preamble planes sig_real/sig_imag float32 [n_pkt][n_rx][len_ltf]; the engine
 * appends P(:,iTx) per pair (massiveMIMO_dataGenerator.py:307-311) and runs both nets. */
MAMIMO_API mamimo_status mamimo_predict_time(mamimo_engine* e, const float* sig_real, const float* sig_imag,
                                             int64_t n_pkt, float* Y_real, float* Y_imag,
                                             mamimo_mem mem, void* stream);

/* ---- next row: OFDM demodulation front-end ------------------
 * Configure: FFT length (power of two, <= 4096), cyclic prefix, symbol sampling offset (0..cp_len; the
 * reference passes cp_len) and the kept carriers (1-based indices into the fftshifted spectrum, i.e.
 * prm.CarriersLocations, n_sc of them).  Replaces the ofdmdemod call at pg/generate_maMIMO_LTF.m:336-338. */
MAMIMO_API mamimo_status mamimo_set_ofdm(mamimo_engine* e, int32_t fft_len, int32_t cp_len, int32_t sym_offset,
                                         const int32_t* carriers_1based);
/* x: time-domain samples, complex [n_pkt][n_rx][n_ltf*(fft_len+cp_len)] (MATLAB inputRXSig [lenLTF x Nr] per
 * packet) -> Y complex64 [n_pkt][n_rx][n_ltf][n_sc], the layout mamimo_ls_estimate / mamimo_estimate take. */
MAMIMO_API mamimo_status mamimo_ofdm_demod(mamimo_engine* e, const void* x, mamimo_ctype x_type, int64_t n_pkt,
                                           void* Y, mamimo_mem mem, void* stream);
/* Full path from the time domain: demod -> LS (+interp) -> both FC nets (H_real/H_imag may be NULL on an
 * engine without an MLP; H_ls complex64 may be NULL). */
MAMIMO_API mamimo_status mamimo_estimate_time(mamimo_engine* e, const void* x, mamimo_ctype x_type, int64_t n_pkt,
                                              void* H_ls, float* H_real, float* H_imag, mamimo_mem mem,
                                              void* stream);

/* ---- next row: LMMSE smoother ---------------------------------
 * Replaces the isMMSE branch of pg/helperMIMOChannelEstimate.m:37-39, i.e. LMMSE_ce(hD(:,j,i), Nsc, Nsc, Nps, tau,
 * SNR(i)) (pg/LMMSE_ce.m:23-39) for every pair of a batch: H_mmse = Rhp * inv(Rpp) * H_ls with
 * Rpp = 1./(1 + j*2*pi*tau_rms/Nsc*Nps*(k - k')) + eye/snr, Rhp = 1./(1 + j*2*pi*tau_rms/Nsc*(k - k'*Nps)).
 * tau_rms [n_pkt] (host) is what LMMSE_ce.m:27-30 derives from its `h` argument (mamimo_tau_rms restates it),
 * snr_db [n_pkt][n_rx] (host) is SNR(i) in dB.  H_ls / H_mmse complex [n_pkt][n_rx][n_tx][n_sc]; the pilot
 * spacing is the engine's n_ps.  One FP64 Cholesky solve per (packet, rx) with the n_tx pairs as right-hand sides.
 * MAMIMO_ERR_RANGE when a matrix is not positive definite in FP64 (device buffers: reported by mamimo_synchronize). */
MAMIMO_API mamimo_status mamimo_lmmse(mamimo_engine* e, const void* H_ls, mamimo_ctype h_type, int64_t n_pkt,
                                      const double* tau_rms, const double* snr_db, void* H_mmse,
                                      mamimo_ctype out_type, mamimo_mem mem, void* stream);
/* tau_rms of LMMSE_ce.m:27-30 for h [n] (real, or complex interleaved when is_complex) */
MAMIMO_API double mamimo_tau_rms(const double* h, int32_t n, int32_t is_complex);

/* ---- next row: per-subcarrier SVD of H-hat, the first step of the hybrid-precoder consumer -------
 * Replaces `H = Hin.'; [~,~,v] = svd(H); Fopt = v(:,1:Ns)` of pg/omphybweights.m:174-176 (per subcarrier, called from
 * pg/BER_test_maMIMO_LTF.m:372) for every (packet, tone) of a batch, restricted to what does not depend on LAPACK's
 * choice of basis: the n_rx singular values (descending) and the n_rx dominant right singular vectors of the
 * [n_rx x n_tx] matrix H(i,j) = H-hat[pkt][i][j][k] (each vector unique up to a phase; V1 V1^H is unique).
 * H complex [n_pkt][n_rx][n_tx][n_sc]; sigma real [n_pkt][n_rx][n_sc] (float for MAMIMO_C64 out_type, double for
 * MAMIMO_C128); V1 complex [n_pkt][n_rx][n_tx][n_sc] = v_r[j] of tone k at [r][j][k], may be NULL.  n_rx <= 8. */
MAMIMO_API mamimo_status mamimo_svd(mamimo_engine* e, const void* H, mamimo_ctype h_type, int64_t n_pkt, void* sigma,
                                    void* V1, mamimo_ctype out_type, mamimo_mem mem, void* stream);

/* ---- next row (continued): orthogonal matching pursuit over a steering dictionary -----------------
 * Replaces, for every (packet, tone) at once, the transmit side of getWeightsForSubcarrier after its SVD
 * (pg/omphybweights.m:178-179: `[Fbb,Frf] = ompdecomp(Fopt,At,'MaxSparsity',NtRF); Fbb = sqrt(Ns)*Fbb/norm(Frf*Fbb,'fro')`)
 * with ompdecomp.m:101-121's greedy loop (identity weight), as pg/BER_test_maMIMO_LTF.m:364-372 calls it: one
 * dictionary At for all subcarriers of the batch.  FP64 arithmetic throughout.
 *   set_steering_dictionary: At complex128 interleaved, HOST memory, [n_rays][n_tx] (row r = MATLAB's At(:, r+1)).
 *   omp: F = Fopt, complex [n_pkt][f_rows][n_tx][n_sc] with rows 0..ns-1 used -- mamimo_svd's V1 as it stands
 *        (f_rows = n_rx) or a tight tensor (f_rows = ns).  Outputs, tone index fastest like every tensor here:
 *        idx  int32 [n_pkt][n_rf][n_sc]   0-based dictionary row chosen in round j; -1 after an early stop
 *                                         (residual norm <= eps, ompdecomp.m:105).  Frf(k, j, :) = At row idx[j].
 *        err  float [n_pkt][n_rf][n_sc]   Frobenius norm of Fopt - atoms*coeff after round j (the last = Errnorm)
 *        Fbb  complex [n_pkt][ns][n_rf][n_sc] = the returned Fbb(k, s, j) (:196, scaled as :179); columns of rounds
 *                                         not run are zero.  Defined up to Fopt's own phase / unitary freedom.
 *   ns <= 8, n_rf <= 8, n_tx <= 100. */
MAMIMO_API mamimo_status mamimo_set_steering_dictionary(mamimo_engine* e, const double* At, int32_t n_rays);
MAMIMO_API mamimo_status mamimo_omp(mamimo_engine* e, const void* F, mamimo_ctype f_type, int32_t f_rows, int64_t n_pkt,
                                    int32_t ns, int32_t n_rf, int32_t* idx, float* err, void* Fbb,
                                    mamimo_ctype fbb_type, mamimo_mem mem, void* stream);

/* Error reporting of DEVICE-buffer calls.  Every entry point taking mem == MAMIMO_MEM_DEVICE only enqueues work on
 * `stream` and returns; conditions its kernels detect (MAMIMO_ERR_RANGE, MAMIMO_ERR_TIMEOUT, LMMSE not-PD) are
 * latched in a device flag word and reported -- and cleared -- by the next mamimo_synchronize (waits for the whole
 * device) or mamimo_poll_flags (waits for `stream` only).  HOST-buffer calls are synchronous and report directly. */
MAMIMO_API mamimo_status mamimo_synchronize(mamimo_engine* e);
MAMIMO_API mamimo_status mamimo_poll_flags(mamimo_engine* e, void* stream);
MAMIMO_API mamimo_status mamimo_get_stats(const mamimo_engine* e, mamimo_stats* out);
/* begin: bracket every engine kernel with a CUDA event pair on its stream; end: synchronise and sum them */
MAMIMO_API mamimo_status mamimo_profile_begin(mamimo_engine* e);
MAMIMO_API mamimo_status mamimo_profile_end(mamimo_engine* e, mamimo_profile* out);

/* Diagnostics (library built with -DMAMIMO_FC_DEBUG_COUNTERS, engine created with MAMIMO_FC_DEBUG=1 in the environment;
 * MAMIMO_ERR_STATE otherwise).  Sums over the CTA-pair FC kernel's cluster-launches since the last reset:
 * [0] producer cycles waiting for a free stage, [1] producer cycles total, [2] MMA-issuer cycles waiting for a drained
 * TMEM buffer, [3] waiting for operands, [4] MMA-issuer cycles total, [5] cluster-launches, [6] MMA-issuer wall time in
 * ns (globaltimer): [4] / [6] = the SM clock in GHz the kernel actually ran at, [7] unused. */
MAMIMO_API mamimo_status mamimo_get_debug_counters(mamimo_engine* e, uint64_t out[8], int32_t reset);

/* pinned host memory helpers for callers that want overlapped copies */
MAMIMO_API void* mamimo_host_alloc(size_t bytes);
MAMIMO_API void mamimo_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MAMIMO_H_ */
