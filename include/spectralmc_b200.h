/*
 * spectralmc_b200 — C ABI of the B200-native batch-generation hot path.
 *
 * One shared library (libspectralmc_b200.so), plain pointers and sizes only; no
 * torch / Python types.  The reference (Tuee22/SpectralMC) is 100 % Python and has no
 * FFI of its own: each entry point below replaces a Python call site of the reference,
 * cited as file:line under /root/reference/src/spectralmc/.  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add at each seam.
 *
 * Conventions
 *   - every function returns 0 on success, an SMC_E* code otherwise; the message is
 *     available from smc_last_error() (thread-local).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All
 *     work is stream-ordered; no function synchronises the host unless its name ends in
 *     `_host`.  Scratch space is a caller-provided workspace whose size comes from the matching
 *     *_workspace_bytes(); the library allocates device memory only in smc_p2p_alloc (peer
 *     exchange buffers have to come from cudaMalloc to be exportable over cudaIpc).
 *   - `dtype`: SMC_F32 / SMC_F64 (the reference's Precision.float32/float64,
 *     models/numerical.py:124-130).  Complex outputs are interleaved (re, im) pairs of the
 *     same width (complex64 / complex128).
 *   - matrices are C-contiguous (row-major), as the reference's CuPy arrays are.
 *   - contracts are rows of 6 doubles in the reference's field order X0, K, T, r, d, v
 *     (gbm.py:267-277).
 */
#ifndef SPECTRALMC_B200_H_
#define SPECTRALMC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 103: the float64 stream draws TWO Box-Muller pairs per Philox block (43-bit radius + 21-bit angle field per pair,
 *      oracle/philox.py); float64 results of 102 and earlier are a different sample set.  The float32 streams are unchanged. */
#define SMC_VERSION 103 /* 0.1.3 */

enum smc_status {
  SMC_OK = 0,
  SMC_EINVAL = 1,      /* bad argument (shape, dtype, enum, alignment)      */
  SMC_ECUDA = 2,       /* CUDA runtime / launch error (message has details) */
  SMC_EWORKSPACE = 3,  /* workspace too small                               */
  SMC_EUNSUPPORTED = 4 /* legal in the reference but not built here         */
};

enum smc_dtype { SMC_F32 = 0, SMC_F64 = 1 };
/* effects/montecarlo.py:24-35 */
enum smc_scheme {
  SMC_LOG_EULER = 0,   /* PathScheme.LOG_EULER  (gbm.py:245-250)                                  */
  SMC_SIMPLE_EULER = 1, /* PathScheme.SIMPLE_EULER (gbm.py:251-257)                                */
  /* fused path only: the same log-Euler mathematics with the exponential taken at EVERY step,
   * X *= exp(a + b z), exactly as the reference kernel iterates (gbm.py:249).  SMC_LOG_EULER
   * sums the log-returns and exponentiates once (prod exp(x_j) == exp(sum x_j)); this variant
   * exists so the two can be measured side by side. */
  SMC_LOG_EULER_STEPWISE = 2
};
enum smc_normalization { SMC_NORMALIZE = 0, SMC_RAW = 1 };
/* Which counter-based generator draws the normals (the stream is this library's own specification,
 * oracle/philox.py; element (i, j) of matrix k is a pure function of (seed, k, i, j, stream_version)).
 * PHILOX10 is the default everywhere and the stream every published number is measured on.  PHILOX7 is an
 * OPT-IN: the same construction with seven Philox rounds (the fewest that pass BigCrush, Salmon et al. SC'11;
 * it passes this repository's own battery, profiles/r2_philox7_evaluation.md) and about 13 % faster in the
 * fused float32 kernel.  It is a DIFFERENT sample set: never selected implicitly. */
enum smc_stream_version { SMC_STREAM_PHILOX10 = 0, SMC_STREAM_PHILOX7 = 1 };
/* how the CF estimate is formed from the [B, N] payoff matrix */
enum smc_cf_method {
  SMC_CF_MEAN_THEN_FFT = 0, /* FFT_n(mean_b mat): one length-N transform (linearity)          */
  SMC_CF_ROW_FFT = 1        /* mean_b FFT_n(mat[b,:]): a shared-memory/shuffle FFT per row,   */
                            /* fused with the batch mean (N a power of two, 32 <= N <= 512)   */
};

int smc_version(void);
const char* smc_last_error(void);
/* SM count / compute capability of the current device (fails loudly without a GPU). */
int smc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------
 * K1  normal supply — replaces
 *     cp.random.default_rng(seed).standard_normal((rows, cols), dtype)
 *     (async_normals.py:214-215; effects/interpreter.py:578-583).
 * Writes the `matrix_index`-th (rows, cols) standard-normal matrix of stream `seed`
 * (Philox4x32-10 + Box–Muller; the stream is specified in oracle/philox.py).  Element
 * (i, j) is a pure function of (seed, matrix_index, i, j).
 */
int smc_philox_normals(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed,
                       uint64_t matrix_index, void* stream);
/* the same with an explicit generator (smc_stream_version); smc_philox_normals is stream_version 0 */
int smc_philox_normals_v(void* out, int64_t rows, int64_t cols, int dtype, uint64_t seed,
                         uint64_t matrix_index, int stream_version, void* stream);

/* ---------------------------------------------------------------------------------------
 * K2  path kernel — replaces the Numba launch
 *     SimulateBlackScholes[blocks, threads, stream](io, timesteps, dt, X0, r, d, v, log_flag)
 *     (gbm.py:224-257, launched at gbm.py:413-426 and effects/interpreter.py:645-654).
 * `io` holds N(0,1) draws on entry, shape (rows = timesteps, cols = paths); on return
 * io[i, j] is path j after step i+1.  threads_per_block in {32,...,1024} is honoured as the
 * CTA size (gbm.py:71).
 */
int smc_gbm_paths_inplace(void* io, int64_t rows, int64_t cols, int dtype, double dt, double X0,
                          double r, double d, double v, int scheme, int threads_per_block,
                          void* stream);

/* Same stepping, read-only: consumes `normals` (rows, cols) and writes only the terminal
 * row to `terminal` (cols).  This is the 1x-read HBM-roofline form of K2 used when the
 * caller needs sims[-1] only (gbm.py:472). */
int smc_gbm_terminal_from_normals(const void* normals, int64_t rows, int64_t cols, int dtype,
                                  double dt, double X0, double r, double d, double v, int scheme,
                                  void* terminal, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4+K5  forward normalisation of every row — replaces
 *     row_means = cp.mean(sims, axis=1); sims *= expand_dims(forwards / row_means, 1)
 *     (gbm.py:437-438).  `forwards` is a device vector (rows) of `dtype`.
 */
size_t smc_normalize_rows_workspace_bytes(int64_t rows, int64_t cols);
int smc_normalize_rows(void* sims, int64_t rows, int64_t cols, int dtype, const void* forwards,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * K6  payoff — replaces gbm.py:467-474:
 *     put = df * max(K - terminal, 0); call = df * max(terminal - K, 0)  (arithmetic in dtype)
 * put / call may be NULL.
 */
int smc_payoff(const void* terminal, int64_t n, int dtype, double K, double df, void* put,
               void* call, void* stream);

/* K11  host-price reduction — replaces the three cp .mean() of gbm.py:496-498 with one
 * fused, deterministic reduction.  out3 = device double[3]: mean(underlying), mean(put),
 * mean(call).  Any input may be NULL (its slot is set to 0). */
size_t smc_means3_workspace_bytes(int64_t n);
int smc_means3(const void* underlying, const void* put, const void* call, int64_t n, int dtype,
               double* out3, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * K7+K8  CF estimate of a materialised payoff matrix — replaces
 *     cp.mean(cp.fft.fft(mat, axis=1), axis=0),  mat = put_price.reshape(B, N)
 *     (gbm_trainer.py:409-412, 814-817; effects/interpreter.py:703).
 * `mat` is (B, N) real of `dtype`; `out` is N complex of matching width.
 */
size_t smc_cf_fft_mean_workspace_bytes(int64_t batches, int64_t network_size, int method);
int smc_cf_fft_mean(const void* mat, int64_t batches, int64_t network_size, int dtype, int method,
                    void* out, void* workspace, size_t workspace_bytes, void* stream);

/* Per-row spectra WITHOUT the batch mean — replaces cp.fft.fft(tensor, axis=-1) of the ComputeFFT
 * effect (effects/interpreter.py:680-712) for a real (B, N) matrix; `out` is (B, N) complex of
 * matching width.  Powers of two in [32, 512] use the register/shuffle FFT of SMC_CF_ROW_FFT;
 * any other N <= 8192 a table-driven DFT (O(N^2) per row).  The training path never needs this
 * (it consumes the batch mean, smc_cf_fft_mean). */
int smc_fft_rows(const void* mat, int64_t batches, int64_t network_size, int dtype, void* out,
                 void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused batch path — replaces the whole per-contract Python loop
 *     [ _simulate_fft(c) for c in sobol_inputs ]  +  cp.asarray(fft_values)
 *     (gbm_trainer.py:1546-1553  ->  gbm_trainer.py:806-817  ->  gbm.py:450-488,400-447
 *      ->  async_normals.py:388-396)
 * with in-register Philox normals, log-Euler / simple-Euler stepping, payoff, column sums and
 * one FFT per contract.  Contract c consumes normal matrix `first_matrix_index + c`, exactly
 * as C successive BlackScholes.price() calls consume C matrices (gbm.py:405).
 *
 * Sharding (multi-GPU): a rank simulates batch rows [batch_begin, batch_end) of every
 * contract; Philox counters use the GLOBAL path index b*N + n, so the union over ranks equals
 * the single-GPU result up to summation order.  Outputs are PARTIAL: cf_out holds
 * (1/batches_total) * sum over the local rows, so a sum-allreduce over ranks gives the mean.
 *
 * smc_cf_fused            RAW, or NORMALIZE on one GPU (internally: terminal pass, mean, payoff
 *                          pass; terminals are staged in the workspace, contracts are
 *                          processed in chunks that fit it).
 * smc_fused_terminal      phase A of NORMALIZE for sharded runs: simulates and stores terminal
 *                          prices (C, P_local) and their per-contract LOCAL sums (double[C]).
 * smc_cf_from_terminal    phase B: payoff + CF from stored terminals, given the GLOBAL terminal
 *                          sums (after the caller's allreduce) — or NULL for RAW.
 */
typedef struct smc_fused_args {
  const double* contracts;   /* device, [n_contracts, 6]                                    */
  int64_t n_contracts;
  int64_t timesteps;
  int64_t network_size;      /* N                                                            */
  int64_t batches_total;     /* B of the whole job (all ranks)                               */
  int64_t batch_begin;       /* local rows [batch_begin, batch_end)                          */
  int64_t batch_end;
  int dtype;
  int scheme;
  int normalization;
  uint64_t seed;             /* SimulationParams.mc_seed (gbm.py:84)                         */
  uint64_t first_matrix_index; /* SimulationParams.skip (gbm.py:86) at the first contract     */
  int stream_version;        /* smc_stream_version; 0 (zero-initialised) = Philox4x32-10      */
} smc_fused_args;

size_t smc_cf_fused_workspace_bytes(const smc_fused_args* args);
/* number of kernels one smc_cf_fused call launches for these arguments (for launch accounting):
 * RAW is ONE kernel (tiles, reduction tree and transform); NORMALIZE two per chunk of contracts. */
int smc_cf_fused_launch_count(const smc_fused_args* args);
int smc_cf_fused(const smc_fused_args* args, void* cf_out /* [n_contracts, N] complex */,
                 void* workspace, size_t workspace_bytes, void* stream);
/* Introspection (no device access): how the simulation of `args` is cut into CTAs.
 * out = { tiles per contract, batch rows per tile, row lanes R, reduction-tree levels, fan-in of the root,
 *         main tiles (the first `main tiles` tiles have `batch rows per tile` rows; the remaining ones, the fine tail a
 *         single-contract launch ends on, have `tail rows per tile`), tail rows per tile }; capacity >= 7.  The cut depends on the problem shape only, so results never depend on the device the job
 * lands on. */
int smc_cf_fused_plan(const smc_fused_args* args, int64_t* out, int capacity);

size_t smc_fused_terminal_workspace_bytes(const smc_fused_args* args);
int smc_fused_terminal(const smc_fused_args* args, void* terminal /* [n_contracts, P_local] */,
                       double* terminal_sum /* device double[n_contracts] */, void* workspace,
                       size_t workspace_bytes, void* stream);

size_t smc_cf_from_terminal_workspace_bytes(const smc_fused_args* args);
int smc_cf_from_terminal(const smc_fused_args* args, const void* terminal,
                         const double* terminal_sum_global /* device double[C] or NULL */,
                         void* cf_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Batch-sharded RAW runs with the all-reduce FUSED into the finalise kernel over peer memory
 * (NVLink / NVSwitch) — the multi-GPU form of smc_cf_fused: same arguments and sharding
 * contract, but cf_out receives the COMPLETE [n_contracts, N] targets on every rank, with no
 * separate collective: each rank stores its float64 partial column sums straight into every
 * peer's exchange buffer, publishes a per-contract flag, waits for the peers' flags, sums the
 * `world` vectors in rank order (identical bits on every rank) and takes the one transform.
 *
 * Exchange buffers: one per rank, allocated by smc_p2p_alloc (cudaMalloc, zeroed, exported as a
 * 64-byte cudaIpc handle), opened by the other ranks with smc_p2p_open; all ranks must have
 * finished allocating and opening before the first call (a barrier of the caller's process group).
 * `epoch` must be > 0, equal on all ranks for one call and strictly increasing from call to call;
 * all ranks must make the same sequence of calls on one stream each.  Waiting for a peer is bounded
 * in time (`timeout_ms`): a peer that never arrives yields NaN in the affected targets and the epoch of
 * the call in this rank's status word (smc_p2p_status) — an error the host can report, with the CUDA
 * context intact, instead of a device trap that would take every rank down in turn.
 */
typedef struct smc_p2p_group {
  int rank;
  int world;                   /* <= 16 */
  void* buffers[16];           /* buffers[q]: rank q's exchange buffer as mapped in THIS process */
  int64_t capacity_contracts;  /* what the buffers were sized for (smc_p2p_buffer_bytes)        */
  int64_t network_size;
  uint32_t epoch;
  uint32_t timeout_ms;         /* how long a kernel waits for a peer; 0 = SMC_P2P_TIMEOUT_MS or two minutes */
} smc_p2p_group;

size_t smc_p2p_buffer_bytes(int64_t capacity_contracts, int64_t network_size, int world);
int smc_p2p_alloc(size_t bytes, void** ptr, void* handle64 /* out: 64 bytes */);
int smc_p2p_open(const void* handle64, void** ptr);
int smc_p2p_close(void* ptr);
int smc_p2p_free(void* ptr);
int smc_cf_fused_p2p(const smc_fused_args* args, const smc_p2p_group* group, void* cf_out,
                     void* workspace /* smc_cf_fused_workspace_bytes(args) */, size_t workspace_bytes,
                     void* stream);
/* Every host-side check smc_cf_fused_p2p makes, without touching the device: call it BEFORE advancing
 * the epoch, so that a failure on one rank (shape, workspace, unsupported N) cannot leave the ranks'
 * epochs out of step. */
int smc_cf_fused_p2p_check(const smc_fused_args* args, const smc_p2p_group* group, size_t workspace_bytes);
/* Epoch of the first call on this rank whose wait for a peer timed out (0 = none).  Synchronises `stream`. */
int smc_p2p_status(const smc_p2p_group* group, uint32_t* timed_out_epoch, void* stream);
/* NORMALIZE over several GPUs on the same buffers (one epoch covers both exchanges of a step):
 *   smc_fused_terminal  ->  smc_p2p_allreduce_sum_f64(terminal_sum, n_contracts)  ->  smc_cf_from_terminal_p2p
 * smc_p2p_allreduce_sum_f64 sums `count` (<= capacity_contracts) device doubles over the ranks in place, in rank
 * order (identical bits everywhere); smc_cf_from_terminal_p2p is smc_cf_from_terminal with the final exchange
 * fused into the finalise kernel (complete targets on every rank). */
int smc_p2p_allreduce_sum_f64(double* inout, int64_t count, const smc_p2p_group* group, void* stream);
int smc_cf_from_terminal_p2p(const smc_fused_args* args, const smc_p2p_group* group, const void* terminal,
                             const double* terminal_sum_global /* device double[C] or NULL */, void* cf_out,
                             void* workspace /* smc_cf_from_terminal_workspace_bytes(args) */,
                             size_t workspace_bytes, void* stream);

/* Host-buffer convenience for the reference-facing call: copies `contracts_host`
 * (ideally pinned) to the device, runs smc_cf_fused, copies the [n_contracts, N] complex
 * result to `cf_host` and synchronises `stream`.  args->contracts is ignored.  The workspace
 * must be smc_cf_fused_host_workspace_bytes() (device memory).  When `cf_host` is pinned
 * (cudaHostAlloc / cudaHostRegister) and at most 64 KiB, the finishing CTAs write it directly through
 * its device alias instead of a staging copy + D2H; pageable buffers take the copy.  The contracts
 * are always copied to the device (every CTA reads them). */
size_t smc_cf_fused_host_workspace_bytes(const smc_fused_args* args);
int smc_cf_fused_host(const smc_fused_args* args, const double* contracts_host, void* cf_host,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * CVNN step — the consumer of the [C, N] targets.  Replaces, for networks
 * that are a (possibly nested) ComplexSequential of ComplexLinear / modReLU / zReLU
 * (cvnn.py:65-146, 149-162, 168-210, 439-452), the op-by-op torch execution of
 *     GbmCVNNPricer._torch_step (gbm_trainer.py:819-835):
 *         pred = cvnn(real_in, imag_in)
 *         loss = mse(pred_r, Re targets) + mse(pred_i, Im targets)
 *         zero_grad(); loss.backward(); adam.step()         (optim.Adam, gbm_trainer.py:1513)
 * and the forward of predict_price (gbm_trainer.py:1722-1728), with a fixed sequence of
 * stream-ordered launches that never touches the host (CUDA-graph capturable; the Adam step
 * counter lives on the device).  ComplexLinear is ONE complex GEMM with the bias and the
 * following activation fused into its epilogue.  Unsupported layers (the batch norms, residual
 * wrappers) are reported by the caller not building a descriptor — those networks stay on torch.
 *
 * Parameters live in ONE flat buffer of the network dtype; each layer's block starts at
 * `param_offset` (elements) and is laid out in the order of torch's module.parameters():
 *     LINEAR : real_weight[out, in], imag_weight[out, in] (, real_bias[out], imag_bias[out])
 *     MODRELU: bias[features]            ZRELU: nothing
 * `grads`, `exp_avg`, `exp_avg_sq` are flat buffers of the same length and layout.
 * Activations are row-major planes (real, imag) of shape [rows, features]; `targets` is
 * [rows, out] interleaved complex of the network dtype.  `loss` is a DEVICE double.
 */
enum smc_cvnn_layer_kind { SMC_LAYER_LINEAR = 0, SMC_LAYER_MODRELU = 1, SMC_LAYER_ZRELU = 2 };

typedef struct smc_cvnn_layer {
  int kind;
  int has_bias;          /* LINEAR only                                                      */
  int64_t in_features;   /* LINEAR: inputs; MODRELU: features                                */
  int64_t out_features;  /* LINEAR only                                                      */
  int64_t param_offset;  /* first element of this layer's parameters in the flat buffer     */
} smc_cvnn_layer;

typedef struct smc_cvnn_net {
  const smc_cvnn_layer* layers; /* host array, flattened in execution order                  */
  int n_layers;
  int dtype;
  int64_t n_inputs;             /* width of (real_in, imag_in)                               */
  int64_t n_params;             /* length of the flat parameter buffer                       */
} smc_cvnn_net;

/* torch.optim.Adam hyper-parameters (no weight decay, no amsgrad — the trainer's defaults) */
typedef struct smc_adam_args {
  double lr, beta1, beta2, eps;
} smc_adam_args;

size_t smc_cvnn_workspace_bytes(const smc_cvnn_net* net, int64_t rows, int training);
/* width of the network output, or -1 if the descriptor is inconsistent */
int64_t smc_cvnn_output_width(const smc_cvnn_net* net);
int smc_cvnn_forward(const smc_cvnn_net* net, const void* params, const void* in_r, const void* in_i,
                     int64_t rows, void* out_r, void* out_i, void* workspace, size_t workspace_bytes,
                     void* stream);
/* forward + loss + backward: writes every element of `grads` and the scalar *loss */
int smc_cvnn_loss_backward(const smc_cvnn_net* net, const void* params, const void* in_r,
                           const void* in_i, const void* targets, int64_t rows, void* grads,
                           double* loss, void* workspace, size_t workspace_bytes, void* stream);
/* one Adam update of n elements; *step (device int64, steps taken so far) is incremented */
int smc_adam_step(void* params, const void* grads, void* exp_avg, void* exp_avg_sq, int64_t n,
                  int dtype, int64_t* step, const smc_adam_args* hyper, void* stream);
/* smc_cvnn_loss_backward followed by smc_adam_step over the whole parameter buffer */
int smc_cvnn_train_step(const smc_cvnn_net* net, void* params, void* grads, void* exp_avg,
                        void* exp_avg_sq, int64_t* step, const smc_adam_args* hyper, const void* in_r,
                        const void* in_i, const void* targets, int64_t rows, double* loss,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Audit of the float32 normal stream on the device, with no matrix in HBM (the stream is this library's own
 * specification, oracle/philox.py, so the library ships the means to test it at 2^33 draws and beyond;
 * tests/test_gpu_stream_battery.py).  Block b of the audit is the Philox block of (column b mod cols, row group
 * b div cols) of matrix `matrix_index`.
 *   radius_hist / angle_hist   device uint32[2^21] each, ZEROED by the caller: counts of the 21-bit fields
 *   tails4                     device uint64[4], zeroed: normals with |z| > 4, 5, 5.5, 6 (refined entries included)
 *   power_sums4                device double[4], zeroed: sum z, z^2, z^3, z^4 over the 6 * n_blocks normals
 * smc_diag_stream_lags_f32: sums7 (device double[7], zeroed) receives sum_i z[i,j] z[i-lag,j] for lag = 1..6 over
 * all columns (rows a multiple of 6: products within AND across the 6-row blocks) and, in sums7[6], the products of
 * horizontally adjacent entries z[i,j] z[i,j+1] (pairs inside one warp of 32 columns). */
int smc_diag_stream_fields_f32(uint64_t seed, uint64_t matrix_index, uint64_t n_blocks, uint32_t cols,
                               uint32_t* radius_hist, uint32_t* angle_hist, uint64_t* tails4, double* power_sums4,
                               int stream_version, void* stream);
int smc_diag_stream_lags_f32(uint64_t seed, uint64_t matrix_index, uint32_t cols, uint32_t rows, double* sums7,
                             int stream_version, void* stream);

/* Pipe-peak calibration microbenchmarks (FP32 FMA issue and MUFU/XU), used by bench.py to
 * state the compute roofline on the box it runs on: each runs `iters` dependent-chain
 * iterations per thread on a full grid and returns executed lane-operations in *ops.
 * kind: 0 = FFMA, 1 = MUFU.EX2, 2 = IMAD.WIDE.U32 + LOP3 (Philox-like mix), 3 = DFMA (the FP64 pipe). */
int smc_pipe_calibrate(int kind, int64_t iters, double* ops, float* sink /* device, >= 1 */,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRALMC_B200_H_ */
