// Fused batch path and CF estimate: in-register Philox normals -> GBM stepping -> payoff ->
// column sums over batches -> one FFT per contract, as ONE kernel per step.
//
// Replaces, per training step, the reference's Python loop
//   [ _simulate_fft(c) for c in sobol_inputs ] + cp.asarray(fft_values)
//   (/root/reference/src/spectralmc/gbm_trainer.py:1546-1553, :806-817; gbm.py:400-488;
//    async_normals.py:388-396)
// and, for materialised inputs, cp.mean(cp.fft.fft(mat, axis=1), axis=0) (gbm_trainer.py:814-817).
//
// Structure (all reductions fixed-order, no float atomics => bit-reproducible):
//   step_kernel      one CTA per (contract, tile of batch rows), handed out by the hardware block
//                    scheduler.  The first CTAs to reach a contract form its constants and publish
//                    them in a table; nobody waits for them.  Thread (r, col) owns column `col` of the
//                    [B, N] payoff matrix and walks rows r, r+R, ... of its tile, accumulating in
//                    float64; the CTA folds the R row-lanes in shared memory and writes one partial
//                    column-sum vector per tile.  (Short paths, timesteps <= 3: grouped_short_tile.)
//   ticket tree      the tile vectors of a contract are the leaves of a radix-16 tree.  A CTA that
//                    completes a vector takes a ticket (atomic counter) on its parent; whoever takes
//                    the LAST ticket of a node folds that node's children in index order (so the sum
//                    never depends on which CTA, or how many SMs, did the work) and moves up.
//   finish_contract  the CTA that completes the root scales by 1/B and takes ONE length-N transform
//                    (mean_b FFT_n(mat) == FFT_n(mean_b mat)) in float64 shared memory (radix-2 for
//                    powers of two, table-driven DFT otherwise) and narrows to the output width —
//                    or, for batch-sharded multi-GPU runs, pushes the real vector into every peer's
//                    exchange buffer; a small collect kernel then sums the peers' vectors in rank order.
//   Large N (transform working set > 16 KiB of shared memory) keeps the transform in a separate
//   one-CTA-per-contract kernel (cf_finalize_kernel / cf_exchange_finalize_kernel).
// Tile and tree shapes are functions of the problem shape only (never of the SM count), and a
// tile's vector does not depend on which CTA computed it, so results do not depend on the device
// the job lands on or on the order in which CTAs run.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "smc_internal.h"  // first: fixes SMC_NS / the Philox round count of this build of the file
#include "smc_device.cuh"

namespace SMC_NS {
using namespace ::smc;  // shared helpers (smc_internal.h); everything that draws normals lives in SMC_NS

constexpr int CF_BLOCK = 256;
constexpr int64_t TARGET_TILES = 12288;  // CTAs per launch aimed at (profiles/r2_codegen_variant_matrix.txt: 4096..24576 within 1 %)
constexpr int64_t STREAM_TILES = 2368;  // 16 x 148
constexpr int64_t TAIL_TILES = 2048;    // single-contract launches end on this many smallest-size tiles (about two waves)
constexpr int64_t TARGET_TILES_FINE_TAIL = 3072;  // ... which lets their main tiles be four times larger (profiles/r2_codegen_variant_matrix.txt)
constexpr int64_t MIN_TILE_PATH_STEPS = 16384;  // a simulated tile is at least 64 path-steps per thread
constexpr int TREE_RADIX = 16;
constexpr int MAX_LEVELS = 8;           // 16^8 > 2^31 tiles
constexpr size_t FUSED_FINALIZE_SMEM_MAX = 16 * 1024;  // transform working set that may live in the step kernel
constexpr int SCHEME_LOG_STEPWISE = 2;  // SMC_LOG_EULER_STEPWISE
constexpr int MAX_PEERS = 16;

#ifndef SMC_TAIL_NOINLINE
#define SMC_TAIL_NOINLINE 0  // codegen knob: the cold per-tile tail (tickets, folds, transform) as out-of-line functions; inlined measured 0.5-1 % faster at c2 (profiles/r2_codegen_variant_matrix.txt)
#endif
#if SMC_TAIL_NOINLINE
#define SMC_COLD __noinline__
#else
#define SMC_COLD __forceinline__
#endif

enum Source { SRC_FUSED = 0, SRC_TERMINAL = 1, SRC_MATRIX = 2 };
enum Output { OUT_COLSUM = 0, OUT_TERMINAL = 1 };
enum Finish { FINISH_NONE = 0, FINISH_TRANSFORM = 1, FINISH_EXCHANGE = 2 };

// Radix-16 reduction tree over the tile vectors of ONE contract.  Level 0 are the tiles; level l has
// count[l] = ceil(count[l-1] / 16) nodes; the root folds the count[levels] <= 16 vectors of the top level.
struct TreePlan {
  int levels;                     // intermediate levels (0 when tiles <= 16)
  int64_t count[MAX_LEVELS + 1];
  int64_t off[MAX_LEVELS + 1];    // first node of level l >= 1 within a contract's node storage
  int64_t nodes;                  // intermediate nodes per contract
};

static TreePlan make_tree(int64_t tiles) {
  TreePlan t{};
  t.count[0] = tiles;
  while (t.count[t.levels] > TREE_RADIX && t.levels < MAX_LEVELS) {
    const int l = t.levels + 1;
    t.count[l] = (t.count[l - 1] + TREE_RADIX - 1) / TREE_RADIX;
    t.off[l] = t.nodes;
    t.nodes += t.count[l];
    t.levels = l;
  }
  return t;
}

struct TilePlan {
  int chunk_w;          // columns covered per pass (min(N, 256))
  int lanes_r;          // R: row lanes per pass (256 / chunk_w)
  int64_t tile_rows;    // multiple of R
  int64_t tiles;        // tiles per contract
  // simulated tiles: the LAST rows of a contract are cut finer (tail_tile_rows = R, one pass per thread), so that the
  // launch ends on small CTAs — the block scheduler hands tiles out in index order — instead of a ragged wave of big ones
  int64_t main_tiles;   // tiles [0, main_tiles) have tile_rows rows and cover rows [0, main_rows)
  int64_t main_rows;
  int64_t tail_tile_rows;
  TreePlan tree;
};

static int64_t env_int(const char* name, int64_t fallback) {
  const char* e = std::getenv(name);
  const long long v = e ? std::atoll(e) : 0;
  return v > 0 ? static_cast<int64_t>(v) : fallback;
}

// `streaming`: the tile's source is an HBM-resident matrix (payoffs / staged terminals), so tiles
// are sized for bandwidth (>= 64 KiB of input each, a few thousand CTAs) instead of for
// load-balancing a compute-bound simulation.  Simulated tiles (`timesteps` > 0) are sized for about
// TARGET_TILES CTAs per launch (TARGET_TILES_FINE_TAIL main tiles + a fine tail when the launch holds one
// contract) but never below MIN_TILE_PATH_STEPS path-steps, so that short paths (the reference's own tests
// run one timestep) do not drown in per-tile bookkeeping.
// One tile = one CTA, handed out by the hardware block scheduler in index order.  (A persistent grid drawing tiles
// from a device counter was measured and rejected: the warp scheduler is not fair between resident
// CTAs — in one 1.4 ms launch some CTAs completed 53 tiles and others 4 — so the last tiles of starved
// CTAs stretched the tail by 100-400 us; freshly launched CTAs rotate through the priorities instead.
// profiles/r2_persistent_grid_experiment.md.)
static TilePlan make_plan(int64_t n_contracts, int64_t rows_local, int64_t n, bool streaming = false,
                          int64_t timesteps = 0) {
  TilePlan p{};
  p.chunk_w = static_cast<int>(std::min<int64_t>(n, CF_BLOCK));
  p.lanes_r = CF_BLOCK / p.chunk_w;
  const int64_t R = p.lanes_r;
  static const int64_t target_env = env_int("SMC_TARGET_TILES", 0);  // tuning knobs
  static const int64_t tail_tiles = env_int("SMC_TAIL_TILES", TAIL_TILES + 1) - 1;  // SMC_TAIL_TILES=1 switches the fine tail off
  // Fine tail (single-contract simulated launches): the last TAIL_TILES smallest-size tiles' worth of rows (at most a quarter of
  // the rows) are cut into tiles of the smallest size, so the launch ends on small CTAs — the block scheduler hands tiles out in
  // index order — and the MAIN tiles can be large (less per-CTA bookkeeping) without a ragged last wave of big ones:
  // 1.300 -> 1.282 ms at config c2.  Shape-only, like everything else here.
  const bool simulated = !streaming && timesteps > 0;
  const int64_t smallest = !simulated ? R : std::max<int64_t>(R, ((MIN_TILE_PATH_STEPS + n * timesteps - 1) / (n * timesteps) + R - 1) / R * R);
  const bool fine_tail = simulated && tail_tiles > 0 && n_contracts == 1 && rows_local / 4 >= smallest;
  const int64_t target = streaming ? STREAM_TILES : (target_env > 0 ? target_env : (fine_tail ? TARGET_TILES_FINE_TAIL : TARGET_TILES));
  int64_t want = (n_contracts * rows_local + target - 1) / target;
  if (streaming) want = std::max<int64_t>(want, (16384 + n - 1) / n);  // >= 16 Ki elements per tile
  if (simulated) want = std::max<int64_t>(want, smallest);
  want = std::max<int64_t>(want, 1);
  p.tile_rows = (want + R - 1) / R * R;
  p.tile_rows = std::min<int64_t>(p.tile_rows, (rows_local + R - 1) / R * R);
  p.main_rows = rows_local;
  p.tail_tile_rows = p.tile_rows;
  if (fine_tail && p.tile_rows > smallest) {
    const int64_t tail_rows = std::min<int64_t>(tail_tiles * smallest, rows_local / 4) / smallest * smallest;
    if (tail_rows > 0) {
      p.main_rows = rows_local - tail_rows;
      p.tail_tile_rows = smallest;
    }
  }
  p.main_tiles = (p.main_rows + p.tile_rows - 1) / p.tile_rows;
  p.tiles = p.main_tiles + (rows_local - p.main_rows + p.tail_tile_rows - 1) / p.tail_tile_rows;
  p.tree = make_tree(p.tiles);
  return p;
}

// Exchange buffers of the peer-memory all-reduce (see the section further down).
struct PeerExchange {
  double* data[MAX_PEERS];     // exchange buffer of rank p as mapped into this process
  int rank, world;
  unsigned epoch;              // > 0, the same on every rank for one call, increasing
  int64_t capacity_contracts;  // contracts the buffers were sized for
  unsigned long long timeout_ns;  // a peer that does not arrive within this time is reported, not waited for
};

// Everything one launch of the step kernel needs (passed by value: constant bank).
struct TileParams {
  const double* contracts;      // [*, 6], indexed by GLOBAL contract
  int64_t contract0;            // first global contract handled by this launch
  int64_t timesteps;
  int64_t n;                    // network_size
  int64_t batches_total;
  int64_t row_begin, row_end;   // local batch rows
  int64_t paths_local;          // (row_end - row_begin) * n
  int64_t tile_rows, tiles;
  int64_t main_tiles, main_rows, tail_tile_rows;  // fine tail (TilePlan)
  int chunk_w, lanes_r;
  int chunk_shift;              // log2(chunk_w) when it is a power of two, else -1
  int scheme;
  int64_t launch_contracts;     // contracts covered by this launch
  int normalize;                // apply scale[c] = F / mean before the payoff
  PhiloxKeys keys;
  uint64_t first_matrix_index;
  const void* terminal_in;      // [launch contracts, paths_local]   (SRC_TERMINAL)
  const void* matrix;           // [rows, n]                         (SRC_MATRIX)
  const double* terminal_sum;   // GLOBAL sums per launch contract   (normalize)
  void* terminal_out;           // [launch contracts, paths_local]   (OUT_TERMINAL)
  double* partial;              // [launch contracts, tiles, n]      (OUT_COLSUM)
  double* term_partial;         // [launch contracts, tiles]         (OUT_TERMINAL)
  double* terminal_sum_out;     // [launch contracts]                (OUT_TERMINAL): local sums
  // per-contract constants: computed by the first CTAs that need them, then shared through this table
  void* consts;                 // SimConsts<Real>[launch contracts]
  unsigned* consts_ready;       // [launch contracts], zeroed before the launch
  // reduction tree
  unsigned* tickets;            // OUT_COLSUM: [launch contracts, tree.nodes + 1]; OUT_TERMINAL: [launch contracts]
  double* nodes;                // [launch contracts, tree.nodes, n]
  TreePlan tree;
  // what the CTA that completes a contract's root does
  int finish;                   // Finish
  int fft_mode, log2n;          // see transform_store
  double scale;                 // 1 / batches_total
  void* out;                    // [*, n] complex
  int64_t out_contract0;
  PeerExchange px;              // FINISH_EXCHANGE
};

// per-contract constants, derived in float64 and narrowed once
template <typename Real>
struct SimConsts {
  Real X0, K, df, scale;
  Real lin0, lin1;  // log-Euler: X_T = X0 * exp(lin0 + lin1 * sum z)   [log2 units for float32]
                    // stepwise   : X *= exp(lin0 + lin1 z) per step
                    // simple     : X = |X + X * (lin0 + lin1 z)|,  lin0 = (r - d) dt
};

template <typename Real>
__device__ __forceinline__ SimConsts<Real> make_consts(const TileParams& p, const int SCHEME, int64_t c_global,
                                                       int64_t c_local) {
  const ContractRow k = load_contract(p.contracts, c_global);
  const double dt = k.T / static_cast<double>(p.timesteps);  // gbm.py:411
  const double sdt = sqrt(dt);                               // gbm.py:243
  const double unit = sizeof(Real) == 4 ? 1.4426950408889634074 : 1.0;  // float32 feeds MUFU.EX2
  SimConsts<Real> s;
  s.X0 = static_cast<Real>(k.X0);
  s.K = static_cast<Real>(k.K);                  // gbm.py:467
  s.df = static_cast<Real>(exp(-k.r * k.T));     // df[-1], gbm.py:431,466
  if (SCHEME == SMC_LOG_EULER) {
    const double drift_dt = (k.r - k.d - 0.5 * k.v * k.v) * dt;  // gbm.py:246,249
    s.lin0 = static_cast<Real>(drift_dt * static_cast<double>(p.timesteps) * unit);
    s.lin1 = static_cast<Real>(k.v * sdt * unit);
  } else if (SCHEME == SCHEME_LOG_STEPWISE) {
    s.lin0 = static_cast<Real>((k.r - k.d - 0.5 * k.v * k.v) * dt * unit);
    s.lin1 = static_cast<Real>(k.v * sdt * unit);
  } else {
    // X += X * (lin0 + lin1 z): the increment is formed at full relative precision; folding the 1
    // into lin0 would round (r - d) dt to an ulp of 1.0 and bias every step the same way.
    s.lin0 = static_cast<Real>((k.r - k.d) * dt);  // gbm.py:252,255
    s.lin1 = static_cast<Real>(k.v * sdt);
  }
  s.scale = Real(1);
  if (p.normalize) {
    // forwards[-1] / row_means[-1], gbm.py:430,437-438 (ratio formed in the engine dtype)
    const Real fwd = static_cast<Real>(k.X0 * exp((k.r - k.d) * k.T));
    const double paths_total = static_cast<double>(p.batches_total) * static_cast<double>(p.n);
    const Real mean = static_cast<Real>(p.terminal_sum[c_local] / paths_total);
    s.scale = fwd / mean;
  }
  return s;
}

// Per-contract constants (float64 exp/sqrt/divide, out of line so the hot kernel's register allocation
// does not see them) are formed by ONE thread of whichever CTAs reach a contract first and shared with
// every later CTA through a table in the workspace: `ready[c]` (zeroed before the launch) is set with
// release semantics once table[c] is written.  A CTA that finds the flag clear does not wait for another
// CTA — its own thread 0 computes the same values and hands them to the CTA through shared memory — so
// nothing depends on the order in which CTAs are scheduled.  The CTA decides once (barrier + vote): threads
// of one CTA may read different flag values, and a thread that read "clear" is ordered behind the publisher
// through the barrier with a thread that acquired "set".
// (Every thread computing the constants itself was measured first: with four tiles per contract, as at the
// reference's test size, all four CTAs start before the first publishes, and 8 warps x ~500 instructions of
// float64 code per CTA were half of the short-path kernel's time.)
template <typename Real>
static __device__ __noinline__ void compute_consts(const TileParams& p, int64_t c_local, SimConsts<Real>* out) {
  *out = make_consts<Real>(p, p.scheme, p.contract0 + c_local, c_local);
}

__device__ __forceinline__ unsigned load_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void store_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// `slot`: shared memory for one SimConsts<Real>, free again on return
template <typename Real>
__device__ __forceinline__ SimConsts<Real> contract_consts(const TileParams& p, int64_t c_local, SimConsts<Real>* slot) {
  Real* table = reinterpret_cast<Real*>(static_cast<SimConsts<Real>*>(p.consts) + c_local);
  SimConsts<Real> k;
  if (__syncthreads_or(load_acquire_gpu(p.consts_ready + c_local) != 0u)) {
    k.X0 = __ldcg(table + 0); k.K = __ldcg(table + 1); k.df = __ldcg(table + 2); k.scale = __ldcg(table + 3);
    k.lin0 = __ldcg(table + 4); k.lin1 = __ldcg(table + 5);
    return k;
  }
  if (threadIdx.x == 0) {
    compute_consts<Real>(p, c_local, slot);
    table[0] = slot->X0; table[1] = slot->K; table[2] = slot->df; table[3] = slot->scale; table[4] = slot->lin0; table[5] = slot->lin1;
    store_release_gpu(p.consts_ready + c_local, 1u);
  }
  __syncthreads();
  k = *slot;
  __syncthreads();  // the slot is scratch the caller reuses
  return k;
}

template <typename Real, int SCHEME>
__device__ __forceinline__ void consume(Real& acc, Real z, const SimConsts<Real>& k) {
  if (SCHEME == SMC_LOG_EULER) {
    acc += z;
  } else if (SCHEME == SCHEME_LOG_STEPWISE) {
    if (sizeof(Real) == 4)
      acc *= mufu_ex2(fmaf(static_cast<float>(k.lin1), static_cast<float>(z), static_cast<float>(k.lin0)));
    else
      acc *= exp(fma(static_cast<double>(k.lin1), static_cast<double>(z), static_cast<double>(k.lin0)));
  } else {
    if (sizeof(Real) == 4)
      acc = static_cast<Real>(fabsf(fmaf(static_cast<float>(acc),
                                         fmaf(static_cast<float>(k.lin1), static_cast<float>(z), static_cast<float>(k.lin0)),
                                         static_cast<float>(acc))));
    else
      acc = static_cast<Real>(fabs(fma(static_cast<double>(acc),
                                       fma(static_cast<double>(k.lin1), static_cast<double>(z), static_cast<double>(k.lin0)),
                                       static_cast<double>(acc))));
  }
}

// terminal price of global path `col` of the matrix (k_lo, k_hi): every one of the `timesteps`
// normals is drawn and consumed.
// RAGGED = false is the specialisation for timesteps % 6 == 0: the hot loop's throughput depends
// on ptxas' interleaving of IMAD.WIDE / MUFU / LOP3 issue, and merely compiling the tail code into
// the same kernel costs 2 % at config c2 (A/B in one run: 1.387 vs 1.414 ms), so kernels that
// cannot have a tail do not contain one.
template <int SCHEME, bool REFINE, bool RAGGED>
__device__ __forceinline__ float simulate_path_f32(const SimConsts<float>& k, uint32_t col, int64_t timesteps,
                                                   const PhiloxKeys& keys, uint32_t k_lo, uint32_t k_hi,
                                                   uint32_t& min_word) {
  float acc = SCHEME == SMC_LOG_EULER ? 0.0f : k.X0;
  const uint32_t nq = static_cast<uint32_t>(timesteps / 6);
#ifndef SMC_F32_UNROLL
#define SMC_F32_UNROLL 2  // codegen knob, see profiles/r1_codegen_variant_matrix.txt
#endif
  constexpr int kUnroll = SMC_F32_UNROLL;
#ifndef SMC_F32X2
#define SMC_F32X2 1  // packed FADD2/FMUL2/FFMA2 Box-Muller over two row groups at a time (log-Euler sum only): 163 vs 181 issue slots per 12 normals
#endif
  uint32_t q = 0;
  if (SMC_F32X2 && SCHEME == SMC_LOG_EULER && !REFINE) {
    float2 acc2 = make_float2(0.0f, 0.0f);
#ifndef SMC_F32X2_UNROLL
#define SMC_F32X2_UNROLL 1
#endif
    constexpr int kUnrollX2 = SMC_F32X2_UNROLL;
#pragma unroll kUnrollX2
    for (; q + 1 < nq; q += 2) acc2 = normals12_sum_f32x2(col, q, k_lo, k_hi, keys, acc2, min_word);
    acc = (acc2.x + acc2.y) * 1.41421356237309505f;  // normals12_sum_f32x2 accumulates (z_even + z_odd) / sqrt(2)
  }
#pragma unroll kUnroll
  for (; q < nq; ++q) {
    float z[6];
    normals6_f32_impl<REFINE>(col, q, k_lo, k_hi, keys, z, min_word);
#ifndef SMC_SUM_TREE
#define SMC_SUM_TREE 0
#endif
    if (SMC_SUM_TREE && SCHEME == SMC_LOG_EULER) {
      acc += ((z[0] + z[1]) + (z[2] + z[3])) + (z[4] + z[5]);
    } else {
#pragma unroll
      for (int u = 0; u < 6; ++u) consume<float, SCHEME>(acc, z[u], k);
    }
  }
  if (RAGGED && timesteps <= 3) {
    // short layout, one path at a time (the general form, for shapes the grouped tile does not cover): the
    // path's normals are numbers lane * T .. lane * T + T - 1 of the block its group of G = 6 / T columns shares
    const uint32_t T = static_cast<uint32_t>(timesteps), G = 6u / T;
    const uint32_t g = col / G, first = (col - g * G) * T;
    float z[6];
    normals6_f32_impl<REFINE, 3>(g, F32_SHORT_BIT, k_lo, k_hi, keys, z, min_word);
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i) {
      if (i < T) {
        const uint32_t idx = first + i;
        const float zi = idx == 0 ? z[0] : idx == 1 ? z[1] : idx == 2 ? z[2] : idx == 3 ? z[3] : idx == 4 ? z[4] : z[5];
        consume<float, SCHEME>(acc, zi, k);
      }
    }
  } else if (RAGGED) {  // last block: evaluate only the pairs that are consumed
    const int rem = static_cast<int>(timesteps - static_cast<int64_t>(nq) * 6);
    if (rem) {
      float z[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (rem <= 2)
        normals6_f32_impl<REFINE, 1>(col, nq, k_lo, k_hi, keys, z, min_word);
      else if (rem <= 4)
        normals6_f32_impl<REFINE, 2>(col, nq, k_lo, k_hi, keys, z, min_word);
      else
        normals6_f32_impl<REFINE, 3>(col, nq, k_lo, k_hi, keys, z, min_word);
#pragma unroll
      for (int u = 0; u < 5; ++u)
        if (u < rem) consume<float, SCHEME>(acc, z[u], k);
    }
  }
  if (SCHEME == SMC_LOG_EULER) return k.X0 * mufu_ex2(fmaf(k.lin1, acc, k.lin0));
  return acc;
}

// the rare re-simulation (some block of the path had a zero radius field): same path with the
// refinement applied.  Arguments by value so the caller keeps its key block in the constant bank.
template <int SCHEME, bool RAGGED>
static __device__ __noinline__ float simulate_path_exact_f32(float X0, float lin0, float lin1, uint32_t col,
                                                             int64_t timesteps, uint32_t seed_lo, uint32_t seed_hi,
                                                             uint32_t k_lo, uint32_t k_hi) {
  PhiloxKeys keys;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    keys.k0[r] = seed_lo + static_cast<uint32_t>(r) * PHILOX_W0;
    keys.k1[r] = seed_hi + static_cast<uint32_t>(r) * PHILOX_W1;
  }
  SimConsts<float> k{};
  k.X0 = X0;
  k.lin0 = lin0;
  k.lin1 = lin1;
  uint32_t unused = 0;
  return simulate_path_f32<SCHEME, true, RAGGED>(k, col, timesteps, keys, k_lo, k_hi, unused);
}

template <int SCHEME, bool RAGGED>
__device__ __forceinline__ float simulate_terminal(const SimConsts<float>& k, uint32_t col, int64_t timesteps,
                                                   const PhiloxKeys& keys, uint32_t k_lo, uint32_t k_hi) {
#ifndef SMC_F32_POSTHOC_REFINE
#define SMC_F32_POSTHOC_REFINE 1
#endif
  uint32_t min_word = 0xffffffffu;
#if SMC_F32_POSTHOC_REFINE
  float v = simulate_path_f32<SCHEME, false, RAGGED>(k, col, timesteps, keys, k_lo, k_hi, min_word);
  if (__builtin_expect(min_word < 2048u, 0))
    v = simulate_path_exact_f32<SCHEME, RAGGED>(k.X0, k.lin0, k.lin1, col, timesteps, keys.k0[0], keys.k1[0], k_lo, k_hi);
  return v;
#else
  return simulate_path_f32<SCHEME, true, RAGGED>(k, col, timesteps, keys, k_lo, k_hi, min_word);
#endif
}

template <int SCHEME, bool RAGGED>
__device__ __forceinline__ double simulate_terminal(const SimConsts<double>& k, uint32_t col, int64_t timesteps,
                                                    const PhiloxKeys& keys, uint32_t k_lo, uint32_t k_hi) {
  double acc = SCHEME == SMC_LOG_EULER ? 0.0 : k.X0;
  const uint32_t nq = static_cast<uint32_t>(timesteps >> 2);  // whole blocks: four normals (two pairs) each
#ifndef SMC_F64_UNROLL
#define SMC_F64_UNROLL 1
#endif
  constexpr int kUnroll = SMC_F64_UNROLL;
#pragma unroll kUnroll
  for (uint32_t q = 0; q < nq; ++q) {
    if (SCHEME == SMC_LOG_EULER) {
      acc = normals4_sum_f64(col, q, k_lo, k_hi, keys, acc);
    } else {
      double z[4];
      normals4_f64(col, q, k_lo, k_hi, keys, z);
#pragma unroll
      for (int u = 0; u < 4; ++u) consume<double, SCHEME>(acc, z[u], k);
    }
  }
  if (SCHEME == SMC_LOG_EULER) acc *= 1.41421356237309514547;  // normals4_sum_f64 accumulates (z0 + z1) / sqrt(2)
  const int rem = static_cast<int>(timesteps & 3);
  if (rem) {
    double z[4];
    normals4_f64(col, nq, k_lo, k_hi, keys, z, rem > 2 ? 2 : 1);
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (u < rem) consume<double, SCHEME>(acc, z[u], k);
  }
  if (SCHEME == SMC_LOG_EULER) return k.X0 * exp(fma(k.lin1, acc, k.lin0));
  return acc;
}

__device__ __forceinline__ unsigned bit_reverse(unsigned x, int bits) { return __brev(x) >> (32 - bits); }

// exp(-2 pi i j / n) for j < ntw into (twr, twi)
__device__ __forceinline__ void fill_twiddles(double* twr, double* twi, int64_t ntw, int64_t n) {
  const double kTwoOverN = 2.0 / static_cast<double>(n);
  for (int64_t j = threadIdx.x; j < ntw; j += CF_BLOCK) {
    double sn, cs;
    sincospi(-static_cast<double>(j) * kTwoOverN, &sn, &cs);
    twr[j] = cs;
    twi[j] = sn;
  }
}

// Length-n forward transform of the real vector in shared memory and narrowing store.  mode 0: `re`
// holds the input in bit-reversed order, `im` zeros (radix-2, in place); mode 1: `re` holds the input
// in natural order (table-driven DFT).  All threads of the CTA; the input must be visible (synced).
template <typename Real>
__device__ __forceinline__ void transform_store(double* re, double* im, const double* twr, const double* twi,
                                                int64_t n, int mode, Real* __restrict__ dst) {
  if (mode == 0) {
    for (int64_t half = 1; half < n; half <<= 1) {
      const int64_t stride = n / (2 * half);
      for (int64_t b = threadIdx.x; b < n / 2; b += CF_BLOCK) {
        const int64_t j = b & (half - 1);
        const int64_t i0 = ((b - j) << 1) + j, i1 = i0 + half;
        const double wr = twr[j * stride], wi = twi[j * stride];
        const double tr = wr * re[i1] - wi * im[i1];
        const double ti = wr * im[i1] + wi * re[i1];
        re[i1] = re[i0] - tr;
        im[i1] = im[i0] - ti;
        re[i0] += tr;
        im[i0] += ti;
      }
      __syncthreads();
    }
    for (int64_t kk = threadIdx.x; kk < n; kk += CF_BLOCK) {
      dst[2 * kk] = static_cast<Real>(re[kk]);
      dst[2 * kk + 1] = static_cast<Real>(im[kk]);
    }
  } else {
    for (int64_t kk = threadIdx.x; kk < n; kk += CF_BLOCK) {
      double ar = 0.0, ai = 0.0;
      int64_t m = 0;
      for (int64_t j = 0; j < n; ++j) {
        ar = fma(re[j], twr[m], ar);
        ai = fma(re[j], twi[m], ai);
        m += kk;
        if (m >= n) m -= n;
      }
      dst[2 * kk] = static_cast<Real>(ar);
      dst[2 * kk + 1] = static_cast<Real>(ai);
    }
  }
}

// s += v[0] + v[stride] + ... (count terms), in index order, with eight loads in flight: these
// second-level reductions read L2-resident vectors written by other CTAs (hence the L2-only loads)
// and are bound by load latency, not bandwidth.
__device__ __forceinline__ double strided_sum(const double* __restrict__ v, int64_t count, int64_t stride) {
  double s = 0.0;
  int64_t i = 0;
  for (; i + 8 <= count; i += 8) {
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = __ldcg(v + (i + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u];
  }
  for (; i < count; ++i) s += __ldcg(v + i * stride);
  return s;
}

// Column-wise sum of `count` length-n vectors (src, src + n, ...), in a fixed order, by the whole CTA;
// emit(col, sum) is called once per column by one thread.  When n divides 128 the CTA's 256 threads
// split into `subs` lanes per column, each summing every subs-th vector; the lanes are folded in
// shared memory in lane order.  `sm` holds CF_BLOCK doubles and is free again on return.
template <typename Emit>
__device__ __forceinline__ void fold_vectors(const double* __restrict__ src, int64_t count, int64_t n, double* sm,
                                             Emit emit) {
  if (n <= CF_BLOCK / 2 && CF_BLOCK % n == 0) {
    const int subs = CF_BLOCK / static_cast<int>(n);
    const int sub = threadIdx.x / static_cast<int>(n), col = threadIdx.x - sub * static_cast<int>(n);
    const int64_t mine = sub < count ? (count - sub + subs - 1) / subs : 0;  // vectors sub, sub + subs, ...
    sm[threadIdx.x] = strided_sum(src + static_cast<int64_t>(sub) * n + col, mine, static_cast<int64_t>(subs) * n);
    __syncthreads();
    if (sub == 0) {
      double tot = 0.0;
      for (int k = 0; k < subs; ++k) tot += sm[k * n + col];
      emit(static_cast<int64_t>(col), tot);
    }
    __syncthreads();
    return;
  }
  for (int64_t col = threadIdx.x; col < n; col += CF_BLOCK) emit(col, strided_sum(src + col, count, n));
}

__device__ __forceinline__ void store_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned load_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Exchange buffer of one rank, in 8-byte cells:
//   [2 slots][world senders][capacity contracts][n]   partial column sums          (exchange-finalise)
//   [2 slots][world senders][capacity contracts]      their per-contract flags
//   [2 slots][world senders][capacity contracts]      one double per contract      (small all-reduce: terminal sums)
//   [2 slots][world senders]                          its per-sender flags
//   [1]                                               status: epoch of the first call that timed out waiting for a peer
__host__ __device__ inline size_t exchange_data_doubles(int64_t capacity_contracts, int64_t n, int world) {
  return static_cast<size_t>(2) * world * capacity_contracts * n;
}
__host__ __device__ inline size_t exchange_small_base(int64_t capacity_contracts, int64_t n, int world) {
  return exchange_data_doubles(capacity_contracts, n, world) + static_cast<size_t>(2) * world * capacity_contracts;
}
__host__ __device__ inline size_t exchange_status_cell(int64_t capacity_contracts, int64_t n, int world) {
  return exchange_small_base(capacity_contracts, n, world) + static_cast<size_t>(2) * world * capacity_contracts +
         static_cast<size_t>(2) * world;
}
__host__ __device__ inline size_t exchange_total_cells(int64_t capacity_contracts, int64_t n, int world) {
  return exchange_status_cell(capacity_contracts, n, world) + 1;
}

// Wait until `flag` (in this rank's own exchange buffer) carries `epoch`.  Returns false when the peer
// did not arrive within px.timeout_ns: the caller then poisons its output with NaN and the epoch is
// recorded in the buffer's status cell (smc_p2p_status) — a reportable error instead of a trap that
// would kill the context of every rank in turn.
__device__ __forceinline__ bool wait_for_flag(const unsigned* flag, const PeerExchange& px, size_t status_cell) {
  if (load_acquire_sys(flag) == px.epoch) return true;
  const unsigned long long t0 = global_timer_ns();
  for (unsigned polls = 1;; ++polls) {
    if (load_acquire_sys(flag) == px.epoch) return true;
    if ((polls & 1023u) == 0u && global_timer_ns() - t0 > px.timeout_ns) {
      atomicCAS(reinterpret_cast<unsigned*>(px.data[px.rank] + status_cell), 0u, px.epoch);
      return false;
    }
  }
}

// ---- what the CTA holding the last ticket of a contract does ---------------------------------
// Shared memory (doubles) behind `dyn`: re[n], im[n], then the twiddle table (n/2 pairs for radix-2,
// n pairs for the table DFT).
template <typename Real>
static __device__ SMC_COLD void finish_contract(const TileParams& p, int64_t c_local, const double* top,
                                                    int64_t top_count, double* sm, double* dyn) {
  const int64_t n = p.n;
  if (p.finish == FINISH_TRANSFORM) {
    double* re = dyn;
    double* im = dyn + n;
    double* twr = dyn + 2 * n;
    double* twi = twr + (p.fft_mode == 0 ? n / 2 : n);
    fill_twiddles(twr, twi, p.fft_mode == 0 ? n / 2 : n, n);
    fold_vectors(top, top_count, n, sm, [&](int64_t col, double v) {
      int64_t where = col;
      if (p.fft_mode == 0 && p.log2n > 0) where = static_cast<int64_t>(bit_reverse(static_cast<unsigned>(col), p.log2n));
      re[where] = v * p.scale;
      im[where] = 0.0;
    });
    __syncthreads();
    transform_store<Real>(re, im, twr, twi, n, p.fft_mode, static_cast<Real*>(p.out) + (p.out_contract0 + c_local) * n * 2);
    __syncthreads();  // re / im may be reused by this CTA's next contract
  } else if (p.finish == FINISH_EXCHANGE) {
    // phase 1 of the peer exchange: store the scaled real vector into every peer's buffer, then publish
    const PeerExchange& px = p.px;
    const size_t cell = ((static_cast<size_t>(px.epoch & 1u) * px.world + px.rank) * px.capacity_contracts + c_local);
    fold_vectors(top, top_count, n, sm, [&](int64_t col, double v) {
      const double x = v * p.scale;
      for (int q = 0; q < px.world; ++q) px.data[q][cell * n + col] = x;
    });
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < px.world)
      store_release_sys(reinterpret_cast<unsigned*>(px.data[threadIdx.x] + exchange_data_doubles(px.capacity_contracts, n, px.world) + cell),
                        px.epoch);
  }
  // FINISH_NONE: the top-level vectors stay where they are for a separate finalise kernel
}

// Called by every thread of a CTA after it has written tile vector `tile` of contract `c_local`:
// walks up the ticket tree for as long as this CTA completes nodes.
template <typename Real>
static __device__ SMC_COLD void tile_done_colsum(const TileParams& p, int64_t c_local, int64_t tile, double* sm,
                                                     double* dyn) {
  __shared__ int s_last;
  unsigned* tick = p.tickets + c_local * (p.tree.nodes + 1);
  double* nodes = p.nodes + c_local * p.tree.nodes * p.n;
  const double* level_src = p.partial + c_local * p.tiles * p.n;  // vectors of the current level
  int64_t node = tile;
  for (int level = 0;; ++level) {
    const bool top = level == p.tree.levels;
    const int64_t parent = top ? 0 : node / TREE_RADIX;
    const int64_t first = parent * TREE_RADIX;
    const int64_t expected = top ? p.tree.count[level] : min(static_cast<int64_t>(TREE_RADIX), p.tree.count[level] - first);
    unsigned* t = top ? tick + p.tree.nodes : tick + p.tree.off[level + 1] + parent;
    // release pattern of a grid barrier: the CTA's writes are ordered before the barrier, thread 0's fence
    // is cumulative over what it has observed, so the vector is visible device-wide before the ticket is
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const bool last = atomicAdd(t, 1u) == static_cast<unsigned>(expected - 1);
      if (last) __threadfence();  // acquire side: the other CTAs' vectors are read next
      s_last = last;
    }
    __syncthreads();
    if (!s_last) return;
    if (top) {
      finish_contract<Real>(p, c_local, level_src, expected, sm, dyn);
      return;
    }
    double* dst = nodes + (p.tree.off[level + 1] + parent) * p.n;
    fold_vectors(level_src + first * p.n, expected, p.n, sm, [&](int64_t col, double v) { dst[col] = v; });
    level_src = nodes + p.tree.off[level + 1] * p.n;
    node = parent;
  }
}

// OUT_TERMINAL: the CTA that stores the last tile sum of a contract folds them (fixed order)
static __device__ SMC_COLD void tile_done_terminal(const TileParams& p, int64_t c_local, double tile_total, int64_t tile,
                                                       double* sm) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    p.term_partial[c_local * p.tiles + tile] = tile_total;
    __threadfence();
    s_last = atomicAdd(p.tickets + c_local, 1u) == static_cast<unsigned>(p.tiles - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double s = 0.0;
  if (threadIdx.x < p.tiles)  // this thread's tiles: threadIdx.x, threadIdx.x + 256, ... (eight loads in flight)
    s = strided_sum(p.term_partial + c_local * p.tiles + threadIdx.x, (p.tiles - threadIdx.x + CF_BLOCK - 1) / CF_BLOCK, CF_BLOCK);
  const double tot = block_sum(s, sm);
  if (threadIdx.x == 0) p.terminal_sum_out[c_local] = tot;
  __syncthreads();
}

// phase 2 of the peer exchange (after this rank has pushed everything it will push): wait for every
// rank's vector of a contract, sum in rank order, transform.
template <typename Real>
static __device__ __noinline__ void exchange_collect(const TileParams& p, double* dyn) {
  __shared__ int s_ok;
  const PeerExchange& px = p.px;
  const int64_t n = p.n, cap = px.capacity_contracts;
  const int64_t slot = px.epoch & 1u;
  const size_t flag_base = exchange_data_doubles(cap, n, px.world);
  const size_t status_cell = exchange_status_cell(cap, n, px.world);
  double* re = dyn;
  double* im = dyn + n;
  double* twr = dyn + 2 * n;
  double* twi = twr + (p.fft_mode == 0 ? n / 2 : n);
  fill_twiddles(twr, twi, p.fft_mode == 0 ? n / 2 : n, n);
  const double* mine = px.data[px.rank];
  for (int64_t c = blockIdx.x; c < p.launch_contracts; c += gridDim.x) {
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < px.world) {
      const unsigned* flag = reinterpret_cast<const unsigned*>(mine + flag_base + ((slot * px.world + threadIdx.x) * cap + c));
      if (!wait_for_flag(flag, px, status_cell)) s_ok = 0;
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    for (int64_t col = threadIdx.x; col < n; col += CF_BLOCK) {
      double s = 0.0;
      for (int q = 0; q < px.world; ++q) s += __ldcv(mine + ((slot * px.world + q) * cap + c) * n + col);
      int64_t where = col;
      if (p.fft_mode == 0 && p.log2n > 0) where = static_cast<int64_t>(bit_reverse(static_cast<unsigned>(col), p.log2n));
      re[where] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
      im[where] = 0.0;
    }
    __syncthreads();
    transform_store<Real>(re, im, twr, twi, n, p.fft_mode, static_cast<Real*>(p.out) + (p.out_contract0 + c) * n * 2);
    __syncthreads();  // re / im are reused by the next contract of this CTA
  }
}

// Short paths (T = timesteps <= 3, float32): G = 6 / T adjacent paths share one Philox block, so a
// thread draws one block and finishes G paths with it — every normal drawn is consumed, where the general
// form would spend a whole block (and three Box-Muller pairs) per path.  Thread t of the CTA walks groups
// g_first + t, g_first + t + 256, ... of the tile's path range and keeps one float64 accumulator per lane
// of the group; the host takes this form only when 256 G is a multiple of N, so that lane u of thread t
// stays in ONE column, (g_first G + t G + u) mod N, for the whole tile.  The 256 G accumulators are then
// folded to the N column sums of the tile in shared memory, in a fixed order.
#ifndef SMC_SHORT_POSTHOC
#define SMC_SHORT_POSTHOC 1
#endif
#ifndef SMC_SHORT_PIN
#define SMC_SHORT_PIN 1
#endif
// the rare group with a zero radius field (6e-6 of the blocks): the same block with the refinement applied, out of line and
// by value, so that the common path carries neither the test-and-call of the refinement nor the moves that set the call up
struct Normals6 {
  float z0, z1, z2, z3, z4, z5;
};
static __device__ __noinline__ Normals6 short_group_exact(uint32_t g, uint32_t k_lo, uint32_t k_hi, uint32_t seed_lo, uint32_t seed_hi) {
  PhiloxKeys keys;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    keys.k0[r] = seed_lo + static_cast<uint32_t>(r) * PHILOX_W0;
    keys.k1[r] = seed_hi + static_cast<uint32_t>(r) * PHILOX_W1;
  }
  float z[6];
  uint32_t unused = 0;
  normals6_f32_impl<true, 3>(g, F32_SHORT_BIT, k_lo, k_hi, keys, z, unused);
  return Normals6{z[0], z[1], z[2], z[3], z[4], z[5]};
}

template <int SCHEME, int T>
__device__ __forceinline__ void grouped_short_tile(const TileParams& p, const SimConsts<float>& k, uint32_t k_lo, uint32_t k_hi,
                                                   int64_t row0, int64_t row1, double* __restrict__ dst, double* sm,
                                                   double* lanes /* shared, 256 G doubles */) {
  constexpr int G = 6 / T;
  const int64_t p0 = row0 * p.n, p1 = row1 * p.n;  // the tile's global paths [p0, p1), p1 <= 2^32 - 1
  const int64_t g_first = p0 / G, g_end = (p1 + G - 1) / G;
  // groups [g_in0, g_in1) lie wholly inside the tile and need no per-path range check; at most the first and the
  // last group of a tile are cut by its boundary (a cut group is drawn by both tiles, each keeping its own lanes)
  const uint32_t g_in0 = static_cast<uint32_t>((p0 + G - 1) / G), g_in1 = static_cast<uint32_t>(p1 / G);
  // Payoffs are summed in float32 for at most SHORT_FLUSH consecutive groups per lane and then folded into the
  // float64 accumulator: the float32 -> float64 conversion is an XU-pipe instruction, and at one timestep this
  // kernel is XU-bound (12 MUFU of Box-Muller + 6 MUFU.EX2 per block; six conversions per block were a quarter of
  // the pipe's work).  Eight non-negative terms add a relative error of at most 7 * 2^-24 to a run — the size of
  // the float32 payoff's own rounding — and the order is fixed, so bit-reproducibility is untouched.
  constexpr int SHORT_FLUSH = 8;
  double acc[G];
  float run[G];
#pragma unroll
  for (int u = 0; u < G; ++u) acc[u] = 0.0, run[u] = 0.0f;
  // lin1 and -X0 multiply a per-path value in FFMAs whose addend is a uniform register too; an FFMA takes one uniform operand,
  // and ptxas re-copies the other into a vector register at EVERY use (12 copies per block).  Pinning the two in vector registers
  // once per tile removes the copies.
  float lin1 = k.lin1, neg_x0 = -k.X0;
#if SMC_SHORT_PIN
  // (an identity shuffle: ptxas allocates uniform registers itself and sees through anything that is not a real per-lane instruction)
  lin1 = __shfl_sync(0xffffffffu, lin1, threadIdx.x & 31);
  neg_x0 = __shfl_sync(0xffffffffu, neg_x0, threadIdx.x & 31);
#endif
  auto put_of = [&](const float (&z)[6], int u) {
    float state = SCHEME == SMC_LOG_EULER ? z[u * T] : k.X0;  // log-Euler: the sum starts AT the first normal (0 + z is not folded: -0)
#pragma unroll
    for (int i = SCHEME == SMC_LOG_EULER ? 1 : 0; i < T; ++i) consume<float, SCHEME>(state, z[u * T + i], k);
    const float diff = SCHEME == SMC_LOG_EULER ? fmaf(mufu_ex2(fmaf(lin1, state, k.lin0)), neg_x0, k.K) : k.K - state;
    return k.df * (diff > 0.0f ? diff : 0.0f);  // gbm.py:473
  };
  const uint32_t g_stop = static_cast<uint32_t>(g_end);  // g_end <= 2^32 / G + 1
  auto add_group = [&](uint32_t g, const float (&z)[6]) {
    if (g >= g_in0 && g < g_in1) {
#pragma unroll
      for (int u = 0; u < G; ++u) run[u] += put_of(z, u);
    } else {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int64_t path = static_cast<int64_t>(g) * G + u;
        if (path >= p0 && path < p1) run[u] += put_of(z, u);
      }
    }
  };
  auto flush = [&]() {
#pragma unroll
    for (int u = 0; u < G; ++u) acc[u] += static_cast<double>(run[u]), run[u] = 0.0f;
  };
  // (drawing two groups per iteration with packed Box-Muller pairs, as the long-path loop does, was measured: no gain —
  // this loop is bound by its instruction count, profiles/r2_short_path_notes.md)
  int pending = 0;
  uint32_t g = static_cast<uint32_t>(g_first) + threadIdx.x;
  for (; g < g_stop; g += CF_BLOCK) {
    float z[6];
#if SMC_SHORT_POSTHOC
    uint32_t min_word = 0xffffffffu;
    normals6_f32_impl<false, 3>(g, F32_SHORT_BIT, k_lo, k_hi, p.keys, z, min_word);
    if (__builtin_expect(min_word < 2048u, 0)) {
      const Normals6 e = short_group_exact(g, k_lo, k_hi, p.keys.k0[0], p.keys.k1[0]);
      z[0] = e.z0, z[1] = e.z1, z[2] = e.z2, z[3] = e.z3, z[4] = e.z4, z[5] = e.z5;
    }
#else
    uint32_t unused = 0;
    normals6_f32_impl<true, 3>(g, F32_SHORT_BIT, k_lo, k_hi, p.keys, z, unused);
#endif
    add_group(g, z);
    if (++pending >= SHORT_FLUSH) {
      pending = 0;
      flush();
    }
  }
#pragma unroll
  for (int u = 0; u < G; ++u) acc[u] += static_cast<double>(run[u]);
#pragma unroll
  for (int u = 0; u < G; ++u) lanes[threadIdx.x * G + u] = acc[u];
  __syncthreads();
  // entry e = t G + u belongs to column (o + e) mod N, o = (g_first G) mod N; column `col` owns entries
  // e0, e0 + N, ... (count per column = 256 G / N).  `subs` threads share a column, then fold in order.
  // (32-bit index arithmetic: this runs once per CTA, and at the reference's test size a CTA is short)
  const uint32_t n = static_cast<uint32_t>(p.n);
  const uint32_t o = (static_cast<uint32_t>(g_first % n) * G) % n;
  const uint32_t per_col = static_cast<uint32_t>(CF_BLOCK) * G / n;
  if (n <= CF_BLOCK) {
    const uint32_t subs = CF_BLOCK / n;
    const uint32_t sub = threadIdx.x / n, col = threadIdx.x - sub * n;
    double s = 0.0;
    if (sub < subs) {
      const uint32_t e0 = col >= o ? col - o : col + n - o;
      for (uint32_t m = sub; m < per_col; m += subs) s += lanes[e0 + m * n];
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    if (sub == 0) {
      double tot = 0.0;
      for (uint32_t q = 0; q < subs; ++q) tot += sm[q * n + col];
      dst[col] = tot;
    }
    __syncthreads();
  } else {
    for (uint32_t col = threadIdx.x; col < n; col += CF_BLOCK) {
      const uint32_t e0 = col >= o ? col - o : col + n - o;
      double s = 0.0;
      for (uint32_t m = 0; m < per_col; ++m) s += lanes[e0 + m * n];
      dst[col] = s;
    }
    __syncthreads();
  }
}

#ifndef SMC_F64_FUSED_MIN_CTAS
#define SMC_F64_FUSED_MIN_CTAS 3  // 80 registers: the coefficient pairs of the float64 loop stay in registers (no LDC in the loop); -2.7 % against 4
#endif
#ifndef SMC_F32_FUSED_MIN_CTAS_OTHER
#define SMC_F32_FUSED_MIN_CTAS_OTHER 5  // simple-Euler / stepwise / terminal-staging instantiations: 5 measured best
#endif
#ifndef SMC_F32_SHORT_MIN_CTAS
#define SMC_F32_SHORT_MIN_CTAS 6  // short-path grouped tile: 6 CTAs x 40 registers measured best of {3,4,5,6} (profiles/r2_short_path_notes.md)
#endif
#ifndef SMC_F32_FUSED_MIN_CTAS_TERMINAL
#define SMC_F32_FUSED_MIN_CTAS_TERMINAL 4  // log-Euler with staged terminals (NORMALIZE pass A): 1.360 ms at c2 vs 1.445 at 5
#endif
#ifndef SMC_F32_FUSED_MIN_CTAS
#define SMC_F32_FUSED_MIN_CTAS 6  // round 2 (ticket tail in the kernel): 6 CTAs x 40 registers measured best, 1.309 ms at c2 vs 1.348 at 5 and 1.365 at 4 (profiles/r2_codegen_variant_matrix.txt); round 1 (separate tail kernels) had 5 best
#endif
// float64 fused instantiations are capped (4 CTAs = 32 warps per SM): uncapped they take 90
// registers and run 2 CTAs per SM
// FORM (float32 fused kernels): 0 = timesteps a multiple of 6, no tail code in the kernel; 1 = general;
// 2 = short paths (timesteps <= 3) in the grouped layout, one Philox block per G = 6 / timesteps paths.
template <typename Real, int SRC, int SCHEME, int OUT, int FORM = 1>
__global__ void __launch_bounds__(CF_BLOCK, SRC != SRC_FUSED ? 0 : (sizeof(Real) == 8 ? SMC_F64_FUSED_MIN_CTAS : (FORM == 2 ? SMC_F32_SHORT_MIN_CTAS : (SCHEME == SMC_LOG_EULER ? (OUT == OUT_COLSUM ? SMC_F32_FUSED_MIN_CTAS : SMC_F32_FUSED_MIN_CTAS_TERMINAL) : SMC_F32_FUSED_MIN_CTAS_OTHER))))
    step_kernel(const __grid_constant__ TileParams p) {
  extern __shared__ double dyn[];
  __shared__ double sm[CF_BLOCK];
  const int64_t c_local = blockIdx.y + static_cast<int64_t>(blockIdx.z) * 65535;
  if (c_local >= p.launch_contracts) return;
  const int64_t tile = blockIdx.x;
  const int64_t c_global = p.contract0 + c_local;
  const bool in_tail = tile >= p.main_tiles;
  const int64_t row0 = p.row_begin + (in_tail ? p.main_rows + (tile - p.main_tiles) * p.tail_tile_rows : tile * p.tile_rows);
  const int64_t row1 = in_tail ? min(row0 + p.tail_tile_rows, p.row_end) : min(row0 + p.tile_rows, p.row_begin + p.main_rows);
  constexpr bool RAGGED = FORM != 0;

  SimConsts<Real> k{};
  if (SRC != SRC_MATRIX) k = contract_consts<Real>(p, c_local, reinterpret_cast<SimConsts<Real>*>(sm));
  const uint64_t mi = p.first_matrix_index + static_cast<uint64_t>(c_global);
  const uint32_t k_lo = static_cast<uint32_t>(mi), k_hi = static_cast<uint32_t>(mi >> 32);

  if constexpr (FORM == 2 && SRC == SRC_FUSED && OUT == OUT_COLSUM && sizeof(Real) == 4) {
    __shared__ double lanes[CF_BLOCK * 6];
    double* dst = p.partial + (c_local * p.tiles + tile) * p.n;
    if (p.timesteps == 1) grouped_short_tile<SCHEME, 1>(p, k, k_lo, k_hi, row0, row1, dst, sm, lanes);
    else if (p.timesteps == 2) grouped_short_tile<SCHEME, 2>(p, k, k_lo, k_hi, row0, row1, dst, sm, lanes);
    else grouped_short_tile<SCHEME, 3>(p, k, k_lo, k_hi, row0, row1, dst, sm, lanes);
    tile_done_colsum<Real>(p, c_local, tile, sm, dyn);
    return;
  }

  const int r = p.chunk_shift >= 0 ? static_cast<int>(threadIdx.x >> p.chunk_shift) : static_cast<int>(threadIdx.x) / p.chunk_w;
  const int lc = threadIdx.x - r * p.chunk_w;
  const int64_t local_base = c_local * p.paths_local - p.row_begin * p.n;  // + global path -> staging index
  double tile_total = 0.0;

  for (int64_t n0 = 0; n0 < p.n; n0 += p.chunk_w) {
    const int64_t col = n0 + lc;
    const bool active = (r < p.lanes_r) && (col < p.n);
    double acc = 0.0;
    if (active) {
      for (int64_t row = row0 + r; row < row1; row += p.lanes_r) {
        const int64_t path = row * p.n + col;  // global path index b*N + n (gbm_trainer.py:814-816)
        Real val;
        if (SRC == SRC_FUSED)
          val = simulate_terminal<SCHEME, RAGGED>(k, static_cast<uint32_t>(path), p.timesteps, p.keys, k_lo, k_hi);
        else if (SRC == SRC_TERMINAL)
          val = __ldcs(static_cast<const Real*>(p.terminal_in) + local_base + path);
        else
          val = __ldcs(static_cast<const Real*>(p.matrix) + (path - p.row_begin * p.n));
        if (OUT == OUT_TERMINAL) {
          static_cast<Real*>(p.terminal_out)[local_base + path] = val;
          acc += static_cast<double>(val);
        } else if (SRC == SRC_MATRIX) {
          acc += static_cast<double>(val);
        } else {
          if (p.normalize) val *= k.scale;                              // gbm.py:438
          const Real diff = k.K - val;
          const Real put = k.df * (diff > Real(0) ? diff : Real(0));    // gbm.py:473
          acc += static_cast<double>(put);
        }
      }
    }
    if (OUT == OUT_COLSUM) {
      sm[threadIdx.x] = acc;
      __syncthreads();
      if (r == 0 && col < p.n) {
        double s = 0.0;
        for (int rr = 0; rr < p.lanes_r; ++rr) s += sm[rr * p.chunk_w + lc];
        p.partial[(c_local * p.tiles + tile) * p.n + col] = s;
      }
      __syncthreads();
    } else {
      tile_total += acc;
    }
  }
  if (OUT == OUT_COLSUM) {
    tile_done_colsum<Real>(p, c_local, tile, sm, dyn);
  } else {
    const double t = block_sum(tile_total, sm);
    tile_done_terminal(p, c_local, t, tile, sm);
  }
}

// Phase 2 of the peer exchange as a kernel of its own (persistent, co-resident), launched right after a
// step kernel whose finishing CTAs pushed (FINISH_EXCHANGE).
template <typename Real>
__global__ void __launch_bounds__(CF_BLOCK) exchange_collect_kernel(const __grid_constant__ TileParams p) {
  extern __shared__ double dyn[];
  exchange_collect<Real>(p, dyn);
}

// Separate finalise kernel for transforms whose working set does not fit beside the step kernel's
// CTAs.  One CTA per contract: fold the `groups` top-level vectors (contract c's start at
// vecs + c * contract_stride), scale, transform, narrow.
// Shared memory (doubles): re[n], im[n], then the twiddle table (n/2 pairs for radix-2, n pairs for
// the DFT).  mode: 0 radix-2 (n power of two), 1 table DFT, 2 DFT with on-the-fly twiddles and the
// folded vector staged in `spill` (global) for n too large for shared memory.
template <typename Real>
__global__ void __launch_bounds__(CF_BLOCK)
    cf_finalize_kernel(const double* __restrict__ vecs, int64_t contract_stride, int64_t groups, int64_t n, double scale,
                       int mode, int log2n, Real* __restrict__ out /* [contracts, n, 2] */, int64_t out_contract0,
                       double* __restrict__ spill) {
  extern __shared__ double smem[];
  const int64_t c = blockIdx.x;
  const double* src = vecs + c * contract_stride;
  Real* dst = out + (out_contract0 + c) * n * 2;

  if (mode == 2) {
    const double kTwoOverN = 2.0 / static_cast<double>(n);
    double* x = spill + c * n;
    for (int64_t col = threadIdx.x; col < n; col += CF_BLOCK) x[col] = strided_sum(src + col, groups, n) * scale;
    __syncthreads();  // x is written and read by this CTA only
    for (int64_t kk = threadIdx.x; kk < n; kk += CF_BLOCK) {
      double re = 0.0, im = 0.0;
      int64_t m = 0;
      for (int64_t j = 0; j < n; ++j) {
        double sn, cs;
        sincospi(-static_cast<double>(m) * kTwoOverN, &sn, &cs);
        re = fma(x[j], cs, re);
        im = fma(x[j], sn, im);
        m += kk;
        if (m >= n) m -= n;
      }
      dst[2 * kk] = static_cast<Real>(re);
      dst[2 * kk + 1] = static_cast<Real>(im);
    }
    return;
  }

  double* re = smem;
  double* im = smem + n;
  double* twr = smem + 2 * n;
  double* twi = twr + (mode == 0 ? n / 2 : n);
  fill_twiddles(twr, twi, mode == 0 ? n / 2 : n, n);
  for (int64_t col = threadIdx.x; col < n; col += CF_BLOCK) {
    const double s = strided_sum(src + col, groups, n);
    int64_t where = col;
    if (mode == 0 && log2n > 0) where = static_cast<int64_t>(bit_reverse(static_cast<unsigned>(col), log2n));
    re[where] = s * scale;
    im[where] = 0.0;
  }
  __syncthreads();
  transform_store<Real>(re, im, twr, twi, n, mode, dst);
}

// ---- all-reduce over peer memory (NVLink / NVSwitch) fused into the step -----------------------
// Multi-GPU form of the contract finish for batch-sharded runs: instead of transforming its local
// partial sums and handing the complex result to ncclAllReduce, every rank
//   phase 1  (finish_contract, FINISH_EXCHANGE) folds its top-level vectors into the real length-n
//            vector x_rank (float64, already scaled by 1 / B_total) and STORES it into a slot of every
//            peer's exchange buffer (its own included), then publishes a per-contract flag on each peer
//            (release, system scope);
//   phase 2  (exchange_collect) waits for the flags of all ranks on its own buffer (acquire), sums the
//            `world` vectors in rank order — every rank forms the identical float64 sum — and takes the
//            ONE transform.
// The exchange therefore moves n doubles per contract and rank (half of what the complex all-reduce
// moves), carries the sum in float64, and costs one small launch and no host involvement.
// Deadlock freedom: phase 1 runs inside the step kernel and never waits; phase 2 is a kernel of its
// own, launched behind it on the same stream and never larger than the number of co-resident CTAs, so
// every flag a peer waits for is written by a kernel that is already running or complete.  Slots
// alternate with the epoch parity: a rank can be at most one call ahead of a peer, because its phase 2
// of call k needs that peer's phase 1 of call k.  A wait is bounded in TIME (PeerExchange::timeout_ns,
// default two minutes): a peer that never arrives yields NaN targets and a status word, not a trap.
//
// The same two phases as a kernel of their own, for transforms too large to sit in the step kernel:
template <typename Real>
__global__ void __launch_bounds__(CF_BLOCK)
    cf_exchange_finalize_kernel(const double* __restrict__ vecs, int64_t contract_stride, int64_t groups, int64_t n,
                                double scale, int mode, int log2n, Real* __restrict__ out, int64_t contracts,
                                const PeerExchange px) {
  extern __shared__ double smem[];
  __shared__ int s_ok;
  double* re = smem;
  double* im = smem + n;
  double* twr = smem + 2 * n;
  double* twi = twr + (mode == 0 ? n / 2 : n);
  fill_twiddles(twr, twi, mode == 0 ? n / 2 : n, n);

  const int64_t cap = px.capacity_contracts;
  const int64_t slot = px.epoch & 1u;
  const size_t flag_base = exchange_data_doubles(cap, n, px.world);  // flags follow the data (as 8-byte cells)
  const size_t status_cell = exchange_status_cell(cap, n, px.world);

  // phase 1: fold, push to every peer, publish
  for (int64_t c = blockIdx.x; c < contracts; c += gridDim.x) {
    const double* src = vecs + c * contract_stride;
    const size_t cell = ((slot * px.world + px.rank) * cap + c);
    for (int64_t col = threadIdx.x; col < n; col += CF_BLOCK) {
      const double x = strided_sum(src + col, groups, n) * scale;
      for (int p = 0; p < px.world; ++p) px.data[p][cell * n + col] = x;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < px.world)
      store_release_sys(reinterpret_cast<unsigned*>(px.data[threadIdx.x] + flag_base + cell), px.epoch);
  }

  // phase 2: wait for every rank's vector of this contract, sum in rank order, transform
  const double* mine = px.data[px.rank];
  for (int64_t c = blockIdx.x; c < contracts; c += gridDim.x) {
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < px.world) {
      const unsigned* flag = reinterpret_cast<const unsigned*>(mine + flag_base + ((slot * px.world + threadIdx.x) * cap + c));
      if (!wait_for_flag(flag, px, status_cell)) s_ok = 0;
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    for (int64_t col = threadIdx.x; col < n; col += CF_BLOCK) {
      double s = 0.0;
      for (int q = 0; q < px.world; ++q) s += __ldcv(mine + ((slot * px.world + q) * cap + c) * n + col);
      int64_t where = col;
      if (mode == 0 && log2n > 0) where = static_cast<int64_t>(bit_reverse(static_cast<unsigned>(col), log2n));
      re[where] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
      im[where] = 0.0;
    }
    __syncthreads();
    transform_store<Real>(re, im, twr, twi, n, mode, out + c * n * 2);
    __syncthreads();  // re / im are reused by the next contract of this CTA
  }
}

// In-place sum over ranks of `count` doubles (count <= capacity contracts) through the small region of the
// exchange buffers: push to every peer, publish one flag per peer, wait for all senders, sum in rank order.
// One CTA (the vector is at most a few thousand doubles); used for the NORMALIZE terminal sums.
__global__ void __launch_bounds__(CF_BLOCK)
    p2p_allreduce_small_kernel(double* __restrict__ inout, int64_t count, int64_t n, const PeerExchange px) {
  __shared__ int s_ok;
  const int64_t cap = px.capacity_contracts;
  const int64_t slot = px.epoch & 1u;
  const size_t base = exchange_small_base(cap, n, px.world);
  const size_t flags = base + static_cast<size_t>(2) * px.world * cap;
  if (threadIdx.x == 0) s_ok = 1;
  for (int64_t i = threadIdx.x; i < count; i += CF_BLOCK) {
    const double v = inout[i];
    for (int p = 0; p < px.world; ++p) px.data[p][base + (slot * px.world + px.rank) * cap + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < px.world) {
    store_release_sys(reinterpret_cast<unsigned*>(px.data[threadIdx.x] + flags + slot * px.world + px.rank), px.epoch);
    const unsigned* flag = reinterpret_cast<const unsigned*>(px.data[px.rank] + flags + slot * px.world + threadIdx.x);
    if (!wait_for_flag(flag, px, exchange_status_cell(cap, n, px.world))) s_ok = 0;
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  const double* mine = px.data[px.rank] + base;
  for (int64_t i = threadIdx.x; i < count; i += CF_BLOCK) {
    double s = 0.0;
    for (int q = 0; q < px.world; ++q) s += __ldcv(mine + (slot * px.world + q) * cap + i);
    inout[i] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
  }
}

// ------------------------------------------------------------------------------------------
// host-side orchestration
// ------------------------------------------------------------------------------------------
constexpr size_t SMEM_LIMIT = 200 * 1024;

struct FinalizePlan {
  int mode, log2n;
  size_t smem;
  bool in_step;  // the transform runs inside the step kernel (finish_contract)
};

static FinalizePlan finalize_plan(int64_t n) {
  FinalizePlan f{};
  const bool pow2 = (n & (n - 1)) == 0;
  f.log2n = 0;
  while ((int64_t(1) << f.log2n) < n) ++f.log2n;
  if (pow2 && static_cast<size_t>(n) * 24 <= SMEM_LIMIT) {
    f.mode = 0;
    f.smem = static_cast<size_t>(n) * 24;
  } else if (static_cast<size_t>(n) * 32 <= SMEM_LIMIT) {
    f.mode = 1;
    f.smem = static_cast<size_t>(n) * 32;
    f.log2n = 0;
  } else {
    f.mode = 2;
    f.smem = 0;
    f.log2n = 0;
  }
  if (f.mode == 0 && n == 1) f.smem = 32;
  static const bool separate = [] {  // diagnostic: SMC_SEPARATE_FINALIZE=1 keeps the transform in its own kernel
    const char* e = std::getenv("SMC_SEPARATE_FINALIZE");
    return e && std::atoi(e) != 0;
  }();
  f.in_step = f.mode != 2 && f.smem <= FUSED_FINALIZE_SMEM_MAX && !separate;
  return f;
}

struct Workspace {
  char* base;
  size_t size, used;
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = align_up(count * sizeof(T));
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

// Control words of one launch: the consts-ready flags followed by the tickets.  They are the only part
// of the workspace that must be zero when the kernel starts (one cudaMemsetAsync per launch).
constexpr size_t CONSTS_STRIDE = 64;  // >= sizeof(SimConsts<double>)

static size_t control_bytes(const TilePlan& plan, int64_t contracts, bool colsum) {
  const size_t tickets = colsum ? static_cast<size_t>(contracts) * (plan.tree.nodes + 1) : static_cast<size_t>(contracts);
  return align_up((static_cast<size_t>(contracts) + tickets) * sizeof(unsigned));
}

// bytes needed by the column-sum pipeline for `contracts` contracts
static size_t colsum_bytes(const TilePlan& plan, int64_t contracts, int64_t n) {
  size_t b = align_up(static_cast<size_t>(contracts) * plan.tiles * n * sizeof(double));
  b += align_up(static_cast<size_t>(contracts) * plan.tree.nodes * n * sizeof(double));
  b += align_up(static_cast<size_t>(contracts) * CONSTS_STRIDE);
  b += control_bytes(plan, contracts, true);
  if (finalize_plan(n).mode == 2) b += align_up(static_cast<size_t>(contracts) * n * sizeof(double));
  return b;
}

static int check_args(const char* fn, const smc_fused_args* a) {
  SMC_REQUIRE(a != nullptr, "%s: args is NULL", fn);
  SMC_REQUIRE(a->n_contracts > 0, "%s: n_contracts must be > 0", fn);
  SMC_REQUIRE(a->timesteps > 0 && a->network_size > 0 && a->batches_total > 0,
              "%s: timesteps, network_size and batches_total must be > 0", fn);
  SMC_REQUIRE(a->batch_begin >= 0 && a->batch_begin < a->batch_end && a->batch_end <= a->batches_total,
              "%s: invalid batch range [%lld, %lld) of %lld", fn, (long long)a->batch_begin,
              (long long)a->batch_end, (long long)a->batches_total);
  SMC_REQUIRE(a->dtype == SMC_F32 || a->dtype == SMC_F64, "%s: invalid dtype %d", fn, a->dtype);
  SMC_REQUIRE(a->scheme >= 0 && a->scheme <= 2, "%s: invalid scheme %d", fn, a->scheme);
  SMC_REQUIRE(a->normalization == SMC_NORMALIZE || a->normalization == SMC_RAW, "%s: invalid normalization %d",
              fn, a->normalization);
  SMC_REQUIRE(static_cast<double>(a->batches_total) * static_cast<double>(a->network_size) <= 4294967295.0,
              "%s: total paths exceed the 32-bit path counter", fn);
  SMC_REQUIRE(a->timesteps <= 0x7fffffffLL, "%s: timesteps too large", fn);
  SMC_REQUIRE(a->stream_version == SMC_STREAM_PHILOX10 || a->stream_version == SMC_STREAM_PHILOX7, "%s: invalid stream_version %d", fn,
              a->stream_version);
  SMC_REQUIRE((a->first_matrix_index >> 62) == 0, "%s: first_matrix_index too large", fn);
  return SMC_OK;
}

static TileParams base_params(const smc_fused_args* a, const TilePlan& plan) {
  TileParams p{};
  p.contracts = a->contracts;
  p.timesteps = a->timesteps;
  p.n = a->network_size;
  p.batches_total = a->batches_total;
  p.row_begin = a->batch_begin;
  p.row_end = a->batch_end;
  p.paths_local = (a->batch_end - a->batch_begin) * a->network_size;
  p.tile_rows = plan.tile_rows;
  p.tiles = plan.tiles;
  p.main_tiles = plan.main_tiles;
  p.main_rows = plan.main_rows;
  p.tail_tile_rows = plan.tail_tile_rows;
  p.chunk_w = plan.chunk_w;
  p.lanes_r = plan.lanes_r;
  p.tree = plan.tree;
  p.scheme = a->scheme;
  p.keys = make_philox_keys(a->seed);
  p.first_matrix_index = a->first_matrix_index;
  p.scale = 1.0 / static_cast<double>(a->batches_total);
  return p;
}

static TilePlan sim_plan(const smc_fused_args* a) {
  return make_plan(a->n_contracts, a->batch_end - a->batch_begin, a->network_size, false, a->timesteps);
}
static TilePlan stream_plan(const smc_fused_args* a) {
  return make_plan(a->n_contracts, a->batch_end - a->batch_begin, a->network_size, true);
}

// co-resident CTAs of a kernel (the peer-exchange collect kernels must never be larger than this)
static int resident_ctas(const void* kernel, size_t smem, int64_t* resident) {
  int per_sm = 0;
  SMC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CF_BLOCK, smem));
  *resident = static_cast<int64_t>(per_sm) * sm_count();
  SMC_REQUIRE(*resident > 0, "the kernel does not fit on an SM (%zu bytes of shared memory)", smem);
  return SMC_OK;
}

// diagnostic: SMC_SHORT_GROUPED=0 sends short paths through the general form (one block per path)
static bool short_grouped_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("SMC_SHORT_GROUPED");
    return !(e && std::atoi(e) == 0);
  }();
  return on;
}

template <typename Kernel>
static int launch_one(Kernel kernel, const TileParams& p, size_t smem, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>(p.tiles), static_cast<unsigned>(std::min<int64_t>(p.launch_contracts, 65535)),
                  static_cast<unsigned>((p.launch_contracts + 65534) / 65535));
  kernel<<<grid, CF_BLOCK, smem, st>>>(p);
  SMC_LAUNCH_OK("step_kernel");
  return SMC_OK;
}

// Launches the step kernel over `contracts` contracts: one CTA per (contract, tile).  `ctl` is
// control_bytes() of workspace, `consts` CONSTS_STRIDE bytes per contract; the rest of `p` (sources,
// outputs, finish) is filled in by the caller.
template <typename Real, int SRC, int OUT>
static int launch_step(TileParams p, int64_t contracts, const FinalizePlan& f, void* ctl, size_t ctl_bytes, void* consts,
                       cudaStream_t st) {
  if (p.tiles > 0x7fffffffLL) return set_error(SMC_EINVAL, "tile grid too large (%lld)", (long long)p.tiles);
  p.launch_contracts = contracts;
  p.chunk_shift = (p.chunk_w & (p.chunk_w - 1)) == 0 ? __builtin_ctz(static_cast<unsigned>(p.chunk_w)) : -1;
  p.consts = consts;
  p.consts_ready = static_cast<unsigned*>(ctl);
  p.tickets = static_cast<unsigned*>(ctl) + contracts;
  p.fft_mode = f.mode;
  p.log2n = f.log2n;
  SMC_CUDA_OK(cudaMemsetAsync(ctl, 0, ctl_bytes, st));
  const size_t smem = (OUT == OUT_COLSUM && p.finish == FINISH_TRANSFORM) ? f.smem : 0;
  // float32 fused kernels exist in two forms: with and without the ragged-tail code (see simulate_path_f32)
  if constexpr (SRC != SRC_FUSED) {
    return launch_one(step_kernel<Real, SRC, SMC_LOG_EULER, OUT, 1>, p, smem, st);
  } else if constexpr (sizeof(Real) == 8) {
    if (p.scheme == SMC_LOG_EULER) return launch_one(step_kernel<Real, SRC, SMC_LOG_EULER, OUT, 1>, p, smem, st);
    if (p.scheme == SMC_SIMPLE_EULER) return launch_one(step_kernel<Real, SRC, SMC_SIMPLE_EULER, OUT, 1>, p, smem, st);
    return launch_one(step_kernel<Real, SRC, SCHEME_LOG_STEPWISE, OUT, 1>, p, smem, st);
  } else {
    // FORM: 0 whole 6-row blocks, 1 general, 2 short paths in the grouped layout (COLSUM only; the staging pass of
    // NORMALIZE and shapes where 256 G is not a multiple of N take the general form, one path at a time)
    const bool grouped = OUT == OUT_COLSUM && p.timesteps <= 3 && (CF_BLOCK * (6 / p.timesteps)) % p.n == 0 && short_grouped_enabled();
    const int form = p.timesteps % 6 == 0 ? 0 : (grouped ? 2 : 1);
    if (p.scheme == SMC_LOG_EULER) {
      if (form == 0) return launch_one(step_kernel<Real, SRC, SMC_LOG_EULER, OUT, 0>, p, smem, st);
      if (form == 2) return launch_one(step_kernel<Real, SRC, SMC_LOG_EULER, OUT == OUT_COLSUM ? OUT : OUT_COLSUM, OUT == OUT_COLSUM ? 2 : 1>, p, smem, st);
      return launch_one(step_kernel<Real, SRC, SMC_LOG_EULER, OUT, 1>, p, smem, st);
    } else if (p.scheme == SMC_SIMPLE_EULER) {
      if (form == 0) return launch_one(step_kernel<Real, SRC, SMC_SIMPLE_EULER, OUT, 0>, p, smem, st);
      if (form == 2) return launch_one(step_kernel<Real, SRC, SMC_SIMPLE_EULER, OUT == OUT_COLSUM ? OUT : OUT_COLSUM, OUT == OUT_COLSUM ? 2 : 1>, p, smem, st);
      return launch_one(step_kernel<Real, SRC, SMC_SIMPLE_EULER, OUT, 1>, p, smem, st);
    }
    if (form == 0) return launch_one(step_kernel<Real, SRC, SCHEME_LOG_STEPWISE, OUT, 0>, p, smem, st);
    if (form == 2) return launch_one(step_kernel<Real, SRC, SCHEME_LOG_STEPWISE, OUT == OUT_COLSUM ? OUT : OUT_COLSUM, OUT == OUT_COLSUM ? 2 : 1>, p, smem, st);
    return launch_one(step_kernel<Real, SRC, SCHEME_LOG_STEPWISE, OUT, 1>, p, smem, st);
  }
}

static PeerExchange make_peer_exchange(const smc_p2p_group* g) {
  PeerExchange px{};
  for (int q = 0; q < g->world; ++q) px.data[q] = static_cast<double*>(g->buffers[q]);
  px.rank = g->rank;
  px.world = g->world;
  px.epoch = g->epoch;
  px.capacity_contracts = g->capacity_contracts;
  static const unsigned long long env_ms = [] {  // SMC_P2P_TIMEOUT_MS overrides the default of two minutes
    const char* e = std::getenv("SMC_P2P_TIMEOUT_MS");
    const long long v = e ? std::atoll(e) : 0;
    return v > 0 ? static_cast<unsigned long long>(v) : 120000ull;
  }();
  px.timeout_ns = (g->timeout_ms > 0 ? static_cast<unsigned long long>(g->timeout_ms) : env_ms) * 1000000ull;
  return px;
}

// Column-sum step over `contracts` contracts of `a` (sources already set in `p`): tiles -> ticket tree ->
// targets in `out` rows out_contract0.., complete (one GPU / NULL group: local sums scaled by 1/B_total;
// with a peer group: summed over the ranks).  Takes partial / nodes / control / spill from `w`.
template <typename Real, int SRC>
static int colsum_step(TileParams p, const TilePlan& plan, int64_t contracts, int64_t n, void* out, int64_t out_contract0,
                       const smc_p2p_group* g, Workspace& w, cudaStream_t st) {
  const FinalizePlan f = finalize_plan(n);
  p.partial = w.take<double>(contracts * plan.tiles * n);
  p.nodes = plan.tree.nodes ? w.take<double>(contracts * plan.tree.nodes * n) : nullptr;
  void* consts = w.take<char>(contracts * CONSTS_STRIDE);
  const size_t ctl_bytes = control_bytes(plan, contracts, true);
  void* ctl = w.take<char>(ctl_bytes);
  double* spill = f.mode == 2 ? w.take<double>(contracts * n) : nullptr;
  p.out = out;
  p.out_contract0 = out_contract0;
  p.finish = !f.in_step ? FINISH_NONE : (g ? FINISH_EXCHANGE : FINISH_TRANSFORM);
  if (g) p.px = make_peer_exchange(g);
  if (int e = launch_step<Real, SRC, OUT_COLSUM>(p, contracts, f, ctl, ctl_bytes, consts, st)) return e;
  if (f.in_step && g == nullptr) return SMC_OK;
  if (f.in_step) {
    // the finishing CTAs have pushed this rank's vectors; a small co-resident grid collects the peers'
    p.launch_contracts = contracts;
    p.fft_mode = f.mode;
    p.log2n = f.log2n;
    int64_t resident = 0;
    if (int e = resident_ctas(reinterpret_cast<const void*>(exchange_collect_kernel<Real>), f.smem, &resident)) return e;
    exchange_collect_kernel<Real><<<static_cast<unsigned>(std::min<int64_t>(contracts, resident)), CF_BLOCK, f.smem, st>>>(p);
    SMC_LAUNCH_OK("exchange_collect_kernel");
    return SMC_OK;
  }
  // large transforms: the top-level vectors of every contract -> separate kernel
  const double* vecs = plan.tree.levels ? p.nodes + plan.tree.off[plan.tree.levels] * n : p.partial;
  const int64_t stride = (plan.tree.levels ? plan.tree.nodes : plan.tiles) * n;
  const int64_t groups = plan.tree.count[plan.tree.levels];
  if (g == nullptr) {
    if (f.smem > 48 * 1024)
      SMC_CUDA_OK(cudaFuncSetAttribute(cf_finalize_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(f.smem)));
    cf_finalize_kernel<Real><<<static_cast<unsigned>(contracts), CF_BLOCK, f.smem, st>>>(
        vecs, stride, groups, n, p.scale, f.mode, f.log2n, static_cast<Real*>(out), out_contract0, spill);
    SMC_LAUNCH_OK("cf_finalize_kernel");
    return SMC_OK;
  }
  if (f.smem > 48 * 1024)
    SMC_CUDA_OK(cudaFuncSetAttribute(cf_exchange_finalize_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(f.smem)));
  int64_t resident = 0;
  if (int e = resident_ctas(reinterpret_cast<const void*>(cf_exchange_finalize_kernel<Real>), f.smem, &resident)) return e;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(contracts, resident));  // co-resident: see the kernel
  cf_exchange_finalize_kernel<Real><<<grid, CF_BLOCK, f.smem, st>>>(
      vecs, stride, groups, n, p.scale, f.mode, f.log2n, static_cast<Real*>(out) + out_contract0 * n * 2, contracts, p.px);
  SMC_LAUNCH_OK("cf_exchange_finalize_kernel");
  return SMC_OK;
}

// Terminal staging step (NORMALIZE pass A): simulate, store terminals, per-contract local sums.
template <typename Real>
static int terminal_step(TileParams p, const TilePlan& plan, int64_t contracts, void* terminal, double* terminal_sum,
                         Workspace& w, cudaStream_t st) {
  p.term_partial = w.take<double>(contracts * plan.tiles);
  void* consts = w.take<char>(contracts * CONSTS_STRIDE);
  const size_t ctl_bytes = control_bytes(plan, contracts, false);
  void* ctl = w.take<char>(ctl_bytes);
  p.terminal_out = terminal;
  p.terminal_sum_out = terminal_sum;
  p.finish = FINISH_NONE;
  return launch_step<Real, SRC_FUSED, OUT_TERMINAL>(p, contracts, FinalizePlan{}, ctl, ctl_bytes, consts, st);
}

static size_t terminal_step_bytes(const TilePlan& plan, int64_t contracts) {
  return align_up(static_cast<size_t>(contracts) * plan.tiles * sizeof(double)) + align_up(static_cast<size_t>(contracts) * CONSTS_STRIDE) +
         control_bytes(plan, contracts, false);
}

// contracts per pass of the single-GPU NORMALIZE path, given the bytes left for staging
static size_t normalize_bytes_per_contract(const smc_fused_args* a, const TilePlan& simp, const TilePlan& pay_plan) {
  const int64_t rows = a->batch_end - a->batch_begin;
  return align_up(static_cast<size_t>(rows) * a->network_size * real_size(a->dtype)) +
         colsum_bytes(pay_plan, 1, a->network_size) + terminal_step_bytes(simp, 1) + 512;
}

static bool valid_shape(const smc_fused_args* a) {
  return a != nullptr && a->n_contracts > 0 && a->network_size > 0 && a->batch_end > a->batch_begin && a->timesteps > 0;
}

}  // namespace SMC_NS

using namespace SMC_NS;

#ifndef SMC_STREAM_P7
extern "C" size_t smc_cf_fused_workspace_bytes(const smc_fused_args* a) {
  if (!valid_shape(a)) return 0;
  const TilePlan plan = sim_plan(a);
  if (a->normalization == SMC_RAW) return colsum_bytes(plan, a->n_contracts, a->network_size) + 256;
  // NORMALIZE: stage terminals; cap the staging at 8 GiB by chunking contracts
  const size_t per = normalize_bytes_per_contract(a, plan, stream_plan(a));
  const size_t cap = size_t(8) << 30;
  int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(a->n_contracts, static_cast<int64_t>(cap / per)));
  return per * chunk + align_up(a->n_contracts * sizeof(double)) + 256;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fused_launch_count(const smc_fused_args* a) {
  if (!valid_shape(a)) return 0;
  const int separate = finalize_plan(a->network_size).in_step ? 0 : 1;
  if (a->normalization == SMC_RAW) return 1 + separate;  // the step kernel (+ the transform kernel for large N)
  return 2 + separate;  // terminal step + payoff step (+ transform), per chunk of contracts
}
#endif

// introspection: how the simulation of these arguments is cut into CTAs (no device access)
#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fused_plan(const smc_fused_args* a, int64_t* out, int capacity) {
  clear_error();
  if (int e = check_args("smc_cf_fused_plan", a)) return e;
  const TilePlan plan = sim_plan(a);
  SMC_REQUIRE(out != nullptr && capacity >= 7, "smc_cf_fused_plan: out needs 7 entries");
  out[0] = plan.tiles;
  out[1] = plan.tile_rows;
  out[2] = plan.lanes_r;
  out[3] = plan.tree.levels;
  out[4] = plan.tree.count[plan.tree.levels];
  out[5] = plan.main_tiles;
  out[6] = plan.tail_tile_rows;
  return SMC_OK;
}
#endif

template <typename Real>
static int cf_fused_impl(const smc_fused_args* a, const smc_p2p_group* g, void* cf_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t rows = a->batch_end - a->batch_begin;
  const int64_t n = a->network_size;
  const TilePlan plan = sim_plan(a);
  Workspace w{static_cast<char*>(ws), ws_bytes, 0};

  if (a->normalization == SMC_RAW) {
    if (ws_bytes < colsum_bytes(plan, a->n_contracts, n))
      return set_error(SMC_EWORKSPACE, "smc_cf_fused: workspace %zu < %zu", ws_bytes,
                       colsum_bytes(plan, a->n_contracts, n));
    return colsum_step<Real, SRC_FUSED>(base_params(a, plan), plan, a->n_contracts, n, cf_out, 0, g, w, st);
  }

  // NORMALIZE on one device: needs the mean over ALL paths of a contract
  if (a->batch_begin != 0 || a->batch_end != a->batches_total)
    return set_error(SMC_EINVAL,
                     "smc_cf_fused: NORMALIZE over a batch shard needs the global terminal mean; use "
                     "smc_fused_terminal + allreduce + smc_cf_from_terminal");
  const TilePlan pay_plan = stream_plan(a);
  double* term_sum = w.take<double>(a->n_contracts);
  const size_t avail = ws_bytes > w.used ? ws_bytes - w.used : 0;
  const int64_t chunk = std::min<int64_t>(
      a->n_contracts, static_cast<int64_t>(avail / normalize_bytes_per_contract(a, plan, pay_plan)));
  if (chunk < 1) return set_error(SMC_EWORKSPACE, "smc_cf_fused: workspace %zu too small for one contract", ws_bytes);
  const size_t mark = w.used;
  for (int64_t c0 = 0; c0 < a->n_contracts; c0 += chunk) {
    const int64_t cc = std::min<int64_t>(chunk, a->n_contracts - c0);
    w.used = mark;
    TileParams p = base_params(a, plan);
    p.contract0 = c0;
    Real* staging = w.take<Real>(cc * rows * n);
    if (int e = terminal_step<Real>(p, plan, cc, staging, term_sum + c0, w, st)) return e;
    TileParams pb = base_params(a, pay_plan);
    pb.contract0 = c0;
    pb.terminal_in = staging;
    pb.terminal_sum = term_sum + c0;
    pb.normalize = 1;
    if (int e = colsum_step<Real, SRC_TERMINAL>(pb, pay_plan, cc, n, cf_out, c0, nullptr, w, st)) return e;
  }
  return SMC_OK;
}

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fused(const smc_fused_args* a, void* cf_out, void* ws, size_t ws_bytes, void* stream) {
  clear_error();
  if (int e = check_args("smc_cf_fused", a)) return e;
  SMC_REQUIRE(a->contracts != nullptr && cf_out != nullptr, "smc_cf_fused: NULL pointer");
  SMC_REQUIRE(ws != nullptr, "smc_cf_fused: workspace is NULL");
  if (a->stream_version == SMC_STREAM_PHILOX7) return smc_p7_cf_fused(a, nullptr, cf_out, ws, ws_bytes, stream);
  return a->dtype == SMC_F32 ? cf_fused_impl<float>(a, nullptr, cf_out, ws, ws_bytes, as_stream(stream))
                             : cf_fused_impl<double>(a, nullptr, cf_out, ws, ws_bytes, as_stream(stream));
}
#endif

// ---- batch-sharded RAW with the all-reduce fused into the step kernel (peer memory) ----------------
#ifndef SMC_STREAM_P7
extern "C" size_t smc_p2p_buffer_bytes(int64_t capacity_contracts, int64_t network_size, int world) {
  if (capacity_contracts <= 0 || network_size <= 0 || world <= 0 || world > MAX_PEERS) return 0;
  return exchange_total_cells(capacity_contracts, network_size, world) * sizeof(double);
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_p2p_alloc(size_t bytes, void** ptr, void* handle64) {
  clear_error();
  SMC_REQUIRE(bytes > 0 && ptr && handle64, "smc_p2p_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  SMC_CUDA_OK(cudaMalloc(ptr, bytes));
  SMC_CUDA_OK(cudaMemset(*ptr, 0, bytes));  // epochs start at 1: a zeroed flag is "not yet"
  SMC_CUDA_OK(cudaDeviceSynchronize());
  SMC_CUDA_OK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), *ptr));
  return SMC_OK;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_p2p_open(const void* handle64, void** ptr) {
  clear_error();
  SMC_REQUIRE(handle64 && ptr, "smc_p2p_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  SMC_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SMC_OK;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_p2p_close(void* ptr) {
  clear_error();
  SMC_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return SMC_OK;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_p2p_free(void* ptr) {
  clear_error();
  SMC_CUDA_OK(cudaFree(ptr));
  return SMC_OK;
}
#endif

static int check_group_only(const char* fn, const smc_p2p_group* g) {
  SMC_REQUIRE(g != nullptr, "%s: group is NULL", fn);
  SMC_REQUIRE(g->world >= 1 && g->world <= MAX_PEERS && g->rank >= 0 && g->rank < g->world, "%s: bad rank %d of %d", fn,
              g->rank, g->world);
  SMC_REQUIRE(g->epoch > 0, "%s: epoch must be > 0 (zero marks an unwritten flag)", fn);
  SMC_REQUIRE(g->capacity_contracts > 0 && g->network_size > 0, "%s: bad buffer shape", fn);
  for (int q = 0; q < g->world; ++q) SMC_REQUIRE(g->buffers[q] != nullptr, "%s: buffer of rank %d is NULL", fn, q);
  return SMC_OK;
}

static int check_group(const char* fn, const smc_fused_args* a, const smc_p2p_group* g) {
  if (int e = check_group_only(fn, g)) return e;
  SMC_REQUIRE(a->n_contracts <= g->capacity_contracts && a->network_size == g->network_size,
              "%s: exchange buffers sized for %lld contracts x %lld, call has %lld x %lld", fn,
              (long long)g->capacity_contracts, (long long)g->network_size, (long long)a->n_contracts,
              (long long)a->network_size);
  return SMC_OK;
}

// Everything smc_cf_fused_p2p would reject, without touching the device: callers validate BEFORE they
// advance the epoch, so that a host-side failure on one rank cannot put the ranks out of step.
#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fused_p2p_check(const smc_fused_args* a, const smc_p2p_group* g, size_t ws_bytes) {
  clear_error();
  if (int e = check_args("smc_cf_fused_p2p", a)) return e;
  SMC_REQUIRE(a->normalization == SMC_RAW,
              "smc_cf_fused_p2p: NORMALIZE needs the global terminal mean first (smc_fused_terminal + allreduce + smc_cf_from_terminal)");
  if (int e = check_group("smc_cf_fused_p2p", a, g)) return e;
  if (finalize_plan(a->network_size).mode == 2)
    return set_error(SMC_EUNSUPPORTED, "smc_cf_fused_p2p: network_size %lld needs the spill transform", (long long)a->network_size);
  const size_t need = colsum_bytes(sim_plan(a), a->n_contracts, a->network_size);
  if (ws_bytes < need) return set_error(SMC_EWORKSPACE, "smc_cf_fused_p2p: workspace %zu < %zu", ws_bytes, need);
  return SMC_OK;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fused_p2p(const smc_fused_args* a, const smc_p2p_group* g, void* cf_out, void* ws, size_t ws_bytes,
                                void* stream) {
  if (int e = smc_cf_fused_p2p_check(a, g, ws_bytes)) return e;
  SMC_REQUIRE(a->contracts != nullptr && cf_out != nullptr && ws != nullptr, "smc_cf_fused_p2p: NULL pointer");
  if (a->stream_version == SMC_STREAM_PHILOX7) return smc_p7_cf_fused(a, g, cf_out, ws, ws_bytes, stream);
  return a->dtype == SMC_F32 ? cf_fused_impl<float>(a, g, cf_out, ws, ws_bytes, as_stream(stream))
                             : cf_fused_impl<double>(a, g, cf_out, ws, ws_bytes, as_stream(stream));
}
#endif

// Epoch of the first call on this rank whose wait for a peer timed out (0: none).  Synchronises `stream`.
#ifndef SMC_STREAM_P7
extern "C" int smc_p2p_status(const smc_p2p_group* g, uint32_t* timed_out_epoch, void* stream) {
  clear_error();
  SMC_REQUIRE(timed_out_epoch != nullptr && g != nullptr, "smc_p2p_status: NULL pointer");
  smc_p2p_group probe = *g;
  if (probe.epoch == 0) probe.epoch = 1;
  if (int e = check_group_only("smc_p2p_status", &probe)) return e;
  const double* cell = static_cast<const double*>(g->buffers[g->rank]) +
                       exchange_status_cell(g->capacity_contracts, g->network_size, g->world);
  SMC_CUDA_OK(cudaMemcpyAsync(timed_out_epoch, cell, sizeof(uint32_t), cudaMemcpyDeviceToHost, as_stream(stream)));
  SMC_CUDA_OK(cudaStreamSynchronize(as_stream(stream)));
  return SMC_OK;
}
#endif

// ---- two-phase API for batch-sharded NORMALIZE -------------------------------------------
#ifndef SMC_STREAM_P7
extern "C" size_t smc_fused_terminal_workspace_bytes(const smc_fused_args* a) {
  if (!valid_shape(a)) return 0;
  return terminal_step_bytes(sim_plan(a), a->n_contracts) + 256;
}
#endif

template <typename Real>
static int fused_terminal_impl(const smc_fused_args* a, void* terminal, double* terminal_sum, void* ws,
                               size_t ws_bytes, cudaStream_t st) {
  const TilePlan plan = sim_plan(a);
  if (ws_bytes < terminal_step_bytes(plan, a->n_contracts))
    return set_error(SMC_EWORKSPACE, "smc_fused_terminal: workspace too small");
  Workspace w{static_cast<char*>(ws), ws_bytes, 0};
  return terminal_step<Real>(base_params(a, plan), plan, a->n_contracts, terminal, terminal_sum, w, st);
}

#ifndef SMC_STREAM_P7
extern "C" int smc_fused_terminal(const smc_fused_args* a, void* terminal, double* terminal_sum, void* ws,
                                  size_t ws_bytes, void* stream) {
  clear_error();
  if (int e = check_args("smc_fused_terminal", a)) return e;
  SMC_REQUIRE(a->contracts && terminal && terminal_sum && ws, "smc_fused_terminal: NULL pointer");
  if (a->stream_version == SMC_STREAM_PHILOX7) return smc_p7_fused_terminal(a, terminal, terminal_sum, ws, ws_bytes, stream);
  return a->dtype == SMC_F32 ? fused_terminal_impl<float>(a, terminal, terminal_sum, ws, ws_bytes, as_stream(stream))
                             : fused_terminal_impl<double>(a, terminal, terminal_sum, ws, ws_bytes, as_stream(stream));
}
#endif

#ifndef SMC_STREAM_P7
extern "C" size_t smc_cf_from_terminal_workspace_bytes(const smc_fused_args* a) {
  if (!valid_shape(a)) return 0;
  return colsum_bytes(stream_plan(a), a->n_contracts, a->network_size) + 256;
}
#endif

template <typename Real>
static int cf_from_terminal_impl(const smc_fused_args* a, const smc_p2p_group* g, const void* terminal, const double* tsum,
                                 void* cf_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t n = a->network_size;
  const TilePlan plan = stream_plan(a);
  if (ws_bytes < colsum_bytes(plan, a->n_contracts, n))
    return set_error(SMC_EWORKSPACE, "smc_cf_from_terminal: workspace too small");
  Workspace w{static_cast<char*>(ws), ws_bytes, 0};
  TileParams p = base_params(a, plan);
  p.terminal_in = terminal;
  p.terminal_sum = tsum;
  p.normalize = tsum != nullptr;
  return colsum_step<Real, SRC_TERMINAL>(p, plan, a->n_contracts, n, cf_out, 0, g, w, st);
}

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_from_terminal(const smc_fused_args* a, const void* terminal, const double* tsum, void* cf_out,
                                    void* ws, size_t ws_bytes, void* stream) {
  clear_error();
  if (int e = check_args("smc_cf_from_terminal", a)) return e;
  SMC_REQUIRE(a->contracts && terminal && cf_out && ws, "smc_cf_from_terminal: NULL pointer");
  SMC_REQUIRE((a->normalization == SMC_NORMALIZE) == (tsum != nullptr),
              "smc_cf_from_terminal: terminal_sum_global must be given iff normalization is NORMALIZE");
  return a->dtype == SMC_F32
             ? cf_from_terminal_impl<float>(a, nullptr, terminal, tsum, cf_out, ws, ws_bytes, as_stream(stream))
             : cf_from_terminal_impl<double>(a, nullptr, terminal, tsum, cf_out, ws, ws_bytes, as_stream(stream));
}
#endif

// NORMALIZE over several GPUs without a collective call: smc_fused_terminal, then this in-place sum over
// ranks of the per-contract terminal sums, then smc_cf_from_terminal_p2p.
#ifndef SMC_STREAM_P7
extern "C" int smc_p2p_allreduce_sum_f64(double* inout, int64_t count, const smc_p2p_group* g, void* stream) {
  clear_error();
  SMC_REQUIRE(inout != nullptr && count > 0, "smc_p2p_allreduce_sum_f64: bad argument");
  if (int e = check_group_only("smc_p2p_allreduce_sum_f64", g)) return e;
  SMC_REQUIRE(count <= g->capacity_contracts, "smc_p2p_allreduce_sum_f64: %lld values, buffers hold %lld",
              (long long)count, (long long)g->capacity_contracts);
  p2p_allreduce_small_kernel<<<1, CF_BLOCK, 0, as_stream(stream)>>>(inout, count, g->network_size, make_peer_exchange(g));
  SMC_LAUNCH_OK("p2p_allreduce_small_kernel");
  return SMC_OK;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_from_terminal_p2p(const smc_fused_args* a, const smc_p2p_group* g, const void* terminal,
                                        const double* tsum, void* cf_out, void* ws, size_t ws_bytes, void* stream) {
  clear_error();
  if (int e = check_args("smc_cf_from_terminal_p2p", a)) return e;
  SMC_REQUIRE(a->contracts && terminal && cf_out && ws, "smc_cf_from_terminal_p2p: NULL pointer");
  SMC_REQUIRE((a->normalization == SMC_NORMALIZE) == (tsum != nullptr),
              "smc_cf_from_terminal_p2p: terminal_sum_global must be given iff normalization is NORMALIZE");
  if (int e = check_group("smc_cf_from_terminal_p2p", a, g)) return e;
  if (finalize_plan(a->network_size).mode == 2)
    return set_error(SMC_EUNSUPPORTED, "smc_cf_from_terminal_p2p: network_size %lld needs the spill transform", (long long)a->network_size);
  return a->dtype == SMC_F32
             ? cf_from_terminal_impl<float>(a, g, terminal, tsum, cf_out, ws, ws_bytes, as_stream(stream))
             : cf_from_terminal_impl<double>(a, g, terminal, tsum, cf_out, ws, ws_bytes, as_stream(stream));
}
#endif

// ---- materialised payoff matrix -> CF -------------------------------------------------------
#ifndef SMC_STREAM_P7
extern "C" size_t smc_cf_fft_mean_workspace_bytes(int64_t batches, int64_t n, int method) {
  if (batches <= 0 || n <= 0) return 0;
  if (method == SMC_CF_ROW_FFT && rowfft_supported(n)) return rowfft_workspace_bytes(batches, n);
  return colsum_bytes(make_plan(1, batches, n, true), 1, n) + 256;
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fft_mean(const void* mat, int64_t batches, int64_t n, int dtype, int method, void* out,
                               void* ws, size_t ws_bytes, void* stream) {
  clear_error();
  SMC_REQUIRE(mat && out && ws, "smc_cf_fft_mean: NULL pointer");
  SMC_REQUIRE(batches > 0 && n > 0, "smc_cf_fft_mean: invalid shape (%lld, %lld)", (long long)batches, (long long)n);
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_cf_fft_mean: invalid dtype %d", dtype);
  SMC_REQUIRE(method == SMC_CF_MEAN_THEN_FFT || method == SMC_CF_ROW_FFT, "smc_cf_fft_mean: invalid method %d", method);
  if (method == SMC_CF_ROW_FFT) return rowfft_mean(mat, batches, n, dtype, out, ws, ws_bytes, as_stream(stream));
  const TilePlan plan = make_plan(1, batches, n, true);
  if (ws_bytes < colsum_bytes(plan, 1, n)) return set_error(SMC_EWORKSPACE, "smc_cf_fft_mean: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace w{static_cast<char*>(ws), ws_bytes, 0};
  TileParams p{};
  p.n = n;
  p.batches_total = batches;
  p.row_begin = 0;
  p.row_end = batches;
  p.paths_local = batches * n;
  p.tile_rows = plan.tile_rows;
  p.tiles = plan.tiles;
  p.main_tiles = plan.main_tiles;
  p.main_rows = plan.main_rows;
  p.tail_tile_rows = plan.tail_tile_rows;
  p.chunk_w = plan.chunk_w;
  p.lanes_r = plan.lanes_r;
  p.tree = plan.tree;
  p.matrix = mat;
  p.scale = 1.0 / static_cast<double>(batches);
  if (dtype == SMC_F32) return colsum_step<float, SRC_MATRIX>(p, plan, 1, n, out, 0, nullptr, w, st);
  return colsum_step<double, SRC_MATRIX>(p, plan, 1, n, out, 0, nullptr, w, st);
}
#endif

// ---- per-row spectra (the ComputeFFT operator) ----------------------------------------------
#ifndef SMC_STREAM_P7
extern "C" int smc_fft_rows(const void* mat, int64_t batches, int64_t n, int dtype, void* out, void* stream) {
  clear_error();
  SMC_REQUIRE(batches > 0 && n > 0, "smc_fft_rows: invalid shape (%lld, %lld)", (long long)batches, (long long)n);
  SMC_REQUIRE(mat && out, "smc_fft_rows: NULL pointer");
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "smc_fft_rows: invalid dtype %d", dtype);
  return fft_rows(mat, batches, n, dtype, out, as_stream(stream));
}
#endif

// ---- host-buffer entry point ----------------------------------------------------------------
#ifndef SMC_STREAM_P7
extern "C" size_t smc_cf_fused_host_workspace_bytes(const smc_fused_args* a) {
  if (a == nullptr || a->n_contracts <= 0) return 0;
  const size_t cbytes = align_up(static_cast<size_t>(a->n_contracts) * 6 * sizeof(double));
  const size_t obytes = align_up(static_cast<size_t>(a->n_contracts) * a->network_size * 2 * real_size(a->dtype));
  return cbytes + obytes + smc_cf_fused_workspace_bytes(a);
}
#endif

#ifndef SMC_STREAM_P7
extern "C" int smc_cf_fused_host(const smc_fused_args* a, const double* contracts_host, void* cf_host, void* ws,
                                 size_t ws_bytes, void* stream) {
  clear_error();
  if (int e = check_args("smc_cf_fused_host", a)) return e;
  SMC_REQUIRE(contracts_host && cf_host && ws, "smc_cf_fused_host: NULL pointer");
  if (ws_bytes < smc_cf_fused_host_workspace_bytes(a))
    return set_error(SMC_EWORKSPACE, "smc_cf_fused_host: workspace %zu < %zu", ws_bytes,
                     smc_cf_fused_host_workspace_bytes(a));
  cudaStream_t st = as_stream(stream);
  const size_t cbytes = static_cast<size_t>(a->n_contracts) * 6 * sizeof(double);
  const size_t obytes = static_cast<size_t>(a->n_contracts) * a->network_size * 2 * real_size(a->dtype);
  char* base = static_cast<char*>(ws);
  double* d_contracts = reinterpret_cast<double*>(base);
  void* d_out = base + align_up(cbytes);
  char* rest = base + align_up(cbytes) + align_up(obytes);
  // The contract rows are copied to the device (every CTA of the first wave reads them; through a host alias
  // that was 36 000 PCIe transactions, +0.29 ms at config c2).  The targets are written ONCE, by the CTA that
  // finishes a contract, so when the output buffer is pinned (device-addressable under unified addressing) and
  // small, the kernel stores them straight into host memory instead of a staging copy + cudaMemcpyAsync.
  auto device_alias = [](const void* host) -> void* {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
      cudaGetLastError();  // unregistered pageable memory reports an error on older drivers: not sticky
      return nullptr;
    }
    return attr.type == cudaMemoryTypeHost ? attr.devicePointer : nullptr;
  };
  constexpr size_t kZeroCopyOut = 64 << 10;
  void* o_alias = obytes <= kZeroCopyOut ? device_alias(cf_host) : nullptr;
  smc_fused_args b = *a;
  SMC_CUDA_OK(cudaMemcpyAsync(d_contracts, contracts_host, cbytes, cudaMemcpyHostToDevice, st));
  b.contracts = d_contracts;
  if (int e = smc_cf_fused(&b, o_alias ? o_alias : d_out, rest, ws_bytes - align_up(cbytes) - align_up(obytes), stream)) return e;
  if (!o_alias) SMC_CUDA_OK(cudaMemcpyAsync(cf_host, d_out, obytes, cudaMemcpyDeviceToHost, st));
  SMC_CUDA_OK(cudaStreamSynchronize(st));
  return SMC_OK;
}
#endif

// ---- the Philox4x32-7 build of this file exports only the two entry points that draw normals -------------
#ifdef SMC_STREAM_P7
extern "C" int smc_p7_cf_fused(const smc_fused_args* a, const smc_p2p_group* g, void* cf_out, void* ws, size_t ws_bytes, void* stream) {
  return a->dtype == SMC_F32 ? cf_fused_impl<float>(a, g, cf_out, ws, ws_bytes, as_stream(stream))
                             : cf_fused_impl<double>(a, g, cf_out, ws, ws_bytes, as_stream(stream));
}
extern "C" int smc_p7_fused_terminal(const smc_fused_args* a, void* terminal, double* terminal_sum, void* ws, size_t ws_bytes,
                                     void* stream) {
  return a->dtype == SMC_F32 ? fused_terminal_impl<float>(a, terminal, terminal_sum, ws, ws_bytes, as_stream(stream))
                             : fused_terminal_impl<double>(a, terminal, terminal_sum, ws, ws_bytes, as_stream(stream));
}
#endif
