// CVNN training step behind the C ABI: the consumer of the CF targets.
//
// Replaces, for networks that are a (possibly nested) ComplexSequential of ComplexLinear /
// modReLU / zReLU (/root/reference/src/spectralmc/cvnn.py:65-146, 149-162, 168-210, 439-452),
// the torch op-by-op execution of GbmCVNNPricer._torch_step (gbm_trainer.py:819-835):
//   pred = cvnn(real_in, imag_in); loss = mse(pred_r, Re t) + mse(pred_i, Im t);
//   zero_grad; backward; Adam step
// with a short, fixed sequence of stream-ordered launches that never touches the host, so the
// whole step can be captured into one CUDA graph:
//   * ComplexLinear = ONE complex GEMM (four real accumulators per output, combined exactly as
//     cvnn.py:137-138 combine the four real matmuls) with bias and the following activation
//     fused into the epilogue;
//   * backward = hand-derived: activation backward (elementwise), weight gradients as a
//     split-M complex GEMM with a fixed-order second pass, input gradients as a complex GEMM
//     with conj(W);
//   * MSE loss + its gradient in one pass; Adam over the flat parameter buffer in one launch
//     with the step counter on the device.
// Arithmetic is in the network dtype (float32 / float64) as torch's is — no TF32 / tensor
// cores: the reference trains in full fp32 (torch's default matmul precision), K is 6..256 and a
// whole step at the reference's sizes is a few tens of microseconds of launch latency.
// All reductions are fixed-order (no atomics): results are bit-reproducible.
#include <algorithm>
#include <vector>

#include "smc_device.cuh"
#include "smc_internal.h"

namespace smc {
namespace {

constexpr int BLK = 256;
constexpr int TI = 32, TJ = 32, TL = 16;     // cgemm tile
// The step is latency-bound at the reference's sizes (a thousand rows, widths 6..256), so the
// reductions over the batch are cut into many short, independent pieces: 64-row weight-gradient
// splits and 128-row column-sum slabs, each followed by a fixed-order second pass.
constexpr int64_t SPLIT_ROWS = 64;            // rows of the batch per weight-gradient split
constexpr int64_t MAX_SPLITS = 256;
constexpr int64_t COLSUM_SLAB = 128;          // rows per column-sum slab
constexpr int COLSUM_PLANES = 3;
constexpr int64_t LOSS_CHUNK = 4096;          // elements per CTA of the loss kernel
constexpr double MODRELU_EPS = 1e-9;          // cvnn.py:205

enum Act { ACT_NONE = 0, ACT_MODRELU = 1, ACT_ZRELU = 2 };

// C[i, j] = sum_l A[i, l] * (CONJ_B ? conj(B[l, j]) : B[l, j]); planes, arbitrary strides.
template <typename Real>
struct GemmParams {
  const Real *ar, *ai;  // A planes
  int64_t sai, sal;
  const Real *br, *bi;  // B planes
  int64_t sbl, sbj;
  Real *cr, *ci;        // C planes, row-major [I, J]; split s writes at + s * split_stride
  int64_t I, J, L;
  int64_t l_per_split;  // == L when not split
  int64_t split_stride;
  // epilogue (forward only): + bias[j], activation, optional store of the pre-activation
  const Real *bias_r, *bias_i;
  int act;
  const Real* act_bias;
  Real *pre_r, *pre_i;
};

template <typename Real>
__device__ __forceinline__ void modrelu_apply(Real zr, Real zi, Real b, Real& yr, Real& yi) {
  const Real mag = sqrt(zr * zr + zi * zi + static_cast<Real>(MODRELU_EPS));  // cvnn.py:205
  const Real thr = max(mag + b, Real(0));                                      // cvnn.py:206
  const Real s = thr / mag;                                                    // cvnn.py:207
  yr = s * zr;
  yi = s * zi;
}

template <typename Real, bool CONJ_B>
__global__ void __launch_bounds__(BLK) cgemm_kernel(const GemmParams<Real> p) {
  __shared__ Real a_r[TL][TI + 1], a_i[TL][TI + 1], b_r[TL][TJ + 1], b_i[TL][TJ + 1];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * TI, j0 = static_cast<int64_t>(blockIdx.x) * TJ;
  const int64_t l_begin = static_cast<int64_t>(blockIdx.z) * p.l_per_split;
  const int64_t l_end = min(l_begin + p.l_per_split, p.L);
  // the fastest-varying index of each operand decides which index neighbouring threads walk
  const bool a_l_fast = p.sal == 1, b_l_fast = p.sbl == 1 && p.sbj != 1;

  Real rr[2][2] = {}, ii[2][2] = {}, ri[2][2] = {}, ir[2][2] = {};
  for (int64_t l0 = l_begin; l0 < l_end; l0 += TL) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int li, ii_;
      if (a_l_fast) { li = t & 15; ii_ = (t >> 4) + 16 * h; } else { ii_ = t & 31; li = (t >> 5) + 8 * h; }
      const int64_t gi = i0 + ii_, gl = l0 + li;
      const bool ok = gi < p.I && gl < l_end;
      a_r[li][ii_] = ok ? p.ar[gi * p.sai + gl * p.sal] : Real(0);
      a_i[li][ii_] = ok ? p.ai[gi * p.sai + gl * p.sal] : Real(0);
      int lj, jj;
      if (b_l_fast) { lj = t & 15; jj = (t >> 4) + 16 * h; } else { jj = t & 31; lj = (t >> 5) + 8 * h; }
      const int64_t gj = j0 + jj, gl2 = l0 + lj;
      const bool ok2 = gj < p.J && gl2 < l_end;
      b_r[lj][jj] = ok2 ? p.br[gl2 * p.sbl + gj * p.sbj] : Real(0);
      b_i[lj][jj] = ok2 ? p.bi[gl2 * p.sbl + gj * p.sbj] : Real(0);
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < TL; ++l) {
      Real xr[2], xi[2], wr[2], wi[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) { xr[a] = a_r[l][ty + 16 * a]; xi[a] = a_i[l][ty + 16 * a]; }
#pragma unroll
      for (int b = 0; b < 2; ++b) { wr[b] = b_r[l][tx + 16 * b]; wi[b] = b_i[l][tx + 16 * b]; }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          rr[a][b] = fma(xr[a], wr[b], rr[a][b]);
          ii[a][b] = fma(xi[a], wi[b], ii[a][b]);
          ri[a][b] = fma(xr[a], wi[b], ri[a][b]);
          ir[a][b] = fma(xi[a], wr[b], ir[a][b]);
        }
    }
    __syncthreads();
  }

  Real* cr = p.cr + static_cast<int64_t>(blockIdx.z) * p.split_stride;
  Real* ci = p.ci + static_cast<int64_t>(blockIdx.z) * p.split_stride;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int64_t i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i >= p.I || j >= p.J) continue;
      // (xr + i xi)(wr + i wi): real = xr wr - xi wi, imag = xr wi + xi wr  (cvnn.py:137-138);
      // with conj(B): real = xr wr + xi wi, imag = xi wr - xr wi
      Real zr = CONJ_B ? rr[a][b] + ii[a][b] : rr[a][b] - ii[a][b];
      Real zi = CONJ_B ? ir[a][b] - ri[a][b] : ri[a][b] + ir[a][b];
      if (p.bias_r) { zr += p.bias_r[j]; zi += p.bias_i[j]; }  // cvnn.py:141-144
      const int64_t e = i * p.J + j;
      if (p.act != ACT_NONE) {
        if (p.pre_r) { p.pre_r[e] = zr; p.pre_i[e] = zi; }
        if (p.act == ACT_MODRELU) {
          modrelu_apply(zr, zi, p.act_bias[j], zr, zi);
        } else {
          const bool keep = zr >= Real(0) && zi >= Real(0);  // cvnn.py:161
          zr = keep ? zr : Real(0);
          zi = keep ? zi : Real(0);
        }
      }
      cr[e] = zr;
      ci[e] = zi;
    }
}

// fixed-order sum of the split partials: dst_r[e] = sum_s part[s][0][e], dst_i[e] = sum_s part[s][1][e]
template <typename Real>
__global__ void __launch_bounds__(BLK)
    reduce_split_kernel(const Real* __restrict__ part, int64_t splits, int64_t count, Real* __restrict__ dst_r,
                        Real* __restrict__ dst_i) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * BLK + threadIdx.x;
  if (e >= count) return;
  Real sr = 0, si = 0;
  for (int64_t s = 0; s < splits; ++s) {
    sr += part[s * 2 * count + e];
    si += part[s * 2 * count + count + e];
  }
  dst_r[e] = sr;
  dst_i[e] = si;
}

// column sums of up to three [rows, cols] planes in one launch: plane z, slab y of COLSUM_SLAB rows
// (slab_rows rows) -> dst[z][y * cols + j]
template <typename Real>
struct ColsumJob {
  const Real* src[COLSUM_PLANES];
  Real* dst[COLSUM_PLANES];
};

template <typename Real>
__global__ void __launch_bounds__(BLK)
    colsum_kernel(const ColsumJob<Real> job, int64_t rows, int64_t cols, int64_t slab_rows) {
  __shared__ Real sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const Real* __restrict__ src = job.src[blockIdx.z];
  const int64_t j = static_cast<int64_t>(blockIdx.x) * 32 + tx;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * slab_rows, r1 = min(r0 + slab_rows, rows);
  Real s = 0;
  if (j < cols) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) s += src[r * cols + j];
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < cols) {
    Real tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sm[k][tx];
    job.dst[blockIdx.z][static_cast<int64_t>(blockIdx.y) * cols + j] = tot;
  }
}

// standalone activation forward (an activation that does not directly follow a ComplexLinear)
template <typename Real>
__global__ void __launch_bounds__(BLK)
    act_forward_kernel(const Real* __restrict__ zr, const Real* __restrict__ zi, int64_t count, int64_t cols, int act,
                       const Real* __restrict__ bias, Real* __restrict__ yr, Real* __restrict__ yi) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * BLK + threadIdx.x;
  if (e >= count) return;
  Real a = zr[e], b = zi[e];
  if (act == ACT_MODRELU) {
    modrelu_apply(a, b, bias[e % cols], a, b);
  } else {
    const bool keep = a >= Real(0) && b >= Real(0);
    a = keep ? a : Real(0);
    b = keep ? b : Real(0);
  }
  yr[e] = a;
  yi[e] = b;
}

// activation backward, in place on (gr, gi): grad wrt the activation output -> grad wrt its input z.
// modReLU: y = s z, s = relu(|z| + b) / |z|  =>  with on = (|z| + b > 0), dot = gr zr + gi zi:
//   gz = g s - on b dot z / |z|^3 ;  d/db = on dot / |z|   (written to bterm for the column sum)
template <typename Real>
__global__ void __launch_bounds__(BLK)
    act_backward_kernel(Real* __restrict__ gr, Real* __restrict__ gi, const Real* __restrict__ zr,
                        const Real* __restrict__ zi, int64_t count, int64_t cols, int act,
                        const Real* __restrict__ bias, Real* __restrict__ bterm) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * BLK + threadIdx.x;
  if (e >= count) return;
  const Real a = zr[e], b = zi[e];
  Real ga = gr[e], gb = gi[e];
  if (act == ACT_MODRELU) {
    const Real bb = bias[e % cols];
    const Real mag = sqrt(a * a + b * b + static_cast<Real>(MODRELU_EPS));
    const bool on = mag + bb > Real(0);
    const Real dot = ga * a + gb * b;
    const Real s = on ? (mag + bb) / mag : Real(0);
    const Real corr = on ? -bb * dot / (mag * mag * mag) : Real(0);
    ga = ga * s + corr * a;
    gb = gb * s + corr * b;
    bterm[e] = on ? dot / mag : Real(0);
  } else {
    const bool keep = a >= Real(0) && b >= Real(0);
    ga = keep ? ga : Real(0);
    gb = keep ? gb : Real(0);
  }
  gr[e] = ga;
  gi[e] = gb;
}

// loss = mean((pr - tr)^2) + mean((pi - ti)^2) (gbm_trainer.py:828-830); g = 2 (p - t) / count
template <typename Real>
__global__ void __launch_bounds__(BLK)
    mse_grad_kernel(const Real* __restrict__ pr, const Real* __restrict__ pi, const Real* __restrict__ target,
                    int64_t count, Real* __restrict__ gr, Real* __restrict__ gi, double* __restrict__ partial) {
  __shared__ double sm[32];
  const int64_t e0 = static_cast<int64_t>(blockIdx.x) * LOSS_CHUNK, e1 = min(e0 + LOSS_CHUNK, count);
  const Real k = static_cast<Real>(2.0 / static_cast<double>(count));
  double sr = 0.0, si = 0.0;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += BLK) {
    const Real dr = pr[e] - target[2 * e], di = pi[e] - target[2 * e + 1];
    gr[e] = k * dr;
    gi[e] = k * di;
    sr += static_cast<double>(dr * dr);
    si += static_cast<double>(di * di);
  }
  const double tr = block_sum(sr, sm);
  const double ti = block_sum(si, sm);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = tr;
    partial[2 * blockIdx.x + 1] = ti;
  }
}

__global__ void __launch_bounds__(BLK)
    loss_finalize_kernel(const double* __restrict__ partial, int64_t chunks, int64_t count, double* __restrict__ loss) {
  __shared__ double sm[32];
  double sr = 0.0, si = 0.0;
  for (int64_t c = threadIdx.x; c < chunks; c += BLK) {
    sr += partial[2 * c];
    si += partial[2 * c + 1];
  }
  const double tr = block_sum(sr, sm);
  const double ti = block_sum(si, sm);
  if (threadIdx.x == 0) *loss = tr / static_cast<double>(count) + ti / static_cast<double>(count);
}

// torch.optim.Adam (defaults: no weight decay, no amsgrad), single-tensor form:
//   m.lerp_(g, 1 - b1); v = b2 v + (1 - b2) g g; denom = sqrt(v) / sqrt(1 - b2^t) + eps;
//   p += -(lr / (1 - b1^t)) * (m / denom)
template <typename Real>
__global__ void __launch_bounds__(BLK)
    adam_kernel(Real* __restrict__ p, const Real* __restrict__ g, Real* __restrict__ m, Real* __restrict__ v,
                int64_t n, const int64_t* __restrict__ step, double lr, double b1, double b2, double eps) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * BLK + threadIdx.x;
  if (e >= n) return;
  const double t = static_cast<double>(*step + 1);
  const Real step_size = static_cast<Real>(lr / (1.0 - pow(b1, t)));
  const Real bc2_sqrt = static_cast<Real>(sqrt(1.0 - pow(b2, t)));
  const Real grad = g[e];
  Real mm = m[e], vv = v[e];
  mm += (grad - mm) * static_cast<Real>(1.0 - b1);
  vv = vv * static_cast<Real>(b2) + static_cast<Real>(1.0 - b2) * grad * grad;
  const Real denom = sqrt(vv) / bc2_sqrt + static_cast<Real>(eps);
  m[e] = mm;
  v[e] = vv;
  p[e] -= step_size * (mm / denom);
}

__global__ void bump_step_kernel(int64_t* step) { *step += 1; }

// ---- host-side plan ---------------------------------------------------------------------------
struct Node {           // a ComplexLinear with the activation fused behind it, or a lone activation
  bool linear;
  int act;
  int64_t in_w, out_w;
  bool has_bias;
  int64_t w_off;        // parameter offsets (elements); -1 = absent
  int64_t b_off, act_off;
  size_t out_r, out_i;  // workspace byte offsets of the node output planes
  size_t pre_r, pre_i;  // pre-activation planes (training, fused activation)
};

struct Plan {
  std::vector<Node> nodes;
  int64_t out_w = 0, max_w = 0;
  size_t g_r[2], g_i[2], bterm, wpart, colpart, losspart, total = 0;
};

int build_plan(const char* fn, const smc_cvnn_net* net, int64_t rows, bool training, bool last_to_user, Plan* plan) {
  SMC_REQUIRE(net != nullptr && net->layers != nullptr && net->n_layers > 0, "%s: empty network", fn);
  SMC_REQUIRE(net->dtype == SMC_F32 || net->dtype == SMC_F64, "%s: invalid dtype %d", fn, net->dtype);
  SMC_REQUIRE(net->n_inputs > 0 && rows > 0, "%s: n_inputs and rows must be > 0", fn);
  const size_t rs = real_size(net->dtype);
  int64_t w = net->n_inputs;
  size_t used = 0;
  auto take = [&](size_t bytes) { const size_t o = used; used += align_up(bytes); return o; };
  plan->max_w = w;
  for (int l = 0; l < net->n_layers; ++l) {
    const smc_cvnn_layer& L = net->layers[l];
    Node nd{};
    nd.w_off = nd.b_off = nd.act_off = -1;
    if (L.kind == SMC_LAYER_LINEAR) {
      SMC_REQUIRE(L.in_features == w, "%s: layer %d expects %lld inputs, the signal has %lld", fn, l,
                  (long long)L.in_features, (long long)w);
      SMC_REQUIRE(L.out_features > 0, "%s: layer %d has no outputs", fn, l);
      nd.linear = true;
      nd.in_w = w;
      nd.out_w = L.out_features;
      nd.has_bias = L.has_bias != 0;
      nd.w_off = L.param_offset;
      nd.b_off = nd.has_bias ? L.param_offset + 2 * nd.in_w * nd.out_w : -1;
      const int64_t end = L.param_offset + 2 * nd.in_w * nd.out_w + (nd.has_bias ? 2 * nd.out_w : 0);
      SMC_REQUIRE(L.param_offset >= 0 && end <= net->n_params, "%s: layer %d parameters fall outside the buffer", fn, l);
      nd.act = ACT_NONE;
      if (l + 1 < net->n_layers && net->layers[l + 1].kind != SMC_LAYER_LINEAR) {  // fuse the activation
        const smc_cvnn_layer& A = net->layers[++l];
        SMC_REQUIRE(A.kind == SMC_LAYER_MODRELU || A.kind == SMC_LAYER_ZRELU, "%s: unknown layer kind %d", fn, A.kind);
        nd.act = A.kind == SMC_LAYER_MODRELU ? ACT_MODRELU : ACT_ZRELU;
        if (nd.act == ACT_MODRELU) {
          SMC_REQUIRE(A.in_features == nd.out_w, "%s: modReLU width %lld after a layer of width %lld", fn,
                      (long long)A.in_features, (long long)nd.out_w);
          SMC_REQUIRE(A.param_offset >= 0 && A.param_offset + nd.out_w <= net->n_params,
                      "%s: layer %d parameters fall outside the buffer", fn, l);
          nd.act_off = A.param_offset;
        }
      }
      w = nd.out_w;
    } else {
      SMC_REQUIRE(L.kind == SMC_LAYER_MODRELU || L.kind == SMC_LAYER_ZRELU, "%s: unknown layer kind %d", fn, L.kind);
      nd.linear = false;
      nd.in_w = nd.out_w = w;
      nd.act = L.kind == SMC_LAYER_MODRELU ? ACT_MODRELU : ACT_ZRELU;
      if (nd.act == ACT_MODRELU) {
        SMC_REQUIRE(L.in_features == w, "%s: modReLU width %lld on a signal of width %lld", fn,
                    (long long)L.in_features, (long long)w);
        SMC_REQUIRE(L.param_offset >= 0 && L.param_offset + w <= net->n_params,
                    "%s: layer %d parameters fall outside the buffer", fn, l);
        nd.act_off = L.param_offset;
      }
    }
    plan->max_w = std::max(plan->max_w, w);
    plan->nodes.push_back(nd);
  }
  plan->out_w = w;
  const size_t n_nodes = plan->nodes.size();
  for (size_t k = 0; k < n_nodes; ++k) {
    Node& nd = plan->nodes[k];
    const size_t plane = static_cast<size_t>(rows) * nd.out_w * rs;
    const bool user_out = last_to_user && k + 1 == n_nodes;
    nd.out_r = user_out ? 0 : take(plane);
    nd.out_i = user_out ? 0 : take(plane);
    if (training && nd.linear && nd.act != ACT_NONE) {
      nd.pre_r = take(plane);
      nd.pre_i = take(plane);
    }
  }
  if (training) {
    const size_t plane = static_cast<size_t>(rows) * plan->max_w * rs;
    for (int k = 0; k < 2; ++k) {
      plan->g_r[k] = take(plane);
      plan->g_i[k] = take(plane);
    }
    plan->bterm = take(plane);
    const int64_t splits = std::min<int64_t>(MAX_SPLITS, (rows + SPLIT_ROWS - 1) / SPLIT_ROWS);
    int64_t max_w_elems = 0;
    for (const Node& nd : plan->nodes)
      if (nd.linear) max_w_elems = std::max(max_w_elems, nd.in_w * nd.out_w);
    plan->wpart = take(static_cast<size_t>(splits) * 2 * max_w_elems * rs);
    plan->colpart = take(static_cast<size_t>(COLSUM_PLANES) * ((rows + COLSUM_SLAB - 1) / COLSUM_SLAB) * plan->max_w * rs);
    plan->losspart = take(static_cast<size_t>((rows * plan->out_w + LOSS_CHUNK - 1) / LOSS_CHUNK) * 2 * sizeof(double));
  }
  plan->total = used + 256;
  return SMC_OK;
}

template <typename Real>
int launch_gemm(const GemmParams<Real>& p, bool conj_b, int64_t splits, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>((p.J + TJ - 1) / TJ), static_cast<unsigned>((p.I + TI - 1) / TI),
                  static_cast<unsigned>(splits));
  SMC_REQUIRE(grid.y <= 65535, "cvnn: too many rows for one launch (%lld)", (long long)p.I);
  if (conj_b)
    cgemm_kernel<Real, true><<<grid, BLK, 0, st>>>(p);
  else
    cgemm_kernel<Real, false><<<grid, BLK, 0, st>>>(p);
  SMC_LAUNCH_OK("cgemm_kernel");
  return SMC_OK;
}

// dst[z][j] = sum over rows of src[z][:, j] for n <= 3 planes (one or two fixed-order passes)
template <typename Real>
int column_sums(const Real* const* src, Real* const* dst, int n, int64_t rows, int64_t cols, Real* scratch,
                cudaStream_t st) {
  if (n == 0) return SMC_OK;
  const int64_t slabs = (rows + COLSUM_SLAB - 1) / COLSUM_SLAB;
  const unsigned gx = static_cast<unsigned>((cols + 31) / 32);
  SMC_REQUIRE(slabs <= 65535, "cvnn: too many rows for the column sum (%lld)", (long long)rows);
  ColsumJob<Real> job{};
  for (int z = 0; z < n; ++z) {
    job.src[z] = src[z];
    job.dst[z] = slabs == 1 ? dst[z] : scratch + static_cast<int64_t>(z) * slabs * cols;
  }
  colsum_kernel<Real><<<dim3(gx, static_cast<unsigned>(slabs), n), BLK, 0, st>>>(job, rows, cols, COLSUM_SLAB);
  SMC_LAUNCH_OK("colsum_kernel");
  if (slabs == 1) return SMC_OK;
  for (int z = 0; z < n; ++z) {
    job.src[z] = job.dst[z];
    job.dst[z] = dst[z];
  }
  // the slab partials of a plane form one [slabs, cols] matrix, summed as a single slab
  colsum_kernel<Real><<<dim3(gx, 1, n), BLK, 0, st>>>(job, slabs, cols, slabs);
  SMC_LAUNCH_OK("colsum_kernel");
  return SMC_OK;
}

template <typename Real>
int run_forward(const Plan& plan, const Real* params, const Real* in_r, const Real* in_i, int64_t rows, bool training,
                Real* user_r, Real* user_i, char* ws, cudaStream_t st) {
  const Real *xr = in_r, *xi = in_i;
  for (size_t k = 0; k < plan.nodes.size(); ++k) {
    const Node& nd = plan.nodes[k];
    const bool user_out = user_r != nullptr && k + 1 == plan.nodes.size();
    Real* yr = user_out ? user_r : reinterpret_cast<Real*>(ws + nd.out_r);
    Real* yi = user_out ? user_i : reinterpret_cast<Real*>(ws + nd.out_i);
    if (nd.linear) {
      GemmParams<Real> p{};
      p.ar = xr; p.ai = xi; p.sai = nd.in_w; p.sal = 1;                       // X [rows, in]
      p.br = params + nd.w_off; p.bi = p.br + nd.in_w * nd.out_w;             // W [out, in] -> B[l, j] = W[j, l]
      p.sbl = 1; p.sbj = nd.in_w;
      p.cr = yr; p.ci = yi;
      p.I = rows; p.J = nd.out_w; p.L = nd.in_w; p.l_per_split = nd.in_w;
      if (nd.has_bias) { p.bias_r = params + nd.b_off; p.bias_i = p.bias_r + nd.out_w; }
      p.act = nd.act;
      if (nd.act == ACT_MODRELU) p.act_bias = params + nd.act_off;
      if (training && nd.act != ACT_NONE) {
        p.pre_r = reinterpret_cast<Real*>(ws + nd.pre_r);
        p.pre_i = reinterpret_cast<Real*>(ws + nd.pre_i);
      }
      if (int rc = launch_gemm(p, false, 1, st)) return rc;
    } else {
      const int64_t count = rows * nd.out_w;
      act_forward_kernel<Real><<<static_cast<unsigned>((count + BLK - 1) / BLK), BLK, 0, st>>>(
          xr, xi, count, nd.out_w, nd.act, nd.act == ACT_MODRELU ? params + nd.act_off : nullptr, yr, yi);
      SMC_LAUNCH_OK("act_forward_kernel");
    }
    xr = yr;
    xi = yi;
  }
  return SMC_OK;
}

template <typename Real>
int run_loss_backward(const Plan& plan, const Real* params, const Real* in_r, const Real* in_i, const Real* targets,
                      int64_t rows, Real* grads, double* loss, char* ws, cudaStream_t st) {
  if (int rc = run_forward<Real>(plan, params, in_r, in_i, rows, true, nullptr, nullptr, ws, st)) return rc;
  const Node& last = plan.nodes.back();
  const int64_t count = rows * plan.out_w;
  const int64_t chunks = (count + LOSS_CHUNK - 1) / LOSS_CHUNK;
  int cur = 0;
  Real* gr = reinterpret_cast<Real*>(ws + plan.g_r[cur]);
  Real* gi = reinterpret_cast<Real*>(ws + plan.g_i[cur]);
  double* losspart = reinterpret_cast<double*>(ws + plan.losspart);
  mse_grad_kernel<Real><<<static_cast<unsigned>(chunks), BLK, 0, st>>>(
      reinterpret_cast<const Real*>(ws + last.out_r), reinterpret_cast<const Real*>(ws + last.out_i), targets, count,
      gr, gi, losspart);
  SMC_LAUNCH_OK("mse_grad_kernel");
  loss_finalize_kernel<<<1, BLK, 0, st>>>(losspart, chunks, count, loss);
  SMC_LAUNCH_OK("loss_finalize_kernel");

  Real* bterm = reinterpret_cast<Real*>(ws + plan.bterm);
  Real* colpart = reinterpret_cast<Real*>(ws + plan.colpart);
  Real* wpart = reinterpret_cast<Real*>(ws + plan.wpart);
  for (size_t k = plan.nodes.size(); k-- > 0;) {
    const Node& nd = plan.nodes[k];
    const Real* xr = k == 0 ? in_r : reinterpret_cast<const Real*>(ws + plan.nodes[k - 1].out_r);
    const Real* xi = k == 0 ? in_i : reinterpret_cast<const Real*>(ws + plan.nodes[k - 1].out_i);
    const int64_t n_out = rows * nd.out_w;
    if (nd.act != ACT_NONE) {
      // pre-activation: saved by the fused epilogue, or this node's input for a lone activation
      const Real* zr = nd.linear ? reinterpret_cast<const Real*>(ws + nd.pre_r) : xr;
      const Real* zi = nd.linear ? reinterpret_cast<const Real*>(ws + nd.pre_i) : xi;
      act_backward_kernel<Real><<<static_cast<unsigned>((n_out + BLK - 1) / BLK), BLK, 0, st>>>(
          gr, gi, zr, zi, n_out, nd.out_w, nd.act, nd.act == ACT_MODRELU ? params + nd.act_off : nullptr, bterm);
      SMC_LAUNCH_OK("act_backward_kernel");
    }
    {  // bias gradients: column sums of gr, gi (ComplexLinear bias) and of the modReLU bias term, one launch
      const Real* src[COLSUM_PLANES];
      Real* dst[COLSUM_PLANES];
      int n = 0;
      if (nd.linear && nd.has_bias) {
        src[n] = gr; dst[n++] = grads + nd.b_off;
        src[n] = gi; dst[n++] = grads + nd.b_off + nd.out_w;
      }
      if (nd.act == ACT_MODRELU) { src[n] = bterm; dst[n++] = grads + nd.act_off; }
      if (int rc = column_sums<Real>(src, dst, n, rows, nd.out_w, colpart, st)) return rc;
    }
    if (!nd.linear) continue;  // (gr, gi) is now the gradient wrt this node's input, same width
    {  // dW[o, k] = sum_m G[m, o] conj(X[m, k]):  dA = gr^T xr + gi^T xi ; dB = gi^T xr - gr^T xi
      const int64_t splits = std::min<int64_t>(MAX_SPLITS, (rows + SPLIT_ROWS - 1) / SPLIT_ROWS);
      const int64_t welems = nd.in_w * nd.out_w;
      GemmParams<Real> p{};
      p.ar = gr; p.ai = gi; p.sai = 1; p.sal = nd.out_w;          // A[i = o, l = m] = G[m, o]
      p.br = xr; p.bi = xi; p.sbl = nd.in_w; p.sbj = 1;           // B[l = m, j = k] = X[m, k]
      p.I = nd.out_w; p.J = nd.in_w; p.L = rows;
      p.l_per_split = (rows + splits - 1) / splits;
      Real* d_a = grads + nd.w_off;
      Real* d_b = d_a + welems;
      if (splits == 1) {
        p.cr = d_a; p.ci = d_b; p.split_stride = 0;
        if (int rc = launch_gemm(p, true, 1, st)) return rc;
      } else {
        p.cr = wpart; p.ci = wpart + welems; p.split_stride = 2 * welems;
        if (int rc = launch_gemm(p, true, splits, st)) return rc;
        reduce_split_kernel<Real><<<static_cast<unsigned>((welems + BLK - 1) / BLK), BLK, 0, st>>>(wpart, splits, welems, d_a, d_b);
        SMC_LAUNCH_OK("reduce_split_kernel");
      }
    }
    if (k > 0) {  // dX = G conj(W):  gxr = gr A + gi B ; gxi = gi A - gr B
      GemmParams<Real> p{};
      p.ar = gr; p.ai = gi; p.sai = nd.out_w; p.sal = 1;          // A[i = m, l = o]
      p.br = params + nd.w_off; p.bi = p.br + nd.in_w * nd.out_w; // B[l = o, j = k] = W[o, k]
      p.sbl = nd.in_w; p.sbj = 1;
      cur ^= 1;
      Real* nr = reinterpret_cast<Real*>(ws + plan.g_r[cur]);
      Real* ni = reinterpret_cast<Real*>(ws + plan.g_i[cur]);
      p.cr = nr; p.ci = ni;
      p.I = rows; p.J = nd.in_w; p.L = nd.out_w; p.l_per_split = nd.out_w;
      if (int rc = launch_gemm(p, true, 1, st)) return rc;
      gr = nr;
      gi = ni;
    }
  }
  return SMC_OK;
}

int check_buffers(const char* fn, const Plan& plan, const void* ws, size_t ws_bytes) {
  SMC_REQUIRE(ws != nullptr || plan.total <= 256, "%s: workspace is NULL", fn);
  if (ws_bytes < plan.total)
    return set_error(SMC_EWORKSPACE, "%s: workspace of %zu bytes, %zu needed", fn, ws_bytes, plan.total);
  SMC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "%s: workspace must be 16-byte aligned", fn);
  return SMC_OK;
}

}  // namespace
}  // namespace smc

using namespace smc;

extern "C" size_t smc_cvnn_workspace_bytes(const smc_cvnn_net* net, int64_t rows, int training) {
  Plan plan;
  if (build_plan("smc_cvnn_workspace_bytes", net, rows, training != 0, training == 0, &plan) != SMC_OK) return 0;
  return plan.total;
}

extern "C" int64_t smc_cvnn_output_width(const smc_cvnn_net* net) {
  Plan plan;
  if (build_plan("smc_cvnn_output_width", net, 1, false, true, &plan) != SMC_OK) return -1;
  return plan.out_w;
}

extern "C" int smc_cvnn_forward(const smc_cvnn_net* net, const void* params, const void* in_r, const void* in_i,
                                int64_t rows, void* out_r, void* out_i, void* workspace, size_t workspace_bytes,
                                void* stream) {
  const char* fn = "smc_cvnn_forward";
  clear_error();
  Plan plan;
  if (int rc = build_plan(fn, net, rows, false, true, &plan)) return rc;
  SMC_REQUIRE(params && in_r && in_i && out_r && out_i, "%s: NULL buffer", fn);
  if (int rc = check_buffers(fn, plan, workspace, workspace_bytes)) return rc;
  char* ws = static_cast<char*>(workspace);
  if (net->dtype == SMC_F32)
    return run_forward<float>(plan, static_cast<const float*>(params), static_cast<const float*>(in_r),
                              static_cast<const float*>(in_i), rows, false, static_cast<float*>(out_r),
                              static_cast<float*>(out_i), ws, as_stream(stream));
  return run_forward<double>(plan, static_cast<const double*>(params), static_cast<const double*>(in_r),
                             static_cast<const double*>(in_i), rows, false, static_cast<double*>(out_r),
                             static_cast<double*>(out_i), ws, as_stream(stream));
}

extern "C" int smc_cvnn_loss_backward(const smc_cvnn_net* net, const void* params, const void* in_r, const void* in_i,
                                      const void* targets, int64_t rows, void* grads, double* loss, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  const char* fn = "smc_cvnn_loss_backward";
  clear_error();
  Plan plan;
  if (int rc = build_plan(fn, net, rows, true, false, &plan)) return rc;
  SMC_REQUIRE(params && in_r && in_i && targets && grads && loss, "%s: NULL buffer", fn);
  if (int rc = check_buffers(fn, plan, workspace, workspace_bytes)) return rc;
  char* ws = static_cast<char*>(workspace);
  if (net->dtype == SMC_F32)
    return run_loss_backward<float>(plan, static_cast<const float*>(params), static_cast<const float*>(in_r),
                                    static_cast<const float*>(in_i), static_cast<const float*>(targets), rows,
                                    static_cast<float*>(grads), loss, ws, as_stream(stream));
  return run_loss_backward<double>(plan, static_cast<const double*>(params), static_cast<const double*>(in_r),
                                   static_cast<const double*>(in_i), static_cast<const double*>(targets), rows,
                                   static_cast<double*>(grads), loss, ws, as_stream(stream));
}

extern "C" int smc_adam_step(void* params, const void* grads, void* exp_avg, void* exp_avg_sq, int64_t n, int dtype,
                             int64_t* step, const smc_adam_args* h, void* stream) {
  const char* fn = "smc_adam_step";
  clear_error();
  SMC_REQUIRE(params && grads && exp_avg && exp_avg_sq && step && h, "%s: NULL argument", fn);
  SMC_REQUIRE(n > 0, "%s: n must be > 0", fn);
  SMC_REQUIRE(dtype == SMC_F32 || dtype == SMC_F64, "%s: invalid dtype %d", fn, dtype);
  SMC_REQUIRE(h->lr >= 0 && h->beta1 >= 0 && h->beta1 < 1 && h->beta2 >= 0 && h->beta2 < 1 && h->eps >= 0,
              "%s: invalid hyper-parameters", fn);
  cudaStream_t st = as_stream(stream);
  const unsigned grid = static_cast<unsigned>((n + BLK - 1) / BLK);
  if (dtype == SMC_F32)
    adam_kernel<float><<<grid, BLK, 0, st>>>(static_cast<float*>(params), static_cast<const float*>(grads),
                                             static_cast<float*>(exp_avg), static_cast<float*>(exp_avg_sq), n, step,
                                             h->lr, h->beta1, h->beta2, h->eps);
  else
    adam_kernel<double><<<grid, BLK, 0, st>>>(static_cast<double*>(params), static_cast<const double*>(grads),
                                              static_cast<double*>(exp_avg), static_cast<double*>(exp_avg_sq), n, step,
                                              h->lr, h->beta1, h->beta2, h->eps);
  SMC_LAUNCH_OK("adam_kernel");
  bump_step_kernel<<<1, 1, 0, st>>>(step);
  SMC_LAUNCH_OK("bump_step_kernel");
  return SMC_OK;
}

extern "C" int smc_cvnn_train_step(const smc_cvnn_net* net, void* params, void* grads, void* exp_avg, void* exp_avg_sq,
                                   int64_t* step, const smc_adam_args* h, const void* in_r, const void* in_i,
                                   const void* targets, int64_t rows, double* loss, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (int rc = smc_cvnn_loss_backward(net, params, in_r, in_i, targets, rows, grads, loss, workspace, workspace_bytes, stream))
    return rc;
  SMC_REQUIRE(net->n_params > 0, "smc_cvnn_train_step: the network has no parameters");
  return smc_adam_step(params, grads, exp_avg, exp_avg_sq, net->n_params, net->dtype, step, h, stream);
}
