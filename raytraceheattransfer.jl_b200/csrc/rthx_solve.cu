// rthx_solve.cu — grey GERT equilibrium solve on the device.
//
// Consumer of the exchange-factor matrix the tracer produces.  Restates the linear-algebra core of
// src/HeatTransfer/equilibrium/equilibriumGrey2D.jl:
//   M = I - Diagonal(coeff) * F'            :148-149   (never formed: M x = x - coeff .* (F' x))
//   gmres!(..., memory = 50, restart = true, rtol = 1e-12)   :152-155   -> restarted GMRES(50), stop at atol + rtol*|r0|
//   g_i = sum_k F[k,i] j[k]                 :168-194   (the receiver-indexed incident power; r = b g, Abs = (1-b) g on the host)
// The reference takes the dense `M \ h` branch (:157) for a dense F; here both cases run the same Krylov iteration,
// whose answer agrees with the LU solve to the requested relative residual (1e-12).
//
// The one heavy operation is y = F' x on a dense n x n Float64 matrix (cfg3: n = 10 605, 900 MB): one read of F per
// Krylov step, HBM-bound at 8 n^2 bytes.  F is read where it already lives — the padded row-major matrix the
// device-side smoothing (rthx_smooth_F) leaves behind — so trace -> smooth -> solve never moves F over PCIe.
//   dense row-major F  : y_j = sum_i F[i][j] x[i]   matvec_t_partial_kernel (column strips x row tiles, double2 loads,
//                                                   fixed summation order) + matvec_finish_kernel (tile sum + epilogue)
//   dense col-major F  : the memory holds A = F' row-major, y = A x: matvec_rows_kernel (one block per row)
//   CSC sparse F       : y_j = sum over column j    matvec_csc_kernel (one warp per column, Julia's SparseMatrixCSC)
// The Krylov vectors (n doubles each) are tiny next to F; their kernels are written for determinism, not speed.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "rthx_internal.h"

namespace rthx {

// ---------------------------------------------------------------------------------------------------------------
// y = F' x
// ---------------------------------------------------------------------------------------------------------------
constexpr int MV_THREADS = 256;
constexpr int MV_UNROLL = 8;

// part[tile][j] = sum_{i in row tile} F[i][j] x[i]; rows are ld doubles apart (ld even, rows 16-byte aligned).
__global__ void __launch_bounds__(MV_THREADS) matvec_t_partial_kernel(const double* __restrict__ F, size_t ld, int n, int rows_per_tile,
                                                                      const double* __restrict__ x, double* __restrict__ part) {
  const size_t ld2 = ld >> 1;
  const size_t j2 = (size_t)blockIdx.x * MV_THREADS + threadIdx.x;
  if (j2 >= ld2) return;
  const int r0 = blockIdx.y * rows_per_tile;
  const int r1 = min(n, r0 + rows_per_tile);
  const double2* p = reinterpret_cast<const double2*>(F) + (size_t)r0 * ld2 + j2;
  double ax = 0.0, ay = 0.0;
  int i = r0;
  for (; i + MV_UNROLL <= r1; i += MV_UNROLL) {
    double2 v[MV_UNROLL];
    double xi[MV_UNROLL];
#pragma unroll
    for (int u = 0; u < MV_UNROLL; ++u) { v[u] = __ldg(p + (size_t)u * ld2); xi[u] = x[i + u]; }
#pragma unroll
    for (int u = 0; u < MV_UNROLL; ++u) { ax = fma(v[u].x, xi[u], ax); ay = fma(v[u].y, xi[u], ay); }
    p += (size_t)MV_UNROLL * ld2;
  }
  for (; i < r1; ++i) {
    const double2 v = __ldg(p);
    const double xi = x[i];
    ax = fma(v.x, xi, ax); ay = fma(v.y, xi, ay);
    p += ld2;
  }
  reinterpret_cast<double2*>(part + (size_t)blockIdx.y * ld)[j2] = make_double2(ax, ay);
}

// y_j = sum_tile part[tile][j] (ascending tile order); out = y (apply_m == 0) or x - coeff .* y (apply_m != 0)
__global__ void __launch_bounds__(256) matvec_finish_kernel(const double* __restrict__ part, int tiles, size_t ld, int n, const double* __restrict__ x,
                                                            const double* __restrict__ coeff, int apply_m, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double y = 0.0;
  for (int t = 0; t < tiles; ++t) y += part[(size_t)t * ld + j];
  out[j] = apply_m ? fma(-coeff[j], y, x[j]) : y;
}

__device__ __forceinline__ double block_sum256(double s, double* red) {
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double a = 0.0;
  if (threadIdx.x == 0) for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += red[k];
  return a;
}

// A = F' stored row-major (a Julia `Matrix` F handed over as it lies in memory): y_i = sum_j A[i][j] x[j], one block per row.
__global__ void __launch_bounds__(MV_THREADS) matvec_rows_kernel(const double* __restrict__ A, size_t ld, int n, const double* __restrict__ x,
                                                                 const double* __restrict__ coeff, int apply_m, double* __restrict__ out) {
  __shared__ double red[32];
  const size_t row = blockIdx.x;
  const double2* a2 = reinterpret_cast<const double2*>(A + row * ld);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  double s = 0.0;
  const int n2 = n >> 1;
  for (int j = threadIdx.x; j < n2; j += MV_THREADS) {
    const double2 v = __ldg(a2 + j);
    const double2 xv = x2[j];
    s = fma(v.x, xv.x, fma(v.y, xv.y, s));
  }
  if ((n & 1) && threadIdx.x == 0) s = fma(A[row * ld + (n - 1)], x[n - 1], s);
  const double y = block_sum256(s, red);
  if (threadIdx.x == 0) out[row] = apply_m ? fma(-coeff[row], y, x[row]) : y;
}

// CSC of F (Julia SparseMatrixCSC fields, 0-based): column j lists the senders k with F[k,j] != 0; one warp per column.
__global__ void __launch_bounds__(256) matvec_csc_kernel(const long long* __restrict__ colptr, const int* __restrict__ rowval, const double* __restrict__ nzval,
                                                         int n, const double* __restrict__ x, const double* __restrict__ coeff, int apply_m,
                                                         double* __restrict__ out) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  double s = 0.0;
  for (long long k = colptr[j] + lane; k < colptr[j + 1]; k += 32) s = fma(nzval[k], x[rowval[k]], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = apply_m ? fma(-coeff[j], s, x[j]) : s;
}

// ---------------------------------------------------------------------------------------------------------------
// Krylov vector kernels (n doubles each; deterministic reductions)
// ---------------------------------------------------------------------------------------------------------------
// d[i] (+)= v_i . w for i < k, one block per basis vector
__global__ void __launch_bounds__(256) dots_kernel(const double* __restrict__ V, size_t nl, int n, const double* __restrict__ w, double* __restrict__ d, int accumulate) {
  __shared__ double red[32];
  const double* v = V + (size_t)blockIdx.x * nl;
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) s = fma(v[j], w[j], s);
  const double a = block_sum256(s, red);
  if (threadIdx.x == 0) d[blockIdx.x] = accumulate ? d[blockIdx.x] + a : a;
}

// w -= sum_{i<k} c[i] v_i
__global__ void __launch_bounds__(256) update_kernel(const double* __restrict__ V, size_t nl, int n, int k, const double* __restrict__ c, double* __restrict__ w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = w[j];
  for (int i = 0; i < k; ++i) s = fma(-c[i], V[(size_t)i * nl + j], s);
  w[j] = s;
}

// x += sum_{i<k} y[i] v_i
__global__ void __launch_bounds__(256) combine_kernel(const double* __restrict__ V, size_t nl, int n, int k, const double* __restrict__ y, double* __restrict__ x) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = x[j];
  for (int i = 0; i < k; ++i) s = fma(y[i], V[(size_t)i * nl + j], s);
  x[j] = s;
}

// out[0] = |w|^2, single block (fixed order)
__global__ void __launch_bounds__(1024) norm2_kernel(const double* __restrict__ w, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) s = fma(w[j], w[j], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double a = 0.0; for (int k = 0; k < 32; ++k) a += red[k]; out[0] = a; }
}

// v = w / sqrt(norm2[0])  (v = 0 on a happy breakdown)
__global__ void __launch_bounds__(256) normalise_kernel(const double* __restrict__ w, int n, const double* __restrict__ norm2, double* __restrict__ v) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double nn = norm2[0];
  v[j] = nn > 0.0 ? w[j] / sqrt(nn) : 0.0;
}

// r = rhs - w   (w = M x)
__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ rhs, const double* __restrict__ w, int n, double* __restrict__ r) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) r[j] = rhs[j] - w[j];
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
// Row tiling of the dense transposed product: column strips x row tiles fill exactly ONE wave of resident blocks
// (SM count x occupancy of matvec_t_partial_kernel), so no block waits for a slot and no tail wave runs half empty.
// The tiling fixes the summation order; it depends on n and the device model only (results are bit-reproducible on B200).
static int mv_row_tiles(int n, size_t ld) {
  static int slots = 0;
  if (slots == 0) {
    int dev = 0, sms = 148, occ = 6;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, matvec_t_partial_kernel, MV_THREADS, 0) != cudaSuccess || occ < 1) occ = 6;
    slots = sms * occ;
  }
  const int col_blocks = (int)(((ld >> 1) + MV_THREADS - 1) / MV_THREADS);
  int tiles = slots / col_blocks;
  tiles = std::max(1, std::min(tiles, std::min(256, (n + 15) / 16)));
  return tiles;
}

size_t solve_part_doubles(int n, size_t ld) { return (size_t)mv_row_tiles(n, ld) * ld; }

static cudaError_t matvec(const SolveMatrix& A, int n, const double* x, const double* coeff, int apply_m, double* out, double* part, cudaStream_t st,
                          int* launches) {
  if (A.kind == 0) {
    const int tiles = mv_row_tiles(n, A.ld);
    const int rpt = (n + tiles - 1) / tiles;
    const dim3 grid((unsigned)(((A.ld >> 1) + MV_THREADS - 1) / MV_THREADS), (unsigned)tiles);
    matvec_t_partial_kernel<<<grid, MV_THREADS, 0, st>>>(A.dense, A.ld, n, rpt, x, part);
    matvec_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, tiles, A.ld, n, x, coeff, apply_m, out);
    *launches += 2;
  } else if (A.kind == 1) {
    matvec_rows_kernel<<<n, MV_THREADS, 0, st>>>(A.dense, A.ld, n, x, coeff, apply_m, out);
    *launches += 1;
  } else {
    matvec_csc_kernel<<<(n + 7) / 8, 256, 0, st>>>(A.colptr, A.rowval, A.nzval, n, x, coeff, apply_m, out);
    *launches += 1;
  }
  return cudaGetLastError();
}

long long matvec_bytes(const SolveMatrix& A, int n) {
  if (A.kind == 2) return A.nnz * 12 + (long long)(n + 1) * 8;
  return 8ll * n * n;
}

// Restarted GMRES(m) for (I - diag(coeff) F') x = rhs, x0 = 0, classical Gram-Schmidt applied twice, Givens rotations on
// the host.  `buf` is device scratch of solve_vec_doubles(n, nl, m) doubles; coeff / rhs are already on the device inside it.
// Layout (nl = padded vector length): V[(m+1)] | w | x | r | coeff | rhs | d[2(m+1)+2].
size_t solve_vec_doubles(size_t nl, int m) { return (size_t)(m + 6) * nl + (size_t)(2 * (m + 1) + 2 + 14); }

cudaError_t run_gmres(const SolveMatrix& A, int n, size_t nl, int m, int max_iters, double rtol, double atol, double* buf, double* part, double* j_host,
                      double* g_host, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, SolveResult* out) {
  cudaError_t e;
  double* V = buf;
  double* w = V + (size_t)(m + 1) * nl;
  double* x = w + nl;
  double* r = x + nl;
  double* coeff = r + nl;
  double* rhs = coeff + nl;
  double* d = rhs + nl;                 // [0..m] pass-1+2 coefficients, [m+1] |w|^2, [m+2 ..] y upload area
  double* yv = d + (m + 2);
  int launches = 0, matvecs = 0;
  const int vb = (n + 255) / 256;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), hcol(m + 2), y(m);
  auto sync_copy = [&](double* dst, const double* src, size_t cnt) -> cudaError_t {
    cudaError_t ee = cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (ee != cudaSuccess) return ee;
    return cudaStreamSynchronize(st);
  };
  if ((e = cudaEventRecord(e0, st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(x, 0, nl * sizeof(double), st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(r, rhs, nl * sizeof(double), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  double nrm2 = 0;
  norm2_kernel<<<1, 1024, 0, st>>>(r, n, d + m + 1); ++launches;
  if ((e = sync_copy(&nrm2, d + m + 1, 1)) != cudaSuccess) return e;
  double beta = std::sqrt(nrm2);
  const double rhs_norm = beta;
  const double eps = atol + rtol * beta;                      // Krylov.jl: ε = atol + rtol * ‖r0‖
  int iters = 0, restarts = 0, stalled = 0;
  bool converged = beta <= eps;
  double res = beta;
  while (!converged && iters < max_iters) {
    const double beta_cycle = beta;
    // v_0 = r / beta
    normalise_kernel<<<vb, 256, 0, st>>>(r, n, d + m + 1, V); ++launches;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int k = 0;
    for (; k < m && iters < max_iters; ++k) {
      double* vk = V + (size_t)k * nl;
      if ((e = matvec(A, n, vk, coeff, 1, w, part, st, &launches)) != cudaSuccess) return e;
      ++matvecs;
      // classical Gram-Schmidt, twice (as stable as the modified form, two reductions instead of k+1)
      dots_kernel<<<k + 1, 256, 0, st>>>(V, nl, n, w, d, 0);
      update_kernel<<<vb, 256, 0, st>>>(V, nl, n, k + 1, d, w);
      dots_kernel<<<k + 1, 256, 0, st>>>(V, nl, n, w, yv, 0);
      update_kernel<<<vb, 256, 0, st>>>(V, nl, n, k + 1, yv, w);
      norm2_kernel<<<1, 1024, 0, st>>>(w, n, d + m + 1);
      normalise_kernel<<<vb, 256, 0, st>>>(w, n, d + m + 1, V + (size_t)(k + 1) * nl);
      launches += 6;
      // one copy: d[0..m], |w|^2, second-pass coefficients
      std::vector<double> tmp((size_t)(m + 2) + (k + 1));
      if ((e = sync_copy(tmp.data(), d, tmp.size())) != cudaSuccess) return e;
      for (int i = 0; i <= k; ++i) hcol[i] = tmp[i] + tmp[(size_t)(m + 2) + i];
      hcol[k + 1] = std::sqrt(tmp[m + 1]);
      // Givens: apply the previous rotations to the new column, then annihilate h[k+1]
      for (int i = 0; i < k; ++i) {
        const double t = cs[i] * hcol[i] + sn[i] * hcol[i + 1];
        hcol[i + 1] = -sn[i] * hcol[i] + cs[i] * hcol[i + 1];
        hcol[i] = t;
      }
      const double rr = std::hypot(hcol[k], hcol[k + 1]);
      cs[k] = rr > 0 ? hcol[k] / rr : 1.0;
      sn[k] = rr > 0 ? hcol[k + 1] / rr : 0.0;
      hcol[k] = rr;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      for (int i = 0; i <= k; ++i) H[(size_t)i * m + k] = hcol[i];
      ++iters;
      res = std::fabs(g[k + 1]);
      if (res <= eps || hcol[k + 1] == 0.0) { ++k; break; }
    }
    // y = H^-1 g (upper triangular, k columns), x += V y
    for (int i = k - 1; i >= 0; --i) {
      double s = g[i];
      for (int c = i + 1; c < k; ++c) s -= H[(size_t)i * m + c] * y[c];
      y[i] = H[(size_t)i * m + i] != 0.0 ? s / H[(size_t)i * m + i] : 0.0;
    }
    if (k > 0) {
      if ((e = cudaMemcpyAsync(yv, y.data(), sizeof(double) * k, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
      combine_kernel<<<vb, 256, 0, st>>>(V, nl, n, k, yv, x); ++launches;
    }
    // true residual for the restart / the final report
    if ((e = matvec(A, n, x, coeff, 1, w, part, st, &launches)) != cudaSuccess) return e;
    ++matvecs;
    residual_kernel<<<vb, 256, 0, st>>>(rhs, w, n, r);
    norm2_kernel<<<1, 1024, 0, st>>>(r, n, d + m + 1);
    launches += 2;
    if ((e = sync_copy(&nrm2, d + m + 1, 1)) != cudaSuccess) return e;
    beta = std::sqrt(nrm2);
    res = beta;
    if (beta <= eps) { converged = true; break; }
    if (iters >= max_iters) break;
    // a purely relative target (atol = 0) can sit below the rounding floor of the products: two restart cycles in a row that
    // no longer reduce the true residual end the solve as "not converged" instead of spinning to max_iters
    stalled = beta > 0.9 * beta_cycle ? stalled + 1 : 0;
    if (stalled >= 2) break;
    ++restarts;
    // Rounding can leave the recurrence residual below eps while the true one sits a hair above it; one more
    // cycle from the true residual settles that, so no special case is needed here.
  }
  // g = F' j
  if ((e = matvec(A, n, x, coeff, 0, w, part, st, &launches)) != cudaSuccess) return e;
  ++matvecs;
  if ((e = cudaMemcpyAsync(j_host, x, sizeof(double) * n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if (g_host && (e = cudaMemcpyAsync(g_host, w, sizeof(double) * n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(e1, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  out->iterations = iters; out->restarts = restarts; out->launches = launches; out->converged = converged ? 1 : 0; out->matvecs = matvecs;
  out->residual = res; out->rhs_norm = rhs_norm; out->total_ms = ms; out->matvec_bytes = matvec_bytes(A, n);
  return cudaGetLastError();
}

// timing helper for the roofline: `reps` back-to-back products y = F' x, returns ms per product
cudaError_t time_matvec(const SolveMatrix& A, int n, size_t nl, int m, double* buf, double* part, int reps, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1,
                        double* ms_per_pass) {
  cudaError_t e;
  double* w = buf + (size_t)(m + 1) * nl;
  double* x = w + nl;
  double* coeff = x + 2 * nl;
  int launches = 0;
  if ((e = matvec(A, n, x, coeff, 1, w, part, st, &launches)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(e0, st)) != cudaSuccess) return e;
  for (int i = 0; i < reps; ++i)
    if ((e = matvec(A, n, x, coeff, 1, w, part, st, &launches)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(e1, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_pass = ms / reps;
  return cudaGetLastError();
}

}  // namespace rthx
