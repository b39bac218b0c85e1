// Stage 4 -- QR measurement compression of the stacked, gated MSCKF Jacobian.
//
// Reference: the `H.rows() > H.cols()` branch of OrcVIO::measurementUpdate_msckf
// (src/orcvio.cpp:1664-1683) and of removeLostFeatures (:2532-2552): SuiteSparseQR
// (natural ordering) on H.sparseView(), H_thin = (Q^T H).topRows(L + 6N), r_thin likewise.
// Any orthogonal Q yields the same posterior, so this file restates the step as a
// two-level Householder factorisation that exploits what SPQR exploits -- the block
// sparsity of H -- in a shape that fits the GPU:
//
//   level 1  k_qr_tiles   one CTA per row tile.  Features are sorted by their first clone,
//            so a tile only touches the clone window [c0, c1); the CTA assembles the
//            (rows x 6(c1-c0)) window densely in shared memory from the compact per-feature
//            blocks written by the gate kernel, triangularises it and emits at most
//            6(c1-c0) rows.  Tiles of all filters of the batch run concurrently.
//   level 2  k_qr_chain   one CTA per filter sweeps the clone blocks left to right,
//            merging the tile factors into a banded front (carry rows + new tile),
//            re-triangularising it and emitting 6 finished rows of R per clone block.
//
// The result R (6N x 6N, upper triangular over the clone columns; the 22 leading IMU
// columns of an MSCKF Jacobian are structurally zero) and r_thin = Q^T r feed the EKF
// update.  When the stack has fewer rows than columns the missing rows of R are zero,
// which is exactly the reference's "no compression" case up to an orthogonal transform.
//
// Roofline: FP64 FMA bound in principle (2 M w^2 flop, w = window width), in practice
// bound by the latency of the w dependent reflections per front; see DESIGN.md.
#include "kernels.h"

namespace ob {

// Row-circular, column-absolute accessor used for both the tile and the chain front.
struct Front {
  double* base;
  int ld;        // leading dimension (doubles)
  int rcap;      // number of physical rows (circular)
  int row0;      // physical index of logical row 0
  int cmask;     // physical column = j & cmask (power-of-two ring) ; ~0 = no wrap
  int rhs;       // physical column of the right-hand side
  __device__ __forceinline__ double& at(int i, int j) const {
    int pr = row0 + i;
    if (pr >= rcap) pr -= rcap;
    const int pc = (j < 0) ? rhs : (j & cmask);   // j = -1 addresses the right-hand side
    return base[(size_t)pr * ld + pc];
  }
};
#define FRONT_RHS (-1)

// CTA-wide Householder triangularisation of the logical rows [0, m) of `f` over the columns
// [c0, c1) plus the right-hand-side column.  On exit the block is upper
// trapezoidal (entries below the diagonal are zeroed).  All threads of the CTA call this.
__device__ void cta_householder(const Front& f, int m, int c0, int c1, double* sh, int nelim = -1) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int ncol = c1 - c0;
  const int steps = min(nelim >= 0 ? nelim : ncol, m - 1);
  for (int k = 0; k < steps; ++k) {
    const int ck = c0 + k;
    // sigma = sum_{i>k} a(i,ck)^2 : warp partial sums -> shared
    double part = 0.0;
    for (int i = k + 1 + tid; i < m; i += nt) {
      double v = f.at(i, ck);
      part += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) sh[warp] = part;
    __syncthreads();
    double sig = 0.0;
    for (int w = 0; w < nw; ++w) sig += sh[w];
    const double akk = f.at(k, ck);
    double t = 0.0, v0 = 1.0, mu = akk;
    if (sig > 0.0) {
      mu = sqrt(akk * akk + sig);
      v0 = (akk <= 0.0) ? (akk - mu) : (-sig / (akk + mu));
      t = 2.0 * v0 * v0 / (sig + v0 * v0);
    }
    __syncthreads();                       // everyone has read sh[] and a(k,ck)
    if (sig > 0.0) {
      const double iv0 = 1.0 / v0;
      for (int i = k + 1 + tid; i < m; i += nt) f.at(i, ck) *= iv0;   // v below the diagonal
      if (tid == 0) f.at(k, ck) = mu;
      __syncthreads();
      // apply to the remaining columns (one warp per column, rows over lanes)
      const int ntrail = (c1 - ck - 1) + 1;
      for (int jj = warp; jj < ntrail; jj += nw) {
        const int cj = (jj < c1 - ck - 1) ? (ck + 1 + jj) : FRONT_RHS;
        double dot = 0.0;
        for (int i = k + 1 + lane; i < m; i += 32) dot += f.at(i, ck) * f.at(i, cj);
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        dot += f.at(k, cj);
        const double s = t * dot;
        __syncwarp();
        for (int i = k + 1 + lane; i < m; i += 32) f.at(i, cj) -= s * f.at(i, ck);
        if (lane == 0) f.at(k, cj) -= s;
      }
      __syncthreads();
      for (int i = k + 1 + tid; i < m; i += nt) f.at(i, ck) = 0.0;
    }
    // (next iteration touches column ck+1 first; the zeroing above cannot race with it)
  }
  __syncthreads();
}

// ---------------------------------------------------------------- level 1: row tiles
__global__ void __launch_bounds__(QR_THREADS) k_qr_tiles(QrArgs a) {
  extern __shared__ double smem[];
  __shared__ double red[32];
  __shared__ int s_rows;
  const Tile tl = a.tiles[blockIdx.x];
  const int W = 6 * (tl.c1_blk - tl.c0_blk);
  const int ld = W + 2;                    // W columns + rhs, padded
  const int tid = threadIdx.x, nt = blockDim.x;
  // zero the tile
  for (int e = tid; e < tl.rows * ld; e += nt) smem[e] = 0.0;
  if (tid == 0) s_rows = 0;
  __syncthreads();
  // assemble: passing candidates only, packed contiguously (thread 0 assigns row bases)
  __shared__ int rowbase[QR_THREADS];
  const int nc = tl.cand_end - tl.cand_begin;      // host guarantees nc <= QR_THREADS
  if (tid == 0) {
    int rws = 0;
    for (int q = 0; q < nc; ++q) {
      const int c = tl.cand_begin + q;
      const bool ok = (a.status[c] & ST_GATE_PASS) != 0;
      rowbase[q] = ok ? rws : -1;
      if (ok) rws += 2 * a.cand[c].jac_m - 3;
    }
    s_rows = rws;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int q = warp; q < nc; q += nw) {
    if (rowbase[q] < 0) continue;
    const int c = tl.cand_begin + q;
    const Cand cd = a.cand[c];
    const int r = 2 * cd.jac_m - 3;
    const int w = 6 * (cd.e_blk - cd.s_blk + 1);
    const int coff = 6 * (cd.s_blk - tl.c0_blk);
    const double* hb = a.hblk + cd.hblk_off;
    for (int e = lane; e < r * w; e += 32) {
      int i = e / w, j = e % w;
      smem[(size_t)(rowbase[q] + i) * ld + coff + j] = hb[e];
    }
    for (int i = lane; i < r; i += 32) smem[(size_t)(rowbase[q] + i) * ld + W] = a.rblk[cd.row_off + i];
  }
  __syncthreads();
  const int m = s_rows;
  Front f{smem, ld, tl.rows > 0 ? tl.rows : 1, 0, 0x7fffffff, W};
  cta_householder(f, m, 0, W, red);
  // emit min(m, W) rows (zero padded to W rows) as W x (W+1)
  double* out = a.tile_out + tl.out_off;
  const int keep = min(m, W);
  for (int e = tid; e < W * (W + 1); e += nt) {
    int i = e / (W + 1), j = e % (W + 1);
    out[e] = (i < keep) ? smem[(size_t)i * ld + j] : 0.0;
  }
}

// ---------------------------------------------------------------- level 2: banded chain
__global__ void __launch_bounds__(QR_THREADS) k_qr_chain(QrArgs a, int front_rows_cap, int wcap, int use_global) {
  extern __shared__ double smem[];
  __shared__ double red[32];
  const int fi = blockIdx.x;
  const FilterWork fw = a.fw[fi];
  if (!fw.active) return;
  const int n = 6 * fw.N;
  const int ld = wcap + 1;                // ring of wcap clone columns + rhs at column wcap
  double* fb = use_global ? (a.front_scratch + (size_t)fi * a.front_stride) : smem;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* Rm = a.Rm + (size_t)fi * a.r_stride;
  double* rth = a.rthin + (size_t)fi * a.ldr;
  const int ldr = a.ldr;
  for (int e = tid; e < front_rows_cap * ld; e += nt) fb[e] = 0.0;
  __syncthreads();
  Front f{fb, ld, front_rows_cap, 0, wcap - 1, wcap};
  int s = 0;            // current clone block (origin of the front)
  int nrows = 0;        // live rows in the front
  int cend = 0;         // one past the last column touched by the front
  int t = fw.tile_begin;
  while (true) {
    const int next_c0 = (t < fw.tile_end) ? a.tiles[t].c0_blk : fw.N;
    // emit finished clone blocks s .. next_c0-1
    while (s < next_c0) {
      for (int e = tid; e < 6 * n; e += nt) {
        int i = e / n, j = e % n;
        double v = 0.0;
        if (i < nrows && j >= 6 * s && j < cend) v = f.at(i, j);
        Rm[(size_t)(6 * s + i) * ldr + j] = v;
      }
      if (tid < 6) rth[6 * s + tid] = (tid < nrows) ? f.at(tid, FRONT_RHS) : 0.0;
      __syncthreads();
      // drop the emitted rows (and clear them for reuse)
      const int drop = min(6, nrows);
      for (int e = tid; e < drop * ld; e += nt) f.at(e / ld, e % ld) = 0.0;
      __syncthreads();
      f.row0 = (f.row0 + drop) % front_rows_cap;
      nrows -= drop;
      ++s;
      if (cend < 6 * s) cend = 6 * s;
    }
    if (t >= fw.tile_end) break;
    // append every tile that starts at block s
    while (t < fw.tile_end && a.tiles[t].c0_blk == s) {
      const Tile tl = a.tiles[t];
      const int W = 6 * (tl.c1_blk - tl.c0_blk);
      const double* src = a.tile_out + tl.out_off;
      // rows actually produced: leading rows with a non-zero (tile emits zero padding)
      const int add = min(W, tl.rows);
      if (nrows + add > front_rows_cap) {   // cannot happen: the host sizes the front
        if (tid == 0 && a.err) *a.err = 1;
        ++t;
        continue;
      }
      for (int e = tid; e < add * (W + 1); e += nt) {
        int i = e / (W + 1), j = e % (W + 1);
        double v = src[(size_t)i * (W + 1) + j];
        if (j < W) f.at(nrows + i, 6 * s + j) = v;
        else f.at(nrows + i, FRONT_RHS) = v;
      }
      __syncthreads();
      nrows += add;
      if (6 * tl.c1_blk > cend) cend = 6 * tl.c1_blk;
      ++t;
      // re-triangularise the front; rows beyond its width become zero
      cta_householder(f, nrows, 6 * s, cend, red);
      const int width = cend - 6 * s;
      if (nrows > width) {
        nrows = width;
      }
    }
  }
}

// Left-nullspace projection of a dense block (removeLostObjects, src/orcvio.cpp:2162 ->
// nullspace_project_inplace_svd, math_utils.hpp:287-312): M = [H_f | H_x | r] (rows x ncols,
// row-major in global memory); Householder-eliminates the first `nelim` columns and applies the
// reflections to every column.  Rows [nelim, rows) of the trailing columns are A^T H_x, A^T r for
// an orthonormal basis A of null(H_f^T) (any basis gives the same gate value and posterior).
//
// k_project_dense_panel (the default): the reflections only depend on the H_f panel (rows x nelim <= 560 x 45), so
// every CTA keeps its own copy of the panel in shared memory, column-major (lanes over rows: conflict-free), factors
// it redundantly -- the same arithmetic in the same order in every CTA -- and carries PD_COLS trailing columns
// through the reflections IN REGISTERS (one warp per two columns, rows over lanes).  Each warp recomputes the
// column norm of step k itself, so a step costs ONE __syncthreads (the panel columns right of k are updated by the
// warps in turn) instead of three plus a pass over global memory: 140 x 154 block 2170 -> ~60 us.  Only the trailing
// columns are written back (the triangular factor of H_f is not used by the update).
constexpr int PD_THREADS = 512, PD_WARPS = PD_THREADS / 32, PD_CPW = 2, PD_COLS = PD_WARPS * PD_CPW;

// PD_RPL = rows per lane (rows <= 32 PD_RPL): 6 / 12 / 18 for blocks of up to 192 / 384 / 576 rows
template <int PD_RPL>
__global__ void __launch_bounds__(PD_THREADS) k_project_dense_panel(double* M, int rows, int ld, int nelim, int ncols) {
  extern __shared__ double pan[];                       // nelim columns of rpad doubles
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rpad = (rows + 31) & ~31;
  for (int e = tid; e < rpad * nelim; e += PD_THREADS) pan[e] = 0.0;
  __syncthreads();
  for (int e = tid; e < rows * nelim; e += PD_THREADS) {  // coalesced along the rows of M
    const int i = e / nelim, c = e - i * nelim;
    pan[(size_t)c * rpad + i] = M[(size_t)i * ld + c];
  }
  double x[PD_CPW][PD_RPL];
  int cj[PD_CPW];
#pragma unroll
  for (int q = 0; q < PD_CPW; ++q) {
    cj[q] = nelim + blockIdx.x * PD_COLS + warp * PD_CPW + q;
#pragma unroll
    for (int j = 0; j < PD_RPL; ++j) {
      const int i = lane + 32 * j;
      x[q][j] = (cj[q] < ncols && i < rows) ? M[(size_t)i * ld + cj[q]] : 0.0;
    }
  }
  __syncthreads();
  const int steps = min(nelim, rows - 1);
  for (int k = 0; k < steps; ++k) {
    const double* ck = pan + (size_t)k * rpad;
    double sig = 0.0;
#pragma unroll
    for (int j = 0; j < PD_RPL; ++j) {
      const int i = lane + 32 * j;
      if (i > k && i < rows) { const double v = ck[i]; sig += v * v; }
    }
    for (int o = 16; o > 0; o >>= 1) sig += __shfl_xor_sync(0xffffffffu, sig, o);
    if (sig > 0.0) {                                   // (uniform over the CTA: every warp computes the same sig)
      const double akk = ck[k];
      const double mu = sqrt(akk * akk + sig);
      const double v0 = (akk <= 0.0) ? (akk - mu) : (-sig / (akk + mu));
      const double t = 2.0 * v0 * v0 / (sig + v0 * v0), iv0 = 1.0 / v0;
      // v = column k below the diagonal / v0: in registers while they last, else re-read from shared memory at each use
      constexpr bool VREG = PD_RPL <= 12;
      double v[VREG ? PD_RPL : 1];
      if (VREG) {
#pragma unroll
        for (int j = 0; j < PD_RPL; ++j) {
          const int i = lane + 32 * j;
          v[j] = (i > k && i < rows) ? ck[i] * iv0 : 0.0;
        }
      }
      auto vj = [&](int j) {
        if (VREG) return v[j];
        const int i = lane + 32 * j;
        return (i > k && i < rows) ? ck[i] * iv0 : 0.0;
      };
      const int jk = k >> 5, lk = k & 31;              // row k lives in lane lk, register jk
      // panel columns right of k, the warps in turn
      for (int c = k + 1 + warp; c < nelim; c += PD_WARPS) {
        double* cc = pan + (size_t)c * rpad;
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < PD_RPL; ++j) {
          const int i = lane + 32 * j;
          if (i > k && i < rows) dot += vj(j) * cc[i];
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const double s = t * (dot + cc[k]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < PD_RPL; ++j) {
          const int i = lane + 32 * j;
          if (i > k && i < rows) cc[i] -= s * vj(j);
        }
        if (lane == lk) cc[k] -= s;
      }
      // own trailing columns
#pragma unroll
      for (int q = 0; q < PD_CPW; ++q) {
        double dot = 0.0, xk = 0.0;
#pragma unroll
        for (int j = 0; j < PD_RPL; ++j) {
          dot += vj(j) * x[q][j];
          if (j == jk) xk = x[q][j];
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        xk = __shfl_sync(0xffffffffu, xk, lk);
        const double s = t * (dot + xk);
#pragma unroll
        for (int j = 0; j < PD_RPL; ++j) {
          x[q][j] -= s * vj(j);
          if (j == jk && lane == lk) x[q][j] -= s;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < PD_CPW; ++q)
    if (cj[q] < ncols) {
#pragma unroll
      for (int j = 0; j < PD_RPL; ++j) {
        const int i = lane + 32 * j;
        if (i < rows) M[(size_t)i * ld + cj[q]] = x[q][j];
      }
    }
}

// fallback for blocks whose panel does not fit in shared memory: one CTA on global memory
__global__ void __launch_bounds__(QR_THREADS) k_project_dense(double* M, int rows, int ld, int nelim, int ncols) {
  __shared__ double red[32];
  Front f{M, ld, rows, 0, 0x7fffffff, ncols - 1};
  cta_householder(f, rows, 0, ncols - 1, red, nelim);
}

void launch_project_dense(double* M, int rows, int ld, int nelim, int ncols, cudaStream_t s) {
  const int rpad = (rows + 31) & ~31;
  const size_t smem = (size_t)rpad * nelim * sizeof(double);
  if (rows <= 32 * 18 && smem <= 220 * 1024 && ncols > nelim) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(k_project_dense_panel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      cudaFuncSetAttribute(k_project_dense_panel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      cudaFuncSetAttribute(k_project_dense_panel<18>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      attr = true;
    }
    const int grid = (ncols - nelim + PD_COLS - 1) / PD_COLS;
    if (rows <= 32 * 6) k_project_dense_panel<6><<<grid, PD_THREADS, smem, s>>>(M, rows, ld, nelim, ncols);
    else if (rows <= 32 * 12) k_project_dense_panel<12><<<grid, PD_THREADS, smem, s>>>(M, rows, ld, nelim, ncols);
    else k_project_dense_panel<18><<<grid, PD_THREADS, smem, s>>>(M, rows, ld, nelim, ncols);
    check_launch("k_project_dense_panel");
    return;
  }
  k_project_dense<<<1, QR_THREADS, 0, s>>>(M, rows, ld, nelim, ncols);
  check_launch("k_project_dense");
}

void launch_qr(const QrArgs& a, size_t tile_smem_doubles, int max_w_blk, int max_n, cudaStream_t s,
               int* launches, cudaEvent_t mid) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_qr_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_qr_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    check_launch("qr attributes");
    attr = true;
  }
  if (a.n_tiles > 0) {
    size_t smem = tile_smem_doubles * sizeof(double);
    if (smem < 1024) smem = 1024;
    k_qr_tiles<<<a.n_tiles, QR_THREADS, smem, s>>>(a);
    check_launch("k_qr_tiles");
    if (launches) ++*launches;
  }
  if (mid) cudaEventRecord(mid, s);
  // chain: the front is a ring of wcap (power of two >= widest window) clone columns and
  // up to 2 * Wmax + 8 rows (carry + one appended tile)
  (void)max_n;
  const int W = 6 * max_w_blk;
  int wcap = 64;
  while (wcap < W + 8) wcap <<= 1;
  const int rows_cap = 2 * W + 8;
  const size_t smem = (size_t)rows_cap * (wcap + 1) * sizeof(double);
  const int use_global = smem > (size_t)QR_SMEM_BYTES ? 1 : 0;
  k_qr_chain<<<a.n_filters, QR_THREADS, use_global ? 0 : smem, s>>>(a, rows_cap, wcap, use_global);
  check_launch("k_qr_chain");
  if (launches) ++*launches;
}

}  // namespace ob
