// One-CTA blocked Cholesky with carried rows, look-ahead teams and FP64 tensor-core trailing updates.
//
// Used for both factorisations of the whitened-form update (info_kernel.cu): the prior
// P = F F^T (k_chol_prior) and W = C C^T with strips of F_1 / v carried through as extra rows
// (k_chol_w_solve).  The reference does the equivalent work with Eigen's ldlt()
// (src/orcvio.cpp:1690 `S.ldlt().solve(H P)`); on the GPU a 200 x 200 factorisation is a latency
// chain of m dependent pivots, so the routine is organised around that chain:
//
//   * the "tall" matrix [A; X] (m x m lower triangle + nx carried rows) lives in shared memory
//     as 8 x 8 row-major tiles (512 contiguous bytes each): every access below is a 16-byte,
//     bank-conflict-free vector access and a tile is exactly one mma.m8n8k4 accumulator;
//   * 8-column panels.  The CTA (512 threads) is split into a PANEL team (4 warps) and a TRAILING
//     team (12 warps).  In iteration p the trailing team first applies panel p-1 to tile column p
//     and signals a named barrier, then applies it to every later tile column with DMMA
//     (two m8n8k4 per tile, operands straight from the transposed panel buffer).  Meanwhile the
//     panel team factors the 8 x 8 diagonal tile (8 lanes, one matrix row per lane, pivots
//     exchanged with shuffles) and solves the rows below it (thread per row).  One CTA barrier per
//     panel.
//   * the factor is left in the tiles; callers copy what they need out with chol_for_rows().
//
// `tol != nullptr`: pivots <= tol[k] are exact zeros (semidefinite prior: the IMU pose duplicates
// the newest clone after augmentation and rows 15..21 are exactly zero) -> zero column.
#pragma once
#include "kernels.h"

namespace ob {

constexpr int CHB = 8;
constexpr int CHOL_THREADS = 512;
constexpr int CHOL_PANEL_THREADS = 128;                           // panel team (warps 0..3)
constexpr int CHOL_TRAIL_WARPS = (CHOL_THREADS - CHOL_PANEL_THREADS) / 32;
constexpr int CHOL_MAXM = ORCVIO_LEG + 6 * ORCVIO_MAX_OBS;        // 214
constexpr int CHOL_MAXR = 244;   // rows incl. carried rows; == 4 (mod 16): the 4 k-rows of an MMA fragment
                                 // load start 8 banks apart -> conflict-free 64-bit fragment loads
constexpr int CHOL_MAXT = (CHOL_MAXM + 7) / 8;                    // tile columns

// tile index of (ti, tj), tj <= min(ti, Tm-1); Tm = tile columns = ceil(m / 8)
__device__ __host__ __forceinline__ int chol_tile(int ti, int tj, int Tm) {
  return ti < Tm ? (ti * (ti + 1) >> 1) + tj : (Tm * (Tm + 1) >> 1) + (ti - Tm) * Tm + tj;
}
// element (i, j) of the tall matrix, j <= min(i, m-1)
__device__ __host__ __forceinline__ int chol_at(int i, int j, int Tm) {
  return chol_tile(i >> 3, j >> 3, Tm) * 64 + (i & 7) * 8 + (j & 7);
}
__device__ __host__ __forceinline__ size_t chol_smem_doubles(int m, int nx) {
  const int Tm = (m + 7) >> 3, Tr = (m + nx + 7) >> 3;
  return (size_t)chol_tile(Tr, 0, Tm) * 64;
}

struct CholShared {
  double PT[2][CHB][CHOL_MAXR];     // solved panels, transposed: PT[buf][c][row]
  double dblk[CHB][CHB];            // factored diagonal block (lower)
  double dinv[CHB];                 // 1 / diagonal (0 for skipped pivots)
  unsigned char tdec[CHOL_MAXT * (CHOL_MAXT + 1) / 2][2];   // triangle index -> (a, b), b <= a
};

__device__ __forceinline__ void chol_team_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(CHOL_PANEL_THREADS) : "memory");
}
__device__ __forceinline__ void chol_col_arrive() { asm volatile("bar.arrive 2, %0;" ::"n"(CHOL_THREADS) : "memory"); }
__device__ __forceinline__ void chol_col_wait() { asm volatile("bar.sync 2, %0;" ::"n"(CHOL_THREADS) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 1/x for a normal positive pivot: hardware seed (2^-23) + two Newton steps, no special-case branch.
__device__ __forceinline__ double chol_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// Zero the tiles and fill the decode table; call before scattering the matrix into `A`.
__device__ __forceinline__ void chol_init(double* __restrict__ A, CholShared& cs, int m, int nx) {
  const int nd2 = (int)(chol_smem_doubles(m, nx) >> 1);
  double2* A2 = reinterpret_cast<double2*>(A);
  for (int e = threadIdx.x; e < nd2; e += blockDim.x) A2[e] = make_double2(0.0, 0.0);
  for (int a = threadIdx.x; a < CHOL_MAXT; a += blockDim.x)
    for (int b = 0; b <= a; ++b) {
      const int t = (a * (a + 1) >> 1) + b;
      cs.tdec[t][0] = (unsigned char)a;
      cs.tdec[t][1] = (unsigned char)b;
    }
}

// 1/sqrt(x) for a normal positive x: hardware seed (2^-22) + two Newton steps, branch-free.
__device__ __forceinline__ double chol_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double xh = 0.5 * x;
  double e = fma(-xh * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-xh * y, y, 0.5);
  return fma(y, e, y);
}

__device__ __forceinline__ void chol_cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void chol_cp_async_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Scatter rows [row_begin, row_end) of the tall matrix into the tiles: element (i, j) comes from the global
// address addr(i, j), j < ncols(i) (ncols(i) = i + 1 for triangle rows, m for carried rows).  Every element is
// one 8-byte cp.async straight into its tile slot, so ALL loads of the CTA are in flight together (one
// memory round trip for the whole matrix instead of one per row batch).  The tiles must have been zeroed
// (chol_init) and made visible (__syncthreads) before; the caller finishes with chol_cp_async_wait() +
// __syncthreads().
template <class Addr>
__device__ __forceinline__ void chol_load_rows(double* __restrict__ A, int m, int row_begin, int row_end, Addr addr) {
  const int Tm = (m + 7) >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = row_begin + warp; i < row_end; i += nw) {
    const int nc = min(i + 1, m);
    const int tbase = chol_tile(i >> 3, 0, Tm) * 64 + (i & 7) * 8;
    for (int j = lane; j < nc; j += 32) chol_cp_async8(A + tbase + (j >> 3) * 64 + (j & 7), addr(i, j));
  }
}

// A: tiled (chol_at) tall matrix: m x m lower triangle followed by nx carried rows.
// On return tile storage holds L (rows < m) and X C^-T (rows >= m).
// PROF: thread 0 (panel team) / thread 128 (trailing team) log clock64() per phase into prof[p][8].
template <bool PROF = false, int LAYOUT = 0>
__device__ void cta_cholesky(double* __restrict__ A, CholShared& cs, const double* __restrict__ tol, int m, int nx,
                             long long* prof = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Teams by warp scheduler: warps 0, 4, 8, 12 (all on scheduler 0) are the panel team, so the
  // pivot chain never queues behind a 16-cycle DMMA of the trailing team on its FP64 pipe.
  // (LAYOUT 1, experiment: panel team = warps 0..3, one per scheduler.)
  const bool panel_team = LAYOUT == 0 ? (warp & 3) == 0 : warp < 4;
  const int tid = panel_team ? (LAYOUT == 0 ? (warp >> 2) : warp) * 32 + lane : -1;   // panel-team thread index
  const int tw = LAYOUT == 0 ? (warp >> 2) * 3 + (warp & 3) - 1 : warp - 4;            // trailing-team warp index
  const bool trail_lead = ((LAYOUT == 0 ? warp == 1 : warp == 4) && lane == 0);
  const int mrows = m + nx;
  const int Tm = (m + 7) >> 3, Tr = (mrows + 7) >> 3;
  __syncthreads();
  for (int p = 0; p < Tm; ++p) {
    const int k0 = p * CHB;
    const int nb = min(CHB, m - k0);
    const double(*PTp)[CHOL_MAXR] = cs.PT[(p + 1) & 1];      // panel p-1
    double(*PTn)[CHOL_MAXR] = cs.PT[p & 1];                  // panel p (written here)
    if (PROF && tid == 0) prof[p * 8 + 0] = clock64();
    if (panel_team) {
      // ------------------------------------------------ panel team
      chol_col_wait();                                       // tile column p carries panels 0..p-1
      if (PROF && tid == 0) prof[p * 8 + 1] = clock64();
      if (PROF) __syncwarp();
      if (warp == 0) {
        // lane a (mod 8) owns row a of the diagonal tile
        const int a = lane & 7;
        double* dt = A + (size_t)chol_tile(p, p, Tm) * 64 + a * 8;
        double d[CHB];
#pragma unroll
        for (int b = 0; b < CHB; b += 2) {
          const double2 t = *reinterpret_cast<const double2*>(dt + b);
          d[b] = t.x;
          d[b + 1] = t.y;
        }
#pragma unroll
        for (int b = 0; b < CHB; ++b)
          if (a >= nb || b >= nb) d[b] = (a == b) ? 1.0 : 0.0;
        // Unscaled (L D L^T) elimination, branch-free.  The dependent chain per pivot is
        //   shuffle -> reciprocal seed -> 2 FMA (one Newton step) -> multiply -> FMA (second Newton step folded
        //   into w = u / pivot) -> FMA into the next pivot;
        // thresholds are preloaded, rejected pivots are handled by selects.  The 8 rsqrt that turn the
        // result into L = U D^-1/2 run afterwards, in parallel.
        double thr[CHB], ivs[CHB];
#pragma unroll
        for (int c = 0; c < CHB; ++c) thr[c] = (tol != nullptr && c < nb) ? tol[k0 + c] : 0.0;
#pragma unroll
        for (int c = 0; c < CHB; ++c) {
          const double pv = __shfl_sync(0xffffffffu, d[c], c, 8);
          const bool ok = pv > thr[c];
          const double u = d[c];
          double y0;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(pv));
          const double e = fma(-pv, y0, 1.0);
          const double y1 = fma(y0, e, y0);                   // 1/pv (1 - e^2)
          const double e2 = e * e;
          const double t = u * y1;
          double w = fma(t, e2, t);                           // u / pv to working precision
          w = ok ? w : 0.0;
          ivs[c] = ok ? pv : 0.0;
#pragma unroll
          for (int b = c + 1; b < CHB; ++b) {
            const double ub = __shfl_sync(0xffffffffu, u, b, 8);
            d[b] = fma(-w, ub, d[b]);                         // meaningful for a >= b
          }
        }
#pragma unroll
        for (int c = 0; c < CHB; ++c) {
          const double rs = chol_rsqrt(ivs[c] > 0.0 ? ivs[c] : 1.0);
          ivs[c] = (ivs[c] > 0.0) ? rs : 0.0;
          d[c] *= ivs[c];
        }
        __syncwarp();   // lanes 8..31 hold redundant copies of the rows: their reads of the tile are done before it is rewritten
        if (lane < CHB) {
#pragma unroll
          for (int b = 0; b < CHB; b += 2) {
            const double2 t = make_double2(b <= a ? d[b] : 0.0, b + 1 <= a ? d[b + 1] : 0.0);
            *reinterpret_cast<double2*>(&cs.dblk[a][b]) = t;
            if (a < nb) *reinterpret_cast<double2*>(dt + b) = t;   // rows >= nb of this tile are carried rows
          }
          if (lane == 0) {
#pragma unroll
            for (int b = 0; b < CHB; ++b) cs.dinv[b] = ivs[b];
          }
        }
      }
      if (PROF && tid == 0) prof[p * 8 + 2] = clock64();
      chol_team_barrier();
      // rows below the block: x <- x L_d^-T  (update form: full ILP across the remaining columns)
      const int ibase = k0 + nb;
      if (ibase + tid < mrows) {
        double dl[CHB][CHB], di[CHB];
#pragma unroll
        for (int c = 0; c < CHB; c += 2) {
          const double2 t = *reinterpret_cast<const double2*>(&cs.dinv[c]);
          di[c] = t.x;
          di[c + 1] = t.y;
        }
#pragma unroll
        for (int c = 1; c < CHB; ++c)
#pragma unroll
          for (int q = 0; q < c; q += 2) {
            const double2 t = *reinterpret_cast<const double2*>(&cs.dblk[c][q]);
            dl[c][q] = t.x;
            if (q + 1 < c) dl[c][q + 1] = t.y;
          }
        // up to two rows per thread (mrows <= 244), both chains in flight together
        const int i0r = ibase + tid, i1r = i0r + CHOL_PANEL_THREADS;
        const bool two = i1r < mrows;
        double* xt0 = A + (size_t)chol_tile(i0r >> 3, p, Tm) * 64 + (i0r & 7) * 8;
        double* xt1 = two ? A + (size_t)chol_tile(i1r >> 3, p, Tm) * 64 + (i1r & 7) * 8 : xt0;
        double x0[CHB], x1[CHB];
#pragma unroll
        for (int c = 0; c < CHB; c += 2) {
          const double2 t0 = *reinterpret_cast<const double2*>(xt0 + c);
          const double2 t1 = *reinterpret_cast<const double2*>(xt1 + c);
          x0[c] = t0.x; x0[c + 1] = t0.y;
          x1[c] = t1.x; x1[c + 1] = t1.y;
        }
#pragma unroll
        for (int c = 0; c < CHB; ++c) {
          if (c >= nb) { x0[c] = 0.0; x1[c] = 0.0; }
          x0[c] *= di[c];
          x1[c] *= di[c];
#pragma unroll
          for (int q = c + 1; q < CHB; ++q) {
            x0[q] -= x0[c] * dl[q][c];
            x1[q] -= x1[c] * dl[q][c];
          }
        }
#pragma unroll
        for (int c = 0; c < CHB; ++c) PTn[c][i0r] = x0[c];
#pragma unroll
        for (int c = 0; c < CHB; c += 2) *reinterpret_cast<double2*>(xt0 + c) = make_double2(x0[c], x0[c + 1]);
        if (two) {
#pragma unroll
          for (int c = 0; c < CHB; ++c) PTn[c][i1r] = x1[c];
#pragma unroll
          for (int c = 0; c < CHB; c += 2) *reinterpret_cast<double2*>(xt1 + c) = make_double2(x1[c], x1[c + 1]);
        }
      }
      if (PROF && tid == 0) prof[p * 8 + 3] = clock64();
    } else {
      // ------------------------------------------------ trailing team: panel p-1 -> tile columns >= p
      const int fr = lane >> 2, fk = lane & 3;               // fragment row / k index
      if (p > 0) {
        // priority: tile column p (at most 3 tiles per warp, all in flight together)
        const double b0 = PTp[fk][k0 + fr], b1 = PTp[fk + 4][k0 + fr];
        double2 cv[3];
        double a0[3], a1[3];
        double2* cp[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int ti = p + tw + u * CHOL_TRAIL_WARPS;
          if (ti < Tr) {
            a0[u] = -PTp[fk][ti * 8 + fr];
            a1[u] = -PTp[fk + 4][ti * 8 + fr];
            cp[u] = reinterpret_cast<double2*>(A + (size_t)chol_tile(ti, p, Tm) * 64) + lane;
            cv[u] = *cp[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (p + tw + u * CHOL_TRAIL_WARPS < Tr) dmma884(cv[u].x, cv[u].y, a0[u], b0);
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (p + tw + u * CHOL_TRAIL_WARPS < Tr) {
            dmma884(cv[u].x, cv[u].y, a1[u], b1);
            *cp[u] = cv[u];
          }
      }
      __threadfence_block();
      chol_col_arrive();
      if (PROF && trail_lead) prof[p * 8 + 6] = clock64();
      if (p > 0 && p + 1 < Tm) {
        const int nT = Tm - (p + 1);                          // tile columns left
        const int full = nT * (nT + 1) >> 1;
        const int total = full + (Tr - Tm) * nT;
        const int per = (total + CHOL_TRAIL_WARPS - 1) / CHOL_TRAIL_WARPS;
        const int tend = min(total, (tw + 1) * per);
        constexpr int IL = 4;                                  // tiles in flight per warp
        // first tile of this warp's contiguous range, then row-major stepping
        int ti, tj;
        {
          const int t = min(tw * per, total - 1);
          if (t < full) {
            ti = p + 1 + cs.tdec[t][0];
            tj = p + 1 + cs.tdec[t][1];
          } else {
            const int q = t - full;
            ti = Tm + q / nT;
            tj = p + 1 + (q - (q / nT) * nT);
          }
        }
        // the A operand depends on the tile row only: it is reloaded when the row changes (row-major
        // stepping), which takes a quarter of the shared-memory traffic of a tile out of the loop
        int ti_a = -1;
        double a0c = 0.0, a1c = 0.0;
        for (int t0 = tw * per; t0 < tend; t0 += IL) {
          double2 cv[IL];
          double a0[IL], a1[IL], b0[IL], b1[IL];
          double2* cp[IL];
#pragma unroll
          for (int u = 0; u < IL; ++u) {
            if (ti != ti_a) {
              a0c = -PTp[fk][ti * 8 + fr];
              a1c = -PTp[fk + 4][ti * 8 + fr];
              ti_a = ti;
            }
            a0[u] = a0c;
            a1[u] = a1c;
            b0[u] = PTp[fk][tj * 8 + fr];
            b1[u] = PTp[fk + 4][tj * 8 + fr];
            cp[u] = reinterpret_cast<double2*>(A + (size_t)chol_tile(ti, tj, Tm) * 64) + lane;
            cv[u] = *cp[u];
            if (t0 + u + 1 < tend) {                          // advance (stays put on the last tile)
              if (tj < min(ti, Tm - 1)) ++tj;
              else { ++ti; tj = p + 1; }
            }
          }
#pragma unroll
          for (int u = 0; u < IL; ++u) dmma884(cv[u].x, cv[u].y, a0[u], b0[u]);
#pragma unroll
          for (int u = 0; u < IL; ++u) dmma884(cv[u].x, cv[u].y, a1[u], b1[u]);
#pragma unroll
          for (int u = 0; u < IL; ++u)
            if (t0 + u < tend) *cp[u] = cv[u];                // (duplicates of the last tile are not stored)
        }
      }
      if (PROF && trail_lead) prof[p * 8 + 4] = clock64();
    }
    __syncthreads();
    if (PROF && tid == 0) prof[p * 8 + 5] = clock64();
  }
}

// Calls f(i, k, value) for every stored entry with row_begin <= i < m + nx, k <= min(i, m-1).
// Thread mapping: a warp covers 8 rows x 4 columns per step (32-byte global segments when the
// caller's output is contiguous in i, mild bank conflicts on the tile reads).
template <class F>
__device__ __forceinline__ void chol_for_rows(const double* __restrict__ A, int m, int nx, int row_begin, F f) {
  const int Tm = (m + 7) >> 3;
  const int mrows = m + nx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int r = lane >> 2, cq = lane & 3;                    // row in tile, column quarter (2 columns)
  const int t_begin = row_begin >> 3, Tr = (mrows + 7) >> 3;
  // flat loop over (row block, tile column)
  int t = warp;
  for (int ti = t_begin; ti < Tr; ++ti) {
    const int ncol = min(ti + 1, Tm);
    for (; t < ncol; t += nw) {
      const double* tp = A + (size_t)chol_tile(ti, t, Tm) * 64 + r * 8 + cq * 2;
      const double2 v = *reinterpret_cast<const double2*>(tp);
      const int i = ti * 8 + r, k = t * 8 + cq * 2;
      if (i >= row_begin && i < mrows) {
        if (k <= min(i, m - 1)) f(i, k, v.x);
        if (k + 1 <= min(i, m - 1)) f(i, k + 1, v.y);
      }
    }
    t -= ncol;
  }
}

}  // namespace ob
