// The BP5 hot kernel for sm_100a: what LocalPoissonOperator::operator()
// (bp5/step-64.cu:147-194) does through CUDAWrappers::FEEvaluation
// (read_dof_values / evaluate / merged coefficient / integrate /
// distribute_local_to_global; API mirrored by bp5/fe_evaluation_gl.h:133-250)
// inside MatrixFree::cell_loop's apply_kernel_shmem [UPSTREAM]; plus the
// Helmholtz variant of step-64/step-64.cu:201-219.
//
// Design:
//  * persistent CTAs, a tile = CPT consecutive cells, tiles dealt round-robin so
//    that the set of cells in flight is one contiguous window of the mesh (L2
//    reuse of the gathered DoFs and of the dst lines being accumulated);
//  * the tile's metric ([cell][planes][n^3] fp64, 75-85 % of all bytes) is
//    fetched by ONE cp.async.bulk (TMA 1D) per tile into shared memory, signalled
//    on an mbarrier, and re-issued for the next tile as soon as the quadrature
//    phase has consumed it: tens of KB in flight per CTA with zero registers;
//  * gathers for the next tile are issued into registers one tile ahead, their
//    indices two tiles ahead;
//  * contractions: n^2 threads per cell.  Each thread alternates between
//    "home" (owns the z-column (i,j,*)), "x-line" and "y-line" roles, holding a
//    whole line in registers, so one 1D contraction costs one shared-memory load
//    and one store per point instead of n loads; the z direction never leaves
//    registers.  Shape matrices are kernel parameters (constant bank) and, with
//    full unrolling, become immediate operands of the DFMAs.
//  * scatter: fire-and-forget fp64 red.global.add.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.h"

namespace bp5 {

template <int N>
struct ApplyParams {
  const double *metric;   // [tile][CPT][PLANES][N^3], tile stride padded to 16 bytes
  const int *l2g;         // [tiles*CPT][N^3]
  const double *src;
  double *dst;
  long long n_cells;
  long long n_tiles;
  const int *skip;        // optional device flag: non-zero => nothing to do (CG already converged)
  ShapeTables<N> tab;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1D bulk copy global -> shared, completion on an mbarrier, L2 evict-first
// (the metric is streamed exactly once per operator application).
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <int P, int CPT, int PLANES>
struct ApplyCfg {
  static constexpr int N = P + 1, N2 = N * N, N3 = N2 * N;
  static constexpr int NP = (N % 2) ? N : N + 1;          // odd x-pitch of the work tiles: conflict-free line access
  static constexpr int WS = N2 * NP;                       // doubles per work array per cell
  static constexpr int ACTIVE = CPT * N2;
  static constexpr int NT = ((ACTIVE + 31) / 32) * 32;
  static constexpr int METRIC_DOUBLES = (CPT * PLANES * N3 + 1) & ~1;  // per tile, padded to 16 bytes
  static constexpr uint32_t METRIC_BYTES = METRIC_DOUBLES * 8;
  static constexpr int WORK_ARRAYS = 3;
  static constexpr size_t SMEM_BYTES = (size_t)METRIC_BYTES + (size_t)WORK_ARRAYS * CPT * WS * 8 + 16;
  static_assert(METRIC_BYTES % 16 == 0, "bulk copy size must be a multiple of 16 bytes");
};

// QUAD: 0 = Gauss (basis nodes != quadrature points: interpolate, then
// collocation derivative), 1 = Gauss-Lobatto collocation (B = identity).
// HELM: 0 = Poisson (6 planes), 1 = Helmholtz (7th plane a(x) JxW on the values).
template <int P, int QUAD, int HELM, int CPT>
__global__ void __launch_bounds__(ApplyCfg<P, CPT, 6 + HELM>::NT)
    bp5_apply_kernel(const __grid_constant__ ApplyParams<P + 1> prm) {
  using Cfg = ApplyCfg<P, CPT, 6 + HELM>;
  constexpr int N = Cfg::N, N2 = Cfg::N2, N3 = Cfg::N3, NP = Cfg::NP, WS = Cfg::WS, PLANES = 6 + HELM;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Gs = reinterpret_cast<double *>(smem_raw);                 // [CPT][PLANES][N3]
  double *S0 = Gs + Cfg::METRIC_DOUBLES;                             // [CPT][WS]
  double *S1 = S0 + CPT * WS;
  double *S2 = S1 + CPT * WS;
  uint64_t *bar = reinterpret_cast<uint64_t *>(S2 + CPT * WS);

  if (prm.skip != nullptr && *prm.skip != 0) return;
  const int tid = threadIdx.x;
  const bool active = tid < Cfg::ACTIVE;
  const int c = active ? tid / N2 : 0;      // cell within the tile
  const int r = tid % N2;
  const int a = r % N, b = r / N;           // the two free indices of this thread's line/column
  double *s0 = S0 + c * WS, *s1 = S1 + c * WS, *s2 = S2 + c * WS;
  const double *gm = Gs + c * PLANES * N3;
  const double *__restrict__ Bm = prm.tab.B;
  const double *__restrict__ Dt = prm.tab.Dt;

  const long long tile0 = blockIdx.x;
  const long long tstride = gridDim.x;
  uint64_t policy = 0;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    policy = make_evict_first_policy();
    if (tile0 < prm.n_tiles) {
      mbar_expect_tx(bar, Cfg::METRIC_BYTES);
      tma_load_1d(Gs, prm.metric + tile0 * (long long)Cfg::METRIC_DOUBLES, Cfg::METRIC_BYTES, bar, policy);
    }
  }
  __syncthreads();

  // software pipeline of the gather: indices two tiles ahead, values one tile ahead
  int idx_cur[N], idx_nxt[N];
  double u_nxt[N];
  {
    const long long cell0 = tile0 * CPT + c, cell1 = (tile0 + tstride) * CPT + c;
    const bool v0 = active && tile0 < prm.n_tiles && cell0 < prm.n_cells;
    const bool v1 = active && (tile0 + tstride) < prm.n_tiles && cell1 < prm.n_cells;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      idx_cur[k] = v0 ? __ldg(prm.l2g + cell0 * N3 + (k * N + b) * N + a) : -1;
      idx_nxt[k] = v1 ? __ldg(prm.l2g + cell1 * N3 + (k * N + b) * N + a) : -1;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) u_nxt[k] = idx_cur[k] >= 0 ? __ldg(prm.src + idx_cur[k]) : 0.0;
  }

  uint32_t parity = 0;
  for (long long tile = tile0; tile < prm.n_tiles; tile += tstride) {
    double u[N];
    int idx_n2[N];
#pragma unroll
    for (int k = 0; k < N; ++k) u[k] = u_nxt[k];
    {
      // issue next tile's gather and the index loads of the tile after it
      const long long cell2 = (tile + 2 * tstride) * CPT + c;
      const bool v2 = active && (tile + 2 * tstride) < prm.n_tiles && cell2 < prm.n_cells;
#pragma unroll
      for (int k = 0; k < N; ++k) u_nxt[k] = idx_nxt[k] >= 0 ? __ldg(prm.src + idx_nxt[k]) : 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) idx_n2[k] = v2 ? __ldg(prm.l2g + cell2 * N3 + (k * N + b) * N + a) : -1;
    }

    double t[N];   // z-direction data that stays in registers across the quadrature phase
    double mv[N];  // Helmholtz: values at the quadrature points (home column)

    if constexpr (QUAD == 1) {
      // ---------------- Gauss-Lobatto collocation: u is already at the q-points
      // (1) home (i=a, j=b): publish the column, z-derivative in registers
      if (active) {
#pragma unroll
        for (int k = 0; k < N; ++k) s0[(k * N + b) * NP + a] = u[k];
#pragma unroll
        for (int k = 0; k < N; ++k) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) s += Dt[k * N + m] * u[m];
          t[k] = s;
        }
        if constexpr (HELM) {
#pragma unroll
          for (int k = 0; k < N; ++k) mv[k] = u[k];
        }
      }
      __syncthreads();
      // (2) x-line (j=a, k=b) and y-line (i=a, k=b): derivative along the line
      if (active) {
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = s0[(b * N + a) * NP + i];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) s += Dt[i * N + m] * v[m];
          s1[(b * N + a) * NP + i] = s;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = s0[(b * N + j) * NP + a];
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) s += Dt[j * N + m] * v[m];
          s2[(b * N + j) * NP + a] = s;
        }
      }
      __syncthreads();
    } else {
      // ---------------- Gauss quadrature: interpolate to the q-points first
      // (1) home (i=a, j=b): z-interpolation in registers
      if (active) {
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < N; ++k) s += Bm[q * N + k] * u[k];
          s0[(q * N + b) * NP + a] = s;
        }
      }
      __syncthreads();
      // (2) x-line (j=a, qz=b): x-interpolation in place
      if (active) {
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = s0[(b * N + a) * NP + i];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) s += Bm[q * N + i] * v[i];
          s0[(b * N + a) * NP + q] = s;
        }
      }
      __syncthreads();
      // (3) y-line (qx=a, qz=b): y-interpolation (values at q-points), then d/dy
      if (active) {
        double v[N], w[N];
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = s0[(b * N + j) * NP + a];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < N; ++j) s += Bm[q * N + j] * v[j];
          w[q] = s;
          s0[(b * N + q) * NP + a] = s;
        }
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double s = 0.0;
#pragma unroll
          for (int rr = 0; rr < N; ++rr) s += Dt[q * N + rr] * w[rr];
          s2[(b * N + q) * NP + a] = s;
        }
      }
      __syncthreads();
      // (4) x-line (qy=a, qz=b): d/dx ; home (qx=a, qy=b): d/dz in registers
      if (active) {
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = s0[(b * N + a) * NP + i];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double s = 0.0;
#pragma unroll
          for (int rr = 0; rr < N; ++rr) s += Dt[q * N + rr] * v[rr];
          s1[(b * N + a) * NP + q] = s;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = s0[(k * N + b) * NP + a];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          double s = 0.0;
#pragma unroll
          for (int rr = 0; rr < N; ++rr) s += Dt[q * N + rr] * v[rr];
          t[q] = s;
        }
        if constexpr (HELM) {
#pragma unroll
          for (int k = 0; k < N; ++k) mv[k] = v[k];
        }
      }
      __syncthreads();
    }

    // ---------------- quadrature-point phase (home): g <- G g  (bp5/step-64.cu:160-188)
    mbar_wait(bar, parity);
    parity ^= 1;
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int q = (k * N + b) * N + a, w = (k * N + b) * NP + a;
        const double ur = s1[w], us = s2[w], ut = t[k];
        const double g0 = gm[q], g1 = gm[N3 + q], g2 = gm[2 * N3 + q];
        const double g3 = gm[3 * N3 + q], g4 = gm[4 * N3 + q], g5 = gm[5 * N3 + q];
        s1[w] = ur * g0 + us * g3 + ut * g4;
        s2[w] = ur * g3 + us * g1 + ut * g5;
        t[k] = ur * g4 + us * g5 + ut * g2;
        if constexpr (HELM) mv[k] *= gm[6 * N3 + q];
      }
    }
    __syncthreads();
    // the metric buffer is free: fetch the next tile's metric behind the remaining work
    if (tid == 0 && tile + tstride < prm.n_tiles) {
      mbar_expect_tx(bar, Cfg::METRIC_BYTES);
      tma_load_1d(Gs, prm.metric + (tile + tstride) * (long long)Cfg::METRIC_DOUBLES, Cfg::METRIC_BYTES, bar, policy);
    }

    double out[N];
    if constexpr (QUAD == 1) {
      // (4) transposed derivative along x- and y-lines, in place
      if (active) {
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = s1[(b * N + a) * NP + i];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) s += Dt[m * N + i] * v[m];
          s1[(b * N + a) * NP + i] = s;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = s2[(b * N + j) * NP + a];
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) s += Dt[m * N + j] * v[m];
          s2[(b * N + j) * NP + a] = s;
        }
      }
      __syncthreads();
      // (5) home: z-transpose in registers, sum the three directions
      if (active) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const int w = (k * N + b) * NP + a;
          double s = s1[w] + s2[w];
#pragma unroll
          for (int m = 0; m < N; ++m) s += Dt[m * N + k] * t[m];
          if constexpr (HELM) s += mv[k];
          out[k] = s;
        }
      }
    } else {
      // (6a) x-line: D^T along x in place ; home: D^T along z (+ mass term) -> S0
      if (active) {
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = s1[(b * N + a) * NP + i];
#pragma unroll
        for (int rr = 0; rr < N; ++rr) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) s += Dt[q * N + rr] * v[q];
          s1[(b * N + a) * NP + rr] = s;
        }
#pragma unroll
        for (int rr = 0; rr < N; ++rr) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) s += Dt[q * N + rr] * t[q];
          if constexpr (HELM) s += mv[rr];
          s0[(rr * N + b) * NP + a] = s;
        }
      }
      __syncthreads();
      // (6b) y-line (qx=a, qz=b): D^T along y, add x and z parts, then B^T along y
      if (active) {
        double v[N], y[N];
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = s2[(b * N + q) * NP + a];
#pragma unroll
        for (int rr = 0; rr < N; ++rr) {
          double s = s1[(b * N + rr) * NP + a] + s0[(b * N + rr) * NP + a];
#pragma unroll
          for (int q = 0; q < N; ++q) s += Dt[q * N + rr] * v[q];
          y[rr] = s;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) s += Bm[q * N + j] * y[q];
          s0[(b * N + j) * NP + a] = s;
        }
      }
      __syncthreads();
      // (7) x-line (j=a, qz=b): B^T along x in place
      if (active) {
        double v[N];
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = s0[(b * N + a) * NP + q];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) s += Bm[q * N + i] * v[q];
          s0[(b * N + a) * NP + i] = s;
        }
      }
      __syncthreads();
      // (8) home (i=a, j=b): B^T along z in registers
      if (active) {
        double v[N];
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = s0[(q * N + b) * NP + a];
#pragma unroll
        for (int k = 0; k < N; ++k) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < N; ++q) s += Bm[q * N + k] * v[q];
          out[k] = s;
        }
      }
    }

    // distribute_local_to_global (bp5/fe_evaluation_gl.h:161-181): atomic add
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k)
        if (idx_cur[k] >= 0) atomicAdd(prm.dst + idx_cur[k], out[k]);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) { idx_cur[k] = idx_nxt[k]; idx_nxt[k] = idx_n2[k]; }
  }
}

}  // namespace bp5
