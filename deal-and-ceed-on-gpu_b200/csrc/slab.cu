// Slab-pipelined merged CG iteration (single block): the three passes of SolverCGFullMerge's iteration
// (bp5/solver.h:413-485) -- vector update, operator application, dot products -- walk through the mesh together,
// a few slabs of cells apart, so that what they hand to each other stays in the 126 MB L2:
//
//   U_g  update the DoF rows FIRST touched by the cells of slab g:  r += alpha h ; x += .. ; p = beta p - D r ; h = 0
//   C_g  the unmodified cell kernel (apply.cuh) on the tiles of slab g:  h += A p
//   D_g  the DoF rows LAST touched by the cells of slab g (complete now): Dirichlet rows h_c = p_c, sums h.h, r.h, ..
//
// Every pass is an ordinary kernel launch on its own stream; the order U_g -> C_g -> D_g, the look-ahead of U
// and the lag of D are CUDA events (edges of the captured graph).  No kernel spins on another one, the cell kernel
// and its read-only gathers are untouched (every C_g is a fresh launch, so the non-coherent path sees U_g's
// stores).  Per iteration the separate kernels of cg.cu move 12 vector passes through HBM (h is zero-filled,
// read-modify-written and re-read, p and r are re-read); here h, p and r are handed over in L2: 7 passes.
// (A single persistent kernel doing the same with in-kernel signals was built first and was slower:
// experiments/fused_iteration/README.md.)
//
// Rows: in the lexicographic cell order the DoFs a cell touches first are those of its upper-inclusive box, the
// DoFs it touches last those of its lower-inclusive box; with all x at once, U_g / D_g are bands of DoF rows
// (j, k), contiguous in memory per k.
#include <algorithm>
#include <cstdlib>

#include "apply.cuh"
#include "cg_state.cuh"
#include "tile_cells.h"

namespace bp5 {

struct SlabParams {
  double *r, *x, *p, *h;
  const double *diag;
  const CgState *st;
  double *partials;           // [n_slabs][grid][kSlabSums] of this kernel kind
  int umode;                  // 0 update_a0, 1 update_a, 3 update_a1 (solver.h:48-140)
  int od0, od1, od2, deg;
  int ncx, nry;               // cells per row, cell rows per layer
  int lc1, lc2;
  int n_cells, cells_per_slab;
};

constexpr int kSlabThreads = 128;
constexpr int kSlabBatch = 4;
constexpr int kSlabUSums = 2, kSlabDSums = 5;

__device__ __forceinline__ double sl_ld_stream(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void sl_st_stream(double *p, double v, uint64_t pol) {
  asm volatile("st.global.cg.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// DoF range of direction d (1: y, 2: z) covered by the cell rows a..b: UPD first-touched, else last-touched
template <bool UPD>
__device__ __forceinline__ void sl_range(int deg, int a, int b, int lc, int &lo, int &hi) {
  if (UPD) { lo = a * deg + (a == 0 ? 0 : 1); hi = b * deg + deg; }
  else { lo = a * deg; hi = b * deg + deg - 1 + (b == lc - 1 ? 1 : 0); }
}

// cell rows that slabs 0..g reach (UPD) / complete (!UPD)
template <bool UPD>
__device__ __forceinline__ int sl_rows_through(const SlabParams &sp, int g) {
  if (g < 0) return 0;
  long long cells = (long long)(g + 1) * sp.cells_per_slab;
  if (cells > sp.n_cells) cells = sp.n_cells;
  return (int)(UPD ? (cells + sp.ncx - 1) / sp.ncx : cells / sp.ncx);
}

// Walks this warp's contiguous block of the items (row, chunk of 32 x) of a band with additions only
struct SlabCursor {
  int rem, idx, x, j, k, row_base;
  int jlo, jend, od0, od1;
  __device__ __forceinline__ void start(const SlabParams &sp, int klo, int nk, int jlo_, int nj, int rot) {
    jlo = jlo_; jend = jlo_ + nj; od0 = sp.od0; od1 = sp.od1;
    const int nxc = (sp.od0 + 31) >> 5;
    const int n_items = nk * nj * nxc;
    const int G = (int)gridDim.x * (kSlabThreads / 32);
    const int per = (n_items + G - 1) / G;
    const int me = (int)((blockIdx.x * (kSlabThreads / 32) + (threadIdx.x >> 5) + (unsigned)rot * 61u) % (unsigned)G);
    const int first = me * per;
    rem = n_items - first;
    if (rem > per) rem = per;
    if (rem <= 0) { rem = 0; return; }
    const int row = first / nxc, xc = first - row * nxc;
    const int kk = row / nj, jj = row - kk * nj;
    j = jlo_ + jj; k = klo + kk;
    row_base = (k * od1 + j) * od0;
    x = xc << 5;
    idx = row_base + x;
  }
  __device__ __forceinline__ bool valid() const { return rem > 0; }
  __device__ __forceinline__ void next() {
    --rem;
    x += 32; idx += 32;
    if (x >= od0) {
      x = 0; ++j; row_base += od0;
      if (j == jend) { j = jlo; ++k; row_base = (k * od1 + j) * od0; }
      idx = row_base;
    }
  }
};

template <int NV>
__device__ __forceinline__ void sl_block_sums(double (&acc)[NV], double *out) {
  __shared__ double wsum[kSlabThreads / 32][NV];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wsum[threadIdx.x >> 5][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = 0.0;
    for (int w = 0; w < kSlabThreads / 32; ++w) v += wsum[w][threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// U_g.  partials[g][cta] = {r.r, r.Dr} of the residual written here.
template <int MODE, bool DIAG>
__global__ void __launch_bounds__(kSlabThreads) slab_update_kernel(const __grid_constant__ SlabParams sp, int g) {
  if (sp.st->state != 0) return;
  double acc[kSlabUSums] = {0.0, 0.0};
  const int ra = sl_rows_through<true>(sp, g - 1), rb = sl_rows_through<true>(sp, g);
  const int lane = threadIdx.x & 31;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  double alpha = 0.0, beta = 0.0, apa = 0.0, aob = 0.0;
  if (MODE != 0) { alpha = sp.st->alpha; beta = sp.st->beta; }
  if (MODE == 3) { aob = sp.st->alpha_old / sp.st->beta_old; apa = alpha + aob; }
  double *__restrict__ rvec = sp.r, *__restrict__ xvec = sp.x, *__restrict__ pvec = sp.p, *__restrict__ hvec = sp.h;
  const double *__restrict__ diag = sp.diag;
  if (rb > ra) {
    const int lz_a = ra / sp.nry, lz_b = (rb - 1) / sp.nry;
#pragma unroll 1
    for (int lz = lz_a; lz <= lz_b; ++lz) {
      const int ja = (lz == lz_a) ? ra - lz * sp.nry : 0, jb = (lz == lz_b) ? (rb - 1) - lz * sp.nry : sp.nry - 1;
      int klo, khi, jlo, jhi;
      sl_range<true>(sp.deg, lz, lz, sp.lc2, klo, khi);
      sl_range<true>(sp.deg, ja, jb, sp.lc1, jlo, jhi);
      SlabCursor cur;
      cur.start(sp, klo, khi - klo + 1, jlo, jhi - jlo + 1, g + lz);
#pragma unroll 1
      while (cur.valid()) {
        int idx[kSlabBatch];
        double rv[kSlabBatch], hv[kSlabBatch], pv[kSlabBatch], xv[kSlabBatch], dv[kSlabBatch];
#pragma unroll
        for (int u = 0; u < kSlabBatch; ++u) {
          idx[u] = (cur.valid() && cur.x + lane < sp.od0) ? cur.idx + lane : -1;
          if (cur.valid()) cur.next();
        }
#pragma unroll
        for (int u = 0; u < kSlabBatch; ++u) {
          rv[u] = hv[u] = pv[u] = xv[u] = 0.0; dv[u] = 1.0;
          if (idx[u] >= 0) {
            rv[u] = sl_ld_stream(rvec + idx[u], pol);
            if (MODE != 0) { hv[u] = sl_ld_stream(hvec + idx[u], pol); pv[u] = sl_ld_stream(pvec + idx[u], pol); }
            if (MODE == 3) xv[u] = sl_ld_stream(xvec + idx[u], pol);
            if (DIAG) dv[u] = sl_ld_stream(diag + idx[u], pol);
          }
        }
#pragma unroll
        for (int u = 0; u < kSlabBatch; ++u) {
          const int i = idx[u];
          if (i < 0) continue;
          double r_new = rv[u];
          if (MODE == 0) {
            __stcg(pvec + i, -dv[u] * r_new);
          } else {
            r_new = rv[u] + alpha * hv[u];
            if (MODE == 3) sl_st_stream(xvec + i, xv[u] + (apa * pv[u] + aob * dv[u] * rv[u]), pol);
            __stcg(rvec + i, r_new);
            __stcg(pvec + i, beta * pv[u] - dv[u] * r_new);
          }
          acc[0] += r_new * r_new;
          if (DIAG) acc[1] += r_new * dv[u] * r_new;
          __stcg(hvec + i, 0.0);
        }
      }
    }
  }
  sl_block_sums<kSlabUSums>(acc, sp.partials + ((size_t)g * gridDim.x + blockIdx.x) * kSlabUSums);
}

// D_g.  partials[g][cta] = {correction of p.h on Dirichlet rows, h.h, r.h, r.Dh, h.Dh}
template <bool DIAG>
__global__ void __launch_bounds__(kSlabThreads) slab_finish_kernel(const __grid_constant__ SlabParams sp, int g) {
  if (sp.st->state != 0) return;
  double acc[kSlabDSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int ra = sl_rows_through<false>(sp, g - 1), rb = sl_rows_through<false>(sp, g);
  const int lane = threadIdx.x & 31;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const double *__restrict__ rvec = sp.r, *__restrict__ pvec = sp.p;
  double *__restrict__ hvec = sp.h;
  const double *__restrict__ diag = sp.diag;
  // zero Dirichlet values on the whole boundary (bp5/step-64.cu:354-357)
  const int xd1 = sp.od0 - 1, jd1 = sp.od1 - 1, kd1 = sp.od2 - 1;
  if (rb > ra) {
    const int lz_a = ra / sp.nry, lz_b = (rb - 1) / sp.nry;
#pragma unroll 1
    for (int lz = lz_a; lz <= lz_b; ++lz) {
      const int ja = (lz == lz_a) ? ra - lz * sp.nry : 0, jb = (lz == lz_b) ? (rb - 1) - lz * sp.nry : sp.nry - 1;
      int klo, khi, jlo, jhi;
      sl_range<false>(sp.deg, lz, lz, sp.lc2, klo, khi);
      sl_range<false>(sp.deg, ja, jb, sp.lc1, jlo, jhi);
      SlabCursor cur;
      cur.start(sp, klo, khi - klo + 1, jlo, jhi - jlo + 1, g + lz);
#pragma unroll 1
      while (cur.valid()) {
        int idx[kSlabBatch];
        bool dir[kSlabBatch];
        double rv[kSlabBatch], hv[kSlabBatch], pv[kSlabBatch], dv[kSlabBatch];
#pragma unroll
        for (int u = 0; u < kSlabBatch; ++u) {
          const int x = cur.x + lane;
          const bool ok = cur.valid() && x < sp.od0;
          idx[u] = ok ? cur.idx + lane : -1;
          dir[u] = ok && (x == 0 || x == xd1 || cur.j == 0 || cur.j == jd1 || cur.k == 0 || cur.k == kd1);
          if (cur.valid()) cur.next();
        }
#pragma unroll
        for (int u = 0; u < kSlabBatch; ++u) {
          rv[u] = hv[u] = pv[u] = 0.0; dv[u] = 1.0;
          if (idx[u] >= 0) {
            rv[u] = sl_ld_stream(rvec + idx[u], pol);
            hv[u] = sl_ld_stream(hvec + idx[u], pol);
            if (DIAG) dv[u] = sl_ld_stream(diag + idx[u], pol);
            if (dir[u]) pv[u] = __ldcg(pvec + idx[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < kSlabBatch; ++u) {
          if (idx[u] < 0) continue;
          double vs = hv[u];
          if (dir[u]) {
            // copy_constrained_values (bp5/step-64.cu:275): h_c = p_c; the cell kernel summed p_c (A p)_c into p.h
            acc[0] += pv[u] * (pv[u] - vs);
            vs = pv[u];
            __stcg(hvec + idx[u], vs);
          }
          acc[1] += vs * vs;
          acc[2] += rv[u] * vs;
          if (DIAG) { const double dvs = dv[u] * vs; acc[3] += rv[u] * dvs; acc[4] += vs * dvs; }
        }
      }
    }
  }
  sl_block_sums<kSlabDSums>(acc, sp.partials + ((size_t)g * gridDim.x + blockIdx.x) * kSlabDSums);
}

// all partial sums of the iteration in a fixed order, then the scalar recurrences (solver.h:497-533)
__global__ void __launch_bounds__(256) slab_final_kernel(CgState *st, double *history, const double *pu, int n_pu,
                                                         const double *pd, int n_pd, const double *pc, int n_pc,
                                                         bool has_diag) {
  if (st->state != 0) return;
  __shared__ double sh[8][8];
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // 0,1: U sums; 2..6: D sums; 7: cell kernels' p.(A p)
  for (int i = threadIdx.x; i < n_pu; i += blockDim.x) { s[0] += pu[2 * i]; s[1] += pu[2 * i + 1]; }
  for (int i = threadIdx.x; i < n_pd; i += blockDim.x)
#pragma unroll
    for (int j = 0; j < 5; ++j) s[2 + j] += pd[5 * i + j];
  for (int i = threadIdx.x; i < n_pc; i += blockDim.x) s[7] += pc[i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double v = s[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w][j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[8];
    for (int j = 0; j < 8; ++j) { t[j] = 0.0; for (int q = 0; q < 8; ++q) t[j] += sh[q][j]; }
    double q[7];
    q[0] = t[7] + t[2]; q[1] = t[3]; q[2] = t[4]; q[3] = t[0];
    if (has_diag) { q[4] = t[5]; q[5] = t[6]; q[6] = t[1]; }
    else { q[4] = q[2]; q[5] = q[1]; q[6] = q[3]; }
    cg_scalar_step(st, q, history);
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct SlabPlan {
  int n_slabs = 0, tiles_per_slab = 0, grid_stream = 0, cell_grid = 0;
  cudaStream_t s_u = nullptr, s_c[2] = {nullptr, nullptr}, s_d = nullptr;
  cudaEvent_t e_fork = nullptr, e_join = nullptr, e_aux[3] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> e_u, e_c;
  double *pu = nullptr, *pd = nullptr, *pc = nullptr;     // partial sums of the U / D / cell kernels
};

// Opt-in (bp5_operator_set_option("slab_pipeline", 1) or BP5_SLAB=1): measured SLOWER than the separate kernels on
// B200 -- p = 6, 148 M DoFs: 6.0 ms per iteration with 48-96 slabs, 6.7 / 9.2 ms with 192 / 384, against 4.30 ms
// (profiles/r2_notes.md section 2).  The cell kernel is co-limited by the shared-memory pipe, so streaming kernels
// that share its SMs slow it down by more than the saved HBM passes give back, and every slab boundary costs ~10 us.
bool slab_supported(bp5_operator_t op) {
  if (op->prob.geometry_mode != BP5_GEOM_STORED || op->metric == nullptr) return false;
  if (op->prob.cell_order != BP5_CELL_ORDER_DEFAULT || op->hanging) return false;
  return op->prob.part_grid[0] * op->prob.part_grid[1] * op->prob.part_grid[2] == 1;
}

void slab_destroy(bp5_operator_t op) {
  SlabPlan *pl = static_cast<SlabPlan *>(op->slab);
  if (!pl) return;
  for (cudaStream_t s : {pl->s_u, pl->s_c[0], pl->s_c[1], pl->s_d}) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (cudaEvent_t e : pl->e_u) cudaEventDestroy(e);
  for (cudaEvent_t e : pl->e_c) cudaEventDestroy(e);
  if (pl->e_fork) cudaEventDestroy(pl->e_fork);
  if (pl->e_join) cudaEventDestroy(pl->e_join);
  for (cudaEvent_t e : pl->e_aux) if (e) cudaEventDestroy(e);
  cudaFree(pl->pu); cudaFree(pl->pd); cudaFree(pl->pc);
  delete pl;
  op->slab = nullptr;
}

static int slab_plan(bp5_operator_t op) {
  if (op->slab) return BP5_OK;
  SlabPlan *pl = new SlabPlan;
  op->slab = pl;
  bp5_context_t ctx = op->ctx;
  // the cell kernel's persistent grid (2-4 CTAs per SM): a slab is a whole number of rounds over it
  if (op->apply_grid_full == 0) {
    set_error("slab plan: the cell kernel has not been launched yet");
    return BP5_ERR_INVALID;
  }
  pl->cell_grid = op->apply_grid_full;
  int rounds = 4;                                        // BP5_SLAB_ROUNDS: tiles per CTA per slab
  if (const char *v = getenv("BP5_SLAB_ROUNDS")) rounds = std::max(1, atoi(v));
  pl->tiles_per_slab = pl->cell_grid * rounds;
  pl->n_slabs = (int)((op->n_tiles + pl->tiles_per_slab - 1) / pl->tiles_per_slab);
  pl->grid_stream = 2 * ctx->sm_count;
  BP5_CUDA(cudaStreamCreateWithFlags(&pl->s_u, cudaStreamNonBlocking));
  BP5_CUDA(cudaStreamCreateWithFlags(&pl->s_c[0], cudaStreamNonBlocking));
  BP5_CUDA(cudaStreamCreateWithFlags(&pl->s_c[1], cudaStreamNonBlocking));
  BP5_CUDA(cudaStreamCreateWithFlags(&pl->s_d, cudaStreamNonBlocking));
  BP5_CUDA(cudaEventCreateWithFlags(&pl->e_fork, cudaEventDisableTiming));
  BP5_CUDA(cudaEventCreateWithFlags(&pl->e_join, cudaEventDisableTiming));
  for (int a = 0; a < 3; ++a) BP5_CUDA(cudaEventCreateWithFlags(&pl->e_aux[a], cudaEventDisableTiming));
  pl->e_u.resize(pl->n_slabs); pl->e_c.resize(pl->n_slabs);
  for (int g = 0; g < pl->n_slabs; ++g) {
    BP5_CUDA(cudaEventCreateWithFlags(&pl->e_u[g], cudaEventDisableTiming));
    BP5_CUDA(cudaEventCreateWithFlags(&pl->e_c[g], cudaEventDisableTiming));
  }
  BP5_CUDA(cudaMalloc(&pl->pu, sizeof(double) * kSlabUSums * pl->n_slabs * pl->grid_stream));
  BP5_CUDA(cudaMalloc(&pl->pd, sizeof(double) * kSlabDSums * pl->n_slabs * pl->grid_stream));
  BP5_CUDA(cudaMalloc(&pl->pc, sizeof(double) * (size_t)pl->n_slabs * pl->cell_grid));
  BP5_CUDA(cudaMemsetAsync(pl->pc, 0, sizeof(double) * (size_t)pl->n_slabs * pl->cell_grid, ctx->stream));
  return BP5_OK;
}

template <int MODE>
static void launch_slab_update(bool has_diag, int grid, cudaStream_t s, const SlabParams &sp, int g) {
  if (has_diag) slab_update_kernel<MODE, true><<<grid, kSlabThreads, 0, s>>>(sp, g);
  else slab_update_kernel<MODE, false><<<grid, kSlabThreads, 0, s>>>(sp, g);
}

// one merged-CG iteration (cur = 1, 2, ...) enqueued as a fork / join on the context's stream
int slab_enqueue_iteration(bp5_operator_t op, int cur, void *state, double *hist_dev, double *g, double *d, double *h,
                           double *x, const double *diag) {
  int rc;
  if ((rc = slab_plan(op))) return rc;
  SlabPlan *pl = static_cast<SlabPlan *>(op->slab);
  bp5_context_t ctx = op->ctx;
  cudaStream_t s = ctx->stream;
  CgState *st = static_cast<CgState *>(state);
  SlabParams sp{};
  sp.r = g; sp.x = x; sp.p = d; sp.h = h; sp.diag = diag; sp.st = st;
  sp.umode = cur == 1 ? 0 : (cur % 2 == 0 ? 1 : 3);
  sp.od0 = op->od[0]; sp.od1 = op->od[1]; sp.od2 = op->od[2]; sp.deg = op->p;
  sp.ncx = op->lc[0]; sp.nry = op->lc[1]; sp.lc1 = op->lc[1]; sp.lc2 = op->lc[2];
  sp.n_cells = (int)op->n_cells;
  sp.cells_per_slab = pl->tiles_per_slab * op->cells_per_tile;
  const bool has_diag = diag != nullptr;
  const int n = pl->n_slabs, G = pl->grid_stream;
  int look = 2;                                          // slabs U runs ahead of the oldest cell kernel in flight
  if (const char *v = getenv("BP5_SLAB_AHEAD")) look = std::max(1, atoi(v));
  BP5_CUDA(cudaEventRecord(pl->e_fork, s));
  for (cudaStream_t a : {pl->s_u, pl->s_c[0], pl->s_c[1], pl->s_d}) BP5_CUDA(cudaStreamWaitEvent(a, pl->e_fork, 0));
  for (int k = 0; k < n; ++k) {
    // U_k: not before the cells of slab k - look - 1 are done (keeps the live window of r, p, h inside L2)
    if (k - look - 1 >= 0) BP5_CUDA(cudaStreamWaitEvent(pl->s_u, pl->e_c[k - look - 1], 0));
    SlabParams su = sp; su.partials = pl->pu;
    if (sp.umode == 0) launch_slab_update<0>(has_diag, G, pl->s_u, su, k);
    else if (sp.umode == 1) launch_slab_update<1>(has_diag, G, pl->s_u, su, k);
    else launch_slab_update<3>(has_diag, G, pl->s_u, su, k);
    BP5_CHECK_LAUNCH();
    BP5_CUDA(cudaEventRecord(pl->e_u[k], pl->s_u));
    // C_k on alternating streams (the tail of one slab overlaps the head of the next)
    cudaStream_t sc = pl->s_c[k & 1];
    BP5_CUDA(cudaStreamWaitEvent(sc, pl->e_u[k], 0));
    op->range_begin = (long long)k * pl->tiles_per_slab;
    op->range_end = std::min<long long>(op->n_tiles, (long long)(k + 1) * pl->tiles_per_slab);
    op->launch_stream = sc;
    rc = apply_cell_loop(op, h, d, true, pl->pc + (size_t)k * pl->cell_grid);
    op->range_begin = op->range_end = -1;
    op->launch_stream = nullptr;
    if (rc) return rc;
    BP5_CUDA(cudaEventRecord(pl->e_c[k], sc));
    // D_k: the rows completed by slab k (its cells and all earlier ones are done)
    BP5_CUDA(cudaStreamWaitEvent(pl->s_d, pl->e_c[k], 0));
    if (k > 0) BP5_CUDA(cudaStreamWaitEvent(pl->s_d, pl->e_c[k - 1], 0));
    SlabParams sd = sp; sd.partials = pl->pd;
    if (has_diag) slab_finish_kernel<true><<<G, kSlabThreads, 0, pl->s_d>>>(sd, k);
    else slab_finish_kernel<false><<<G, kSlabThreads, 0, pl->s_d>>>(sd, k);
    BP5_CHECK_LAUNCH();
  }
  // everything that ran on the other streams joins s_d (a captured graph must have no loose ends)
  for (int a = 0; a < 3; ++a) {
    cudaStream_t sa = a == 0 ? pl->s_u : pl->s_c[a - 1];
    BP5_CUDA(cudaEventRecord(pl->e_aux[a], sa));
    BP5_CUDA(cudaStreamWaitEvent(pl->s_d, pl->e_aux[a], 0));
  }
  slab_final_kernel<<<1, 256, 0, pl->s_d>>>(st, hist_dev, pl->pu, n * G, pl->pd, n * G, pl->pc, n * pl->cell_grid, has_diag);
  BP5_CHECK_LAUNCH();
  BP5_CUDA(cudaEventRecord(pl->e_join, pl->s_d));
  BP5_CUDA(cudaStreamWaitEvent(s, pl->e_join, 0));
  ctx->launches += 3 * n + 1;
  return BP5_OK;
}

}  // namespace bp5
