// Host side of the fused per-iteration kernel (fused.cuh): grid sizing (all CTAs must be co-resident: they
// synchronise through counters in global memory), macro-step plan, launch.
#include <cstdlib>

#include "apply.cuh"
#include "tile_cells.h"

namespace bp5 {

bool apply_fused_supported(bp5_operator_t op) {
  if (op->prob.geometry_mode != BP5_GEOM_STORED || op->metric == nullptr) return false;
  // every interior cell row must exist: at least one interior cell in each direction
  for (int d = 0; d < 3; ++d)
    if (op->lc[d] - op->has_lo[d] < 1) return false;
  static const bool off = getenv("BP5_NO_FUSE") != nullptr;
  return !off;
}

template <int P, int QUAD, int HELM>
static int launch_fused(bp5_operator_t op, double *dst, const double *src, const FusedCall &call) {
  constexpr int CPT = TileCells<P>::value;
  using Cfg = ApplyCfg<P, CPT, 6 + HELM, 0>;
  constexpr int N = P + 1;
  auto kernel = bp5_fused_kernel<P, QUAD, HELM, CPT>;
  static int blocks_per_sm_of[64] = {0};
  int &blocks_per_sm = blocks_per_sm_of[op->ctx->device & 63];
  if (blocks_per_sm == 0) {
    BP5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES_FUSED));
    BP5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int nb = 0;
    BP5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, Cfg::NT_FUSED, Cfg::SMEM_BYTES_FUSED));
    BP5_REQUIRE(nb > 0, "fused kernel does not fit on an SM");
    if (const char *cv = getenv("BP5_FUSE_BLOCKS")) nb = std::min(nb, std::max(1, atoi(cv)));
    blocks_per_sm = nb;
  }
  const long long n_int_tiles = op->n_tiles - op->n_boundary_tiles;
  BP5_REQUIRE(n_int_tiles > 0, "no interior tiles");
  long long grid = (long long)blocks_per_sm * op->ctx->sm_count;
  if (grid > n_int_tiles) grid = n_int_tiles;
  if (!op->fz_sync) {
    BP5_CUDA(cudaMalloc(&op->fz_sync, sizeof(unsigned) * kFusedSyncWords));
    BP5_CUDA(cudaMemsetAsync(op->fz_sync, 0, sizeof(unsigned) * kFusedSyncWords, op->ctx->stream));
  }
  if (op->fz_partials_cap < grid) {
    if (op->fz_partials) { BP5_CUDA(cudaStreamSynchronize(op->ctx->stream)); cudaFree(op->fz_partials); op->fz_partials = nullptr; }
    BP5_CUDA(cudaMalloc(&op->fz_partials, sizeof(double) * kFusedPartials * grid));
    op->fz_partials_cap = (int)grid;
  }
  // tiles per CTA per macro step: the window of r, p, h kept in L2 grows with it, the number of grid-wide
  // barriers shrinks with it (BP5_FUSE_S overrides for tuning)
  int S = 1;
  if (const char *sv = getenv("BP5_FUSE_S")) S = std::max(1, atoi(sv));
  const long long rounds = (n_int_tiles + grid - 1) / grid;
  ApplyParams<N> prm;
  prm.metric = op->metric; prm.cell_base = op->cell_base; prm.l2g_irr = op->l2g_irr;
  prm.src = src; prm.dst = dst;
  prm.tile_begin = op->n_boundary_tiles;
  prm.n_tiles = op->n_tiles;
  prm.sy = op->od[0]; prm.sz = op->od[0] * op->od[1];
  prm.skip = op->skip_flag;
  prm.dot_partials = nullptr;
  FusedParams &fz = prm.fz;
  fz.r = call.r; fz.x = call.x; fz.diag = call.diag;
  fz.st = static_cast<CgState *>(call.state);
  fz.history = call.history;
  fz.partials = op->fz_partials;
  fz.sums_out = call.sums_out;
  fz.sync = op->fz_sync;
  fz.umode = call.umode; fz.dmode = call.dmode;
  fz.tiles_per_step = S;
  fz.n_steps = (int)((rounds + S - 1) / S);
  fz.ua = 2; fz.dl = 2;
  if (const char *v = getenv("BP5_FUSE_UA")) fz.ua = std::max(2, atoi(v));
  if (const char *v = getenv("BP5_FUSE_DL")) fz.dl = std::max(1, atoi(v));
  BP5_REQUIRE(fz.ua + fz.dl <= kFzRing, "update look-ahead + finish lag must fit the counter ring");
  fz.debug = 0;
#ifdef BP5_FZ_DEBUG
  if (const char *v = getenv("BP5_FUSE_DEBUG")) fz.debug = atoi(v);
#endif
  fz.od0 = op->od[0]; fz.od1 = op->od[1]; fz.od2 = op->od[2]; fz.p = op->p;
  fz.ncx = op->lc[0] - op->has_lo[0]; fz.nry = op->lc[1] - op->has_lo[1]; fz.nrz = op->lc[2] - op->has_lo[2];
  for (int d = 0; d < 3; ++d) { fz.lo[d] = op->has_lo[d]; fz.hi[d] = op->has_hi[d]; }
  fz.lc1 = op->lc[1]; fz.lc2 = op->lc[2];
  const long long n_inner = (long long)fz.ncx * fz.nry * fz.nrz, cps = (long long)S * grid * CPT;
  BP5_REQUIRE(n_inner + cps < 2147483647LL, "block too large for the fused kernel's 32-bit cell counters");
  fz.n_inner = (int)n_inner;
  fz.cells_per_step = (int)cps;
  fill_kernel_tables<N>(prm.tab, op->tab.B, op->tab.Dt);
  op->apply_grid = (int)grid;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (op->profile) {
    if (op->prof_used + 2 > op->prof_events.size()) {
      for (int i = 0; i < 64; ++i) { cudaEvent_t e; BP5_CUDA(cudaEventCreate(&e)); op->prof_events.push_back(e); }
    }
    e0 = op->prof_events[op->prof_used++]; e1 = op->prof_events[op->prof_used++];
    BP5_CUDA(cudaEventRecord(e0, op->ctx->stream));
  }
  kernel<<<(unsigned)grid, Cfg::NT_FUSED, Cfg::SMEM_BYTES_FUSED, op->ctx->stream>>>(prm);
  BP5_CHECK_LAUNCH();
  if (e1) BP5_CUDA(cudaEventRecord(e1, op->ctx->stream));
  op->ctx->launches++;
  return BP5_OK;
}

template <int P>
static int launch_fused_p(bp5_operator_t op, double *dst, const double *src, const FusedCall &call) {
  const bool gll = op->prob.quadrature == BP5_QUAD_GLL;
  const bool helm = op->prob.operator_kind == BP5_OP_HELMHOLTZ;
  return gll ? (helm ? launch_fused<P, 1, 1>(op, dst, src, call) : launch_fused<P, 1, 0>(op, dst, src, call))
             : (helm ? launch_fused<P, 0, 1>(op, dst, src, call) : launch_fused<P, 0, 0>(op, dst, src, call));
}

int apply_fused(bp5_operator_t op, double *dst, const double *src, const FusedCall &call) {
  BP5_REQUIRE(apply_fused_supported(op), "fused kernel not available for this operator");
  switch (op->p) {
#ifdef BP5_FZ_ONLY_P6    // tuning builds: one instantiation, fast to compile
    case 6: return launch_fused<6, 1, 0>(op, dst, src, call);
#else
    case 1: return launch_fused_p<1>(op, dst, src, call);
    case 2: return launch_fused_p<2>(op, dst, src, call);
    case 3: return launch_fused_p<3>(op, dst, src, call);
    case 4: return launch_fused_p<4>(op, dst, src, call);
    case 5: return launch_fused_p<5>(op, dst, src, call);
    case 6: return launch_fused_p<6>(op, dst, src, call);
    case 7: return launch_fused_p<7>(op, dst, src, call);
    case 8: return launch_fused_p<8>(op, dst, src, call);
#endif
  }
  set_error("unsupported degree %d", op->p);
  return BP5_ERR_UNSUPPORTED;
}

#ifdef BP5_FZ_DEBUG
// tuning builds: print and reset CTA 0's tick counters
extern "C" void bp5_debug_fused_ticks(bp5_operator_t op) {
  if (!op->fz_sync) return;
  unsigned long long t[8];
  cudaStreamSynchronize(op->ctx->stream);
  cudaMemcpy(t, op->fz_sync + kFzDbg, sizeof(t), cudaMemcpyDeviceToHost);
  cudaMemset(op->fz_sync + kFzDbg, 0, sizeof(t));
  fprintf(stderr, "fused ticks (CTA 0): cell wait %llu, cell signal %llu, stream wait %llu, stream U %llu, stream D %llu, "
          "stream signal %llu, kernel %llu\n", t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
}
#endif

int apply_fused_check(bp5_operator_t op) {
  if (!op->fz_sync) return BP5_OK;
  unsigned err = 0;
  BP5_CUDA(cudaMemcpyAsync(&err, op->fz_sync + kFzErr, sizeof(unsigned), cudaMemcpyDeviceToHost, op->ctx->stream));
  BP5_CUDA(cudaStreamSynchronize(op->ctx->stream));
  if (err != 0) {
    // re-arm: counters and latch (the launch that timed out left them in an undefined state)
    BP5_CUDA(cudaMemsetAsync(op->fz_sync, 0, sizeof(unsigned) * kFusedSyncWords, op->ctx->stream));
    set_error("fused CG kernel: a grid-wide barrier timed out (not all CTAs were co-resident)");
    return BP5_ERR_CUDA;
  }
  return BP5_OK;
}

}  // namespace bp5
