// Host side of the fused iteration (fused.cuh): two kernels on two streams -- the cell kernel
// (bp5_fused_kernel) and the streaming kernel (bp5_stream_kernel) -- synchronising through counters in global
// memory.  Grid sizing makes every CTA of both kernels co-resident: one streaming CTA per SM plus as many cell
// CTAs as the registers, threads and shared memory left over by it allow.
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

static int ensure_aux_stream(bp5_context_t ctx) {
  if (ctx->stream2) return BP5_OK;
  BP5_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  BP5_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  BP5_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  return BP5_OK;
}

using StreamKernel = void (*)(const FusedParams);
static StreamKernel stream_kernel_for(int umode, bool cg, bool diag) {
  if (!cg) return bp5_stream_kernel<FUSE_U_ZERO, false, false>;
  switch (umode) {
    case FUSE_U_CG0: return diag ? bp5_stream_kernel<FUSE_U_CG0, true, true> : bp5_stream_kernel<FUSE_U_CG0, true, false>;
    case FUSE_U_CG1: return diag ? bp5_stream_kernel<FUSE_U_CG1, true, true> : bp5_stream_kernel<FUSE_U_CG1, true, false>;
    case FUSE_U_CG3: return diag ? bp5_stream_kernel<FUSE_U_CG3, true, true> : bp5_stream_kernel<FUSE_U_CG3, true, false>;
  }
  return nullptr;
}

// registers a CTA takes out of the SM's file: per-warp allocation in units of 256 registers
static int cta_regs(int regs_per_thread, int threads) {
  const int per_warp = ((regs_per_thread * 32 + 255) / 256) * 256;
  return per_warp * ((threads + 31) / 32);
}

template <int P, int QUAD, int HELM>
static int launch_fused(bp5_operator_t op, double *dst, const double *src, const FusedCall &call) {
  constexpr int CPT = TileCells<P>::value;
  using Cfg = ApplyCfg<P, CPT, 6 + HELM, 0>;
  constexpr int N = P + 1;
  bp5_context_t ctx = op->ctx;
  int rc;
  if ((rc = ensure_aux_stream(ctx))) return rc;
  auto kernel = bp5_fused_kernel<P, QUAD, HELM, CPT>;
  const bool cg = call.dmode == FUSE_D_CG;
  StreamKernel skernel = stream_kernel_for(call.umode, cg, call.diag != nullptr);
  BP5_REQUIRE(skernel != nullptr, "bad fused update mode");
  static int blocks_per_sm_of[64] = {0};
  int &blocks_per_sm = blocks_per_sm_of[ctx->device & 63];
  if (blocks_per_sm == 0) {
    BP5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    BP5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int nb = 0;
    BP5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, Cfg::NT, Cfg::SMEM_BYTES));
    // room for one streaming CTA next to the cell CTAs of an SM (worst case over the streaming instantiations)
    cudaFuncAttributes fa, fs;
    BP5_CUDA(cudaFuncGetAttributes(&fa, kernel));
    int stream_regs = 0;
    size_t stream_smem = 0;
    const StreamKernel all[] = {stream_kernel_for(FUSE_U_ZERO, false, false), stream_kernel_for(FUSE_U_CG0, true, false),
                                stream_kernel_for(FUSE_U_CG1, true, false), stream_kernel_for(FUSE_U_CG3, true, false),
                                stream_kernel_for(FUSE_U_CG0, true, true), stream_kernel_for(FUSE_U_CG1, true, true),
                                stream_kernel_for(FUSE_U_CG3, true, true)};
    for (StreamKernel k : all) {
      BP5_CUDA(cudaFuncGetAttributes(&fs, k));
      stream_regs = std::max(stream_regs, fs.numRegs);
      stream_smem = std::max(stream_smem, fs.sharedSizeBytes);
    }
    cudaDeviceProp prop;
    BP5_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    const int regs_left = prop.regsPerMultiprocessor - cta_regs(stream_regs, kFzStreamThreads);
    const int by_regs = regs_left / cta_regs(fa.numRegs, Cfg::NT);
    const int by_threads = (prop.maxThreadsPerMultiProcessor - kFzStreamThreads) / Cfg::NT;
    const long long smem_left = (long long)prop.sharedMemPerMultiprocessor - (long long)(stream_smem + 1024);
    const int by_smem = (int)(smem_left / (long long)(Cfg::SMEM_BYTES + fa.sharedSizeBytes + 1024));
    nb = std::min(std::min(nb, by_regs), std::min(by_threads, by_smem));
    if (const char *cv = getenv("BP5_FUSE_BLOCKS")) nb = std::min(nb, std::max(1, atoi(cv)));
    BP5_REQUIRE(nb > 0, "fused cell kernel does not fit on an SM next to the streaming kernel");
    blocks_per_sm = nb;
  }
  const long long n_int_tiles = op->n_tiles - op->n_boundary_tiles;
  BP5_REQUIRE(n_int_tiles > 0, "no interior tiles");
  long long grid = (long long)blocks_per_sm * ctx->sm_count;
  if (grid > n_int_tiles) grid = n_int_tiles;
  const int sgrid = ctx->sm_count;                 // one streaming CTA per SM
  if (!op->fz_sync) {
    BP5_CUDA(cudaMalloc(&op->fz_sync, sizeof(unsigned) * kFusedSyncWords));
    BP5_CUDA(cudaMemsetAsync(op->fz_sync, 0, sizeof(unsigned) * kFusedSyncWords, ctx->stream));
  }
  if (op->fz_partials_cap < grid + sgrid) {
    if (op->fz_partials) { BP5_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(op->fz_partials); op->fz_partials = nullptr; }
    BP5_CUDA(cudaMalloc(&op->fz_partials, sizeof(double) * kFusedPartials * (grid + sgrid)));
    op->fz_partials_cap = (int)(grid + sgrid);
  }
  // tiles per cell CTA per macro step: the window of r, p, h kept in L2 grows with it, the number of signals
  // shrinks with it (BP5_FUSE_S overrides for tuning)
  int S = 4;
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
  fz.n_cell_ctas = (int)grid; fz.n_stream_ctas = sgrid;
  fz.n_cell_arrivals = (int)grid * (Cfg::NT / 32);      // every warp of a cell CTA signals for itself
  fz.pvec = const_cast<double *>(src); fz.hvec = dst;
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
    BP5_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  // fork: the cell kernel on the context's stream, the streaming kernel on the auxiliary stream; join.
  // The cell kernel goes first: its CTAs fill every SM up to the computed number, which leaves room for exactly
  // one streaming CTA per SM.
  BP5_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  BP5_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  kernel<<<(unsigned)grid, Cfg::NT, Cfg::SMEM_BYTES, ctx->stream>>>(prm);
  BP5_CHECK_LAUNCH();
  skernel<<<(unsigned)sgrid, kFzStreamThreads, 0, ctx->stream2>>>(fz);
  BP5_CHECK_LAUNCH();
  BP5_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream2));
  BP5_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  if (e1) BP5_CUDA(cudaEventRecord(e1, ctx->stream));
  ctx->launches += 2;
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
    set_error("fused CG iteration: a grid-wide signal timed out (the cell and streaming kernels were not co-resident)");
    return BP5_ERR_CUDA;
  }
  return BP5_OK;
}

}  // namespace bp5
