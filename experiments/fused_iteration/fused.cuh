// One CG iteration (or one vmult) as ONE persistent kernel: the vector updates and the dot products of
// SolverCGFullMerge (bp5/solver.h:413-485: update_a0/a/a1, update_b) travel through the mesh together with
// the cell loop, a few cell rows apart, so that everything the three phases hand to each other stays in L2:
//
//   U  update rows      r += alpha h ; x += ... ; p = beta p - D r ; h = 0     (streams r, h, p, x from HBM once)
//   C  cell tiles       h += A p                                             (p, h hit L2; the metric streams via TMA)
//   D  finished rows    h_c = p_c on Dirichlet rows ; h.h, r.h, r.Dh, h.Dh    (r, h hit L2; h goes to HBM once)
//
// Separate kernels move 12 vector passes per iteration through HBM (h is zero-filled, read-modify-written and
// re-read; p and r are re-read): here it is 7.
//
// Two kernels, running concurrently on two streams of the same device.  The cell kernel is the contraction code
// of apply.cuh with its registers and shared memory untouched (bp5_fused_kernel: 2-4 CTAs per SM); the streaming
// kernel (bp5_stream_kernel: one small CTA per SM, 8 warps, ~60 registers, no shared memory) does nothing but U
// and D.  They never share a CTA or a barrier: they talk through counters in global memory, so the streaming
// latency hides behind the contractions of the same SM instead of stalling them.  (A first version put a
// streaming warp into every cell CTA: one register budget and one instruction stream for both roles made the
// cell path 1.5x and the streaming path 10x slower -- profiles/r2_notes.md.)
//
// Schedule.  The cell tiles of the launch are dealt round-robin to the persistent CTAs as before; S consecutive
// rounds form a macro step.  In the lexicographic processing order the DoFs a cell touches FIRST are those of its
// upper-inclusive box and the DoFs it touches LAST those of its lower-inclusive box, so (all x at once)
//   U(K) = the DoF rows first touched by the cell rows that macro step K reaches,
//   D(K) = the DoF rows last touched by the cell rows that are complete after macro step K.
// Streaming CTAs, tick t:  U(t), signal; wait until all cell CTAs finished C(t - UA - DL), then D(t - UA - DL).
// Cell CTAs, step K:       wait until all streaming CTAs finished U(K + 1) (the gather prefetch reaches one
//                          tile into the next step); tiles of step K; signal.
// Signals are arrivals on a ring of monotone counters (slot = step mod 8, target = #CTAs of the signalling
// kernel per lap): nobody can run more than UA + DL <= 8 steps ahead of the slowest CTA, so laps never mix.
// All CTAs of both kernels must be co-resident (the host sizes the grids for that); a wait that cannot be
// satisfied gives up after ~4 s and latches an error word.
// The live window of r, p, h is ~(UA + DL + 1) macro steps of rows plus one DoF plane per cell layer -- tens of
// MB, inside the 126 MB L2; the read-once streams (metric via TMA, U's loads, D's loads) carry L2::evict_first.
//
// Partitioned blocks: the shell (DoFs touched by the cells on the lower ghost layers, and the upper faces that
// are sent to the neighbours) is updated and finished by separate small kernels around the halo exchange
// (cg.cu); this kernel then covers the interior cells and the non-shell rows (same rules, x range clipped).
#pragma once
#include "cg_state.cuh"

namespace bp5 {

enum : int { FUSE_U_CG0 = 0,    // update_a0 (solver.h:48-72):   p = -D r ; h = 0                    (iteration 1)
             FUSE_U_CG1 = 1,    // update_a  (solver.h:74-104):  r += alpha h ; p = beta p - D r ; h = 0
             FUSE_U_CG3 = 3,    // update_a1 (solver.h:106-140): the same and the two-step x update
             FUSE_U_ZERO = 4 }; // vmult: dst = 0 (bp5/step-64.cu:270-271), just ahead of the cells
enum : int { FUSE_D_CG = 0,     // Dirichlet rows + the sums of update_b (solver.h:142-311)
             FUSE_D_COPY = 1 }; // vmult: copy_constrained_values (bp5/step-64.cu:275)

constexpr int kFzRing = 8;                       // counter slots per direction
constexpr int kFzTicket = 2 * kFzRing, kFzErr = 2 * kFzRing + 1;
constexpr int kFzDbg = 2 * kFzRing + 2;           // debug tick counters (8 x 64 bit), tuning builds
constexpr int kFusedSyncWords = 2 * kFzRing + 2 + 16; // [0,8) update arrivals, [8,16) cell arrivals, ticket, error latch
// per-CTA partial sums, grouped by the phase that forms them:
//   U: 0 r.r  1 r.Dr      D: 2 correction of p.h on Dirichlet rows  3 h.h  4 r.h  5 r.Dh  6 h.Dh      C: 7 p.(A p)
constexpr int kFusedPartials = 8;
constexpr int kFzSlotU = 0, kFzSlotD = 2, kFzSlotC = 7;
constexpr int kFzStreamThreads = 192;            // streaming kernel: 6 warps per CTA, one CTA per SM

// -DBP5_FZ_DEBUG: tuning builds.  CTA 0 accumulates clock64 ticks per activity in sync[kFzDbg + i]
// (0 cell-side waits, 1 cell-side signal, 2 stream waits, 3 stream U, 4 stream D, 5 stream signals, 6 whole kernel)
// and `debug` bits switch parts off (1 no U work, 2 no D work, 4 cell signal without fence, 8 cells do not wait).
#ifdef BP5_FZ_DEBUG
#define FZ_DBG_T0() const long long dbg_t0 = clock64()
#define FZ_DBG_ADD(fz, i) do { if (blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long *>((fz).sync + kFzDbg) + (i), (unsigned long long)(clock64() - dbg_t0)); } while (0)
#else
#define FZ_DBG_T0() do {} while (0)
#define FZ_DBG_ADD(fz, i) do {} while (0)
#endif

struct FusedParams {
  double *r, *x;              // CG residual and solution (p = ApplyParams::src, h = ApplyParams::dst)
  const double *diag;         // DiagonalMatrix vector or nullptr (identity)
  CgState *st;
  double *history;
  double *partials;           // [gridDim.x][kFusedPartials], layout above
  double *sums_out;           // != nullptr: leave the 7 local sums here (partitioned blocks) instead of the scalar step
  unsigned *sync;
  int umode, dmode;
  int tiles_per_step;         // S
  int n_steps;                // macro steps of this launch
  int ua, dl;                 // update look-ahead / finish lag in macro steps (ua >= 2, dl >= 1, ua + dl <= kFzRing)
  int n_cell_ctas, n_stream_ctas;   // grids of the two kernels
  int n_cell_arrivals;        // cell-side arrivals per macro step: n_cell_ctas x warps per cell CTA
  double *pvec, *hvec;        // the streaming kernel's view of p (cell kernel: src) and h (dst)
  int debug;                  // tuning builds only
  int od0, od1, od2, p;
  int ncx, nry, nrz;          // interior cells per row, interior cell rows per layer, interior layers
  int lo[3], hi[3];           // a lower / upper neighbour block exists in direction d
  int lc1, lc2;               // local cells in y, z
  int n_inner;                // interior cells
  int cells_per_step;         // S * gridDim.x * CPT
};

__device__ __forceinline__ unsigned fz_ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// arrival of this CTA's group (cell warps or streaming warp) for macro step `step`; called by ONE thread after a
// barrier over the group (fences are cumulative: the group's earlier stores / reductions are ordered before it)
__device__ __forceinline__ void fz_signal(unsigned *ring, int step) {
  // release at gpu scope (MEMBAR.ALL.GPU + RED); __threadfence() would be the sequentially consistent fence
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ring + (step & (kFzRing - 1))) : "memory");
}

// one thread: wait until every CTA's group has arrived for `step`.  A CTA that never arrives (launch larger than
// the resident capacity, a crashed CTA) must not hang the GPU: after ~4 s the wait gives up and latches the error
// word; the host reports it when the solve ends.
__device__ __forceinline__ void fz_wait(const unsigned *ring, int step, int n_arrivals, unsigned *err) {
  const unsigned *ctr = ring + (step & (kFzRing - 1));
  const unsigned target = (unsigned)(step / kFzRing + 1) * (unsigned)n_arrivals;
  if (fz_ld_acquire(ctr) >= target) return;
  if (*reinterpret_cast<volatile unsigned *>(err) != 0) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  while (fz_ld_acquire(ctr) < target) {
    __nanosleep(32);
    if ((++spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull || *reinterpret_cast<volatile unsigned *>(err) != 0) {
        atomicExch(err, 1u);
        return;
      }
    }
  }
}

// streaming accesses of the U / D phases: L2 only (the data is written by other CTAs of the same launch, so the
// non-coherent L1 path is off limits), read-once streams with L2::evict_first
__device__ __forceinline__ double fz_ld_stream(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t make_evict_first_policy_fz() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void fz_st_stream(double *p, double v, uint64_t pol) {
  asm volatile("st.global.cg.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// DoF range of direction d covered by the interior cell rows a..b (relative to the first interior cell):
// UPD: first-touched (upper-inclusive boxes), else last-touched (lower-inclusive boxes); shell entries excluded
template <bool UPD>
__device__ __forceinline__ void fz_range(const FusedParams &fz, int d, int a, int b, int lc, int &lo_out, int &hi_out) {
  const int ca = a + fz.lo[d], cb = b + fz.lo[d];
  const int base_a = ca * fz.p - fz.lo[d], base_b = cb * fz.p - fz.lo[d];
  if (UPD) {
    lo_out = base_a + 1 - ((a == 0 && !fz.lo[d]) ? 1 : 0);
    hi_out = base_b + fz.p - ((cb == lc - 1 && fz.hi[d]) ? 1 : 0);
  } else {
    lo_out = base_a + ((a == 0 && fz.lo[d]) ? 1 : 0);
    hi_out = base_b + fz.p - 1 + ((cb == lc - 1 && !fz.hi[d]) ? 1 : 0);
  }
}

// number of interior cell rows that macro steps 0..K reach (UPD) / complete (!UPD)
template <bool UPD>
__device__ __forceinline__ int fz_rows_through(const FusedParams &fz, int K) {
  if (K < 0) return 0;
  if (K >= fz.n_steps) K = fz.n_steps - 1;
  unsigned cells = (unsigned)(K + 1) * (unsigned)fz.cells_per_step;      // < 2^31: checked by the host
  if (cells > (unsigned)fz.n_inner) cells = (unsigned)fz.n_inner;
  return (int)(UPD ? (cells + fz.ncx - 1) / (unsigned)fz.ncx : cells / (unsigned)fz.ncx);
}

// A band = the DoF rows (j, k) in [jlo, jlo+nj) x [klo, klo+nk), each cut into chunks of 32 consecutive x.
// Items (row, chunk), chunk fastest, are dealt to the streaming warps of the grid in contiguous blocks (a warp
// streams a few KB of consecutive memory per vector); the cursor walks a block with additions only.
struct FzCursor {
  int rem;                 // items left in this warp's block
  int idx;                 // vector index of lane 0's element of the current item
  int x;                   // its x
  int j, k;                // its row
  int row_base;            // index of (xlo, j, k)
  // band constants
  int xlo, xhi, jlo, jend, od0, od1, klo;
  __device__ __forceinline__ void start(const FusedParams &fz, int klo_, int nk, int jlo_, int nj, int xlo_, int xhi_,
                                        int rot) {
    xlo = xlo_; xhi = xhi_; jlo = jlo_; jend = jlo_ + nj; od0 = fz.od0; od1 = fz.od1; klo = klo_;
    const int nxc = (xhi_ - xlo_ + 32) >> 5;
    const int n_items = nk * nj * nxc;
    const int G = (int)gridDim.x * (kFzStreamThreads / 32);      // agents = warps of the streaming kernel
    const int per = (n_items + G - 1) / G;
    // rotate the dealing with the band so that the same warps do not always get the short block
    const int me = (int)((blockIdx.x * (kFzStreamThreads / 32) + (threadIdx.x >> 5) + (unsigned)rot * 61u) % (unsigned)G);
    const int first = me * per;
    rem = n_items - first;
    if (rem > per) rem = per;
    if (rem <= 0) { rem = 0; return; }
    const int row = first / nxc, xc = first - row * nxc;
    const int kk = row / nj, jj = row - kk * nj;
    j = jlo_ + jj; k = klo_ + kk;
    row_base = (k * od1 + j) * od0 + xlo_;
    x = xlo_ + (xc << 5);
    idx = row_base + (xc << 5);
  }
  __device__ __forceinline__ bool valid() const { return rem > 0; }
  __device__ __forceinline__ void next() {
    --rem;
    x += 32; idx += 32;
    if (x > xhi) {
      x = xlo; ++j; row_base += od0;
      if (j == jend) { j = jlo; ++k; row_base = (k * od1 + j) * od0 + xlo; }
      idx = row_base;
    }
  }
};

#ifndef BP5_FZ_BATCH
#define BP5_FZ_BATCH 4    // items in flight per warp of the streaming kernel (x 8 warps per SM)
#endif
constexpr int kFzBatch = BP5_FZ_BATCH;

// U(K): see the header comment.  acc[0] += r.r, acc[1] += r.Dr of the residual written here.  Called by every
// warp of the streaming kernel; MODE = FUSE_U_*.
template <int MODE, bool DIAG>
static __device__ __forceinline__ void fz_update_phase(const FusedParams &fz, int K, uint64_t pol, double (&acc)[2]) {
  const int ra = fz_rows_through<true>(fz, K - 1), rb = fz_rows_through<true>(fz, K);
  if (rb <= ra) return;
  const int lane = threadIdx.x & 31;
  double alpha = 0.0, beta = 0.0, apa = 0.0, aob = 0.0;
  if (MODE == FUSE_U_CG1 || MODE == FUSE_U_CG3) { alpha = fz.st->alpha; beta = fz.st->beta; }
  if (MODE == FUSE_U_CG3) { aob = fz.st->alpha_old / fz.st->beta_old; apa = alpha + aob; }
  double *__restrict__ rvec = fz.r, *__restrict__ xvec = fz.x, *__restrict__ pvec = fz.pvec, *__restrict__ hvec = fz.hvec;
  const double *__restrict__ diag = fz.diag;
  constexpr bool ld_r = MODE != FUSE_U_ZERO, ld_hp = MODE == FUSE_U_CG1 || MODE == FUSE_U_CG3, ld_x = MODE == FUSE_U_CG3;
  const int xlo = fz.lo[0] ? fz.p : 0, xhi = fz.od0 - 1 - fz.hi[0];
  if (xhi < xlo) return;
  const int lz_a = ra / fz.nry, lz_b = (rb - 1) / fz.nry;
#pragma unroll 1
  for (int lz = lz_a; lz <= lz_b; ++lz) {
    const int ja = (lz == lz_a) ? ra - lz * fz.nry : 0, jb = (lz == lz_b) ? (rb - 1) - lz * fz.nry : fz.nry - 1;
    int klo, khi, jlo, jhi;
    fz_range<true>(fz, 2, lz, lz, fz.lc2, klo, khi);
    fz_range<true>(fz, 1, ja, jb, fz.lc1, jlo, jhi);
    const int nj = jhi - jlo + 1, nk = khi - klo + 1;
    if (nj <= 0 || nk <= 0) continue;
    FzCursor cur;
    cur.start(fz, klo, nk, jlo, nj, xlo, xhi, K + lz);
#pragma unroll 1
    while (cur.valid()) {
      int idx[kFzBatch];
      double rv[kFzBatch], hv[kFzBatch], pv[kFzBatch], xv[kFzBatch], dv[kFzBatch];
#pragma unroll
      for (int u = 0; u < kFzBatch; ++u) {
        idx[u] = (cur.valid() && cur.x + lane <= xhi) ? cur.idx + lane : -1;
        if (cur.valid()) cur.next();
      }
#pragma unroll
      for (int u = 0; u < kFzBatch; ++u) {
        rv[u] = hv[u] = pv[u] = xv[u] = 0.0; dv[u] = 1.0;
        if (ld_r && idx[u] >= 0) {
          rv[u] = fz_ld_stream(rvec + idx[u], pol);
          if (ld_hp) { hv[u] = fz_ld_stream(hvec + idx[u], pol); pv[u] = fz_ld_stream(pvec + idx[u], pol); }
          if (ld_x) xv[u] = fz_ld_stream(xvec + idx[u], pol);
          if (DIAG) dv[u] = fz_ld_stream(diag + idx[u], pol);
        }
      }
#pragma unroll
      for (int u = 0; u < kFzBatch; ++u) {
        const int i = idx[u];
        if (i < 0) continue;
        if (ld_r) {
          double r_new = rv[u];
          if (!ld_hp) {
            __stcg(pvec + i, -dv[u] * r_new);
          } else {
            r_new = rv[u] + alpha * hv[u];
            if (ld_x) fz_st_stream(xvec + i, xv[u] + (apa * pv[u] + aob * dv[u] * rv[u]), pol);
            __stcg(rvec + i, r_new);
            __stcg(pvec + i, beta * pv[u] - dv[u] * r_new);
          }
          acc[0] += r_new * r_new;
          if (DIAG) acc[1] += r_new * dv[u] * r_new;
        }
        __stcg(hvec + i, 0.0);
      }
    }
  }
}

// D(K): see the header comment.  acc: 0 correction of p.h, 1 h.h, 2 r.h, 3 r.Dh, 4 h.Dh.
template <bool CG, bool DIAG>
static __device__ __forceinline__ void fz_finish_phase(const FusedParams &fz, int K, uint64_t pol, double (&acc)[5]) {
  const int ra = fz_rows_through<false>(fz, K - 1), rb = fz_rows_through<false>(fz, K);
  if (rb <= ra) return;
  const int lane = threadIdx.x & 31;
  const double *__restrict__ rvec = fz.r, *__restrict__ pvec = fz.pvec;
  double *__restrict__ hvec = fz.hvec;
  const double *__restrict__ diag = fz.diag;
  const int xlo = fz.lo[0] ? fz.p : 0, xhi = fz.od0 - 1 - fz.hi[0];
  if (xhi < xlo) return;
  const int lz_a = ra / fz.nry, lz_b = (rb - 1) / fz.nry;
  // zero Dirichlet values on the whole global boundary (bp5/step-64.cu:354-357): faces without a neighbour block
  const int xd0 = fz.lo[0] ? -1 : 0, xd1 = fz.hi[0] ? -1 : fz.od0 - 1;
  const int jd0 = fz.lo[1] ? -1 : 0, jd1 = fz.hi[1] ? -1 : fz.od1 - 1;
  const int kd0 = fz.lo[2] ? -1 : 0, kd1 = fz.hi[2] ? -1 : fz.od2 - 1;
#pragma unroll 1
  for (int lz = lz_a; lz <= lz_b; ++lz) {
    const int ja = (lz == lz_a) ? ra - lz * fz.nry : 0, jb = (lz == lz_b) ? (rb - 1) - lz * fz.nry : fz.nry - 1;
    int klo, khi, jlo, jhi;
    fz_range<false>(fz, 2, lz, lz, fz.lc2, klo, khi);
    fz_range<false>(fz, 1, ja, jb, fz.lc1, jlo, jhi);
    const int nj = jhi - jlo + 1, nk = khi - klo + 1;
    if (nj <= 0 || nk <= 0) continue;
    FzCursor cur;
    cur.start(fz, klo, nk, jlo, nj, xlo, xhi, K + lz);
#pragma unroll 1
    while (cur.valid()) {
      int idx[kFzBatch];
      bool dir[kFzBatch];
      double rv[kFzBatch], hv[kFzBatch], pv[kFzBatch], dv[kFzBatch];
#pragma unroll
      for (int u = 0; u < kFzBatch; ++u) {
        const int x = cur.x + lane;
        const bool ok = cur.valid() && x <= xhi;
        idx[u] = ok ? cur.idx + lane : -1;
        dir[u] = ok && (x == xd0 || x == xd1 || cur.j == jd0 || cur.j == jd1 || cur.k == kd0 || cur.k == kd1);
        if (cur.valid()) cur.next();
      }
#pragma unroll
      for (int u = 0; u < kFzBatch; ++u) {
        rv[u] = hv[u] = pv[u] = 0.0; dv[u] = 1.0;
        if (idx[u] >= 0) {
          if (CG) {
            rv[u] = fz_ld_stream(rvec + idx[u], pol);
            hv[u] = fz_ld_stream(hvec + idx[u], pol);
            if (DIAG) dv[u] = fz_ld_stream(diag + idx[u], pol);
          }
          if (dir[u]) pv[u] = __ldcg(pvec + idx[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kFzBatch; ++u) {
        if (idx[u] < 0) continue;
        double vs = hv[u];
        if (dir[u]) {
          // copy_constrained_values: h_c = p_c; the cell kernel summed p_c (A p)_c into p.h
          acc[0] += pv[u] * (pv[u] - vs);
          vs = pv[u];
          __stcg(hvec + idx[u], vs);
        }
        if (CG) {
          acc[1] += vs * vs;
          acc[2] += rv[u] * vs;
          if (DIAG) { const double dvs = dv[u] * vs; acc[3] += rv[u] * dvs; acc[4] += vs * dvs; }
        }
      }
    }
  }
}

__device__ __forceinline__ void fz_finalize(const FusedParams &fz);

// The streaming kernel: U runs `ua` macro steps ahead of the cells, D `dl` behind.
// Tick t: U(t), signal; then D(t - ua - dl) once every cell CTA has finished that step.
template <int UMODE, bool CG, bool DIAG>
__global__ void __launch_bounds__(kFzStreamThreads, 4) bp5_stream_kernel(const __grid_constant__ FusedParams fz) {
  if (fz.st != nullptr && fz.st->state != 0) return;         // CG already converged: no-op like every other kernel
  unsigned *u_ring = fz.sync, *c_ring = fz.sync + kFzRing, *err = fz.sync + kFzErr;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t pol = make_evict_first_policy_fz();
  double acc_u[2] = {0.0, 0.0}, acc_d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int n_steps = fz.n_steps, lag = fz.ua + fz.dl;
#pragma unroll 1
  for (int t = 0; t < n_steps + lag; ++t) {
    if (t < n_steps + fz.ua) {
      { FZ_DBG_T0(); if (!(fz.debug & 1)) fz_update_phase<UMODE, DIAG>(fz, t, pol, acc_u); __syncthreads(); if (tid == 0) FZ_DBG_ADD(fz, 3); }
      { FZ_DBG_T0(); if (tid == 0) { fz_signal(u_ring, t); FZ_DBG_ADD(fz, 5); } }
    }
    const int kd = t - lag;
    if (kd >= 0 && !(fz.debug & 32)) {
      { FZ_DBG_T0(); if (tid == 0) { fz_wait(c_ring, kd, fz.n_cell_arrivals, err); FZ_DBG_ADD(fz, 2); } }
      __syncthreads();
      { FZ_DBG_T0(); if (!(fz.debug & 2)) fz_finish_phase<CG, DIAG>(fz, kd, pol, acc_d); if (tid == 0) FZ_DBG_ADD(fz, 4); }
    }
  }
  // per-CTA partial sums of the two phases: lanes in butterfly order, then the warps in order
  __shared__ double wsum[kFzStreamThreads / 32][8];
  double v7[7] = {acc_u[0], acc_u[1], acc_d[0], acc_d[1], acc_d[2], acc_d[3], acc_d[4]};
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    double v = v7[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wsum[tid >> 5][j] = v;
  }
  __syncthreads();
  if (tid < 7) {
    double v = 0.0;
    for (int w = 0; w < kFzStreamThreads / 32; ++w) v += wsum[w][tid];
    fz.partials[(size_t)(fz.n_cell_ctas + blockIdx.x) * kFusedPartials + tid] = v;      // slots 0..6 = U | D order
  }
  fz_finalize(fz);
}

// End of the iteration: the last CTA of EITHER kernel to get here adds the per-CTA sums in CTA order
// (deterministic for given grids), runs the scalar recurrences (or leaves the local sums for the caller's all-rank
// sum) and re-arms the counters.  partials: rows [0, n_cell_ctas) hold the cell kernel's p.(A p) in slot 7, rows
// [n_cell_ctas, n_cell_ctas + n_stream_ctas) the streaming kernel's seven sums in slots 0..6.
__device__ __forceinline__ void fz_finalize(const FusedParams &fz) {
  __shared__ bool last;
  __shared__ double tot[kFusedPartials];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(fz.sync + kFzTicket, 1u) == (unsigned)(fz.n_cell_ctas + fz.n_stream_ctas) - 1);
    if (last) __threadfence();
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < kFusedPartials) {
    double s = 0.0;
    if (threadIdx.x == kFzSlotC)
      for (int b = 0; b < fz.n_cell_ctas; ++b) s += __ldcg(fz.partials + (size_t)b * kFusedPartials + kFzSlotC);
    else
      for (int b = 0; b < fz.n_stream_ctas; ++b)
        s += __ldcg(fz.partials + (size_t)(fz.n_cell_ctas + b) * kFusedPartials + threadIdx.x);
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2 * kFzRing + 1) fz.sync[threadIdx.x] = 0;      // every CTA is past its last wait: re-arm
  if (threadIdx.x == 0) {
    if (fz.dmode == FUSE_D_CG) {
      double q[7];
      q[0] = tot[kFzSlotC] + tot[kFzSlotD];                // p.h = sum_q g^T G g + Dirichlet correction
      q[1] = tot[kFzSlotD + 1]; q[2] = tot[kFzSlotD + 2]; q[3] = tot[kFzSlotU];
      if (fz.diag) { q[4] = tot[kFzSlotD + 3]; q[5] = tot[kFzSlotD + 4]; q[6] = tot[kFzSlotU + 1]; }
      else { q[4] = q[2]; q[5] = q[1]; q[6] = q[3]; }
      if (fz.sums_out) {
#pragma unroll
        for (int j = 0; j < 7; ++j) fz.sums_out[j] = q[j];
      } else {
        cg_scalar_step(fz.st, q, fz.history);
      }
    }
  }
}

}  // namespace bp5
