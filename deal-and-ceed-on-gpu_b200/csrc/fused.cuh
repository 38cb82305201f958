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
// Warp specialisation.  Every CTA has its cell warps (the n^2-threads-per-cell contraction code of apply.cuh,
// unchanged) and ONE extra streaming warp that does nothing but U and D.  The two never meet at a CTA barrier
// (the cell warps use a named barrier): they talk through counters in global memory, so the streaming latency
// is hidden behind the contractions of the same SM instead of stalling them.
//
// Schedule.  The cell tiles of the launch are dealt round-robin to the persistent CTAs as before; S consecutive
// rounds form a macro step.  In the lexicographic processing order the DoFs a cell touches FIRST are those of its
// upper-inclusive box and the DoFs it touches LAST those of its lower-inclusive box, so (all x at once)
//   U(K) = the DoF rows first touched by the cell rows that macro step K reaches,
//   D(K) = the DoF rows last touched by the cell rows that are complete after macro step K.
// Streaming warp, step K:  U(K + UA), signal; wait until all cell groups finished C(K - DL), D(K - DL).
// Cell warps, step K:      wait until all streaming warps finished U(K + 1) (the gather prefetch reaches one
//                          tile into the next step); tiles of step K; signal.
// Signals are arrivals on a ring of monotone counters (slot = step mod 8, target = G per lap): a CTA can only
// run UA + DL <= 8 steps ahead of the slowest one, so laps never mix.  All CTAs must be co-resident (grid <=
// occupancy x SMs); a wait that cannot be satisfied gives up after ~4 s and latches an error word.
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
constexpr int kFzStreamThreads = 32;             // one streaming warp per CTA

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
  __threadfence();
  atomicAdd(ring + (step & (kFzRing - 1)), 1u);
}

// one thread: wait until every CTA's group has arrived for `step`.  A CTA that never arrives (launch larger than
// the resident capacity, a crashed CTA) must not hang the GPU: after ~4 s the wait gives up and latches the error
// word; the host reports it when the solve ends.
__device__ __forceinline__ void fz_wait(const unsigned *ring, int step, unsigned *err) {
  const unsigned *ctr = ring + (step & (kFzRing - 1));
  const unsigned target = (unsigned)(step / kFzRing + 1) * gridDim.x;
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
    const int G = (int)gridDim.x;
    const int per = (n_items + G - 1) / G;
    // rotate the dealing with the band so that the same CTAs do not always get the short block
    const int me = (int)((blockIdx.x + (unsigned)rot * 61u) % (unsigned)G);
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
#define BP5_FZ_BATCH 8    // items per stage of the streaming warp (memory-level parallelism without occupancy)
#endif
constexpr int kFzBatch = BP5_FZ_BATCH;
constexpr int kFzStageVecs = 5;                                              // r, h, p, x, diag
constexpr int kFzStageDoubles = (kFzStageVecs * kFzBatch + kFzBatch / 2) * 32;   // one stage: [vec][item][lane] + int idx[item][lane]
constexpr size_t kFzStageBytes = 2 * (size_t)kFzStageDoubles * sizeof(double);  // double-buffered

// 8-byte asynchronous copy global -> shared (LDGSTS): the update phase keeps two stages of kFzBatch items x up to
// five vectors in flight per warp without holding a single register for them.  Through L1 (.ca is the only
// 8-byte form): safe for what U reads -- r, h, p, x, diag of a DoF are not written by anybody else in this launch
// before this thread reads them, and L1 starts every launch empty.
__device__ __forceinline__ void fz_cp_async8(double *smem_dst, const double *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void fz_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fz_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The loops below are deliberately ROLLED (#pragma unroll 1) and every phase has a single call site: the streaming
// warp's code must stay a few KB.  A first version with unrolled register batches compiled to 140 KB of SASS
// (the cell loop is 18 KB), thrashed the instruction cache and made both roles 2-3x slower.

// U(K): see the header comment.  acc[0] += r.r, acc[1] += r.Dr of the residual written here.  One warp;
// `stage` = this warp's 2 x kFzStageDoubles staging buffer in shared memory.
static __device__ __noinline__ void fz_update_phase(const FusedParams &fz, double *__restrict__ pvec, double *__restrict__ hvec,
                                             int K, uint64_t pol, double *stage, double (&acc)[2]) {
  const int ra = fz_rows_through<true>(fz, K - 1), rb = fz_rows_through<true>(fz, K);
  if (rb <= ra) return;
  const int lane = threadIdx.x & 31;
  const int umode = fz.umode;
  double alpha = 0.0, beta = 0.0, apa = 0.0, aob = 0.0;
  if (umode == FUSE_U_CG1 || umode == FUSE_U_CG3) { alpha = fz.st->alpha; beta = fz.st->beta; }
  if (umode == FUSE_U_CG3) { aob = fz.st->alpha_old / fz.st->beta_old; apa = alpha + aob; }
  double *__restrict__ rvec = fz.r, *__restrict__ xvec = fz.x;
  const double *__restrict__ diag = fz.diag;
  const bool ld_r = umode != FUSE_U_ZERO, ld_hp = umode == FUSE_U_CG1 || umode == FUSE_U_CG3, ld_x = umode == FUSE_U_CG3;
  const int xlo = fz.lo[0] ? fz.p : 0, xhi = fz.od0 - 1 - fz.hi[0];
  if (xhi < xlo) return;
  const int lz_a = ra / fz.nry, lz_b = (rb - 1) / fz.nry;
  constexpr int VS = kFzBatch * 32;          // doubles per vector per stage
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
    // software pipeline over stages of kFzBatch items: issue the copies of stage s+1, then consume stage s
    int s = 0;
    bool have = false;
#pragma unroll 1
    while (true) {
      const bool more = cur.valid();
      if (more) {
        double *buf = stage + (have ? (s ^ 1) : s) * kFzStageDoubles + lane;
        int *ibuf = reinterpret_cast<int *>(stage + (have ? (s ^ 1) : s) * kFzStageDoubles + kFzStageVecs * VS) + lane;
#pragma unroll 1
        for (int u = 0; u < kFzBatch; ++u) {
          const int i = (cur.valid() && cur.x + lane <= xhi) ? cur.idx + lane : -1;
          if (cur.valid()) cur.next();
          ibuf[u * 32] = i;
          if (i >= 0 && ld_r) {
            double *slot = buf + u * 32;
            fz_cp_async8(slot, rvec + i);
            if (ld_hp) { fz_cp_async8(slot + VS, hvec + i); fz_cp_async8(slot + 2 * VS, pvec + i); }
            if (ld_x) fz_cp_async8(slot + 3 * VS, xvec + i);
            if (diag) fz_cp_async8(slot + 4 * VS, diag + i);
          }
        }
        fz_cp_commit();
      }
      if (!have) {
        if (!more) break;
        have = true;
        continue;                           // first stage issued: go and issue the second before consuming
      }
      if (more) fz_cp_wait<1>(); else fz_cp_wait<0>();
      const double *buf = stage + s * kFzStageDoubles + lane;
      const int *ibuf = reinterpret_cast<const int *>(stage + s * kFzStageDoubles + kFzStageVecs * VS) + lane;
#pragma unroll 1
      for (int u = 0; u < kFzBatch; ++u) {
        const int i = ibuf[u * 32];
        if (i < 0) continue;
        if (ld_r) {
          const double *slot = buf + u * 32;
          const double rv = slot[0], dv = diag ? slot[4 * VS] : 1.0;
          double r_new = rv;
          if (!ld_hp) {
            __stcg(pvec + i, -dv * r_new);
          } else {
            const double hv = slot[VS], pv = slot[2 * VS];
            r_new = rv + alpha * hv;
            if (ld_x) fz_st_stream(xvec + i, slot[3 * VS] + (apa * pv + aob * dv * rv), pol);
            __stcg(rvec + i, r_new);
            __stcg(pvec + i, beta * pv - dv * r_new);
          }
          acc[0] += r_new * r_new;
          if (diag) acc[1] += r_new * dv * r_new;
        }
        __stcg(hvec + i, 0.0);
      }
      s ^= 1;
      if (!more) break;
    }
  }
}

// D(K): see the header comment.  acc: 0 correction of p.h, 1 h.h, 2 r.h, 3 r.Dh, 4 h.Dh.  One warp.
static __device__ __noinline__ void fz_finish_phase(const FusedParams &fz, const double *__restrict__ pvec,
                                             double *__restrict__ hvec, int K, uint64_t pol, double (&acc)[5]) {
  const int ra = fz_rows_through<false>(fz, K - 1), rb = fz_rows_through<false>(fz, K);
  if (rb <= ra) return;
  const int lane = threadIdx.x & 31;
  const bool cg = fz.dmode == FUSE_D_CG;
  const double *__restrict__ rvec = fz.r;
  const double *__restrict__ diag = fz.diag;
  const int xlo = fz.lo[0] ? fz.p : 0, xhi = fz.od0 - 1 - fz.hi[0];
  if (xhi < xlo) return;
  const int lz_a = ra / fz.nry, lz_b = (rb - 1) / fz.nry;
  // zero Dirichlet values on the whole global boundary (bp5/step-64.cu:354-357): faces without a neighbour block
  const int xd0 = fz.lo[0] ? -1 : 0, xd1 = fz.hi[0] ? -1 : fz.od0 - 1;
  const int jd0 = fz.lo[1] ? -1 : 0, jd1 = fz.hi[1] ? -1 : fz.od1 - 1;
  const int kd0 = fz.lo[2] ? -1 : 0, kd1 = fz.hi[2] ? -1 : fz.od2 - 1;
  constexpr int DB = 4;                      // items per (unrolled) batch: loads of a batch are issued together
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
      int idx[DB];
      bool dir[DB];
      double rv[DB], hv[DB], pv[DB], dv[DB];
#pragma unroll
      for (int u = 0; u < DB; ++u) {
        const int x = cur.x + lane;
        const bool ok = cur.valid() && x <= xhi;
        idx[u] = ok ? cur.idx + lane : -1;
        dir[u] = ok && (x == xd0 || x == xd1 || cur.j == jd0 || cur.j == jd1 || cur.k == kd0 || cur.k == kd1);
        if (cur.valid()) cur.next();
      }
#pragma unroll
      for (int u = 0; u < DB; ++u) {
        rv[u] = hv[u] = pv[u] = 0.0; dv[u] = 1.0;
        if (idx[u] >= 0) {
          if (cg) {
            rv[u] = fz_ld_stream(rvec + idx[u], pol);
            hv[u] = fz_ld_stream(hvec + idx[u], pol);
            if (diag) dv[u] = fz_ld_stream(diag + idx[u], pol);
          }
          if (dir[u]) pv[u] = __ldcg(pvec + idx[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < DB; ++u) {
        if (idx[u] < 0) continue;
        double vs = hv[u];
        if (dir[u]) {
          // copy_constrained_values: h_c = p_c; the cell kernel summed p_c (A p)_c into p.h
          acc[0] += pv[u] * (pv[u] - vs);
          vs = pv[u];
          __stcg(hvec + idx[u], vs);
        }
        if (cg) {
          acc[1] += vs * vs;
          acc[2] += rv[u] * vs;
          if (diag) { const double dvs = dv[u] * vs; acc[3] += rv[u] * dvs; acc[4] += vs * dvs; }
        }
      }
    }
  }
}

// The streaming warp's whole life (one warp per CTA): U runs `ua` macro steps ahead of the cells, D `dl` behind.
// Tick t: U(t), signal; then D(t - ua - dl) once every cell group has finished that step.
static __device__ __noinline__ void fz_stream_role(const FusedParams &fz, double *pvec, double *hvec, uint64_t pol,
                                            double *stage) {
  unsigned *u_ring = fz.sync, *c_ring = fz.sync + kFzRing, *err = fz.sync + kFzErr;
  const int lane = threadIdx.x & 31;
  double acc_u[2] = {0.0, 0.0}, acc_d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int n_steps = fz.n_steps, lag = fz.ua + fz.dl;
  if (fz.debug & 16) return;                       // tuning: idle streaming warp (with bit 8)
#pragma unroll 1
  for (int t = 0; t < n_steps + lag; ++t) {
    if (t < n_steps + fz.ua) {
      { FZ_DBG_T0(); if (!(fz.debug & 1)) fz_update_phase(fz, pvec, hvec, t, pol, stage, acc_u); __syncwarp(); if (lane == 0) FZ_DBG_ADD(fz, 3); }
      { FZ_DBG_T0(); if (lane == 0) { fz_signal(u_ring, t); FZ_DBG_ADD(fz, 5); } }
    }
    const int kd = t - lag;
    if (kd >= 0 && !(fz.debug & 32)) {
      { FZ_DBG_T0(); if (lane == 0) { fz_wait(c_ring, kd, err); FZ_DBG_ADD(fz, 2); } }
      __syncwarp();
      { FZ_DBG_T0(); if (!(fz.debug & 2)) fz_finish_phase(fz, pvec, hvec, kd, pol, acc_d); __syncwarp(); if (lane == 0) FZ_DBG_ADD(fz, 4); }
    }
  }
  // per-CTA partial sums of the two phases, lanes in butterfly order
  double *out = fz.partials + (size_t)blockIdx.x * kFusedPartials;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    double v = acc_u[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[kFzSlotU + j] = v;
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    double v = acc_d[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[kFzSlotD + j] = v;
  }
}

// End of the launch: the last CTA to get here adds the per-CTA sums in CTA order (deterministic for a given grid),
// runs the scalar recurrences (or leaves the local sums for the caller's all-rank sum) and re-arms the counters.
__device__ __forceinline__ void fz_finalize(const FusedParams &fz) {
  __shared__ bool last;
  __shared__ double tot[kFusedPartials];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(fz.sync + kFzTicket, 1u) == gridDim.x - 1);
    if (last) __threadfence();
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < kFusedPartials) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(fz.partials + b * kFusedPartials + threadIdx.x);
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
