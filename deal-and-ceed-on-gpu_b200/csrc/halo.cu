// Halo exchange support for the domain-partitioned layout: what
// LinearAlgebra::distributed::Vector::update_ghost_values / compress(add) /
// zero_out_ghosts do through Utilities::MPI::Partitioner and CUDA-aware MPI
// inside MatrixFree::cell_loop [UPSTREAM] (requested at bp5/step-64.cu:241;
// smoke-tested by tests/cuda_aware_mpi.cc:29-46).
//
// Layout (abi.cu, bp5_operator_create): a block owns its box of DoFs except its
// lower faces, which belong to the lower neighbour.  Ghosts are grouped by
// direction mask m = 1..7 (bit d set: the owner is the lower neighbour in
// dimension d); each group is one contiguous segment of the ghost region, so
//   update_ghost_values: the OWNER packs its upper face/edge/corner into a send
//                        buffer; the receiver lands it directly in the segment;
//   compress(add):       the ghost segment itself is the send buffer; the owner
//                        adds the received values onto its upper face.
// The transport (NCCL send/recv, or peer-mapped stores) is the caller's; this
// file only provides the device-side pack / add kernels and the message shapes.
#include "common.h"

namespace bp5 {

struct HaloGeom {
  int od[3], ld[3], hlo[3];
  long long count[8];    // entries sent to the upper neighbour in direction m
  long long offset[8];   // start of group m in the packed send buffer
};

// owned index of entry t of send group m: top layer in the bit dimensions, the
// whole owned range (x fastest) in the others
__device__ __forceinline__ long long send_index(const HaloGeom &g, int m, long long t) {
  int q[3];
  long long rem = t;
  for (int d = 0; d < 3; ++d) {
    if (m & (1 << d)) q[d] = g.od[d] - 1;
    else { q[d] = (int)(rem % g.od[d]); rem /= g.od[d]; }
  }
  return q[0] + (long long)g.od[0] * (q[1] + (long long)g.od[1] * q[2]);
}

__global__ void halo_pack_kernel(HaloGeom g, const double *__restrict__ vec, double *__restrict__ sendbuf,
                                 long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int m = 1;
    while (m < 7 && i >= g.offset[m] + g.count[m]) ++m;
    sendbuf[i] = vec[send_index(g, m, i - g.offset[m])];
  }
}

__global__ void halo_unpack_add_kernel(HaloGeom g, double *__restrict__ vec, const double *__restrict__ recvbuf,
                                       long long total) {
  // different groups can hit the same owned DoF (a corner belongs to the face,
  // edge and corner groups of different neighbours): accumulate atomically
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int m = 1;
    while (m < 7 && i >= g.offset[m] + g.count[m]) ++m;
    atomicAdd(&vec[send_index(g, m, i - g.offset[m])], recvbuf[i]);
  }
}

static HaloGeom make_halo(bp5_operator_t op) {
  HaloGeom g{};
  for (int d = 0; d < 3; ++d) { g.od[d] = op->od[d]; g.ld[d] = op->ld[d]; g.hlo[d] = op->has_lo[d]; }
  long long off = 0;
  for (int m = 1; m < 8; ++m) {
    bool exists = true;
    long long cnt = 1;
    for (int d = 0; d < 3; ++d) {
      if (m & (1 << d)) { if (!op->has_hi[d]) exists = false; }
      else cnt *= op->od[d];
    }
    g.count[m] = exists ? cnt : 0;
    g.offset[m] = off;
    off += g.count[m];
  }
  return g;
}

int halo_info(bp5_operator_t op, int64_t *send_count, int64_t *send_offset, int64_t *recv_count,
              int64_t *recv_offset) {
  const HaloGeom g = make_halo(op);
  for (int m = 0; m < 8; ++m) {
    send_count[m] = m ? g.count[m] : 0;
    send_offset[m] = m ? g.offset[m] : 0;
    recv_count[m] = m ? op->ghost_size[m] : 0;
    recv_offset[m] = m ? op->ghost_offset[m] : 0;
  }
  return BP5_OK;
}

static long long halo_total(const HaloGeom &g) { return g.offset[7] + g.count[7]; }

int halo_pack(bp5_operator_t op, const double *vec, double *sendbuf) {
  const HaloGeom g = make_halo(op);
  const long long total = halo_total(g);
  if (total == 0) return BP5_OK;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  halo_pack_kernel<<<(unsigned)grid, 256, 0, op->ctx->stream>>>(g, vec, sendbuf, total);
  BP5_CHECK_LAUNCH();
  op->ctx->launches++;
  return BP5_OK;
}

int halo_unpack_add(bp5_operator_t op, double *vec, const double *recvbuf) {
  const HaloGeom g = make_halo(op);
  const long long total = halo_total(g);
  if (total == 0) return BP5_OK;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  halo_unpack_add_kernel<<<(unsigned)grid, 256, 0, op->ctx->stream>>>(g, vec, recvbuf, total);
  BP5_CHECK_LAUNCH();
  op->ctx->launches++;
  return BP5_OK;
}

}  // namespace bp5
