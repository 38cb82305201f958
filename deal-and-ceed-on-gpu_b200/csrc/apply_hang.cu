// Instantiations of the cell kernel that resolve hanging-node constraints (HANG = 1, apply.cuh) for locally refined
// meshes.  A separate translation unit only so that it compiles in parallel with apply.cu.
#include "apply_launch.cuh"

namespace bp5 {

template <int P>
static int launch_hanging_p(bp5_operator_t op, double *dst, const double *src, int mode, double *dp, int which) {
  const bool gll = op->prob.quadrature == BP5_QUAD_GLL;
  const bool helm = op->prob.operator_kind == BP5_OP_HELMHOLTZ;
  if (mode == 2) return BP5_LAUNCH_QH(P, 2, 0, 1);
  if (mode == 1) return BP5_LAUNCH_QH(P, 1, 0, 1);
  if (mode == 0) return BP5_LAUNCH_QH(P, 0, 0, 1);
  set_error("the coloured cell order is not available on locally refined meshes");
  return BP5_ERR_UNSUPPORTED;
}

int launch_hanging(bp5_operator_t op, double *dst, const double *src, int mode, double *dp, int which) {
  switch (op->p) {
    case 1: return launch_hanging_p<1>(op, dst, src, mode, dp, which);
    case 2: return launch_hanging_p<2>(op, dst, src, mode, dp, which);
    case 3: return launch_hanging_p<3>(op, dst, src, mode, dp, which);
    case 4: return launch_hanging_p<4>(op, dst, src, mode, dp, which);
    case 5: return launch_hanging_p<5>(op, dst, src, mode, dp, which);
    case 6: return launch_hanging_p<6>(op, dst, src, mode, dp, which);
    case 7: return launch_hanging_p<7>(op, dst, src, mode, dp, which);
    case 8: return launch_hanging_p<8>(op, dst, src, mode, dp, which);
  }
  set_error("unsupported degree %d", op->p);
  return BP5_ERR_UNSUPPORTED;
}

}  // namespace bp5
