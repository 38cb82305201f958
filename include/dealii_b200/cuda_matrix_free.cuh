// cuda_matrix_free.cuh -- the DEVICE-side half of the drop-in boundary (compile with nvcc).
//
// The reference's operators are written as device functors against
//   CUDAWrappers::MatrixFree<dim,double>  (reinit / cell_loop / evaluate_coefficients /
//                                          copy_constrained_values / initialize_dof_vector,
//                                          bp5/step-64.cu:235-248,258,274-275,214)  [UPSTREAM]
//   CUDAWrappers::MatrixFree<dim,double>::Data, CUDAWrappers::SharedData<dim,double>
//                                          (bp5/step-64.cu:69-75,128-138)          [UPSTREAM]
//   CUDAWrappers::FEEvaluationGL / FEEvaluation   (bp5/fe_evaluation_gl.h:31-98)
//   q_point_id_in_cell, local_q_point_id, get_quadrature_point, internal::compute_index
//                                          (bp5/step-64.cu:90,163; step-64/step-64.cu:105-111)
// This header re-creates exactly that surface on top of the C ABI
// (bp5_operator_matrix_free_data, include/bp5_b200.h), so LocalPoissonOperator,
// JacobianFunctor, LocalHelmholtzOperator, VaryingCoefficientFunctor and the operator
// classes wrapping them compile unchanged (examples/bp5_functors.cu).
//
// It is the GENERAL path: one thread per DoF / quadrature point, one cell per CTA, indices
// through local_to_global, geometry through inv_jacobian / JxW, like the reference.  The
// tuned hot kernel behind BP5::PoissonOperator (csrc/apply.cuh) implements the same
// mathematics for the two operators the benchmark uses; tests/test_gpu_functor_api.py
// checks that both paths and the CPU oracle agree to 1e-12.
//
// Evaluation scheme (evaluate / integrate): values are interpolated to the quadrature
// points (skipped for Gauss-Lobatto collocation, where shape_values is the identity), then
// differentiated there with the collocation derivative matrix -- 3 or 6 one-dimensional
// contractions each way instead of the 9 of the general evaluator the reference calls.
#pragma once
#include <cuda_runtime.h>

#include "dealii_b200.h"

namespace dealii {

namespace Utilities {
template <typename T> DEAL_II_B200_HOST_DEVICE constexpr T pow(const T base, const int iexp) {
  return iexp <= 0 ? T(1) : base * pow(base, iexp - 1);
}
}  // namespace Utilities

// ------------------------------------------------------------------ Tensor<rank,dim>
template <int rank, int dim, typename Number = double> class Tensor;
template <int dim, typename Number> class Tensor<1, dim, Number> {
 public:
  DEAL_II_B200_HOST_DEVICE Tensor() { for (int d = 0; d < dim; ++d) v[d] = Number(0); }
  DEAL_II_B200_HOST_DEVICE Number &operator[](unsigned d) { return v[d]; }
  DEAL_II_B200_HOST_DEVICE const Number &operator[](unsigned d) const { return v[d]; }

 private:
  Number v[dim];
};
template <int dim, typename Number> class Tensor<2, dim, Number> {
 public:
  DEAL_II_B200_HOST_DEVICE Tensor<1, dim, Number> &operator[](unsigned d) { return v[d]; }
  DEAL_II_B200_HOST_DEVICE const Tensor<1, dim, Number> &operator[](unsigned d) const { return v[d]; }

 private:
  Tensor<1, dim, Number> v[dim];
};

namespace b200 {
inline void check_cuda(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw ExcMessage(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace b200

// ------------------------------------------------------------------ plain device array
// LinearAlgebra::CUDAWrappers::Vector<double>: holds the coefficient (bp5/step-64.cu:219,253)
namespace LinearAlgebra { namespace CUDAWrappers {
template <typename Number> class Vector {
 public:
  Vector() = default;
  Vector(const Vector &) = delete;
  Vector &operator=(const Vector &) = delete;
  ~Vector() { cudaFree(data); }
  void reinit(std::size_t n) {
    cudaFree(data); data = nullptr; n_elements = n;
    b200::check_cuda(cudaMalloc(&data, sizeof(Number) * (n ? n : 1)), "cudaMalloc");
    // zero on the library's (non-blocking) stream: everything that touches this array later runs there, and a
    // memset on the legacy default stream would not be ordered with it
    cudaStream_t stream = static_cast<cudaStream_t>(bp5_context_stream(b200::Context::get()));
    b200::check_cuda(cudaMemsetAsync(data, 0, sizeof(Number) * (n ? n : 1), stream), "cudaMemsetAsync");
    b200::check_cuda(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  }
  Number *get_values() const { return data; }
  std::size_t size() const { return n_elements; }

 private:
  Number *data = nullptr;
  std::size_t n_elements = 0;
};
} }  // namespace LinearAlgebra::CUDAWrappers

namespace CUDAWrappers {
constexpr int warp_size = 32;      // base/cuda_size.h [UPSTREAM], used at bp5/solver.h:61,65
constexpr int block_size = 512;
constexpr int chunk_size = 1;

// ------------------------------------------------------------------ SharedData
template <int dim, typename Number> struct SharedData {
  __device__ SharedData(Number *vd, Number *gq[dim]) : values(vd) {
    for (int d = 0; d < dim; ++d) gradients[d] = gq[d];
  }
  Number *values;            // n^dim: dof values, then values at the quadrature points
  Number *gradients[dim];    // n^dim each: reference-cell gradient at the quadrature points
  // the CTA's shared-memory copies of the 1D tables [q*n + i], filled by MatrixFree's kernels for the
  // one-thread-per-output evaluator (DEALII_B200_POINTWISE_EVALUATE); null: the tables are read from the kernel
  // parameters.  deal.II keeps them in __constant__ memory, where a warp whose lanes want different rows is served one
  // row at a time; the default evaluator reads them uniformly (line-owner contractions) and needs no copy.
  const Number *shape_values = nullptr;
  const Number *co_shape_gradients = nullptr;
};

// ------------------------------------------------------------------ MatrixFree
template <int dim, typename Number = double> class MatrixFree {
  static_assert(dim == 3 && sizeof(Number) == sizeof(double), "the hot path is 3D fp64");
 public:
  enum ParallelizationScheme { parallel_in_elem, parallel_over_elem };

  struct AdditionalData {
    AdditionalData(const ParallelizationScheme s = parallel_in_elem,
                   const UpdateFlags f = update_gradients | update_JxW_values | update_quadrature_points,
                   const bool coloring = false, const bool overlap = false)
        : parallelization_scheme(s), mapping_update_flags(f), use_coloring(coloring),
          overlap_communication_computation(overlap) {}
    ParallelizationScheme parallelization_scheme;
    UpdateFlags mapping_update_flags;
    bool use_coloring;
    bool overlap_communication_computation;     // bp5/step-64.cu:241
  };

  // what a device functor sees (fe_evaluation_gl.h:112-120, bp5/step-64.cu:94-109)
  struct Data {
    Point<dim, Number> *q_points;
    types::global_dof_index *local_to_global;
    Number *inv_jacobian;
    Number *JxW;
    unsigned int n_cells;
    unsigned int padding_length;
    unsigned int row_start;
    unsigned int *constraint_mask;
    bool use_coloring;
    // 1D tables [q*n + i]; deal.II keeps them in __constant__ memory, here they ride in the
    // kernel parameter bank with the rest of the struct
    bool collocation;
    Number shape_values[81];
    Number co_shape_gradients[81];
    // locally refined meshes: 1D parent-to-child interpolation [s][a*n + b] (bp5_matrix_free_data_t)
    Number hanging_interpolation[2][81];
  };

  MatrixFree() = default;
  MatrixFree(const MatrixFree &) = delete;
  ~MatrixFree() { bp5_operator_destroy(op); }

  template <typename IteratorFiltersType = int>
  void reinit(const MappingQGeneric<dim> &mapping, const DoFHandler<dim> &dof_handler,
              const AffineConstraints<Number> &, const Quadrature<1> &quad,
              const AdditionalData &additional_data = AdditionalData()) {
    if (mapping.degree != dof_handler.degree) throw ExcMessage("MappingQGeneric degree must equal fe_degree");
    if (quad.n_points_1d != dof_handler.degree + 1) throw ExcMessage("n_q_points_1d must be fe_degree + 1");
    // the ghost exchange of a partitioned mesh happens behind the library operator (BP5::PoissonOperator on a
    // communicator); this generic path drives ONE block, where overlap_communication_computation (bp5/step-64.cu:241)
    // has nothing to overlap.  On a triangulation that is split over several ranks it must not silently build the
    // whole mesh on every rank: fail loudly.
    const Triangulation<dim> &t = dof_handler.get_triangulation();
    if (t.communicator != nullptr && t.communicator->size() > 1)
      throw ExcMessage("CUDAWrappers::MatrixFree (user-written functors) drives one block; on a partitioned mesh use "
                       "BP5::PoissonOperator, whose cell loop overlaps the ghost exchange");
    use_coloring = additional_data.use_coloring;
    bp5_operator_destroy(op); op = nullptr;
    bp5_problem_t pr{};
    pr.degree = (int32_t)dof_handler.degree; pr.quadrature = quad.abi_kind; pr.operator_kind = BP5_OP_POISSON;
    pr.geometry_mode = BP5_GEOM_STORED;
    for (int d = 0; d < 3; ++d) {
      pr.cells[d] = (int32_t)t.cells(d); pr.lower[d] = t.p1[d]; pr.upper[d] = t.p2[d];
      pr.part_grid[d] = 1; pr.part_coord[d] = 0;
    }
    pr.deformation = t.deformation; pr.deformation_eps = t.deformation_eps;
    for (int d = 0; d < 3; ++d) { pr.refine_lo[d] = t.refine_lo[d]; pr.refine_hi[d] = t.refine_hi[d]; }
    if (t.locally_refined() && additional_data.use_coloring)
      throw ExcMessage("use_coloring is implemented for conforming meshes");
    b200::check(bp5_operator_create(b200::Context::get(), &pr, &op));
    bp5_matrix_free_data_t md;
    b200::check(bp5_operator_matrix_free_data(op, &md));
    data.q_points = reinterpret_cast<Point<dim, Number> *>(md.q_points);
    data.local_to_global = md.local_to_global;
    data.inv_jacobian = md.inv_jacobian;
    data.JxW = md.JxW;
    data.n_cells = md.n_cells;
    data.padding_length = md.padding_length;
    data.row_start = 0;
    data.constraint_mask = md.constraint_mask;
    data.use_coloring = false;
    data.collocation = md.collocation != 0;
    n_q_points_1d = md.n_q_points_1d;
    for (int i = 0; i < 81; ++i) {
      data.shape_values[i] = md.shape_values[i];
      data.co_shape_gradients[i] = md.co_shape_gradients[i];
      data.hanging_interpolation[0][i] = md.hanging_interpolation[0][i];
      data.hanging_interpolation[1][i] = md.hanging_interpolation[1][i];
    }
    n_colors = 1;
    if (use_coloring) {
      // eight parity colours; each one has its own arrays like deal.II's per-colour Data, row_start counts
      // the cells of the colours before it (local_q_point_id adds it, step-64/step-64.cu:208)
      n_colors = 8;
      unsigned int row_start = 0;
      for (int c = 0; c < 8; ++c) {
        b200::check(bp5_operator_matrix_free_data_colored(op, c, &md));
        color_data[c] = data;
        color_data[c].q_points = reinterpret_cast<Point<dim, Number> *>(md.q_points);
        color_data[c].local_to_global = md.local_to_global;
        color_data[c].inv_jacobian = md.inv_jacobian;
        color_data[c].JxW = md.JxW;
        color_data[c].n_cells = md.n_cells;
        color_data[c].constraint_mask = md.constraint_mask;
        color_data[c].row_start = row_start;
        color_data[c].use_coloring = true;
        row_start += md.n_cells * md.padding_length;
      }
    }
  }

  unsigned int n_colors_used() const { return (unsigned int)n_colors; }
  unsigned long long n_dofs() const {
    int64_t n = 0;
    b200::check(bp5_operator_sizes(op, &n, nullptr, nullptr, nullptr));
    return (unsigned long long)n;
  }
  Data get_data(unsigned int color = 0) const { return use_coloring ? color_data[color] : data; }
  bp5_operator_t handle() const { return op; }
  cudaStream_t stream() const { return static_cast<cudaStream_t>(bp5_context_stream(b200::Context::get())); }

  template <typename Functor, typename VectorType>
  void cell_loop(const Functor &func, const VectorType &src, VectorType &dst) const;
  template <typename Functor> void evaluate_coefficients(Functor func) const;

  template <typename VectorType> void copy_constrained_values(const VectorType &src, VectorType &dst) const {
    b200::check(bp5_operator_copy_constrained_values(op, dst.handle(), src.handle()));
    dst.mark_modified();
  }
  template <typename VectorType> void initialize_dof_vector(VectorType &vec) const {
    bp5_vector_t v = nullptr;
    b200::check(bp5_operator_initialize_dof_vector(op, &v));
    vec.adopt(v);
  }

 private:
  bp5_operator_t op = nullptr;
  Data data{};
  Data color_data[8]{};
  bool use_coloring = false;
  int n_colors = 1;
  int n_q_points_1d = 0;
};

// ------------------------------------------------------------------ index helpers
template <int dim> __device__ inline unsigned int q_point_id_in_cell(const unsigned int n_q_points_1d) {
  return dim == 1 ? threadIdx.x % n_q_points_1d
       : dim == 2 ? threadIdx.x % n_q_points_1d + n_q_points_1d * threadIdx.y
                  : threadIdx.x % n_q_points_1d + n_q_points_1d * (threadIdx.y + n_q_points_1d * threadIdx.z);
}
template <int dim, typename Number>
__device__ inline unsigned int local_q_point_id(const unsigned int cell, const typename MatrixFree<dim, Number>::Data *data,
                                                const unsigned int n_q_points_1d, const unsigned int n_q_points) {
  return (data->row_start / data->padding_length + cell) * n_q_points + q_point_id_in_cell<dim>(n_q_points_1d);
}
template <int dim, typename Number>
__device__ inline Point<dim, Number> &get_quadrature_point(const unsigned int cell,
                                                           const typename MatrixFree<dim, Number>::Data *data,
                                                           const unsigned int n_q_points_1d) {
  return *(data->q_points + data->padding_length * cell + q_point_id_in_cell<dim>(n_q_points_1d));
}
namespace internal {
template <int dim, int n_points_1d> __device__ inline unsigned int compute_index() {
  return q_point_id_in_cell<dim>(n_points_1d);
}

// the two kernels of MatrixFree [UPSTREAM apply_kernel_shmem / evaluate_coeff]: one CTA per
// cell, one thread per DoF = quadrature point, block (n, n, n)
template <int dim, typename Number, typename Functor>
__global__ void __launch_bounds__(Functor::n_q_points)
    apply_kernel_shmem(Functor func, const __grid_constant__ typename MatrixFree<dim, Number>::Data gpu_data,
                       const Number *src, Number *dst) {
  __shared__ Number values[Functor::n_local_dofs];
  __shared__ Number gradients[dim][Functor::n_q_points];
  Number *gq[dim];
  for (int d = 0; d < dim; ++d) gq[d] = gradients[d];
  SharedData<dim, Number> shared_data(values, gq);
#ifdef DEALII_B200_POINTWISE_EVALUATE      // the one-thread-per-output evaluator reads table rows per lane: stage them
  __shared__ Number tables[2][Functor::n_dofs_1d * Functor::n_dofs_1d];
  {
    const unsigned int t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if (t < Functor::n_dofs_1d * Functor::n_dofs_1d) {
      tables[0][t] = gpu_data.shape_values[t];
      tables[1][t] = gpu_data.co_shape_gradients[t];
    }
    shared_data.shape_values = tables[0];
    shared_data.co_shape_gradients = tables[1];
    __syncthreads();
  }
#endif
  const unsigned int cell = blockIdx.x;      // whole CTAs only: the functor synchronises
  func(cell, &gpu_data, &shared_data, src, dst);
}
template <int dim, typename Number, typename Functor>
__global__ void __launch_bounds__(Functor::n_q_points)
    evaluate_coeff(Functor func, const __grid_constant__ typename MatrixFree<dim, Number>::Data gpu_data) {
  func(blockIdx.x, &gpu_data);
}
}  // namespace internal

template <int dim, typename Number>
template <typename Functor, typename VectorType>
void MatrixFree<dim, Number>::cell_loop(const Functor &func, const VectorType &src, VectorType &dst) const {
  if ((int)Functor::n_dofs_1d != n_q_points_1d) throw ExcMessage("functor degree does not match MatrixFree::reinit");
  const dim3 block(Functor::n_dofs_1d, Functor::n_dofs_1d, Functor::n_dofs_1d);
  // no ghost exchange around the kernel: this path handles one block (one GPU)
  if (use_coloring) {
    // one launch per colour, in order: within a colour no two cells touch the same DoF (plain += in
    // distribute_local_to_global), the colours add in a fixed order -> bitwise reproducible results
    for (int c = 0; c < 8; ++c) {
      if (color_data[c].n_cells == 0) continue;
      internal::apply_kernel_shmem<dim, Number, Functor>
          <<<color_data[c].n_cells, block, 0, stream()>>>(func, color_data[c], src.get_values(), dst.get_values());
      b200::check_cuda(cudaGetLastError(), "apply_kernel_shmem launch");
    }
  } else {
    internal::apply_kernel_shmem<dim, Number, Functor>
        <<<data.n_cells, block, 0, stream()>>>(func, data, src.get_values(), dst.get_values());
    b200::check_cuda(cudaGetLastError(), "apply_kernel_shmem launch");
  }
  dst.mark_modified();
}

template <int dim, typename Number>
template <typename Functor>
void MatrixFree<dim, Number>::evaluate_coefficients(Functor func) const {
  if ((int)Functor::n_dofs_1d != n_q_points_1d) throw ExcMessage("functor degree does not match MatrixFree::reinit");
  const dim3 block(Functor::n_dofs_1d, Functor::n_dofs_1d, Functor::n_dofs_1d);
  if (use_coloring) {
    for (int c = 0; c < 8; ++c) {
      if (color_data[c].n_cells == 0) continue;
      internal::evaluate_coeff<dim, Number, Functor><<<color_data[c].n_cells, block, 0, stream()>>>(func, color_data[c]);
      b200::check_cuda(cudaGetLastError(), "evaluate_coeff launch");
    }
  } else {
    internal::evaluate_coeff<dim, Number, Functor><<<data.n_cells, block, 0, stream()>>>(func, data);
    b200::check_cuda(cudaGetLastError(), "evaluate_coeff launch");
  }
  b200::check_cuda(cudaStreamSynchronize(stream()), "evaluate_coefficients");
}

// ------------------------------------------------------------------ FEEvaluationGL
// bp5/fe_evaluation_gl.h:31-98, member for member.
template <int dim, int fe_degree, int n_q_points_1d = fe_degree + 1, int n_components_ = 1, typename Number = double>
class FEEvaluationGL {
  static_assert(dim == 3 && n_components_ == 1 && n_q_points_1d == fe_degree + 1,
                "scalar 3D elements with fe_degree+1 quadrature points per direction");
 public:
  using value_type = Number;
  using gradient_type = Tensor<1, dim, Number>;
  using data_type = typename MatrixFree<dim, Number>::Data;
  static constexpr unsigned int dimension = dim;
  static constexpr unsigned int n_components = n_components_;
  static constexpr unsigned int n_q_points = Utilities::pow(n_q_points_1d, dim);
  static constexpr unsigned int tensor_dofs_per_cell = Utilities::pow(fe_degree + 1, dim);

  __device__ FEEvaluationGL(const unsigned int cell_id, const data_type *data, SharedData<dim, Number> *shdata)
      : n_cells(data->n_cells), padding_length(data->padding_length), constraint_mask(data->constraint_mask[cell_id]),
        use_coloring(data->use_coloring), values(shdata->values), mf(data),
        tab_B(shdata->shape_values ? shdata->shape_values : data->shape_values),
        tab_D(shdata->co_shape_gradients ? shdata->co_shape_gradients : data->co_shape_gradients) {
    local_to_global = data->local_to_global + padding_length * cell_id;
    inv_jac = data->inv_jacobian + padding_length * cell_id;
    JxW = data->JxW + padding_length * cell_id;
    for (unsigned int i = 0; i < dim; ++i) gradients[i] = shdata->gradients[i];
    ix = threadIdx.x % n_q_points_1d; iy = threadIdx.y; iz = threadIdx.z;
    idx = ix + n_q_points_1d * (iy + n_q_points_1d * iz);
  }

  // values[idx] = src[local_to_global[idx]], then the hanging-node constraints of this cell
  // (fe_evaluation_gl.h:133-152: read, __syncthreads, resolve_hanging_nodes<false>(constraint_mask, values))
  __device__ void read_dof_values(const Number *src) {
    values[idx] = __ldg(&src[local_to_global[idx]]);
    __syncthreads();
    if (constraint_mask != 0) resolve_hanging_nodes<false>();
  }

  // transposed constraints, then dst[local_to_global[idx]] += values[idx] (fe_evaluation_gl.h:161-181)
  __device__ void distribute_local_to_global(Number *dst) {
    if (constraint_mask != 0) resolve_hanging_nodes<true>();
    const types::global_dof_index j = local_to_global[idx];
    if (use_coloring) dst[j] += values[idx];
    else atomicAdd(&dst[j], values[idx]);     // red.global.add.f64
  }

  // fe_evaluation_gl.h:190-214.  Afterwards gradients[d] hold the reference-cell gradient at the quadrature points
  // (if evaluate_grad) and values the function values there (with Gauss quadrature always, with collocation they are
  // the DoF values anyway).
  // Every 1D contraction is done by the n^2 threads whose index along the direction is zero: each reads its line
  // once, multiplies by the n x n table in registers (uniform reads of the kernel parameters) and writes the line --
  // 2 shared accesses per point and direction instead of the n + 1 of the one-thread-per-output form of deal.II's
  // evaluator (define DEALII_B200_POINTWISE_EVALUATE to get that form back; same results up to summation order).
#ifndef DEALII_B200_POINTWISE_EVALUATE
  __device__ void evaluate(const bool evaluate_val, const bool evaluate_grad) {
    const Number *B = mf->shape_values, *D = mf->co_shape_gradients;
    (void)evaluate_val;
    if (!mf->collocation) {
      line_pass<0, false>(B, values, gradients[0]); __syncthreads();
      line_pass<1, false>(B, gradients[0], gradients[1]); __syncthreads();
      line_pass<2, false>(B, gradients[1], values); __syncthreads();
    }
    if (evaluate_grad) {
      line_pass<0, false>(D, values, gradients[0]);
      line_pass<1, false>(D, values, gradients[1]);
      line_pass<2, false>(D, values, gradients[2]);
      __syncthreads();
    }
  }

  // fe_evaluation_gl.h:223-250: values[dof] = sum over quadrature points of the submitted values /
  // gradients tested with the basis (the transpose of evaluate)
  __device__ void integrate(const bool integrate_val, const bool integrate_grad) {
    const Number *B = mf->shape_values, *D = mf->co_shape_gradients;
    if (integrate_grad) {
      line_pass<0, true>(D, gradients[0], gradients[0]);       // in place: a line has one owner
      line_pass<1, true>(D, gradients[1], gradients[1]);
      line_pass<2, true>(D, gradients[2], gradients[2]);
      __syncthreads();
      Number w = integrate_val ? values[idx] : Number(0);
      w += gradients[0][idx] + gradients[1][idx] + gradients[2][idx];
      values[idx] = w;
      __syncthreads();
    }
    if (mf->collocation) return;
    line_pass<0, true>(B, values, gradients[0]); __syncthreads();
    line_pass<1, true>(B, gradients[0], gradients[1]); __syncthreads();
    line_pass<2, true>(B, gradients[1], values); __syncthreads();
  }
#else
  __device__ void evaluate(const bool evaluate_val, const bool evaluate_grad) {
    const Number *B = tab_B, *D = tab_D;
    constexpr int n = n_q_points_1d, n2 = n * n;
    Number v = values[idx];
    if (!mf->collocation) {
      // interpolate to the quadrature points: x, y, z passes through the gradient arrays
      v = line<0>(B, values, ix); gradients[0][idx] = v; __syncthreads();
      v = line<1>(B, gradients[0], iy); gradients[1][idx] = v; __syncthreads();
      v = line<2>(B, gradients[1], iz); gradients[2][idx] = v; __syncthreads();
    }
    const Number *at_q = mf->collocation ? values : gradients[2];
    Number g[dim];
    if (evaluate_grad) {
      g[0] = line<0>(D, at_q, ix);
      g[1] = line<1>(D, at_q, iy);
      g[2] = line<2>(D, at_q, iz);
    }
    if (!mf->collocation) __syncthreads();      // gradients[2] is about to be overwritten
    if (evaluate_grad)
      for (int d = 0; d < dim; ++d) gradients[d][idx] = g[d];
    if (evaluate_val) values[idx] = v;
    __syncthreads();
    (void)n2;
  }

  // fe_evaluation_gl.h:223-250: values[dof] = sum over quadrature points of the submitted values /
  // gradients tested with the basis (the transpose of evaluate)
  __device__ void integrate(const bool integrate_val, const bool integrate_grad) {
    const Number *B = tab_B, *D = tab_D;
    Number w = integrate_val ? values[idx] : Number(0);
    if (integrate_grad) {
      w += line_t<0>(D, gradients[0], ix);
      w += line_t<1>(D, gradients[1], iy);
      w += line_t<2>(D, gradients[2], iz);
    }
    __syncthreads();
    if (mf->collocation) {
      values[idx] = w;
      __syncthreads();
      return;
    }
    gradients[0][idx] = w; __syncthreads();
    w = line_t<0>(B, gradients[0], ix); gradients[1][idx] = w; __syncthreads();
    w = line_t<1>(B, gradients[1], iy); gradients[2][idx] = w; __syncthreads();
    w = line_t<2>(B, gradients[2], iz); values[idx] = w; __syncthreads();
  }

#endif

  __device__ value_type get_value(const unsigned int q_point) const { return values[q_point]; }
  __device__ value_type get_dof_value(const unsigned int dof) const { return values[dof]; }
  __device__ void submit_value(const value_type &val_in, const unsigned int q_point) {
    values[q_point] = val_in * JxW[q_point];
  }
  __device__ void submit_dof_value(const value_type &val_in, const unsigned int dof) { values[dof] = val_in; }

  // J^-T grad_ref (fe_evaluation_gl.h:328-346)
  __device__ gradient_type get_gradient(const unsigned int q_point) const {
    const Number *inv_jacobian = &inv_jac[q_point];
    const std::size_t plane = (std::size_t)padding_length * n_cells;
    gradient_type grad;
    for (int d_1 = 0; d_1 < dim; ++d_1) {
      Number tmp = 0.;
      for (int d_2 = 0; d_2 < dim; ++d_2) tmp += inv_jacobian[plane * (dim * d_2 + d_1)] * gradients[d_2][q_point];
      grad[d_1] = tmp;
    }
    return grad;
  }
  // J^-1 grad * JxW (fe_evaluation_gl.h:355-369)
  __device__ void submit_gradient(const gradient_type &grad_in, const unsigned int q_point) {
    const Number *inv_jacobian = &inv_jac[q_point];
    const std::size_t plane = (std::size_t)padding_length * n_cells;
    for (int d_1 = 0; d_1 < dim; ++d_1) {
      Number tmp = 0.;
      for (int d_2 = 0; d_2 < dim; ++d_2) tmp += inv_jacobian[plane * (dim * d_1 + d_2)] * grad_in[d_2];
      gradients[d_1][q_point] = tmp * JxW[q_point];
    }
  }
  // the one-argument forms the non-merged BP5 branch uses (bp5/step-64.cu:190)
  __device__ gradient_type get_gradient() const { return get_gradient(idx); }
  __device__ void submit_gradient(const gradient_type &grad_in) { submit_gradient(grad_in, idx); }

  // func(this, q) for this thread's quadrature point, then a barrier (fe_evaluation_gl.h:379-393)
  template <typename Functor> __device__ void apply_quad_point_operations(const Functor &func) {
    func(this, idx);
    __syncthreads();
  }

 private:
  // Hanging-node constraints of a child cell of a locally refined mesh (the slot of deal.II's
  // internal::resolve_hanging_nodes [UPSTREAM], called at fe_evaluation_gl.h:150,167).  The nodes on a constrained
  // face were loaded from the unrefined neighbour's face DoFs (same local index in the parent); the child's values
  // are the parent's face polynomial at the child's nodes: a 1D interpolation along each of the face's two tangential
  // directions.  Direction by direction, every line that lies in a constrained face is interpolated once -- a line on
  // the edge between two constrained faces too.  transpose: the adjoint, for the scatter.
  // constraint_mask: bit d = face normal to d constrained, bit 3+d = child position s_d (face at node 0 or p; matrix).
  template <bool transpose> __device__ void resolve_hanging_nodes() {
    const unsigned int pos[3] = {ix, iy, iz};
    bool on_face[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      on_face[d] = ((constraint_mask >> d) & 1u) != 0 && pos[d] == (((constraint_mask >> (3 + d)) & 1u) ? fe_degree : 0);
    resolve_direction<0, transpose>(on_face[1] || on_face[2]);
    resolve_direction<1, transpose>(on_face[0] || on_face[2]);
    resolve_direction<2, transpose>(on_face[0] || on_face[1]);
  }
  template <int DIR, bool transpose> __device__ void resolve_direction(const bool in_constrained_face) {
    // is any line along DIR constrained at all?  (uniform over the cell: no divergent barrier)
    const unsigned int others = (constraint_mask & 7u) & ~(1u << DIR);
    if (others == 0) return;
    const Number *M = mf->hanging_interpolation[(constraint_mask >> (3 + DIR)) & 1u];
    const unsigned int row = DIR == 0 ? ix : DIR == 1 ? iy : iz;
    Number v = values[idx];
    if (in_constrained_face) v = transpose ? line_t<DIR>(M, values, row) : line<DIR>(M, values, row);
    __syncthreads();
    values[idx] = v;
    __syncthreads();
  }

  // dst(line) = M src(line) (or M^T) for the line along DIR through this thread's point, done by the thread whose own
  // index along DIR is zero; src == dst is allowed (the line is read completely before it is written)
  template <int DIR, bool transpose> __device__ void line_pass(const Number *M, const Number *src, Number *dst) const {
    constexpr int n = n_q_points_1d;
    constexpr int stride = DIR == 0 ? 1 : DIR == 1 ? n : n * n;
    if ((DIR == 0 ? ix : DIR == 1 ? iy : iz) != 0) return;
    Number v[n];
#pragma unroll
    for (int m = 0; m < n; ++m) v[m] = src[idx + m * stride];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      Number sum = 0.;
#pragma unroll
      for (int m = 0; m < n; ++m) sum += (transpose ? M[m * n + i] : M[i * n + m]) * v[m];
      dst[idx + i * stride] = sum;
    }
  }

  // sum_m M[row][m] * a(..m..) along direction DIR through this thread's point
  template <int DIR> __device__ Number line(const Number *M, const Number *a, const unsigned int row) const {
    constexpr int n = n_q_points_1d;
    constexpr int stride = DIR == 0 ? 1 : DIR == 1 ? n : n * n;
    const unsigned int base = idx - (DIR == 0 ? ix : DIR == 1 ? iy : iz) * stride;
    Number s = 0.;
#pragma unroll
    for (int m = 0; m < n; ++m) s += M[row * n + m] * a[base + m * stride];
    return s;
  }
  // transpose: sum_m M[m][col] * a(..m..)
  template <int DIR> __device__ Number line_t(const Number *M, const Number *a, const unsigned int col) const {
    constexpr int n = n_q_points_1d;
    constexpr int stride = DIR == 0 ? 1 : DIR == 1 ? n : n * n;
    const unsigned int base = idx - (DIR == 0 ? ix : DIR == 1 ? iy : iz) * stride;
    Number s = 0.;
#pragma unroll
    for (int m = 0; m < n; ++m) s += M[m * n + col] * a[base + m * stride];
    return s;
  }

  types::global_dof_index *local_to_global;
  unsigned int n_cells;
  unsigned int padding_length;
  const unsigned int constraint_mask;
  const bool use_coloring;
  Number *inv_jac;
  Number *JxW;
  Number *values;
  Number *gradients[dim];
  const data_type *mf;
  const Number *tab_B, *tab_D;     // 1D tables: the CTA's shared-memory copies if the kernel made them
  unsigned int ix, iy, iz, idx;
};

// the reference's operators instantiate deal.II's FEEvaluation (bp5/step-64.cu:156,
// step-64/step-64.cu:211); fe_evaluation_gl.h is its structural copy -- one class here
template <int dim, int fe_degree, int n_q_points_1d = fe_degree + 1, int n_components_ = 1, typename Number = double>
using FEEvaluation = FEEvaluationGL<dim, fe_degree, n_q_points_1d, n_components_, Number>;

}  // namespace CUDAWrappers
}  // namespace dealii
