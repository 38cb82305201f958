// The reference's operators, written the way the reference writes them -- as device functors on
// CUDAWrappers::MatrixFree / FEEvaluationGL -- against include/dealii_b200/cuda_matrix_free.cuh:
//
//   UserBP5::JacobianFunctor, LocalPoissonOperator, PoissonOperator   after bp5/step-64.cu:60-276
//   UserStep64::VaryingCoefficientFunctor, HelmholtzOperatorQuad,
//               LocalHelmholtzOperator, HelmholtzOperator             after step-64/step-64.cu:69-322
//
// Shared by examples/bp5_functors.cu (conforming meshes, against the tuned kernel) and examples/bp5_hanging.cu
// (locally refined meshes with hanging nodes).
#pragma once
#include "dealii_b200/cuda_matrix_free.cuh"

using namespace dealii;
using VectorType = LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>;

namespace UserBP5 {
// merged coefficient G = JxW * J^-1 J^-T, upper triangle in planes xx,yy,zz,xy,xz,yz
template <int dim, int fe_degree> class JacobianFunctor {
 public:
  JacobianFunctor(double *coefficient, const unsigned int n_cells) : coef(coefficient), n_cells(n_cells) {}
  __device__ void operator()(const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data);
  static const unsigned int n_dofs_1d = fe_degree + 1;
  static const unsigned int n_q_points = Utilities::pow(n_dofs_1d, dim);

 private:
  double *coef;
  const unsigned int n_cells;
};

template <int dim, int fe_degree>
__device__ void JacobianFunctor<dim, fe_degree>::operator()(
    const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data) {
  const unsigned int q = CUDAWrappers::q_point_id_in_cell<dim>(fe_degree + 1);
  const std::size_t plane = (std::size_t)gpu_data->n_cells * gpu_data->padding_length;
  const std::size_t at = (std::size_t)cell * gpu_data->padding_length + q;
  Tensor<2, dim> inv_jac;
  for (unsigned int d = 0; d < dim; ++d)
    for (unsigned int e = 0; e < dim; ++e) inv_jac[d][e] = gpu_data->inv_jacobian[at + plane * (d * dim + e)];
  const double JxW = gpu_data->JxW[at];
  const std::size_t stride = (std::size_t)n_cells * n_q_points, out = (std::size_t)cell * n_q_points + q;
  unsigned int c = dim;
  for (unsigned int d = 0; d < dim; ++d)
    for (unsigned int e = d; e < dim; ++e) {
      double sum = 0.;
      for (unsigned int f = 0; f < dim; ++f) sum += inv_jac[d][f] * inv_jac[e][f];
      coef[out + (d == e ? d : c++) * stride] = JxW * sum;
    }
}

template <int dim, int fe_degree> class LocalPoissonOperator {
 public:
  LocalPoissonOperator(double *coefficient, const unsigned int n_cells) : n_cells(n_cells), coef(coefficient) {}
  __device__ void operator()(const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data,
                             CUDAWrappers::SharedData<dim, double> *shared_data, const double *src, double *dst) const;
  static const unsigned int n_dofs_1d = fe_degree + 1;
  static const unsigned int n_local_dofs = Utilities::pow(fe_degree + 1, dim);
  static const unsigned int n_q_points = Utilities::pow(fe_degree + 1, dim);

 private:
  const unsigned int n_cells;
  double *coef;
};

template <int dim, int fe_degree>
__device__ void LocalPoissonOperator<dim, fe_degree>::operator()(
    const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data,
    CUDAWrappers::SharedData<dim, double> *shared_data, const double *src, double *dst) const {
  CUDAWrappers::FEEvaluationGL<dim, fe_degree, fe_degree + 1, 1, double> fe_eval(cell, gpu_data, shared_data);
  fe_eval.read_dof_values(src);
  fe_eval.evaluate(false, true);
  // g <- G g with THIS cell's metric (the shipped kernel reads cell 0's, bp5/step-64.cu:161-177)
  const std::size_t offset = (std::size_t)n_q_points * n_cells;
  const unsigned int q = CUDAWrappers::internal::compute_index<dim, fe_degree + 1>();
  const double *G = coef + (std::size_t)cell * n_q_points + q;
  const double g0 = shared_data->gradients[0][q], g1 = shared_data->gradients[1][q], g2 = shared_data->gradients[2][q];
  shared_data->gradients[0][q] = g0 * G[0] + g1 * G[3 * offset] + g2 * G[4 * offset];
  shared_data->gradients[1][q] = g0 * G[3 * offset] + g1 * G[1 * offset] + g2 * G[5 * offset];
  shared_data->gradients[2][q] = g0 * G[4 * offset] + g1 * G[5 * offset] + g2 * G[2 * offset];
  __syncthreads();
  fe_eval.integrate(false, true);
  fe_eval.distribute_local_to_global(dst);
}

// the non-merged branch of the reference (#else at bp5/step-64.cu:189-191): J^-1 and JxW at every point
template <int dim, int fe_degree> class LocalPoissonOperatorPlain {
 public:
  __device__ void operator()(const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data,
                             CUDAWrappers::SharedData<dim, double> *shared_data, const double *src, double *dst) const {
    CUDAWrappers::FEEvaluationGL<dim, fe_degree, fe_degree + 1, 1, double> fe_eval(cell, gpu_data, shared_data);
    fe_eval.read_dof_values(src);
    fe_eval.evaluate(false, true);
    fe_eval.submit_gradient(fe_eval.get_gradient());
    __syncthreads();
    fe_eval.integrate(false, true);
    fe_eval.distribute_local_to_global(dst);
  }
  static const unsigned int n_dofs_1d = fe_degree + 1;
  static const unsigned int n_local_dofs = Utilities::pow(fe_degree + 1, dim);
  static const unsigned int n_q_points = Utilities::pow(fe_degree + 1, dim);
};

// b_i = int phi_i on the device: what assemble_rhs does with FEValues on the host (bp5/step-64.cu:372-418);
// distribute_local_to_global applies the hanging-node constraints like constraints.distribute_local_to_global there
template <int dim, int fe_degree> class LocalRhsOperator {
 public:
  __device__ void operator()(const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data,
                             CUDAWrappers::SharedData<dim, double> *shared_data, const double *, double *dst) const {
    CUDAWrappers::FEEvaluationGL<dim, fe_degree, fe_degree + 1, 1, double> fe_eval(cell, gpu_data, shared_data);
    fe_eval.submit_value(1., CUDAWrappers::internal::compute_index<dim, fe_degree + 1>());
    __syncthreads();
    fe_eval.integrate(true, false);
    fe_eval.distribute_local_to_global(dst);
  }
  static const unsigned int n_dofs_1d = fe_degree + 1;
  static const unsigned int n_local_dofs = Utilities::pow(fe_degree + 1, dim);
  static const unsigned int n_q_points = Utilities::pow(fe_degree + 1, dim);
};

template <int dim, int fe_degree> class PoissonOperator {
 public:
  PoissonOperator(const DoFHandler<dim> &dof_handler, const AffineConstraints<double> &constraints, bool collocation,
                  bool use_coloring = false);
  void vmult(VectorType &dst, const VectorType &src) const;
  void assemble_rhs(VectorType &b) const {
    VectorType zero;
    zero.reinit(b);
    zero = 0.; b = 0.;
    mf_data.cell_loop(LocalRhsOperator<dim, fe_degree>(), zero, b);
    mf_data.copy_constrained_values(zero, b);          // constrained rows: 0
  }
  const CUDAWrappers::MatrixFree<dim, double> &matrix_free() const { return mf_data; }
  void vmult_plain(VectorType &dst, const VectorType &src) const;
  void initialize_dof_vector(VectorType &vec) const { mf_data.initialize_dof_vector(vec); }
  const double *coefficients() const { return coef.get_values(); }
  std::size_t n_coefficients() const { return coef.size(); }

 private:
  CUDAWrappers::MatrixFree<dim, double> mf_data;
  LinearAlgebra::CUDAWrappers::Vector<double> coef;
  unsigned int n_owned_cells;

 public:
  bool do_zero_out;
};

template <int dim, int fe_degree>
PoissonOperator<dim, fe_degree>::PoissonOperator(const DoFHandler<dim> &dof_handler,
                                                 const AffineConstraints<double> &constraints, bool collocation,
                                                 bool use_coloring)
    : do_zero_out(true) {
  MappingQGeneric<dim> mapping(fe_degree);
  typename CUDAWrappers::MatrixFree<dim, double>::AdditionalData additional_data;
  additional_data.mapping_update_flags = update_values | update_gradients | update_JxW_values | update_quadrature_points;
  additional_data.overlap_communication_computation = true;
  // colouring: only vmult_plain may be used then -- like the reference's, JacobianFunctor / LocalPoissonOperator
  // index the coefficient by the colour-local cell number without the colour's row offset
  additional_data.use_coloring = use_coloring;
  if (collocation) mf_data.reinit(mapping, dof_handler, constraints, QGaussLobatto<1>(fe_degree + 1), additional_data);
  else mf_data.reinit(mapping, dof_handler, constraints, QGauss<1>(fe_degree + 1), additional_data);
  n_owned_cells =
      dynamic_cast<const parallel::Triangulation<dim> *>(&dof_handler.get_triangulation())->n_locally_owned_active_cells();
  coef.reinit(Utilities::pow(fe_degree + 1, dim) * n_owned_cells * dim * (dim + 1) / 2);
  const JacobianFunctor<dim, fe_degree> functor(coef.get_values(), n_owned_cells);
  mf_data.evaluate_coefficients(functor);
}

template <int dim, int fe_degree> void PoissonOperator<dim, fe_degree>::vmult(VectorType &dst, const VectorType &src) const {
  if (do_zero_out) dst = 0.;
  LocalPoissonOperator<dim, fe_degree> local_poisson_operator(coef.get_values(), n_owned_cells);
  mf_data.cell_loop(local_poisson_operator, src, dst);
  mf_data.copy_constrained_values(src, dst);
}
template <int dim, int fe_degree>
void PoissonOperator<dim, fe_degree>::vmult_plain(VectorType &dst, const VectorType &src) const {
  if (do_zero_out) dst = 0.;
  mf_data.cell_loop(LocalPoissonOperatorPlain<dim, fe_degree>(), src, dst);
  mf_data.copy_constrained_values(src, dst);
}
}  // namespace UserBP5

namespace UserStep64 {
template <int dim, int fe_degree> class VaryingCoefficientFunctor {
 public:
  VaryingCoefficientFunctor(double *coefficient) : coef(coefficient) {}
  __device__ void operator()(const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data) {
    const unsigned int pos = CUDAWrappers::local_q_point_id<dim, double>(cell, gpu_data, n_dofs_1d, n_q_points);
    const Point<dim> q_point = CUDAWrappers::get_quadrature_point<dim, double>(cell, gpu_data, n_dofs_1d);
    double p_square = 0.;
    for (unsigned int i = 0; i < dim; ++i) p_square += q_point[i] * q_point[i];
    coef[pos] = 10. / (0.05 + 2. * p_square);        // a(x), step-64/step-64.cu:100-118
  }
  static const unsigned int n_dofs_1d = fe_degree + 1;
  static const unsigned int n_local_dofs = Utilities::pow(n_dofs_1d, dim);
  static const unsigned int n_q_points = Utilities::pow(n_dofs_1d, dim);

 private:
  double *coef;
};

template <int dim, int fe_degree> class HelmholtzOperatorQuad {
 public:
  __device__ HelmholtzOperatorQuad(double coef) : coef(coef) {}
  __device__ void operator()(CUDAWrappers::FEEvaluation<dim, fe_degree> *fe_eval, const unsigned int q) const {
    fe_eval->submit_value(coef * fe_eval->get_value(q), q);
    fe_eval->submit_gradient(fe_eval->get_gradient(q), q);
  }

 private:
  double coef;
};

template <int dim, int fe_degree> class LocalHelmholtzOperator {
 public:
  LocalHelmholtzOperator(double *coefficient) : coef(coefficient) {}
  __device__ void operator()(const unsigned int cell, const typename CUDAWrappers::MatrixFree<dim, double>::Data *gpu_data,
                             CUDAWrappers::SharedData<dim, double> *shared_data, const double *src, double *dst) const {
    const unsigned int pos = CUDAWrappers::local_q_point_id<dim, double>(cell, gpu_data, n_dofs_1d, n_q_points);
    CUDAWrappers::FEEvaluation<dim, fe_degree, fe_degree + 1, 1, double> fe_eval(cell, gpu_data, shared_data);
    fe_eval.read_dof_values(src);
    fe_eval.evaluate(true, true);
    fe_eval.apply_quad_point_operations(HelmholtzOperatorQuad<dim, fe_degree>(coef[pos]));
    fe_eval.integrate(true, true);
    fe_eval.distribute_local_to_global(dst);
  }
  static const unsigned int n_dofs_1d = fe_degree + 1;
  static const unsigned int n_local_dofs = Utilities::pow(fe_degree + 1, dim);
  static const unsigned int n_q_points = Utilities::pow(fe_degree + 1, dim);

 private:
  double *coef;
};

template <int dim, int fe_degree> class HelmholtzOperator {
 public:
  HelmholtzOperator(const DoFHandler<dim> &dof_handler, const AffineConstraints<double> &constraints,
                    bool use_coloring = false) {
    MappingQGeneric<dim> mapping(fe_degree);
    typename CUDAWrappers::MatrixFree<dim, double>::AdditionalData additional_data;
    additional_data.mapping_update_flags = update_values | update_gradients | update_JxW_values | update_quadrature_points;
    additional_data.use_coloring = use_coloring;      // local_q_point_id adds the colour's row offset
    const QGauss<1> quad(fe_degree + 1);
    mf_data.reinit(mapping, dof_handler, constraints, quad, additional_data);
    const unsigned int n_owned_cells =
        dynamic_cast<const parallel::Triangulation<dim> *>(&dof_handler.get_triangulation())->n_locally_owned_active_cells();
    coef.reinit(Utilities::pow(fe_degree + 1, dim) * n_owned_cells);
    const VaryingCoefficientFunctor<dim, fe_degree> functor(coef.get_values());
    mf_data.evaluate_coefficients(functor);
  }
  void vmult(VectorType &dst, const VectorType &src) const {
    dst = 0.;
    LocalHelmholtzOperator<dim, fe_degree> helmholtz_operator(coef.get_values());
    mf_data.cell_loop(helmholtz_operator, src, dst);
    mf_data.copy_constrained_values(src, dst);
  }
  void initialize_dof_vector(VectorType &vec) const { mf_data.initialize_dof_vector(vec); }

 private:
  CUDAWrappers::MatrixFree<dim, double> mf_data;
  LinearAlgebra::CUDAWrappers::Vector<double> coef;
};
}  // namespace UserStep64
