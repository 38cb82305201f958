// The reference's operators, written the way the reference writes them -- as device functors on
// CUDAWrappers::MatrixFree / FEEvaluationGL -- and compiled against this library's re-creation of
// that interface (include/dealii_b200/cuda_matrix_free.cuh):
//
//   UserBP5::JacobianFunctor, LocalPoissonOperator, PoissonOperator   after bp5/step-64.cu:60-276
//   UserStep64::VaryingCoefficientFunctor, HelmholtzOperatorQuad,
//               LocalHelmholtzOperator, HelmholtzOperator             after step-64/step-64.cu:69-322
//
// The program applies each user-written operator and the library's tuned operator
// (BP5::PoissonOperator / Step64::HelmholtzOperator, csrc/apply.cuh) to the same vectors, solves
// with SolverCGFullMerge around both, and prints norms that tests/test_gpu_functor_api.py compares
// with the CPU oracle.
//
//   bp5_functors <degree 2..6> <gauss|gll> <cells_x> <cells_y> <cells_z> <deformation eps> [cell edge = 1]
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <iomanip>

#include "user_operators.cuh"


// ------------------------------------------------------------------------------------------------
static double rel_diff(const std::vector<double> &a, const std::vector<double> &b) {
  double num = 0., den = 0.;
  for (std::size_t i = 0; i < a.size(); ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
  return std::sqrt(num / (den > 0. ? den : 1.));
}

static void dump(const std::string &prefix, const char *name, const std::vector<double> &v) {
  if (prefix.empty()) return;
  FILE *f = std::fopen((prefix + name + ".f64").c_str(), "wb");
  if (!f) throw ExcMessage("cannot write " + prefix + name);
  std::fwrite(v.data(), sizeof(double), v.size(), f);
  std::fclose(f);
}
static bool same_bits(const std::vector<double> &a, const std::vector<double> &b) {
  return a.size() == b.size() && std::memcmp(a.data(), b.data(), sizeof(double) * a.size()) == 0;
}

template <int dim, int fe_degree>
int run(bool collocation, const std::vector<unsigned int> &cells, double eps, double cell_edge, const std::string &dump_prefix) {
  parallel::distributed::Triangulation<dim> triangulation;
  Point<dim> p2;
  for (int d = 0; d < dim; ++d) p2[d] = cells[d] * cell_edge;
  GridGenerator::subdivided_hyper_rectangle(triangulation, cells, Point<dim>(), p2);
  if (eps != 0.) { triangulation.deformation = 1; triangulation.deformation_eps = eps; }
  FE_Q<dim> fe(fe_degree);
  DoFHandler<dim> dof_handler(triangulation);
  dof_handler.distribute_dofs(fe);
  AffineConstraints<double> constraints;
  std::cout << std::setprecision(15);
  std::cout << "n_dofs " << dof_handler.n_dofs() << std::endl;
  int failures = 0;
  auto report = [&](const char *what, double err, double tol) {
    std::cout << what << " " << err << (err <= tol ? "" : "   <-- FAIL") << std::endl;
    if (!(err <= tol)) ++failures;
  };

  {  // ---- BP5
    UserBP5::PoissonOperator<dim, fe_degree> user_op(dof_handler, constraints, collocation);
    BP5::PoissonOperator<dim, fe_degree> lib_op(dof_handler, constraints, collocation ? BP5_QUAD_GLL : BP5_QUAD_GAUSS);
    // the coefficient the user functor computed from inv_jacobian / JxW vs the library's stored metric
    std::vector<double> cu(user_op.n_coefficients()), cl(user_op.n_coefficients());
    cudaMemcpy(cu.data(), user_op.coefficients(), sizeof(double) * cu.size(), cudaMemcpyDeviceToHost);
    b200::check(bp5_operator_export_coefficients(lib_op.handle(), cl.data()));
    report("bp5_coefficient_rel_diff", rel_diff(cu, cl), 1e-13);

    VectorType b, y_user, y_lib, y_plain, z_user, x_user, x_lib;
    lib_op.initialize_dof_vector(b);
    lib_op.assemble_rhs(b);
    VectorType bu;                       // same values in a vector that belongs to the user operator
    user_op.initialize_dof_vector(bu);
    bu.equ(1., b);
    y_user.reinit(bu); y_plain.reinit(bu); z_user.reinit(bu); x_user.reinit(bu);
    y_lib.reinit(b); x_lib.reinit(b);
    user_op.vmult(y_user, bu);
    user_op.vmult_plain(y_plain, bu);
    lib_op.vmult(y_lib, b);
    user_op.vmult(z_user, y_user);
    std::vector<double> a, c;
    y_user.copy_to_host(a); y_lib.copy_to_host(c);
    report("bp5_vmult_user_vs_library", rel_diff(a, c), 1e-12);
    dump(dump_prefix, "bp5_Ab", a);
    y_plain.copy_to_host(a);
    report("bp5_vmult_plain_vs_library", rel_diff(a, c), 1e-12);
    {  // use_coloring: eight colour passes with plain += (fe_evaluation_gl.h:176-177): same operator, and
       // bitwise reproducible from run to run (atomics add in whatever order the hardware schedules)
      UserBP5::PoissonOperator<dim, fe_degree> col_op(dof_handler, constraints, collocation, /*use_coloring=*/true);
      VectorType bc, y1, y2;
      col_op.initialize_dof_vector(bc);
      bc.equ(1., b);
      y1.reinit(bc); y2.reinit(bc);
      col_op.vmult_plain(y1, bc);
      col_op.vmult_plain(y2, bc);
      std::vector<double> a1, a2;
      y1.copy_to_host(a1); y2.copy_to_host(a2);
      report("bp5_vmult_colored_vs_library", rel_diff(a1, c), 1e-12);
      std::cout << "bp5_colored_bitwise_reproducible " << (same_bits(a1, a2) ? 1 : 0) << std::endl;
      if (!same_bits(a1, a2)) ++failures;
    }
    std::cout << "bp5_norm_b " << b.l2_norm() << std::endl;
    std::cout << "bp5_norm_Ab " << y_user.l2_norm() << std::endl;
    std::cout << "bp5_norm_AAb " << z_user.l2_norm() << std::endl;

    DiagonalMatrix<VectorType> preconditioner;
    preconditioner.get_vector().reinit(b);
    preconditioner.get_vector() = 1.;
    const double tol = 1e-6 * b.l2_norm();
    IterationNumberControl c_user(200, tol), c_lib(200, tol), c_std(200, tol);
    user_op.do_zero_out = false;         // bp5/step-64.cu:483
    SolverCGFullMerge<VectorType>(c_user).solve(user_op, x_user, bu, preconditioner);
    lib_op.do_zero_out = false;
    SolverCGFullMerge<VectorType>(c_lib).solve(lib_op, x_lib, b, preconditioner);
    std::cout << "bp5_merged_its_user " << c_user.last_step() << std::endl;
    std::cout << "bp5_merged_its_library " << c_lib.last_step() << std::endl;
    std::cout << "bp5_norm_x " << x_user.l2_norm() << std::endl;
    x_user.copy_to_host(a); x_lib.copy_to_host(c);
    report("bp5_x_user_vs_library", rel_diff(a, c), 1e-8);
    dump(dump_prefix, "bp5_x", a);
    if (std::abs((int)c_user.last_step() - (int)c_lib.last_step()) > 1) { ++failures; std::cout << "iteration counts differ <-- FAIL\n"; }
    user_op.do_zero_out = true;
    x_user = 0.;
    SolverCG<VectorType>(c_std).solve(user_op, x_user, bu, preconditioner);
    std::cout << "bp5_standard_its_user " << c_std.last_step() << std::endl;
    x_user.copy_to_host(a);
    report("bp5_x_standard_vs_library", rel_diff(a, c), 1e-8);
  }
  if (!collocation) {  // ---- step-64 Helmholtz (QGauss only in the reference)
    UserStep64::HelmholtzOperator<dim, fe_degree> user_op(dof_handler, constraints);
    Step64::HelmholtzOperator<dim, fe_degree> lib_op(dof_handler, constraints);
    VectorType b, bu, y_user, y_lib, x_user, x_lib;
    lib_op.initialize_dof_vector(b);
    lib_op.assemble_rhs(b);
    user_op.initialize_dof_vector(bu);
    bu.equ(1., b);
    y_user.reinit(bu); x_user.reinit(bu); y_lib.reinit(b); x_lib.reinit(b);
    user_op.vmult(y_user, bu);
    lib_op.vmult(y_lib, b);
    std::vector<double> a, c;
    y_user.copy_to_host(a); y_lib.copy_to_host(c);
    report("helmholtz_vmult_user_vs_library", rel_diff(a, c), 1e-12);
    dump(dump_prefix, "helmholtz_Ab", a);
    {  // the same operator with use_coloring (the coefficient is indexed through local_q_point_id: colour-safe)
      UserStep64::HelmholtzOperator<dim, fe_degree> col_op(dof_handler, constraints, /*use_coloring=*/true);
      VectorType bc, y1, y2;
      col_op.initialize_dof_vector(bc);
      bc.equ(1., b);
      y1.reinit(bc); y2.reinit(bc);
      col_op.vmult(y1, bc);
      col_op.vmult(y2, bc);
      std::vector<double> a1, a2;
      y1.copy_to_host(a1); y2.copy_to_host(a2);
      report("helmholtz_vmult_colored_vs_library", rel_diff(a1, c), 1e-12);
      std::cout << "helmholtz_colored_bitwise_reproducible " << (same_bits(a1, a2) ? 1 : 0) << std::endl;
      if (!same_bits(a1, a2)) ++failures;
    }
    std::cout << "helmholtz_norm_Ab " << y_user.l2_norm() << std::endl;
    DiagonalMatrix<VectorType> preconditioner;
    preconditioner.get_vector().reinit(b);
    preconditioner.get_vector() = 1.;
    SolverControl c_user(b.size(), 1e-12 * b.l2_norm()), c_lib(b.size(), 1e-12 * b.l2_norm());   // step-64/step-64.cu:513
    SolverCGFullMerge<VectorType>(c_user).solve(user_op, x_user, bu, preconditioner);
    SolverCG<VectorType>(c_lib).solve(lib_op, x_lib, b, preconditioner);
    std::cout << "helmholtz_merged_its_user " << c_user.last_step() << std::endl;
    std::cout << "helmholtz_standard_its_library " << c_lib.last_step() << std::endl;
    std::cout << "helmholtz_norm_x " << x_user.l2_norm() << std::endl;
    x_user.copy_to_host(a); x_lib.copy_to_host(c);
    report("helmholtz_x_user_vs_library", rel_diff(a, c), 1e-8);
  }
  std::cout << (failures ? "FAILED" : "OK") << std::endl;
  return failures;
}

// `bp5_functors bench <degree> <cells per direction>`: what staying on the reference's functor API costs --
// user-written LocalPoissonOperator (one CTA per cell, deal.II-layout arrays, atomics) against the library's
// tuned kernel on the same mesh, GLL collocation, CUDA events on the library's stream.
template <int dim, int fe_degree> int run_bench(unsigned int nc) {
  parallel::distributed::Triangulation<dim> triangulation;
  Point<dim> p2;
  for (int d = 0; d < dim; ++d) p2[d] = nc;
  GridGenerator::subdivided_hyper_rectangle(triangulation, std::vector<unsigned int>(dim, nc), Point<dim>(), p2);
  FE_Q<dim> fe(fe_degree);
  DoFHandler<dim> dof_handler(triangulation);
  dof_handler.distribute_dofs(fe);
  AffineConstraints<double> constraints;
  UserBP5::PoissonOperator<dim, fe_degree> user_op(dof_handler, constraints, true);
  BP5::PoissonOperator<dim, fe_degree> lib_op(dof_handler, constraints, BP5_QUAD_GLL);
  VectorType bu, yu, bl, yl;
  user_op.initialize_dof_vector(bu); yu.reinit(bu);
  lib_op.initialize_dof_vector(bl); yl.reinit(bl);
  lib_op.assemble_rhs(bl);
  bu.equ(1., bl);
  cudaStream_t stream = static_cast<cudaStream_t>(bp5_context_stream(b200::Context::get()));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 10;
  float ms_user = 0.f, ms_lib = 0.f;
  for (int w = 0; w < 2; ++w) { user_op.vmult(yu, bu); lib_op.vmult(yl, bl); }
  cudaEventRecord(e0, stream);
  for (int i = 0; i < reps; ++i) user_op.vmult(yu, bu);
  cudaEventRecord(e1, stream); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_user, e0, e1);
  cudaEventRecord(e0, stream);
  for (int i = 0; i < reps; ++i) lib_op.vmult(yl, bl);
  cudaEventRecord(e1, stream); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_lib, e0, e1);
  double bytes_vmult = 0., bytes_cg = 0.;
  b200::check(bp5_operator_algorithmic_bytes(lib_op.handle(), &bytes_vmult, &bytes_cg));
  std::vector<double> a, c;
  yu.copy_to_host(a); yl.copy_to_host(c);
  std::cout << std::setprecision(8) << "{\"degree\": " << fe_degree << ", \"cells\": " << nc << ", \"dofs\": " << dof_handler.n_dofs()
            << ", \"user_functor_vmult_ms\": " << ms_user / reps << ", \"library_vmult_ms\": " << ms_lib / reps
            << ", \"algorithmic_bytes_per_vmult\": " << bytes_vmult << ", \"rel_diff\": " << rel_diff(a, c) << "}" << std::endl;
  return rel_diff(a, c) <= 1e-12 ? 0 : 1;
}

int main(int argc, char *argv[]) {
  try {
    if (argc > 3 && std::strcmp(argv[1], "bench") == 0) {
      const unsigned int nc = std::atoi(argv[3]);
      switch (std::atoi(argv[2])) {
        case 4: return run_bench<3, 4>(nc);
        case 6: return run_bench<3, 6>(nc);
        default: throw ExcMessage("bench: degree 4 or 6");
      }
    }
    const int degree = argc > 1 ? std::atoi(argv[1]) : 4;
    const bool collocation = argc > 2 && std::strcmp(argv[2], "gll") == 0;
    std::vector<unsigned int> cells(3, 3);
    for (int d = 0; d < 3; ++d)
      if (argc > 3 + d) cells[d] = std::atoi(argv[3 + d]);
    const double eps = argc > 6 ? std::atof(argv[6]) : 0.1;
    const double cell_edge = argc > 7 ? std::atof(argv[7]) : 1.0;
    const std::string dump_prefix = argc > 8 ? argv[8] : "";     // write the vectors as raw doubles: <prefix><name>.f64
    switch (degree) {
      case 1: return run<3, 1>(collocation, cells, eps, cell_edge, dump_prefix);
      case 2: return run<3, 2>(collocation, cells, eps, cell_edge, dump_prefix);
      case 3: return run<3, 3>(collocation, cells, eps, cell_edge, dump_prefix);
      case 4: return run<3, 4>(collocation, cells, eps, cell_edge, dump_prefix);
      case 5: return run<3, 5>(collocation, cells, eps, cell_edge, dump_prefix);
      case 6: return run<3, 6>(collocation, cells, eps, cell_edge, dump_prefix);
      case 7: return run<3, 7>(collocation, cells, eps, cell_edge, dump_prefix);
      case 8: return run<3, 8>(collocation, cells, eps, cell_edge, dump_prefix);
      default: throw ExcMessage("degree must be 1..8");
    }
  } catch (std::exception &exc) {
    std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl;
    return 1;
  }
}
