// Locally refined mesh with hanging nodes through the reference's evaluator interface: the operators of
// user_operators.cuh (bp5/step-64.cu:60-276, step-64/step-64.cu:69-322 written as device functors) on a mesh whose
// cells in a box are refined once.  FEEvaluationGL::read_dof_values / distribute_local_to_global resolve the
// hanging-node constraints from MatrixFree::Data::constraint_mask (bp5/fe_evaluation_gl.h:88,150,167) -- the slot no
// mesh of the reference exercises.  tests/test_gpu_hanging_nodes.py compares the printed numbers and the dumped vectors
// with oracle/hanging_oracle.py.
//
//   bp5_hanging <degree> <gauss|gll> <cx> <cy> <cz> <lo_x> <lo_y> <lo_z> <hi_x> <hi_y> <hi_z> <eps> [dump prefix]
// With a dump prefix: reads <prefix>u.f64 (n_dofs doubles) if present and writes <prefix>{coords,b,Ab,Au,Hu,x}.f64.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>

#include "user_operators.cuh"

static void dump(const std::string &prefix, const char *name, const std::vector<double> &v) {
  if (prefix.empty()) return;
  FILE *f = std::fopen((prefix + name + ".f64").c_str(), "wb");
  if (!f) throw ExcMessage("cannot write " + prefix + name);
  std::fwrite(v.data(), sizeof(double), v.size(), f);
  std::fclose(f);
}
static bool load(const std::string &prefix, const char *name, std::vector<double> &v) {
  if (prefix.empty()) return false;
  FILE *f = std::fopen((prefix + name + ".f64").c_str(), "rb");
  if (!f) return false;
  const std::size_t got = std::fread(v.data(), sizeof(double), v.size(), f);
  std::fclose(f);
  if (got != v.size()) throw ExcMessage("short read of " + prefix + name);
  return true;
}

template <int dim, int fe_degree>
int run(bool collocation, const std::vector<unsigned int> &cells, const std::array<unsigned int, 3> &lo,
        const std::array<unsigned int, 3> &hi, double eps, const std::string &prefix) {
  parallel::distributed::Triangulation<dim> triangulation;
  Point<dim> p2;
  for (int d = 0; d < dim; ++d) p2[d] = 1.;
  GridGenerator::subdivided_hyper_rectangle(triangulation, cells, Point<dim>(), p2);
  triangulation.refine_cells_in_box(lo, hi);
  if (eps != 0.) { triangulation.deformation = 1; triangulation.deformation_eps = eps; }
  FE_Q<dim> fe(fe_degree);
  DoFHandler<dim> dof_handler(triangulation);
  dof_handler.distribute_dofs(fe);
  AffineConstraints<double> constraints;
  std::cout << std::setprecision(15);
  std::cout << "n_dofs " << dof_handler.n_dofs() << std::endl;
  std::cout << "n_cells " << triangulation.n_global_active_cells() << std::endl;

  UserBP5::PoissonOperator<dim, fe_degree> op(dof_handler, constraints, collocation);
  if (op.matrix_free().n_dofs() != dof_handler.n_dofs()) throw ExcMessage("DoFHandler::n_dofs disagrees with the numbering");
  VectorType b, y, z, x, u;
  op.initialize_dof_vector(b);
  y.reinit(b); z.reinit(b); x.reinit(b); u.reinit(b);
  if (collocation) {
    // the reference integrates the right-hand side with QGauss(p+1) in either mode (bp5/step-64.cu:380)
    UserBP5::PoissonOperator<dim, fe_degree> gauss_op(dof_handler, constraints, false);
    VectorType bg;
    gauss_op.initialize_dof_vector(bg);
    gauss_op.assemble_rhs(bg);
    b.equ(1., bg);
  } else {
    op.assemble_rhs(b);
  }
  op.vmult(y, b);
  op.vmult_plain(z, b);
  std::vector<double> h, h2;
  {
    std::vector<double> xyz(3 * dof_handler.n_dofs());
    b200::check(bp5_operator_export_dof_coordinates(op.matrix_free().handle(), xyz.data()));
    dump(prefix, "coords", xyz);
  }
  b.copy_to_host(h); dump(prefix, "b", h);
  y.copy_to_host(h); dump(prefix, "Ab", h);
  z.copy_to_host(h2);
  double num = 0., den = 0.;
  for (std::size_t i = 0; i < h.size(); ++i) { num += (h[i] - h2[i]) * (h[i] - h2[i]); den += h[i] * h[i]; }
  std::cout << "merged_vs_plain_rel_diff " << std::sqrt(num / den) << std::endl;
  {  // the library's tuned operator on the same mesh (same numbering): right-hand side, vmult, merged CG
    BP5::PoissonOperator<dim, fe_degree> lib_op(dof_handler, constraints, collocation ? BP5_QUAD_GLL : BP5_QUAD_GAUSS);
    VectorType bl, yl, xl;
    lib_op.initialize_dof_vector(bl);
    yl.reinit(bl); xl.reinit(bl);
    lib_op.assemble_rhs(bl);
    lib_op.vmult(yl, bl);
    std::vector<double> hb, hl;
    b.copy_to_host(hb); bl.copy_to_host(hl);
    double nb = 0., db = 0., ny = 0., dy = 0.;
    for (std::size_t i = 0; i < hb.size(); ++i) { nb += (hb[i] - hl[i]) * (hb[i] - hl[i]); db += hb[i] * hb[i]; }
    yl.copy_to_host(hl);
    for (std::size_t i = 0; i < h.size(); ++i) { ny += (h[i] - hl[i]) * (h[i] - hl[i]); dy += h[i] * h[i]; }
    std::cout << "library_rhs_rel_diff " << std::sqrt(nb / db) << std::endl;
    std::cout << "library_vmult_rel_diff " << std::sqrt(ny / dy) << std::endl;
    DiagonalMatrix<VectorType> pl;
    pl.get_vector().reinit(bl);
    pl.get_vector() = 1.;
    SolverControl cl(1000, 1e-8 * bl.l2_norm());
    lib_op.do_zero_out = false;
    xl = 0.;
    SolverCGFullMerge<VectorType> sl(cl);
    sl.solve(lib_op, xl, bl, pl);
    std::cout << "library_merged_its " << cl.last_step() << std::endl;
    std::cout << "library_norm_x " << xl.l2_norm() << std::endl;
  }
  std::cout << "norm_b " << b.l2_norm() << std::endl;
  std::cout << "norm_Ab " << y.l2_norm() << std::endl;
  std::vector<double> uh(dof_handler.n_dofs());
  if (load(prefix, "u", uh)) {
    u.import_from_host(uh);
    op.vmult(y, u);
    y.copy_to_host(h); dump(prefix, "Au", h);
    if (!collocation) {
      UserStep64::HelmholtzOperator<dim, fe_degree> helm(dof_handler, constraints);
      VectorType uu, yy;
      helm.initialize_dof_vector(uu);
      yy.reinit(uu);
      uu.import_from_host(uh);
      helm.vmult(yy, uu);
      yy.copy_to_host(h); dump(prefix, "Hu", h);
    }
  }
  // merged CG around the user-written operator, as the reference runs it (bp5/step-64.cu:428-492)
  DiagonalMatrix<VectorType> preconditioner;
  preconditioner.get_vector().reinit(b);
  preconditioner.get_vector() = 1.;
  SolverControl control(1000, 1e-8 * b.l2_norm());
  op.do_zero_out = false;
  x = 0.;
  SolverCGFullMerge<VectorType> solver(control);
  solver.solve(op, x, b, preconditioner);
  std::cout << "merged_its " << control.last_step() << std::endl;
  std::cout << "norm_x " << x.l2_norm() << std::endl;
  x.copy_to_host(h); dump(prefix, "x", h);
  std::cout << "OK" << std::endl;
  return 0;
}

int main(int argc, char *argv[]) {
  try {
    if (argc < 13) throw ExcMessage("usage: bp5_hanging <degree> <gauss|gll> <cells x3> <refine lo x3> <refine hi x3> <eps> [dump prefix]");
    const int degree = std::atoi(argv[1]);
    const bool collocation = std::strcmp(argv[2], "gll") == 0;
    std::vector<unsigned int> cells(3);
    std::array<unsigned int, 3> lo, hi;
    for (int d = 0; d < 3; ++d) {
      cells[d] = std::atoi(argv[3 + d]); lo[d] = std::atoi(argv[6 + d]); hi[d] = std::atoi(argv[9 + d]);
    }
    const double eps = std::atof(argv[12]);
    const std::string prefix = argc > 13 ? argv[13] : "";
    switch (degree) {
      case 1: return run<3, 1>(collocation, cells, lo, hi, eps, prefix);
      case 2: return run<3, 2>(collocation, cells, lo, hi, eps, prefix);
      case 3: return run<3, 3>(collocation, cells, lo, hi, eps, prefix);
      case 4: return run<3, 4>(collocation, cells, lo, hi, eps, prefix);
      case 5: return run<3, 5>(collocation, cells, lo, hi, eps, prefix);
      case 6: return run<3, 6>(collocation, cells, lo, hi, eps, prefix);
      case 7: return run<3, 7>(collocation, cells, lo, hi, eps, prefix);
      case 8: return run<3, 8>(collocation, cells, lo, hi, eps, prefix);
      default: throw ExcMessage("degree must be 1..8");
    }
  } catch (std::exception &exc) {
    std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl;
    return 1;
  }
}
