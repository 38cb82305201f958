// BP5 benchmark driver on the dealii_b200 facade: the same experiment as
// PoissonProblem<dim,degree>::run in the reference (bp5/step-64.cu:621-678) --
// mesh ladder :633-654, three timed blocks :434-548 ("pcg-standard", "pcg-merged",
// "vmult"), best of n_repetitions, throughput = n_dofs * iterations / wall time
// (:458-461), greppable "<tag> <DoFs> <throughput>" lines (:470-474,512-516,543-547).
// Host code only; every device operation goes through the C ABI (include/bp5_b200.h).
//
//   bp5_step64 [--degree 5] [--cycle-min 7] [--cycle-max 24] [--iterations 200]
//              [--repetitions 10] [--min-run 0] [--quadrature gauss|gll]
//              [--refine-corner 1]   (not in the reference: the octant of cells at the origin is refined once more,
//                                     a locally refined mesh with hanging nodes on the octant's inner faces)
#include <cstdlib>
#include <cstring>
#include <limits>

#include "dealii_b200/dealii_b200.h"

using namespace dealii;
using VectorType = LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>;

struct Options {
  unsigned degree = 5, cycle_min = 7, cycle_max = 24, n_iterations = 200, n_repetitions = 10, min_run = 0;
  int quadrature = BP5_QUAD_GAUSS;
  int refine_corner = 0;
};

template <int dim, int fe_degree>
class PoissonProblem {
 public:
  explicit PoissonProblem(const Options &o) : opt(o), fe(fe_degree), dof_handler(triangulation), pcout(std::cout, true) {}

  void run() {
    for (unsigned cycle = opt.cycle_min; cycle <= opt.cycle_max; ++cycle) {
      pcout << "Cycle " << cycle << std::endl;
      // ladder: cells of edge 2^-n_refine on a box of `subdivisions` unit cubes
      unsigned n_refine = cycle / 6;
      const unsigned remainder = cycle % 6;
      std::vector<unsigned> subdivisions(dim, 1);
      if (remainder == 1 && cycle > 1) { subdivisions = {3, 2, 2}; n_refine -= 1; }
      if (remainder == 2) subdivisions[0] = 2;
      else if (remainder == 3) subdivisions[0] = 3;
      else if (remainder == 4) subdivisions[0] = subdivisions[1] = 2;
      else if (remainder == 5) { subdivisions[0] = 3; subdivisions[1] = 2; }
      Point<dim> p2;
      for (unsigned d = 0; d < dim; ++d) p2[d] = subdivisions[d];
      triangulation.clear();
      GridGenerator::subdivided_hyper_rectangle(triangulation, subdivisions, Point<dim>(), p2);
      triangulation.refine_global(n_refine);
      if (opt.refine_corner) {      // what set_refine_flag() on those cells + execute_coarsening_and_refinement() does
        std::array<unsigned, 3> hi;
        for (unsigned d = 0; d < dim; ++d) hi[d] = std::max(1u, triangulation.cells(d) / 2);
        triangulation.refine_cells_in_box({0u, 0u, 0u}, hi);
      }

      setup_system();
      pcout << "   Number of active cells:       " << triangulation.n_global_active_cells() << std::endl
            << "   Number of degrees of freedom: " << dof_handler.n_dofs() << std::endl
            << std::endl;
      system_matrix_dev->assemble_rhs(system_rhs_dev);   // assemble_rhs(), on the device
      solve();
      // output_results (bp5/step-64.cu:565-616): VTU output is disabled there (:569); the L2 norm is printed
      pcout << "  solution norm: " << system_matrix_dev->l2_norm(solution_dev) << std::endl;
      pcout << std::endl;
    }
  }

 private:
  void setup_system() {
    dof_handler.distribute_dofs(fe);
    constraints.clear();
    constraints.close();
    system_matrix_dev.reset(new BP5::PoissonOperator<dim, fe_degree>(dof_handler, constraints, opt.quadrature));
    system_matrix_dev->initialize_dof_vector(solution_dev);
    system_rhs_dev.reinit(solution_dev);
  }

  template <typename Solver>
  void timed_solves(const char *tag, bool zero_out) {
    DiagonalMatrix<VectorType> preconditioner;
    preconditioner.get_vector().reinit(system_rhs_dev);
    preconditioner.get_vector() = 1.;
    double throughput_max = std::numeric_limits<double>::min();
    for (unsigned i = 0; i < opt.n_repetitions; ++i) {
      system_matrix_dev->do_zero_out = zero_out;
      Timer time;
      IterationNumberControl solver_control(opt.n_iterations, 1e-6 * system_rhs_dev.l2_norm());
      Solver cg(solver_control);
      solution_dev = 0;
      cg.solve(*system_matrix_dev, solution_dev, system_rhs_dev, preconditioner);
      b200::Context::synchronize();
      const double measured_time = time.wall_time();
      const double measured_throughput = static_cast<double>(dof_handler.n_dofs()) * solver_control.last_step() / measured_time;
      throughput_max = std::max(throughput_max, measured_throughput);
      pcout << "   Solved in " << solver_control.last_step() << " iterations with time " << measured_time
            << " and DoFs/s " << measured_throughput << " norm " << solution_dev.l2_norm() << std::endl;
    }
    pcout << tag << " " << dof_handler.n_dofs() << " " << throughput_max << std::endl << std::endl;
  }

  void solve() {
    if (opt.min_run == 0) timed_solves<SolverCG<VectorType>>("pcg-standard", true);
    timed_solves<SolverCGFullMerge<VectorType>>("pcg-merged", false);
    if (opt.min_run == 0) {
      double throughput_max = std::numeric_limits<double>::min();
      system_matrix_dev->do_zero_out = true;   // the reference leaves it false here and accumulates garbage
      for (unsigned i = 0; i < opt.n_repetitions; ++i) {
        Timer time;
        for (unsigned t = 0; t < opt.n_iterations; ++t) system_matrix_dev->vmult(system_rhs_dev, solution_dev);
        b200::Context::synchronize();
        const double measured_time = time.wall_time();
        const double measured_throughput = static_cast<double>(dof_handler.n_dofs()) * opt.n_iterations / measured_time;
        throughput_max = std::max(throughput_max, measured_throughput);
        pcout << "   " << opt.n_iterations << " mat-vecs in time " << measured_time << " and DoFs/s "
              << measured_throughput << std::endl;
      }
      pcout << "vmult " << dof_handler.n_dofs() << " " << throughput_max << std::endl << std::endl;
      system_matrix_dev->assemble_rhs(system_rhs_dev);
    }
  }

  Options opt;
  parallel::distributed::Triangulation<dim> triangulation;
  FE_Q<dim> fe;
  DoFHandler<dim> dof_handler;
  AffineConstraints<double> constraints;
  std::unique_ptr<BP5::PoissonOperator<dim, fe_degree>> system_matrix_dev;
  VectorType solution_dev, system_rhs_dev;
  ConditionalOStream pcout;
};

template <int degree> void run_degree(const Options &o) { PoissonProblem<3, degree> p(o); p.run(); }

int main(int argc, char *argv[]) {
  try {
    Options o;
    for (int i = 1; i + 1 < argc; i += 2) {
      const std::string k = argv[i];
      const char *v = argv[i + 1];
      if (k == "--degree") o.degree = std::atoi(v);
      else if (k == "--cycle-min") o.cycle_min = std::atoi(v);
      else if (k == "--cycle-max") o.cycle_max = std::atoi(v);
      else if (k == "--iterations") o.n_iterations = std::atoi(v);
      else if (k == "--repetitions") o.n_repetitions = std::atoi(v);
      else if (k == "--min-run") o.min_run = std::atoi(v);
      else if (k == "--refine-corner") o.refine_corner = std::atoi(v);
      else if (k == "--quadrature") o.quadrature = std::strcmp(v, "gll") == 0 ? BP5_QUAD_GLL : BP5_QUAD_GAUSS;
      else throw ExcMessage("unknown option " + k);
    }
    std::cout << std::endl << "bp5_b200 info:" << std::endl << std::endl << "  " << bp5_version() << std::endl << std::endl;
    switch (o.degree) {
      case 1: run_degree<1>(o); break; case 2: run_degree<2>(o); break; case 3: run_degree<3>(o); break;
      case 4: run_degree<4>(o); break; case 5: run_degree<5>(o); break; case 6: run_degree<6>(o); break;
      case 7: run_degree<7>(o); break; case 8: run_degree<8>(o); break;
      default: throw ExcMessage("degree must be 1..8");
    }
  } catch (std::exception &exc) {
    std::cerr << std::endl << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  }
  return 0;
}
