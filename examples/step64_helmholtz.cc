// step-64 (variable-coefficient Helmholtz, Q3) on the dealii_b200 facade: what
// HelmholtzProblem<3,3>::run does in the reference (step-64/step-64.cu:605-688),
// first with SolverCG, then with the merged solver; SolverControl(n_dofs, 1e-12 |b|) (:513-514).
#include <cstdlib>

#include "dealii_b200/dealii_b200.h"

using namespace dealii;
using VectorType = LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>;

template <int dim, int fe_degree> void run(bool use_merged, unsigned n_cycles) {
  Triangulation<dim> triangulation;
  FE_Q<dim> fe(fe_degree);
  DoFHandler<dim> dof_handler(triangulation);
  AffineConstraints<double> constraints;
  GridGenerator::hyper_cube(triangulation, 0., 1.);
  for (unsigned cycle = 0; cycle < n_cycles; ++cycle) {
    std::cout << "Cycle " << cycle << std::endl;
    triangulation.refine_global(1);
    dof_handler.distribute_dofs(fe);
    Step64::HelmholtzOperator<dim, fe_degree> system_matrix_dev(dof_handler, constraints);
    VectorType solution_dev, system_rhs_dev;
    system_matrix_dev.initialize_dof_vector(solution_dev);
    system_rhs_dev.reinit(solution_dev);
    std::cout << "   Number of active cells:       " << triangulation.n_global_active_cells() << std::endl
              << "   Number of degrees of freedom: " << dof_handler.n_dofs() << std::endl;
    system_matrix_dev.assemble_rhs(system_rhs_dev);
    DiagonalMatrix<VectorType> preconditioner;
    preconditioner.get_vector().reinit(system_rhs_dev);
    preconditioner.get_vector() = 1.;
    SolverControl solver_control(system_rhs_dev.size(), 1e-12 * system_rhs_dev.l2_norm());
    if (use_merged) {
      system_matrix_dev.do_zero_out = false;
      SolverCGFullMerge<VectorType> cg(solver_control);
      cg.solve(system_matrix_dev, solution_dev, system_rhs_dev, preconditioner);
    } else {
      SolverCG<VectorType> cg(solver_control);
      cg.solve(system_matrix_dev, solution_dev, system_rhs_dev, preconditioner);
    }
    std::cout << "  Solved in " << solver_control.last_step() << " iterations." << std::endl;
    std::cout << "  solution l2 norm: " << solution_dev.l2_norm() << std::endl;
    std::cout << "  solution norm: " << system_matrix_dev.l2_norm(solution_dev) << std::endl << std::endl;   // output_results, :590-601
  }
}

int main(int argc, char *argv[]) {
  try {
    const unsigned n_cycles = argc > 1 ? std::atoi(argv[1]) : 1;   // the reference breaks after one cycle (:630)
    run<3, 3>(false, n_cycles);
    run<3, 3>(true, n_cycles);
  } catch (std::exception &exc) {
    std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl;
    return 1;
  }
  return 0;
}
