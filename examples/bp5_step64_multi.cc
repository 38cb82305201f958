// The BP5 driver of the reference on a PARTITIONED mesh, host code in C++ only: what bp5/step-64.cu does with
// MPI_COMM_WORLD, p4est and CUDA-aware MPI (:310,346-367,704-707,720) -- one rank per GPU, device = rank %
// n_devices, owned + ghost vectors, ghost exchange inside vmult, all-rank sums in the solver -- with the
// library's peer-memory transport as the data plane.  No MPI in this image, so the ranks are forked processes
// and the "communicator" (dealii::b200::Communicator: allgather of a few hundred bytes + barrier, used for the
// set-up only) lives in a shared-memory page; an MPI program implements the same two calls with
// MPI_Allgather / MPI_Barrier.
//
//   bp5_step64_multi --ranks 2 [--degree 5] [--cycle-min 7] [--cycle-max 8] [--iterations 200]
//                    [--repetitions 2] [--quadrature gauss|gll] [--devices N]   (N: GPUs to spread the ranks over)
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>

#include "dealii_b200/dealii_b200.h"

using namespace dealii;
using VectorType = LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>;

// ---- shared-memory communicator between forked ranks ------------------------------------------------------
struct SharedPage {
  std::atomic<int> arrived, generation, failed;
  unsigned char slots[64][1024];
};

class ForkCommunicator : public b200::Communicator {
 public:
  ForkCommunicator(SharedPage *page, int rank, int size) : page(page), r(rank), n(size) {}
  int rank() const override { return r; }
  int size() const override { return n; }
  void allgather(const void *send, void *recv_all, std::size_t bytes) override {
    if (bytes > sizeof(page->slots[0])) throw ExcMessage("ForkCommunicator: message too large");
    std::memcpy(page->slots[r], send, bytes);
    barrier();
    for (int i = 0; i < n; ++i) std::memcpy(static_cast<unsigned char *>(recv_all) + i * bytes, page->slots[i], bytes);
    barrier();
  }
  void barrier() override {
    const int gen = page->generation.load();
    if (page->arrived.fetch_add(1) + 1 == n) {
      page->arrived.store(0);
      page->generation.fetch_add(1);
    } else {
      while (page->generation.load() == gen) {
        if (page->failed.load()) throw ExcMessage("another rank failed");
        std::this_thread::sleep_for(std::chrono::microseconds(50));
      }
    }
  }

 private:
  SharedPage *page;
  int r, n;
};

struct Options {
  unsigned degree = 5, cycle_min = 7, cycle_max = 8, n_iterations = 200, n_repetitions = 2, ranks = 2;
  int quadrature = BP5_QUAD_GAUSS, devices = 0;
};

template <int dim, int fe_degree>
class PoissonProblem {
 public:
  PoissonProblem(const Options &o, b200::Communicator &comm)
      : opt(o), triangulation(&comm), fe(fe_degree), dof_handler(triangulation), pcout(std::cout, comm.rank() == 0) {}

  void run() {
    for (unsigned cycle = opt.cycle_min; cycle <= opt.cycle_max; ++cycle) {
      pcout << "Cycle " << cycle << std::endl;
      unsigned n_refine = cycle / 6;                       // the ladder of bp5/step-64.cu:633-654
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

      dof_handler.distribute_dofs(fe);
      constraints.clear();
      constraints.close();
      system_matrix_dev.reset(new BP5::PoissonOperator<dim, fe_degree>(dof_handler, constraints, opt.quadrature));
      system_matrix_dev->initialize_dof_vector(solution_dev);       // owned + ghost entries (:363-366)
      system_rhs_dev.reinit(solution_dev);
      pcout << "   Number of active cells:       " << triangulation.n_global_active_cells() << std::endl
            << "   Number of degrees of freedom: " << dof_handler.n_dofs() << std::endl
            << std::endl;
      system_matrix_dev->assemble_rhs(system_rhs_dev);
      timed_solves<SolverCG<VectorType>>("pcg-standard", true);
      timed_solves<SolverCGFullMerge<VectorType>>("pcg-merged", false);
      // a partitioned operator application by hand, the way cell_loop brackets it (bp5/step-64.cu:272-275):
      // ghost values in, local cells, contributions back to their owners -- must equal vmult
      check_ghost_semantics();
      pcout << "  solution norm: " << system_matrix_dev->l2_norm(solution_dev) << std::endl << std::endl;
      solution_dev.reinit(0); system_rhs_dev.reinit(0);             // vectors go before their operator
      system_matrix_dev.reset();
    }
  }

 private:
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

  void check_ghost_semantics() {
    VectorType by_hand, by_vmult;
    by_hand.reinit(solution_dev); by_vmult.reinit(solution_dev);
    system_matrix_dev->do_zero_out = true;
    system_matrix_dev->vmult(by_vmult, solution_dev);
    solution_dev.update_ghost_values();
    by_hand = 0.;
    b200::check(bp5_operator_cell_loop(system_matrix_dev->handle(), by_hand.handle(), solution_dev.handle()));
    by_hand.compress(VectorOperation::add);
    b200::check(bp5_operator_copy_constrained_values(system_matrix_dev->handle(), by_hand.handle(), solution_dev.handle()));
    by_hand.mark_modified();
    by_hand.add(-1., by_vmult);
    pcout << "  ghost semantics: |cell_loop with update_ghost_values/compress - vmult| / |vmult| = "
          << by_hand.l2_norm() / by_vmult.l2_norm() << std::endl;
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

template <int degree> void run_degree(const Options &o, b200::Communicator &c) { PoissonProblem<3, degree> p(o, c); p.run(); }

static int rank_main(const Options &o, SharedPage *page, int rank) {
  try {
    ForkCommunicator comm(page, rank, (int)o.ranks);
    // cudaSetDevice(rank % n_devices), bp5/step-64.cu:704-707
    b200::Context::set_device(o.devices > 0 ? rank % o.devices : 0);
    if (rank == 0) std::cout << std::endl << "bp5_b200 info:" << std::endl << std::endl << "  " << bp5_version() << ", "
                             << o.ranks << " ranks" << std::endl << std::endl;
    switch (o.degree) {
      case 1: run_degree<1>(o, comm); break; case 2: run_degree<2>(o, comm); break; case 3: run_degree<3>(o, comm); break;
      case 4: run_degree<4>(o, comm); break; case 5: run_degree<5>(o, comm); break; case 6: run_degree<6>(o, comm); break;
      case 7: run_degree<7>(o, comm); break; case 8: run_degree<8>(o, comm); break;
      default: throw ExcMessage("degree must be 1..8");
    }
  } catch (std::exception &exc) {
    page->failed.store(1);
    std::cerr << std::endl << "----------------------------------------------------" << std::endl
              << "Exception on rank " << rank << ": " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  }
  return 0;
}

int main(int argc, char *argv[]) {
  Options o;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    const char *v = argv[i + 1];
    if (k == "--degree") o.degree = std::atoi(v);
    else if (k == "--cycle-min") o.cycle_min = std::atoi(v);
    else if (k == "--cycle-max") o.cycle_max = std::atoi(v);
    else if (k == "--iterations") o.n_iterations = std::atoi(v);
    else if (k == "--repetitions") o.n_repetitions = std::atoi(v);
    else if (k == "--ranks") o.ranks = std::atoi(v);
    else if (k == "--devices") o.devices = std::atoi(v);
    else if (k == "--quadrature") o.quadrature = std::strcmp(v, "gll") == 0 ? BP5_QUAD_GLL : BP5_QUAD_GAUSS;
    else { std::cerr << "unknown option " << k << std::endl; return 1; }
  }
  if (o.ranks < 1 || o.ranks > 64) { std::cerr << "--ranks must be 1..64" << std::endl; return 1; }
  void *mem = mmap(nullptr, sizeof(SharedPage), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (mem == MAP_FAILED) { std::perror("mmap"); return 1; }
  SharedPage *page = new (mem) SharedPage;
  page->arrived.store(0); page->generation.store(0); page->failed.store(0);
  // fork BEFORE anything touches CUDA: every rank creates its own context
  std::vector<pid_t> children;
  for (unsigned r = 1; r < o.ranks; ++r) {
    const pid_t pid = fork();
    if (pid < 0) { std::perror("fork"); return 1; }
    if (pid == 0) _exit(rank_main(o, page, (int)r));
    children.push_back(pid);
  }
  int rc = rank_main(o, page, 0);
  for (pid_t pid : children) {
    int status = 0;
    waitpid(pid, &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
  }
  return rc;
}
