// oracle/trg_driver.cc -- TEST INFRASTRUCTURE ONLY.
// BASELINE configs[0]: the reference's own examples/z2_ising_trg.cpp, compiled from where it
// lies (nothing is copied; its main() is renamed away), run at a chosen bond dimension instead of the hard-coded 128.
// Prints, per beta, the converged free energy per site and the time the example itself attributes to Contract and SVD.
#define main z2_ising_trg_example_main
#include "z2_ising_trg.cpp"
#undef main

#include <cstdlib>

int main(int argc, char **argv) {
  const size_t chi = argc > 1 ? (size_t) std::atoi(argv[1]) : 32;
  const int threads = argc > 2 ? std::atoi(argv[2]) : 4;
  hp_numeric::SetTensorManipulationThreads(threads);
  TRGParams params;
  params.max_iterations = 30;
  params.truncation_error = 1e-14;
  params.min_bond_dim = 4;
  params.max_bond_dim = chi;
  params.convergence_threshold = 1e-15;
  params.verbose = argc > 3 && std::atoi(argv[3]) != 0;     // per-scale lines with the example's own SVD / Contract timers
  std::cout << "# z2_ising_trg (reference example), chi = " << chi << ", threads = " << threads << "\n";
  std::cout << "# beta free_energy_per_site iterations final_chi wall_s\n";
  for (double beta : {0.2, 0.4, 0.7, 1.0}) {
    Timer t("trg");
    TRGIterationResult r = RunTRG(beta, 1.0, params);
    std::cout << std::setprecision(15) << beta << " " << r.free_energy << " " << r.scale << " " << r.bond_dim << " "
              << std::setprecision(4) << t.Elapsed() << "\n";
  }
  return 0;
}
