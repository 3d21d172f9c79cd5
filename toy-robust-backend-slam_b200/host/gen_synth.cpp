// gen_synth — writes a synthetic Manhattan-world .g2o the reader consumes (see synth.h).
//   ./gen_synth N_POSES N_LOOPS OUT.g2o [SEED]
#include <cstdlib>
#include <iostream>
#include "synth.h"

int main(int argc, char** argv) {
  if (argc < 4) { std::cout << "Usage: " << argv[0] << " N_POSES N_LOOPS OUT.g2o [SEED=20260101]\n"; return -1; }
  ReadG2O g;
  const uint64_t seed = argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 20260101ull;
  const long long made = synth::generate_manhattan(std::atoi(argv[1]), std::atoll(argv[2]), seed, &g);
  if (made < 0 || !synth::write_g2o(g, argv[3])) return 1;
  std::cout << "poses " << g.nNodes.size() << " odometry " << g.nEdgesOdometry.size() << " loops " << made << "\n";
  return 0;
}
