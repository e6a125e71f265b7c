#ifndef U_PERTURBATIVE_TRIPLES_GPU_DEFINED
#define U_PERTURBATIVE_TRIPLES_GPU_DEFINED

// Drop-in sisi4s algorithm class for the spin-orbital (T) step of libsisi4s_pt.so
// (include/sisi4s_pt.h: pt_spin_orbital_triples).  Argument keys of UPerturbativeTriples
// (reference src/algorithms/UPerturbativeTriples.cxx:19-27,305); registered under a new name because
// AlgorithmFactory silently overwrites duplicate registrations (src/algorithms/Algorithm.hpp:158-160).

//
// Plan arguments (all real tensors, spin-orbital, antisymmetrised where the reference expects it):
//   HoleEigenEnergies[o], ParticleEigenEnergies[v], CcsdSinglesAmplitudes[v,o], CcsdDoublesAmplitudes[v,v,o,o],
//   PPHHCoulombIntegrals[v,v,o,o], HHHPCoulombIntegrals[o,o,o,v], PPPHCoulombIntegrals[v,v,v,o];
//   device (integer, optional, default 0): the GPU rank 0 evaluates the statements on.
// Output: PerturbativeTriplesEnergy = the triples energy alone.
// Memory: three v^3 o^3 FP64 tensors on the device (the reference holds the same three in CTF).

#include <algorithms/Algorithm.hpp>

namespace sisi4s {
class UPerturbativeTriplesGpu : public Algorithm {
public:
  ALGORITHM_REGISTRAR_DECLARATION(UPerturbativeTriplesGpu);
  UPerturbativeTriplesGpu(std::vector<Argument> const &argumentList);
  virtual ~UPerturbativeTriplesGpu();
  virtual void run();
};
} // namespace sisi4s

#endif
