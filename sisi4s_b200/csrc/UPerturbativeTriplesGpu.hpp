#ifndef U_PERTURBATIVE_TRIPLES_GPU_DEFINED
#define U_PERTURBATIVE_TRIPLES_GPU_DEFINED

// Drop-in sisi4s algorithm class for the spin-orbital (T) step of libsisi4s_pt.so
// (include/sisi4s_pt.h: pt_spin_orbital_triples).  Argument keys of UPerturbativeTriples
// (reference src/algorithms/UPerturbativeTriples.cxx:19-27,305); registered under a new name because
// AlgorithmFactory silently overwrites duplicate registrations (src/algorithms/Algorithm.hpp:158-160).

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
