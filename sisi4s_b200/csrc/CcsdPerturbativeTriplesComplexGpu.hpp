#ifndef CCSD_PERTURBATIVE_TRIPLES_COMPLEX_GPU_DEFINED
#define CCSD_PERTURBATIVE_TRIPLES_COMPLEX_GPU_DEFINED

// Drop-in sisi4s algorithm class for the complex closed-shell (T) step of libsisi4s_pt.so
// (include/sisi4s_pt.h: pt_complex_triples).  Argument keys of CcsdPerturbativeTriplesComplex
// (reference src/algorithms/CcsdPerturbativeTriplesComplex.cxx:32-84); registered under a new name because
// AlgorithmFactory silently overwrites duplicate registrations (src/algorithms/Algorithm.hpp:158-160).

#include <algorithms/Algorithm.hpp>

namespace sisi4s {
class CcsdPerturbativeTriplesComplexGpu : public Algorithm {
public:
  ALGORITHM_REGISTRAR_DECLARATION(CcsdPerturbativeTriplesComplexGpu);
  CcsdPerturbativeTriplesComplexGpu(std::vector<Argument> const &argumentList);
  virtual ~CcsdPerturbativeTriplesComplexGpu();
  virtual void run();
};
} // namespace sisi4s

#endif
