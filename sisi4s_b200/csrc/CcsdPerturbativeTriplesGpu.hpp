#ifndef CCSD_PERTURBATIVE_TRIPLES_GPU_DEFINED
#define CCSD_PERTURBATIVE_TRIPLES_GPU_DEFINED

// Drop-in sisi4s algorithm class for the B200 (T) library.  Add this file and
// CcsdPerturbativeTriplesGpu.cxx to sisi4s_SOURCES (reference src/Makefile.am:12-100)
// and link libsisi4s_pt.so (+ nccl, cudart); see INTEGRATION.md.
//
// It registers under a NEW name because AlgorithmFactory silently overwrites duplicate
// registrations (reference src/algorithms/Algorithm.hpp:158-160); the argument keys are
// those of CcsdPerturbativeTriples (src/algorithms/CcsdPerturbativeTriples.cxx:32-79,240-247)
// and of PerturbativeTriples (src/algorithms/PerturbativeTriples.cxx:172-239).

#include <algorithms/Algorithm.hpp>

namespace sisi4s {
class CcsdPerturbativeTriplesGpu : public Algorithm {
public:
  ALGORITHM_REGISTRAR_DECLARATION(CcsdPerturbativeTriplesGpu);
  CcsdPerturbativeTriplesGpu(std::vector<Argument> const &argumentList);
  virtual ~CcsdPerturbativeTriplesGpu();
  /** gathers the CTF tensors once, runs the (T) loop on the GPU(s), sets the energy */
  virtual void run();
  /** device-memory estimate instead of the reference's DryTensor bookkeeping */
  virtual void dryRun();
};
} // namespace sisi4s

#endif
