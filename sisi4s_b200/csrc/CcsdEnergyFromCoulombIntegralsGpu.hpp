#ifndef CCSD_ENERGY_FROM_COULOMB_INTEGRALS_GPU_DEFINED
#define CCSD_ENERGY_FROM_COULOMB_INTEGRALS_GPU_DEFINED

// Drop-in sisi4s algorithm class for the device CCSD solver of libsisi4s_pt.so (include/sisi4s_ccsd.h).
// Same argument keys as CcsdEnergyFromCoulombIntegralsReference / ClusterSinglesDoublesAlgorithm
// (reference src/algorithms/CcsdEnergyFromCoulombIntegralsReference.cxx:49,136-140,
// ClusterSinglesDoublesAlgorithm.cxx:37-128); registered under a new name because AlgorithmFactory
// silently overwrites duplicate registrations (src/algorithms/Algorithm.hpp:158-160).

#include <algorithms/Algorithm.hpp>

namespace sisi4s {
class CcsdEnergyFromCoulombIntegralsGpu : public Algorithm {
public:
  ALGORITHM_REGISTRAR_DECLARATION(CcsdEnergyFromCoulombIntegralsGpu);
  CcsdEnergyFromCoulombIntegralsGpu(std::vector<Argument> const &argumentList);
  virtual ~CcsdEnergyFromCoulombIntegralsGpu();
  /** gathers the integral blocks (or the vertex) once, solves the amplitude equations on the GPU */
  virtual void run();
};
} // namespace sisi4s

#endif
