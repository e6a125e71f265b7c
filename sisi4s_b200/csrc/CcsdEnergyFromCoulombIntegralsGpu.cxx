// CcsdEnergyFromCoulombIntegralsGpu.cxx -- sisi4s::Algorithm subclass in front of the device CCSD solver.
//
// Replaces, for real closed-shell inputs, CcsdEnergyFromCoulombIntegralsReference driven by
// ClusterSinglesDoublesAlgorithm::run (reference src/algorithms/CcsdEnergyFromCoulombIntegralsReference.cxx:29-295,
// ClusterSinglesDoublesAlgorithm.cxx:37-128): the integral blocks are gathered ONCE with Tensor::read_all and
// handed to the C ABI (include/sisi4s_ccsd.h); the ~100 CTF statements per iteration run on one GPU.
// Rank 0's device does the work (the amplitude equations of the sizes sisi4s targets fit one B200); the
// converged amplitudes are written back into CTF tensors by rank 0 (Tensor::write is collective: the other
// ranks contribute zero elements).
//
// Written against the reference headers; run in tests/test_plugin_harness.py (one rank, stand-ins for MPI / CTF),
// see INTEGRATION.md section 4.
#include "CcsdEnergyFromCoulombIntegralsGpu.hpp"

#include <Sisi4s.hpp>
#include <util/Exception.hpp>
#include <util/Log.hpp>
#include <util/Tensor.hpp>

#include <mpi.h>

#include <cstdint>
#include <string>
#include <vector>

#include <sisi4s_ccsd.h>
#include <sisi4s_tn.h>

using namespace sisi4s;

ALGORITHM_REGISTRAR_DEFINITION(CcsdEnergyFromCoulombIntegralsGpu);

CcsdEnergyFromCoulombIntegralsGpu::CcsdEnergyFromCoulombIntegralsGpu(std::vector<Argument> const &argumentList)
    : Algorithm(argumentList) {}

CcsdEnergyFromCoulombIntegralsGpu::~CcsdEnergyFromCoulombIntegralsGpu() {}

namespace {

#define CCSD_CHECK(call)                                                                 \
  do {                                                                                   \
    if ((call) != 0) throw new EXCEPTION(std::string(#call ": ") + tn_last_error());     \
  } while (0)

struct CcsdGuard {
  ccsd_handle_t h = nullptr;
  ~CcsdGuard() { if (h) ccsd_destroy(h); }
};

std::vector<double> gatherDense(Tensor<double> *t) {
  int64_t n = 1;
  for (int d = 0; d < t->order; ++d) n *= t->lens[d];
  std::vector<double> dense(static_cast<size_t>(n));
  t->read_all(dense.data());   // collective
  return dense;
}

// dense column-major host array -> new CTF tensor; rank 0 writes all elements
Tensor<double> *scatterDense(std::vector<double> const &dense, std::vector<int64_t> const &lens, CTF::World *world,
                             char const *name) {
  std::vector<int> syms(lens.size(), NS);
  Tensor<double> *t(new Tensor<double>(static_cast<int>(lens.size()), lens.data(), syms.data(), *world, name));
  std::vector<int64_t> idx;
  if (world->rank == 0) {
    idx.resize(dense.size());
    for (size_t q(0); q < dense.size(); ++q) idx[q] = static_cast<int64_t>(q);
  }
  t->write(static_cast<int64_t>(idx.size()), idx.data(), dense.data());
  return t;
}

} // namespace

void CcsdEnergyFromCoulombIntegralsGpu::run() {
  Tensor<double> *epsi(getTensorArgument("HoleEigenEnergies"));
  Tensor<double> *epsa(getTensorArgument("ParticleEigenEnergies"));
  const int No(epsi->lens[0]), Nv(epsa->lens[0]);
  CTF::World *world(epsi->wrld);
  const bool root(world->rank == 0);

  // ClusterSinglesDoublesAlgorithm::run (:47-63): mixer and convergence arguments, reference defaults
  CcsdOptions opt;
  ccsd_default_options(&opt);
  const std::string mixerName(getTextArgument("mixer", "LinearMixer"));
  if (mixerName == "LinearMixer") opt.mixer = CCSD_LINEAR_MIXER;
  else if (mixerName == "DiisMixer") opt.mixer = CCSD_DIIS_MIXER;
  else throw new EXCEPTION("Mixer not implemented: " + mixerName);
  opt.max_residua = getIntegerArgument("maxResidua", opt.max_residua);
  opt.mixing_ratio = getRealArgument("mixingRatio", opt.mixing_ratio);
  opt.max_iterations = getIntegerArgument("maxIterations", opt.max_iterations);
  opt.energy_convergence = getRealArgument("energyConvergence", opt.energy_convergence);
  opt.amplitudes_convergence = getRealArgument("amplitudesConvergence", opt.amplitudes_convergence);
  opt.level_shift = getRealArgument("levelShift", opt.level_shift);

  // gather (collective on all ranks); only rank 0 talks to the GPU
  std::vector<double> ei(gatherDense(epsi)), ea(gatherDense(epsa));
  CcsdGuard guard;
  if (root) {
    CCSD_CHECK(ccsd_create(&guard.h, No, Nv, getIntegerArgument("device", 0)));
    CCSD_CHECK(ccsd_set_eigenenergies(guard.h, ei.data(), ea.data()));
  }
  static char const *blocks[] = {"PPHH", "PHPH", "HHHH", "HHHP", "PPPH", "PPPP"};
  bool needVertex(false);
  for (char const *b : blocks) needVertex = needVertex || !isArgumentGiven(std::string(b) + "CoulombIntegrals");
  if (needVertex) {
    // blocks the plan does not pass are built on the device from the vertex
    // (CoulombIntegralsFromVertex.cxx:395-431), Re / Im split as fromComplexTensor does
    Tensor<complex> *GammaFqr(getTensorArgument<complex>("CoulombVertex"));
    const int NF(GammaFqr->lens[0]), Np(GammaFqr->lens[1]);
    const int64_t n(static_cast<int64_t>(NF) * Np * Np);
    std::vector<complex> g(static_cast<size_t>(n));
    GammaFqr->read_all(g.data());
    if (root) {
      std::vector<double> re(static_cast<size_t>(n)), im(static_cast<size_t>(n));
      for (int64_t q(0); q < n; ++q) {
        re[q] = std::real(g[q]);
        im[q] = std::imag(g[q]);
      }
      CCSD_CHECK(ccsd_set_vertex(guard.h, NF, Np, re.data(), im.data()));
    }
  }
  for (char const *b : blocks) {
    const std::string key(std::string(b) + "CoulombIntegrals");
    if (!isArgumentGiven(key)) continue;
    std::vector<double> dense(gatherDense(getTensorArgument(key)));
    if (root) CCSD_CHECK(ccsd_set_integrals(guard.h, b, dense.data()));
  }
  // createAmplitudes (:207-237): optional initial amplitudes
  if (isArgumentGiven("initialSinglesAmplitudes") || isArgumentGiven("initialDoublesAmplitudes")) {
    std::vector<double> t1, t2;
    if (isArgumentGiven("initialSinglesAmplitudes")) t1 = gatherDense(getTensorArgument("initialSinglesAmplitudes"));
    if (isArgumentGiven("initialDoublesAmplitudes")) t2 = gatherDense(getTensorArgument("initialDoublesAmplitudes"));
    if (root) CCSD_CHECK(ccsd_set_amplitudes(guard.h, t1.empty() ? nullptr : t1.data(), t2.empty() ? nullptr : t2.data()));
  }

  CcsdResult res = {};
  std::vector<double> t1(static_cast<size_t>(Nv) * No), t2(static_cast<size_t>(Nv) * Nv * No * No);
  if (root) {
    CCSD_CHECK(ccsd_solve(guard.h, &opt, &res));
    CCSD_CHECK(ccsd_get_amplitudes(guard.h, t1.data(), t2.data()));
  }
  // every rank needs the scalars (setRealArgument runs on all ranks)
  double scalars[4] = {res.energy, res.direct, res.exchange, static_cast<double>(res.iterations * 2 + res.converged)};
  if (MPI_Bcast(scalars, 4 * sizeof(double), MPI_BYTE, 0, world->comm) != MPI_SUCCESS)
    throw new EXCEPTION("CcsdEnergyFromCoulombIntegralsGpu: MPI_Bcast failed");
  const double e(scalars[0]);
  const int iterations(static_cast<int>(scalars[3]) / 2);
  const bool converged(static_cast<int>(scalars[3]) % 2 == 1);

  LOG(1, "CcsdGpu") << "iterations=" << iterations << std::endl;
  LOG(1, "CcsdGpu") << "dir= " << scalars[1] << std::endl;
  LOG(1, "CcsdGpu") << "exc= " << scalars[2] << std::endl;
  LOG(0, "CcsdGpu") << "energy= " << e << std::endl;
  if (!converged && opt.max_iterations > 0)
    LOG(0, "CcsdGpu") << "WARNING: energy or amplitudes convergence not reached." << std::endl;   // :120-124

  // storeAmplitudes (:289-300) and the energy (:34)
  if (isArgumentGiven("CcsdSinglesAmplitudes"))
    allocatedTensorArgument<double>("CcsdSinglesAmplitudes", scatterDense(t1, {Nv, No}, world, "Tai"));
  if (isArgumentGiven("CcsdDoublesAmplitudes"))
    allocatedTensorArgument<double>("CcsdDoublesAmplitudes", scatterDense(t2, {Nv, Nv, No, No}, world, "Tabij"));
  setRealArgument("CcsdEnergy", e);
}
