// CcsdPerturbativeTriplesComplexGpu.cxx -- sisi4s::Algorithm subclass in front of pt_complex_triples.
//
// Replaces CcsdPerturbativeTriplesComplex::run (reference
// src/algorithms/CcsdPerturbativeTriplesComplex.cxx:32-84, Calculator<F>::calculate :166-271) for both of
// its instantiations: real amplitudes / integrals (the imaginary parts are then zero) and complex ones.
// The tensors are gathered once with Tensor::read_all, split into real and imaginary parts
// (fromComplexTensor) and handed to the C ABI; rank 0's GPU does the work and the scalar is broadcast.
// Written against the reference headers; run in tests/test_plugin_harness.py, see INTEGRATION.md section 4.
#include "CcsdPerturbativeTriplesComplexGpu.hpp"

#include <Sisi4s.hpp>
#include <util/Exception.hpp>
#include <util/Log.hpp>
#include <util/Tensor.hpp>

#include <mpi.h>

#include <cstdint>
#include <string>
#include <vector>

#include <sisi4s_pt.h>

using namespace sisi4s;

ALGORITHM_REGISTRAR_DEFINITION(CcsdPerturbativeTriplesComplexGpu);

CcsdPerturbativeTriplesComplexGpu::CcsdPerturbativeTriplesComplexGpu(std::vector<Argument> const &argumentList)
    : Algorithm(argumentList) {}

CcsdPerturbativeTriplesComplexGpu::~CcsdPerturbativeTriplesComplexGpu() {}

namespace {

int64_t elements(int order, int64_t const *lens) {
  int64_t n = 1;
  for (int d = 0; d < order; ++d) n *= lens[d];
  return n;
}

// real and imaginary parts of a tensor argument that is stored either as real or as complex data
void gatherParts(Algorithm *alg, std::string const &name, bool isComplex, std::vector<double> &re, std::vector<double> &im) {
  if (isComplex) {
    Tensor<complex> *t(alg->getTensorArgument<complex>(name));
    const int64_t n(elements(t->order, t->lens));
    std::vector<complex> z(static_cast<size_t>(n));
    t->read_all(z.data());
    re.resize(n);
    im.resize(n);
    for (int64_t q(0); q < n; ++q) {
      re[q] = std::real(z[q]);
      im[q] = std::imag(z[q]);
    }
  } else {
    Tensor<double> *t(alg->getTensorArgument<double>(name));
    const int64_t n(elements(t->order, t->lens));
    re.resize(n);
    im.assign(n, 0.0);
    t->read_all(re.data());
  }
}

} // namespace

void CcsdPerturbativeTriplesComplexGpu::run() {
  Tensor<double> *epsi(getTensorArgument<double>("HoleEigenEnergies"));
  Tensor<double> *epsa(getTensorArgument<double>("ParticleEigenEnergies"));
  const int No(epsi->lens[0]), Nv(epsa->lens[0]);
  CTF::World *world(epsi->wrld);
  const double eCcsd(getRealArgument("CcsdEnergy"));   // mandatory (:78)

  // real or complex amplitudes / integrals: decided by the type of PPHHCoulombIntegrals, as the reference does (:52-53)
  const bool isComplex(dynamic_cast<TensorData<double> *>(getArgumentData("PPHHCoulombIntegrals")) == nullptr);

  std::vector<double> ei(No), ea(Nv);
  epsi->read_all(ei.data());
  epsa->read_all(ea.data());
  std::vector<double> t1r, t1i, t2r, t2i, pr, pi, ur, ui, gr, gi;
  gatherParts(this, "CcsdSinglesAmplitudes", isComplex, t1r, t1i);
  gatherParts(this, "CcsdDoublesAmplitudes", isComplex, t2r, t2i);
  gatherParts(this, "PPHHCoulombIntegrals", isComplex, pr, pi);
  gatherParts(this, "PHHHCoulombIntegrals", isComplex, ur, ui);
  Tensor<complex> *GammaFqr(getTensorArgument<complex>("CoulombVertex"));
  const int NF(GammaFqr->lens[0]), Np(GammaFqr->lens[1]);
  gatherParts(this, "CoulombVertex", true, gr, gi);

  double eTriples(0.0);
  int failed(0);
  std::string message;
  if (world->rank == 0) {
    if (pt_complex_triples(No, Nv, getIntegerArgument("device", 0), ei.data(), ea.data(), t1r.data(), t1i.data(),
                           t2r.data(), t2i.data(), pr.data(), pi.data(), ur.data(), ui.data(), NF, Np, gr.data(),
                           gi.data(), &eTriples, nullptr) != PT_OK) {
      failed = 1;
      message = pt_last_error();
    }
  }
  double packet[2] = {eTriples, static_cast<double>(failed)};
  if (MPI_Bcast(packet, 2 * sizeof(double), MPI_BYTE, 0, world->comm) != MPI_SUCCESS)
    throw new EXCEPTION("CcsdPerturbativeTriplesComplexGpu: MPI_Bcast failed");
  if (packet[1] != 0.0) throw new EXCEPTION("pt_complex_triples: " + message);
  eTriples = packet[0];

  LOG(1, "CcsdPerturbativeTriplesComplexGpu") << "triples=" << eTriples << std::endl;
  LOG(1, "CcsdPerturbativeTriplesComplexGpu") << "ccsd=" << eCcsd << std::endl;
  const double e(eCcsd + eTriples);
  LOG(0, "CcsdPerturbativeTriplesComplexGpu") << "e=" << e << std::endl;
  setRealArgument("CcsdPerturbativeTriplesComplexEnergy", e);   // :82
}
