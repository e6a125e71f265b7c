// UPerturbativeTriplesGpu.cxx -- sisi4s::Algorithm subclass in front of pt_spin_orbital_triples.
//
// Replaces UPerturbativeTriples::run (reference src/algorithms/UPerturbativeTriples.cxx:19-305).  The seven
// tensors are gathered once with Tensor::read_all and handed to the C ABI; rank 0's GPU evaluates the
// statements on the device tensor engine (csrc/upt.cu) and the scalar is broadcast.
// Written against the reference headers; run in tests/test_plugin_harness.py, see INTEGRATION.md section 4.
#include "UPerturbativeTriplesGpu.hpp"

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

ALGORITHM_REGISTRAR_DEFINITION(UPerturbativeTriplesGpu);

UPerturbativeTriplesGpu::UPerturbativeTriplesGpu(std::vector<Argument> const &argumentList)
    : Algorithm(argumentList) {}

UPerturbativeTriplesGpu::~UPerturbativeTriplesGpu() {}

namespace {

// the whole tensor on every rank, after checking its shape against the one the statements need
std::vector<double> gather(Tensor<double> *t, std::string const &name, std::vector<int64_t> const &shape) {
  bool ok(t->order == static_cast<int>(shape.size()));
  int64_t n(1);
  for (size_t d(0); ok && d < shape.size(); ++d) {
    ok = t->lens[d] == shape[d];
    n *= shape[d];
  }
  if (!ok) throw new EXCEPTION("UPerturbativeTriplesGpu: unexpected shape of " + name);
  std::vector<double> data(static_cast<size_t>(n));
  t->read_all(data.data());
  return data;
}

} // namespace

void UPerturbativeTriplesGpu::run() {
  Tensor<double> *epsi(getTensorArgument<double>("HoleEigenEnergies"));
  Tensor<double> *epsa(getTensorArgument<double>("ParticleEigenEnergies"));
  const int64_t No(epsi->lens[0]), Nv(epsa->lens[0]);
  CTF::World *world(epsi->wrld);

  const std::vector<double> ei(gather(epsi, "HoleEigenEnergies", {No}));
  const std::vector<double> ea(gather(epsa, "ParticleEigenEnergies", {Nv}));
  const std::vector<double> tai(gather(getTensorArgument<double>("CcsdSinglesAmplitudes"), "CcsdSinglesAmplitudes", {Nv, No}));
  const std::vector<double> tabij(
      gather(getTensorArgument<double>("CcsdDoublesAmplitudes"), "CcsdDoublesAmplitudes", {Nv, Nv, No, No}));
  const std::vector<double> vabij(
      gather(getTensorArgument<double>("PPHHCoulombIntegrals"), "PPHHCoulombIntegrals", {Nv, Nv, No, No}));
  const std::vector<double> vijka(
      gather(getTensorArgument<double>("HHHPCoulombIntegrals"), "HHHPCoulombIntegrals", {No, No, No, Nv}));
  const std::vector<double> vabci(
      gather(getTensorArgument<double>("PPPHCoulombIntegrals"), "PPPHCoulombIntegrals", {Nv, Nv, Nv, No}));

  double eTriples(0.0);
  int failed(0);
  std::string message;
  if (world->rank == 0) {
    if (pt_spin_orbital_triples(static_cast<int>(No), static_cast<int>(Nv), getIntegerArgument("device", 0), ei.data(),
                                ea.data(), tai.data(), tabij.data(), vabij.data(), vijka.data(), vabci.data(),
                                &eTriples) != PT_OK) {
      failed = 1;
      message = pt_last_error();
    }
  }
  double packet[2] = {eTriples, static_cast<double>(failed)};
  if (MPI_Bcast(packet, 2 * sizeof(double), MPI_BYTE, 0, world->comm) != MPI_SUCCESS)
    throw new EXCEPTION("UPerturbativeTriplesGpu: MPI_Bcast failed");
  if (packet[1] != 0.0) throw new EXCEPTION("pt_spin_orbital_triples: " + message);
  eTriples = packet[0];

  // the reference reports the triples energy alone (its CcsdEnergy sum is commented out, :299-303)
  LOG(0, "PerturbativeTriples") << "triples=" << eTriples << std::endl;
  setRealArgument("PerturbativeTriplesEnergy", eTriples);   // :305
}
