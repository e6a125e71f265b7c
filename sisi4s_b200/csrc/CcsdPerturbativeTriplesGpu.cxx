// CcsdPerturbativeTriplesGpu.cxx -- sisi4s::Algorithm subclass in front of libsisi4s_pt.
//
// Replaces CcsdPerturbativeTriples::run (reference
// src/algorithms/CcsdPerturbativeTriples.cxx:119-248): instead of slicing the CTF tensors
// into O(o^2) distributed sub-tensors and issuing ~150 collective CTF operations per sorted
// hole triple, every rank gathers the inputs ONCE with Tensor::read_all (dense,
// column-major, same call as reference ParenthesisTriples.cxx:794-797), hands them to
// the C ABI (include/sisi4s_pt.h) and runs its share of the i<=j<=k triples on its GPU.
// The only communication on the path is one ncclAllReduce of the scalar energy.
//
// Written against the reference headers; it cannot be linked in the build container of this
// repository (no MPI / Cyclops CTF there), see INTEGRATION.md for how it is compile-checked.
#include "CcsdPerturbativeTriplesGpu.hpp"

#include <DryTensor.hpp>
#include <Sisi4s.hpp>
#include <util/Exception.hpp>
#include <util/Log.hpp>
#include <util/Tensor.hpp>

#include <cuda_runtime.h>
#include <mpi.h>
#include <nccl.h>

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include <sisi4s_pt.h>

using namespace sisi4s;

ALGORITHM_REGISTRAR_DEFINITION(CcsdPerturbativeTriplesGpu);

CcsdPerturbativeTriplesGpu::CcsdPerturbativeTriplesGpu(std::vector<Argument> const &argumentList)
    : Algorithm(argumentList) {}

CcsdPerturbativeTriplesGpu::~CcsdPerturbativeTriplesGpu() {}

namespace {

// errors of the C layer surface the reference's way: throw new EXCEPTION(msg)
#define PT_CHECK(call)                                                              \
  do {                                                                              \
    if ((call) != PT_OK) throw new EXCEPTION(std::string(#call ": ") + pt_last_error()); \
  } while (0)

// dense column-major copy of a CTF tensor; collective, executed by all ranks
std::vector<double> gather(Tensor<double> *t) {
  int64_t n = 1;
  for (int d = 0; d < t->order; ++d) n *= t->lens[d];
  std::vector<double> dense(static_cast<size_t>(n));
  t->read_all(dense.data());
  return dense;
}

void expectShape(Tensor<double> *t, std::vector<int64_t> const &lens, std::string const &name) {
  bool ok = t->order == static_cast<int>(lens.size());
  for (size_t d = 0; ok && d < lens.size(); ++d) ok = t->lens[d] == lens[d];
  if (!ok) throw new EXCEPTION("Incompatible shape of argument: " + name);
}

} // namespace

void CcsdPerturbativeTriplesGpu::run() {
  Tensor<double> *epsi(getTensorArgument("HoleEigenEnergies"));
  Tensor<double> *epsa(getTensorArgument("ParticleEigenEnergies"));
  const int No(epsi->lens[0]);
  const int Nv(epsa->lens[0]);
  CTF::World *world(epsi->wrld);
  const int rank(world->rank), np(world->np);

  // one rank <-> one GPU of the node (ranks beyond the GPU count share devices round-robin)
  int deviceCount(0);
  if (cudaGetDeviceCount(&deviceCount) != cudaSuccess || deviceCount == 0)
    throw new EXCEPTION("CcsdPerturbativeTriplesGpu: no CUDA device (there is no CPU fallback)");
  const int device(getIntegerArgument("device", rank % deviceCount));

  pt_handle_t h(nullptr);
  PT_CHECK(pt_create(&h, No, Nv, device));
  // optional: hole-blocked residency of V_abci for shapes whose v^3 o tensor exceeds the GPU's
  // memory (slabSlots >= 3 slabs resident; the rest is rebuilt from the vertex / re-uploaded)
  const int64_t slabSlots(getIntegerArgument("slabSlots", 0));
  if (slabSlots > 0) PT_CHECK(pt_set_option(h, "slab_slots", slabSlots));
  std::vector<double> hostPpph; // must outlive pt_run in the blocked PPPH contract

  {
    std::vector<double> ei(gather(epsi)), ea(gather(epsa));
    PT_CHECK(pt_set_eigenenergies(h, ei.data(), ea.data()));
  }
  {
    Tensor<double> *Tai(getTensorArgument("CcsdSinglesAmplitudes"));
    expectShape(Tai, {Nv, No}, "CcsdSinglesAmplitudes");
    std::vector<double> t1(gather(Tai));
    PT_CHECK(pt_set_singles(h, t1.data()));
  }
  {
    Tensor<double> *Tabij(getTensorArgument("CcsdDoublesAmplitudes"));
    expectShape(Tabij, {Nv, Nv, No, No}, "CcsdDoublesAmplitudes");
    std::vector<double> t2(gather(Tabij));
    PT_CHECK(pt_set_doubles(h, t2.data()));
  }
  {
    Tensor<double> *Vabij(getTensorArgument("PPHHCoulombIntegrals"));
    expectShape(Vabij, {Nv, Nv, No, No}, "PPHHCoulombIntegrals");
    std::vector<double> v(gather(Vabij));
    PT_CHECK(pt_set_pphh(h, v.data()));
  }
  {
    Tensor<double> *Vijka(getTensorArgument("HHHPCoulombIntegrals"));
    expectShape(Vijka, {No, No, No, Nv}, "HHHPCoulombIntegrals");
    std::vector<double> v(gather(Vijka));
    PT_CHECK(pt_set_hhhp(h, v.data()));
  }
  if (isArgumentGiven("PPPHCoulombIntegrals")) {
    // contract of PerturbativeTriples (PerturbativeTriples.cxx:176): one hole slab at a time,
    // so that no rank ever holds more than v^3 doubles of V_abci on the host
    Tensor<double> *Vabci(getTensorArgument("PPPHCoulombIntegrals"));
    expectShape(Vabci, {Nv, Nv, Nv, No}, "PPPHCoulombIntegrals");
    if (slabSlots > 0 && slabSlots < No) {
      hostPpph = gather(Vabci);
      PT_CHECK(pt_set_ppph_host(h, hostPpph.data()));
    } else
    for (int k(0); k < No; ++k) {
      int start[] = {0, 0, 0, k}, end[] = {Nv, Nv, Nv, k + 1};
      Tensor<double> slab(Vabci->slice(start, end));
      std::vector<double> v(gather(&slab));
      PT_CHECK(pt_set_ppph_slabs(h, k, k + 1, v.data()));
    }
  } else {
    // contract of the compiled CcsdPerturbativeTriples (:48-78): Coulomb vertex, Re/Im split;
    // V_abci is then built on the device as CoulombIntegralsFromVertex.cxx:430-431
    Tensor<complex> *GammaFqr(getTensorArgument<complex>("CoulombVertex"));
    const int NF(GammaFqr->lens[0]), Np(GammaFqr->lens[1]);
    const int64_t n(static_cast<int64_t>(NF) * Np * Np);
    std::vector<complex> g(static_cast<size_t>(n));
    GammaFqr->read_all(g.data());
    std::vector<double> re(static_cast<size_t>(n)), im(static_cast<size_t>(n));
    for (int64_t q(0); q < n; ++q) {
      re[q] = std::real(g[q]);
      im[q] = std::imag(g[q]);
    }
    PT_CHECK(pt_set_vertex(h, NF, Np, re.data(), im.data()));
  }

  // this rank's share of the sorted triples (reference loop order, :156-158)
  int64_t begin(0), end(0);
  PT_CHECK(pt_partition(No, np, rank, &begin, &end));
  double eLocal(0.0);
  PT_CHECK(pt_run(h, begin, end, &eLocal, nullptr));

  // the single collective of the path: all-reduce of the scalar energy over NVLink
  double eTriples(eLocal);
  if (np > 1) {
    ncclUniqueId id;
    if (rank == 0) ncclGetUniqueId(&id);
    MPI_Bcast(&id, sizeof(id), MPI_BYTE, 0, world->comm);
    ncclComm_t comm;
    if (ncclCommInitRank(&comm, np, id, rank) != ncclSuccess)
      throw new EXCEPTION("CcsdPerturbativeTriplesGpu: ncclCommInitRank failed");
    double *dE(nullptr);
    cudaSetDevice(device);
    cudaMalloc(&dE, sizeof(double));
    cudaMemcpy(dE, &eLocal, sizeof(double), cudaMemcpyHostToDevice);
    ncclAllReduce(dE, dE, 1, ncclDouble, ncclSum, comm, 0);
    cudaStreamSynchronize(0);
    cudaMemcpy(&eTriples, dE, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dE);
    ncclCommDestroy(comm);
  }

  PtStats stats;
  PT_CHECK(pt_get_stats(h, &stats));
  PT_CHECK(pt_destroy(h));

  double eCcsd(getRealArgument("CcsdEnergy", 0.0));
  double e(eCcsd + eTriples);
  LOG(0, "CcsdPerturbativeTriplesGpu") << "e=" << e << std::endl;
  LOG(1, "CcsdPerturbativeTriplesGpu") << "ccsd=" << eCcsd << std::endl;
  LOG(1, "CcsdPerturbativeTriplesGpu") << "triples=" << eTriples << std::endl;
  LOG(1, "CcsdPerturbativeTriplesGpu")
      << "device seconds=" << stats.seconds_run << ", TFLOP/s (rank 0 share)="
      << stats.flops_algorithmic / stats.seconds_run * 1e-12 << std::endl;

  // whichever spelling of the output the plan asks for
  bool any(false);
  if (isArgumentGiven("CcsdPerturbativeTriplesEnergy")) {
    setRealArgument("CcsdPerturbativeTriplesEnergy", e);
    any = true;
  }
  if (isArgumentGiven("PerturbativeTriplesEnergy")) {
    setRealArgument("PerturbativeTriplesEnergy", e);
    any = true;
  }
  if (!any) throw new EXCEPTION("Missing argument: CcsdPerturbativeTriplesEnergy");
}

void CcsdPerturbativeTriplesGpu::dryRun() {
  DryTensor<> *epsi(getTensorArgument<double, DryTensor<double>>("HoleEigenEnergies"));
  DryTensor<> *epsa(getTensorArgument<double, DryTensor<double>>("ParticleEigenEnergies"));
  const double No(epsi->lens[0]), Nv(epsa->lens[0]);
  const double nr(std::ceil(Nv / 16.0));
  // packed PPPH + two packed copies of T2 + PPHH and its pre-added pair sums + one staging slab,
  // all FP64, per GPU
  const double bytes(8.0 * (No * nr * nr * std::ceil(Nv / 4.0) * 1024.0 + 4.0 * Nv * Nv * No * No
                            + Nv * Nv * Nv));
  LOG(0, "CcsdPerturbativeTriplesGpu") << "device memory per GPU=" << bytes / 1e9 << " GB" << std::endl;
}
