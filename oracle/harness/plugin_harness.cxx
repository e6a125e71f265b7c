// plugin_harness.cxx -- runs a sisi4s Algorithm subclass of this repository through the REFERENCE's own
// argument machinery, on one rank.  TEST INFRASTRUCTURE (tests/test_plugin_harness.py), never shipped.
//
// Linked from: the reference's src/algorithms/Algorithm.cxx, src/Data.cxx, src/DryTensor.cxx, src/util/Log.cxx,
// src/util/Emitter.cxx compiled where they lie (oracle/Makefile, objects under oracle/_ref/), the plugin classes
// sisi4s_b200/csrc/*Gpu.cxx, libsisi4s_pt.so, and single-process stand-ins for the three absent third-party
// headers (oracle/harness/stubs: <ctf.hpp>, <mpi.h>, <yaml-cpp/yaml.h>).  What it checks is the part of the
// drop-in the syntax check cannot: AlgorithmFactory registration, getTensorArgument / getRealArgument /
// setRealArgument with real Data objects, Tensor::read_all order, shapes, option handling, error propagation.
//
//   plugin_harness <AlgorithmName> <plan.txt> [--dry]
// plan.txt, one entry per line:
//   tensor  <Argument> <file> <len0> <len1> ...     raw FP64, column-major
//   ctensor <Argument> <file> <len0> <len1> ...     raw complex128 (re, im interleaved), column-major
//   real    <Argument> <value>
//   integer <Argument> <value>
//   text    <Argument> <value>
//   out     <Argument>                              real output; printed as "<Argument> = <%.17g>"
//   tout    <Argument> <file>                       real tensor output, written raw (column-major) to <file>
#include <Data.hpp>
#include <DryTensor.hpp>
#include <Sisi4s.hpp>
#include <algorithms/Algorithm.hpp>
#include <util/Emitter.hpp>
#include <util/Exception.hpp>
#include <util/Log.hpp>
#include <util/Tensor.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace sisi4s;

CTF::World *Sisi4s::world = nullptr;
Options *Sisi4s::options = nullptr;

namespace {

template <typename F>
Tensor<F> *loadTensor(std::string const &file, std::vector<int64_t> const &lens, CTF::World &world, std::string const &name) {
  std::vector<int> sym(lens.size() ? lens.size() : 1, NS);
  Tensor<F> *t = new Tensor<F>(static_cast<int>(lens.size()), lens.data(), sym.data(), world, name.c_str());
  int64_t n = 1;
  for (int64_t l : lens) n *= l;
  std::vector<F> buf(static_cast<size_t>(n));
  std::ifstream in(file, std::ios::binary);
  if (!in.read(reinterpret_cast<char *>(buf.data()), sizeof(F) * buf.size())) throw new EXCEPTION("cannot read " + file);
  std::vector<int64_t> idx(static_cast<size_t>(n));
  for (int64_t q = 0; q < n; ++q) idx[q] = q;
  t->write(n, idx.data(), buf.data());   // CTF's global-index write: column-major positions
  return t;
}

} // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <AlgorithmName> <plan.txt> [--dry]\n", argv[0]);
    return 2;
  }
  const std::string algorithmName(argv[1]);
  const bool dry(argc > 3 && std::string(argv[3]) == "--dry");
  CTF::World world(argc, argv);
  Sisi4s::world = &world;
  Log::setRank(world.rank);
  Log::setFileName("/dev/null");
  Log::setLogLevel(1);
  Emitter::setRank(world.rank);
  Emitter::setFileName("/dev/null");
  try {
    std::vector<Argument> arguments;
    std::vector<std::string> outputs;
    std::vector<std::pair<std::string, std::string>> tensorOutputs;
    std::ifstream plan(argv[2]);
    if (!plan) throw new EXCEPTION(std::string("cannot open ") + argv[2]);
    std::string line;
    while (std::getline(plan, line)) {
      std::istringstream s(line);
      std::string kind, name;
      if (!(s >> kind >> name)) continue;
      const std::string dataName(name + "Data");
      if (kind == "tensor" || kind == "ctensor") {
        std::string file;
        s >> file;
        std::vector<int64_t> lens;
        for (int64_t l; s >> l;) lens.push_back(l);
        if (dry) {
          std::vector<int> l32(lens.begin(), lens.end()), sym(lens.size(), NS);
          if (kind == "tensor")
            new TensorData<double, DryTensor<double>>(dataName, new DryTensor<double>(static_cast<int>(lens.size()), l32.data(), sym.data(), SOURCE_LOCATION));
          else
            new TensorData<complex, DryTensor<complex>>(dataName, new DryTensor<complex>(static_cast<int>(lens.size()), l32.data(), sym.data(), SOURCE_LOCATION));
        } else if (kind == "tensor") {
          new TensorData<double>(dataName, loadTensor<double>(file, lens, world, name));
        } else {
          new TensorData<complex>(dataName, loadTensor<complex>(file, lens, world, name));
        }
      } else if (kind == "real") {
        double v;
        s >> v;
        new RealData(dataName, v);
      } else if (kind == "integer") {
        int64_t v;
        s >> v;
        new IntegerData(dataName, v);
      } else if (kind == "text") {
        std::string v;
        s >> v;
        new TextData(dataName, v);
      } else if (kind == "tout") {
        std::string file;
        s >> file;
        tensorOutputs.push_back({name, file});
        new Data(dataName);   // a "mentioned" symbol, as the reference's parser creates for outputs
      } else if (kind == "out") {
        outputs.push_back(name);
        new Data(dataName);
      } else {
        throw new EXCEPTION("unknown plan entry: " + kind);
      }
      arguments.push_back(Argument(name, dataName));
    }
    Algorithm *algorithm(AlgorithmFactory::create(algorithmName, arguments));
    if (!algorithm) throw new EXCEPTION("unknown algorithm: " + algorithmName);
    if (dry) algorithm->dryRun();
    else algorithm->run();
    for (std::string const &name : outputs) {
      RealData *r(dynamic_cast<RealData *>(Data::get(name + "Data")));
      if (r) std::printf("%s = %.17g\n", name.c_str(), r->value);
      else if (!dry) throw new EXCEPTION("output was not set: " + name);
    }
    for (auto const &out : tensorOutputs) {
      TensorData<double> *t(dynamic_cast<TensorData<double> *>(Data::get(out.first + "Data")));
      if (!t) {
        if (dry) continue;
        throw new EXCEPTION("tensor output was not set: " + out.first);
      }
      int64_t n(1);
      for (int d(0); d < t->value->order; ++d) n *= t->value->lens[d];
      std::vector<double> dense(static_cast<size_t>(n));
      t->value->read_all(dense.data());
      std::ofstream(out.second, std::ios::binary).write(reinterpret_cast<char const *>(dense.data()), sizeof(double) * dense.size());
    }
    delete algorithm;
  } catch (DetailedException *e) {
    std::printf("EXCEPTION: %s\n", e->getMessage().c_str());
    return 1;
  }
  return 0;
}
