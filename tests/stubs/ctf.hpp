// Stand-in for Cyclops CTF's <ctf.hpp> (un-vendored dependency of sisi4s, CTF_COMMIT 53ae5daa,
// absent from this environment).  TEST INFRASTRUCTURE: declares just enough of the CTF interface
// for the reference's own headers (src/util/Tensor.hpp, src/Data.hpp, src/algorithms/Algorithm.hpp)
// and for CcsdPerturbativeTriplesGpu.cxx to be SYNTAX-checked (g++ -fsyntax-only).  Nothing here
// is ever linked or executed.
#pragma once
#include <mpi.h>
#include <algorithm>
#include <complex>
#include <functional>
#include <map>
#include <vector>
#include <cstdint>
#include <string>

enum { NS = 0, SY = 1, AS = 2, SH = 3 };

namespace CTF {
class World {
public:
  MPI_Comm comm;
  int rank, np;
  World() : comm(0), rank(0), np(1) {}
  World(int, char **) : comm(0), rank(0), np(1) {}
};

template <typename F> class Idx_Tensor;

template <typename F = double>
class Tensor {
public:
  int order;
  int64_t *lens;
  int *sym;
  World *wrld;
  Tensor();
  Tensor(int order, int const *lens, int const *sym, World &w, char const *name = nullptr);
  Tensor(int order, int64_t const *lens, int const *sym, World &w, char const *name = nullptr);
  Tensor(Tensor const &other);
  ~Tensor();
  void read_all(F *data, bool unpack = false);
  void read_all(int64_t *n, F **data, bool unpack = false);
  void read(int64_t n, int64_t const *idx, F *data);
  void write(int64_t n, int64_t const *idx, F const *data);
  Tensor<F> slice(int const *begin, int const *end);
  Tensor<F> slice(int64_t const *begin, int64_t const *end);
  char const *get_name() const;
  void set_name(char const *);
  Idx_Tensor<F> operator[](char const *idx);
};
template <typename F = double> class Matrix : public Tensor<F> {};
template <typename F = double> class Vector : public Tensor<F> {};
template <typename F = double> class Scalar : public Tensor<F> { public: F get_val(); };
template <typename F> class Idx_Tensor {};
template <typename F = double> class Univar_Function {};
template <typename F = double> class Bivar_Function {};
template <typename F = double> class Transform {};
}  // namespace CTF
