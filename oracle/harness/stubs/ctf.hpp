// Single-process stand-in for Cyclops CTF's <ctf.hpp> (un-vendored dependency of sisi4s, CTF_COMMIT 53ae5daa,
// absent from this environment).  TEST INFRASTRUCTURE of oracle/harness: a dense column-major tensor with the
// handful of members the reference's own Algorithm / Data sources and the plugin classes of
// sisi4s_b200/csrc/*.cxx touch (order, lens, wrld, read_all, write, slice, names).  No index expressions, no
// distribution: enough to RUN the plugin classes through the reference's argument machinery on one rank.
#pragma once
#include <mpi.h>

#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

enum { NS = 0, SY = 1, AS = 2, SH = 3 };

namespace CTF {
class World {
public:
  MPI_Comm comm;
  int rank, np;
  World() : comm(MPI_COMM_WORLD), rank(0), np(1) {}
  World(int, char **) : comm(MPI_COMM_WORLD), rank(0), np(1) {}
};

template <typename F> class Idx_Tensor {};

template <typename F = double>
class Tensor {
public:
  int order = 0;
  int64_t *lens = nullptr;
  int *sym = nullptr;
  World *wrld = nullptr;
  std::vector<F> data;
  std::string name;

  Tensor() {}
  template <typename L>
  void init(int order_, L const *lens_, World &w, char const *name_) {
    order = order_;
    lens = new int64_t[order_ > 0 ? order_ : 1];
    sym = new int[order_ > 0 ? order_ : 1];
    int64_t n = 1;
    for (int d = 0; d < order; ++d) {
      lens[d] = lens_[d];
      sym[d] = NS;
      n *= lens[d];
    }
    data.assign(static_cast<size_t>(n), F(0));
    wrld = &w;
    if (name_) name = name_;
  }
  Tensor(int order_, int const *lens_, int const *, World &w, char const *name_ = nullptr) { init(order_, lens_, w, name_); }
  Tensor(int order_, int64_t const *lens_, int const *, World &w, char const *name_ = nullptr) { init(order_, lens_, w, name_); }
  Tensor(Tensor const &other) {
    init(other.order, other.lens, *other.wrld, other.name.c_str());
    data = other.data;
  }
  Tensor &operator=(Tensor const &) = delete;
  ~Tensor() {
    delete[] lens;
    delete[] sym;
  }
  int64_t size() const { return static_cast<int64_t>(data.size()); }
  void read_all(F *out, bool = false) { std::copy(data.begin(), data.end(), out); }
  void read_all(int64_t *n, F **out, bool = false) {
    *n = size();
    *out = static_cast<F *>(malloc(sizeof(F) * data.size()));
    std::copy(data.begin(), data.end(), *out);
  }
  void read(int64_t n, int64_t const *idx, F *out) {
    for (int64_t q = 0; q < n; ++q) out[q] = data[idx[q]];
  }
  void write(int64_t n, int64_t const *idx, F const *in) {
    for (int64_t q = 0; q < n; ++q) data[idx[q]] = in[q];
  }
  template <typename L>
  Tensor<F> sliceOf(L const *begin, L const *end) {
    std::vector<int64_t> l(order);
    for (int d = 0; d < order; ++d) l[d] = end[d] - begin[d];
    Tensor<F> out(order, l.data(), sym, *wrld, name.c_str());
    std::vector<int64_t> x(order, 0);
    for (int64_t q = 0; q < out.size(); ++q) {
      int64_t src = 0, stride = 1;
      for (int d = 0; d < order; ++d) {
        src += (begin[d] + x[d]) * stride;
        stride *= lens[d];
      }
      out.data[q] = data[src];
      for (int d = 0; d < order; ++d) {
        if (++x[d] < l[d]) break;
        x[d] = 0;
      }
    }
    return out;
  }
  Tensor<F> slice(int const *begin, int const *end) { return sliceOf(begin, end); }
  Tensor<F> slice(int64_t const *begin, int64_t const *end) { return sliceOf(begin, end); }
  char const *get_name() const { return name.c_str(); }
  void set_name(char const *n) { name = n; }
  Idx_Tensor<F> operator[](char const *) { return Idx_Tensor<F>(); }
};
template <typename F = double> class Matrix : public Tensor<F> {};
template <typename F = double> class Vector : public Tensor<F> {};
inline World &get_universe() {
  static World universe;
  return universe;
}
template <typename F = double> class Scalar : public Tensor<F> {
public:
  Scalar() { this->init(0, static_cast<int64_t const *>(nullptr), get_universe(), nullptr); }
  Scalar(World &w) { this->init(0, static_cast<int64_t const *>(nullptr), w, nullptr); }
  Scalar(F value, World &w = get_universe()) {
    this->init(0, static_cast<int64_t const *>(nullptr), w, nullptr);
    this->data[0] = value;
  }
  F get_val() { return this->data.empty() ? F(0) : this->data[0]; }
};
template <typename F = double> class Univar_Function {};
template <typename F = double> class Bivar_Function {};
template <typename F = double> class Transform {};
}  // namespace CTF
