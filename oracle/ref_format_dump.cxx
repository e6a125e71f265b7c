// ref_format_dump.cxx -- compiled against the reference's OWN headers
//   /root/reference/src/util/BinaryTensorFormat.hpp        (TENS binary tensor files)
//   /root/reference/src/algorithms/CoulombVertexReader.hpp (legacy FTODDUMP files)
// (include paths given by oracle/Makefile; the headers are not copied into this repo; the absent
// third-party <ctf.hpp> / <mpi.h> are the declaration-only stand-ins of tests/stubs).  Prints the bytes
// of the header structs the reference writes, so tests/test_tensor_io.py can pin sisi4s_b200/tensor_io.py
// against the real struct layouts.
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <util/BinaryTensorFormat.hpp>
#include <algorithms/CoulombVertexReader.hpp>

using namespace sisi4s;

static void hex(const char *tag, const void *p, size_t n) {
  std::printf("%s", tag);
  for (size_t i = 0; i < n; ++i) std::printf(" %02x", ((const unsigned char *)p)[i]);
  std::printf("\n");
}

template <typename F>
static void tens_header(const char *tag, int order) {
  // BinaryTensorHeader only reads T.order: a zeroed stand-in object is enough (nothing of CTF is linked)
  alignas(Tensor<F>) static unsigned char storage[sizeof(Tensor<F>)];
  std::memset(storage, 0, sizeof storage);
  Tensor<F> *t = reinterpret_cast<Tensor<F> *>(storage);
  t->order = order;
  BinaryTensorHeader h(*t);
  hex(tag, &h, sizeof h);
}

int main() {
  std::printf("sizeof_header %zu sizeof_dim %zu\n", sizeof(BinaryTensorHeader), sizeof(BinaryTensorDimensionHeader));
  tens_header<double>("tens_real_order4", 4);
  tens_header<complex>("tens_complex_order3", 3);
  // writeBinary (TensorIo.cxx:138-147): BinaryTensorDimensionHeader(A.lens[dim], 'a' + dim); its constructor
  // leaves flags / reserved uninitialised, so they are cleared here before the struct is printed
  const int lens[4] = {5, 7, 3, 2};
  for (int dim = 0; dim < 4; ++dim) {
    BinaryTensorDimensionHeader d(lens[dim], 'a' + dim);
    d.flags = 0;
    d.reserved = 0;
    char tag[32];
    std::snprintf(tag, sizeof tag, "tens_dim%d", dim);
    hex(tag, &d, sizeof d);
  }
  typedef CoulombVertexReader::Header H;
  typedef CoulombVertexReader::Chunk C;
  std::printf("ftod_header size %zu magic %zu No %zu Nv %zu NG %zu NSpins %zu kPoints %zu reserved %zu\n", sizeof(H),
              offsetof(H, magic), offsetof(H, No), offsetof(H, Nv), offsetof(H, NG), offsetof(H, NSpins),
              offsetof(H, kPoints), offsetof(H, reserved_));
  std::printf("ftod_chunk size %zu magic %zu size_field %zu magic_len %zu\n", sizeof(C), offsetof(C, magic),
              offsetof(C, size), sizeof(((C *)0)->magic));
  return 0;
}
