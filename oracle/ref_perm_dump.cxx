// ref_perm_dump.cxx -- compiled against the reference's OWN header
// /root/reference/src/math/Permutation.hpp (include path given by oracle/Makefile; the
// header is not copied into this repo).  Prints the tables the hot path depends on
// (CcsdPerturbativeTriples.cxx:22-30,143,161-212) so tests/test_oracle.py can pin the
// oracle's restated permutation algebra against the real reference code.
#include <cstdint>
#include <cstdio>
#include <string>
#include <math/Permutation.hpp>

using namespace sisi4s;

// std::string * Permutation, restated from CcsdPerturbativeTriples.cxx:22-30
// (that operator lives in the .cxx, which needs CTF and cannot be compiled here)
static std::string after(const std::string &s, const Permutation<3> &pi) {
  std::string r(s);
  for (int i = 0; i < 3; ++i) r[i] = s[pi(i)];
  return r;
}

int main() {
  for (int p = 0; p < Permutation<3>::ORDER; ++p) {
    Permutation<3> pi(p);
    std::printf("perm %d %d %d %d inv %d str %s\n", p, pi(0), pi(1), pi(2),
                pi.invariantElementsCount(), after("abc", pi).c_str());
  }
  // which permutations give a distinct (i,j,k) o pi, for the four degeneracy classes
  const int classes[4][3] = {{0, 1, 2}, {0, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  for (int c = 0; c < 4; ++c) {
    Map<3> i;
    for (int m = 0; m < 3; ++m) i(m) = classes[c][m];
    std::printf("distinct %d%d%d", i(0), i(1), i(2));
    for (int p = 0; p < 6; ++p) {
      int q;
      for (q = 0; q < p; ++q)
        if (i * Permutation<3>(q) == i * Permutation<3>(p)) break;
      std::printf(" %d", q < p ? 0 : 1);
    }
    std::printf("\n");
  }
  // composed index strings ("abc" * sigma) * pi used at :205-211
  for (int p = 0; p < 6; ++p)
    for (int s = 0; s < 6; ++s)
      std::printf("compose %d %d %s\n", s, p, after(after("abc", Permutation<3>(s)), Permutation<3>(p)).c_str());
  return 0;
}
