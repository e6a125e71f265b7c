// Stand-in for <yaml-cpp/yaml.h> (un-vendored dependency of sisi4s, absent here).  TEST INFRASTRUCTURE of
// oracle/harness: the emitter interface src/util/Emitter.{hpp,cxx} and src/algorithms/Algorithm.cxx use, writing
// "key: value" lines.
#pragma once
#include <ostream>
#include <string>

namespace YAML {
enum EMITTER_MANIP { Key, Value, BeginMap, EndMap, BeginSeq, EndSeq, Flow, Newline };
class Emitter {
public:
  explicit Emitter(std::ostream &s) : out(&s) {}
  Emitter &operator<<(EMITTER_MANIP m) {
    if (m == Value) *out << ": ";
    if (m == Key && started) *out << "\n";
    started = true;
    return *this;
  }
  template <typename T> Emitter &operator<<(T const &v) {
    *out << v;
    return *this;
  }
private:
  std::ostream *out;
  bool started = false;
};
}  // namespace YAML
