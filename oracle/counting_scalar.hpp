// ORACLE (test infrastructure, not product): an instrumented scalar that counts floating-point operations.
// Sim<Cnt> runs the same restatement as Sim<double>; the counters give the ALGORITHMIC FLOPs per env-step that
// SURVEY.md section 8(d) asks to publish next to the byte count.
#pragma once
#include <cmath>

namespace orc {

struct FlopCounters { long long add = 0, mul = 0, div = 0, sqrt = 0, trig = 0, cmp = 0; };
inline FlopCounters& flop_counters() { static FlopCounters c; return c; }

struct Cnt {
  double v;
  Cnt() : v(0) {}
  Cnt(double x) : v(x) {}
  Cnt(float x) : v(x) {}
  Cnt(int x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return float(v); }
  explicit operator int() const { return int(v); }
  Cnt operator-() const { return Cnt(-v); }
  Cnt& operator+=(Cnt o) { flop_counters().add++; v += o.v; return *this; }
  Cnt& operator-=(Cnt o) { flop_counters().add++; v -= o.v; return *this; }
  Cnt& operator*=(Cnt o) { flop_counters().mul++; v *= o.v; return *this; }
  Cnt& operator/=(Cnt o) { flop_counters().div++; v /= o.v; return *this; }
};
inline Cnt operator+(Cnt a, Cnt b) { flop_counters().add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { flop_counters().add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { flop_counters().mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { flop_counters().div++; return Cnt(a.v / b.v); }
#define ORC_MIXED(op) \
  inline Cnt operator op(Cnt a, double b) { return a op Cnt(b); } \
  inline Cnt operator op(double a, Cnt b) { return Cnt(a) op b; } \
  inline Cnt operator op(Cnt a, int b) { return a op Cnt(b); }    \
  inline Cnt operator op(int a, Cnt b) { return Cnt(a) op b; }
ORC_MIXED(+) ORC_MIXED(-) ORC_MIXED(*) ORC_MIXED(/)
#undef ORC_MIXED
#define ORC_CMP(op) \
  inline bool operator op(Cnt a, Cnt b) { flop_counters().cmp++; return a.v op b.v; }       \
  inline bool operator op(Cnt a, double b) { flop_counters().cmp++; return a.v op b; }      \
  inline bool operator op(double a, Cnt b) { flop_counters().cmp++; return a op b.v; }      \
  inline bool operator op(Cnt a, int b) { flop_counters().cmp++; return a.v op double(b); } \
  inline bool operator op(int a, Cnt b) { flop_counters().cmp++; return double(a) op b.v; }
ORC_CMP(<) ORC_CMP(>) ORC_CMP(<=) ORC_CMP(>=) ORC_CMP(==) ORC_CMP(!=)
#undef ORC_CMP

}  // namespace orc

// the restatement calls std::sqrt / sin / cos / fabs / pow / max by qualified name
namespace std {
inline orc::Cnt sqrt(orc::Cnt a) { orc::flop_counters().sqrt++; return orc::Cnt(std::sqrt(a.v)); }
inline orc::Cnt sin(orc::Cnt a) { orc::flop_counters().trig++; return orc::Cnt(std::sin(a.v)); }
inline orc::Cnt cos(orc::Cnt a) { orc::flop_counters().trig++; return orc::Cnt(std::cos(a.v)); }
inline orc::Cnt fabs(orc::Cnt a) { return orc::Cnt(std::fabs(a.v)); }
inline orc::Cnt pow(orc::Cnt a, orc::Cnt b) { orc::flop_counters().trig++; return orc::Cnt(std::pow(a.v, b.v)); }
inline bool isfinite(orc::Cnt a) { return std::isfinite(a.v); }
inline orc::Cnt floor(orc::Cnt a) { return orc::Cnt(std::floor(a.v)); }
}  // namespace std
