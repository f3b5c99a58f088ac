// O2DTypes.h -- small value types of the Omega2D influence interface, mirrored for the
// B200 host layer.  Names, enumerators and numeric values follow the reference so code written
// against it compiles unchanged:
//   Int, Dimensions, elem_t, move_t      reference src/Omega2D.h:23-45
//   Vector<S>                            reference src/VectorHelper.h:56 (USE_VC off)
//   results_t, ResultsType               reference src/ResultsType.h:25-94
//   summation_t, accel_t, ExecEnv        reference src/ExecEnv.h:26-109
// When the reference's own headers are already included (OMEGA2D_REFERENCE_HEADERS defined by
// the includer) this file defines nothing.
#pragma once
#ifndef OMEGA2D_REFERENCE_HEADERS

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

using Int = uint32_t;
const size_t Dimensions = 2;

enum elem_t { active = 1, reactive = 2, inert = 3, hybrid = 4 };
enum move_t { lagrangian = 1, bodybound = 2, fixed = 3 };

template <class S> using Vector = std::vector<S>;

enum results_t { velonly = 1, velandgrad = 2, psionly = 3, velandvort = 4, velandshear = 5 };

class ResultsType {
public:
  ResultsType(const results_t _type) : m_rtype(_type) {}
  ResultsType() : ResultsType(velonly) {}
  bool compute_psi() const { return m_rtype == psionly; }
  bool compute_vel() const { return m_rtype != psionly; }
  bool compute_grad() const { return m_rtype == velandgrad; }
  bool compute_vort() const { return m_rtype == velandvort; }
  bool compute_shear() const { return m_rtype == velandshear; }
  results_t get_type() const { return m_rtype; }
  std::string to_string() const {
    switch (m_rtype) {
      case psionly: return " for psi";
      case velonly: return " for vel";
      case velandgrad: return " for vel and grad";
      case velandvort: return " for vel and vort";
      case velandshear: return " for vel and shear";
    }
    return "";
  }
protected:
  results_t m_rtype;
};

enum summation_t { direct = 1, barneshut = 2, vic = 3, fmm = 4 };
// gpu_cuda is the value the reference reserves and never uses ("unsupported internally",
// reference src/ExecEnv.h:38); it is the run-time selector of this library.
enum accel_t { cpu_x86 = 1, cpu_vc = 2, gpu_opengl = 3, gpu_cuda = 4 };

class ExecEnv {
public:
  ExecEnv(const bool _internal, const summation_t _sumtype, const accel_t _acceltype)
    : m_internal(_internal), m_summ(_sumtype), m_accel(_acceltype) {}
  // this host layer has exactly one execution path
  ExecEnv() : ExecEnv(true, direct, gpu_cuda) {}
  void use_internal() { m_internal = true; }
  void use_external() { m_internal = false; }
  void set_internal(const bool _isint) { m_internal = _isint; }
  bool is_internal() const { return m_internal; }
  void set_summation(const summation_t _newsumm) { m_summ = _newsumm; }
  void set_instrs(const accel_t _newaccel) { m_accel = _newaccel; }
  accel_t get_instrs() const { return m_accel; }
  std::string to_string() const {
    if (!m_internal) return " external solver";
    std::string s = (m_accel == gpu_cuda) ? " B200 CUDA" : " unknown acceleration";
    s += (m_summ == direct) ? " direct sums" : " unknown algorithm";
    return s;
  }
protected:
  bool m_internal;
  summation_t m_summ;
  accel_t m_accel;
};

#endif  // OMEGA2D_REFERENCE_HEADERS
