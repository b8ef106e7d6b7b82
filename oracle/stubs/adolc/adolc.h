/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the third-party
 * ADOL-C 2.6.3 headers, so that the reference's own pl_calc_parallel.cpp can be
 * compiled from /root/reference/src without building the ADOL-C tarball.
 *
 * It provides a NON-TAPING `adouble` (a plain double wrapper): the reference's
 * calc_pl / calc_gradient / RAD calc_function_gradient paths do not touch it at
 * run time; the ADOL-C entry points (calc_*_function_gradient_adolc) compile but
 * `gradient()` reports failure (-1).  The RAD tape (reference src/myrad.h) is the
 * AD oracle; SURVEY.md §8 a14 measured it bit-identical to ADOL-C.
 * Written from scratch for this repo; contains no ADOL-C code. */
#ifndef TREEPL_B200_ORACLE_ADOLC_STUB_H
#define TREEPL_B200_ORACLE_ADOLC_STUB_H
#include <cmath>
#include <omp.h>

class adouble {
    double v_;
public:
    adouble() : v_(0.0) {}
    adouble(double v) : v_(v) {}
    double value() const { return v_; }
    double getValue() const { return v_; }
    adouble& operator<<=(double v) { v_ = v; return *this; }
    adouble& operator>>=(double& out) { out = v_; return *this; }
    adouble& operator+=(const adouble& o) { v_ += o.v_; return *this; }
    adouble& operator-=(const adouble& o) { v_ -= o.v_; return *this; }
    adouble& operator*=(const adouble& o) { v_ *= o.v_; return *this; }
    adouble& operator/=(const adouble& o) { v_ /= o.v_; return *this; }
    adouble operator++(int) { adouble t(*this); v_ += 1.0; return t; }
    adouble& operator++() { v_ += 1.0; return *this; }
    adouble operator-() const { return adouble(-v_); }
};
inline adouble operator+(const adouble& a, const adouble& b) { return adouble(a.value() + b.value()); }
inline adouble operator-(const adouble& a, const adouble& b) { return adouble(a.value() - b.value()); }
inline adouble operator*(const adouble& a, const adouble& b) { return adouble(a.value() * b.value()); }
inline adouble operator/(const adouble& a, const adouble& b) { return adouble(a.value() / b.value()); }
inline bool operator>(const adouble& a, const adouble& b) { return a.value() > b.value(); }
inline bool operator<(const adouble& a, const adouble& b) { return a.value() < b.value(); }
inline bool operator>=(const adouble& a, const adouble& b) { return a.value() >= b.value(); }
inline bool operator<=(const adouble& a, const adouble& b) { return a.value() <= b.value(); }
inline bool operator==(const adouble& a, const adouble& b) { return a.value() == b.value(); }
inline adouble log(const adouble& a) { return adouble(std::log(a.value())); }
inline adouble sqrt(const adouble& a) { return adouble(std::sqrt(a.value())); }
inline adouble exp(const adouble& a) { return adouble(std::exp(a.value())); }

inline int trace_on(short, int = 0) { return 0; }
inline void trace_off(int = 0) {}
/* no tape exists: tell the caller the driver failed */
inline int gradient(short, int, const double*, double*) { return -1; }
#endif
