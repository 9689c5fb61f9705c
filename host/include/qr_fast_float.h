// quickrank_b200 host layer — decimal text to float, correctly rounded, for the SVMLight reader.
//
// The reference reads feature values with sscanf("%f") (src/io/svml.cc:108-110), i.e. the correctly rounded float of
// the decimal string.  strtof does the same but costs ~100 ns per value, most of a multi-threaded parse.  This is the
// classic exact fast path: a decimal significand below 2^53 and a power of ten up to 10^22 are both exact doubles, so
// ONE IEEE multiplication or division gives the correctly rounded DOUBLE d of the exact value v.  Rounding d to float
// gives the correctly rounded float of v unless a float rounding boundary lies between v and d; boundaries are
// doubles too and |d - v| <= ulp(d)/2, so that can only happen when d IS a boundary (the 29 bits below a float's
// significand are exactly 1000...0) — then, and for anything else unusual (more than 19 digits, exponents beyond
// +-22, values outside the normal float range, inf / nan / hex), strtof decides.  host/float_check.cc compares the two
// on random and adversarial strings.
#ifndef QR_FAST_FLOAT_H
#define QR_FAST_FLOAT_H

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace quickrank {
namespace host {

inline float parse_float(const char *s, char **end) {
  static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char *p = s;
  while (*p == ' ' || *p == '\t') ++p;
  const bool neg = *p == '-';
  if (*p == '-' || *p == '+') ++p;
  uint64_t mant = 0;
  int ndigits = 0, exp10 = 0;
  bool any = false, ok = true;
  for (; *p >= '0' && *p <= '9'; ++p) {
    any = true;
    if (mant || *p != '0') {
      if (++ndigits > 19) { ok = false; break; }
      mant = mant * 10 + (uint64_t) (*p - '0');
    }
  }
  if (ok && *p == '.') {
    ++p;
    for (; *p >= '0' && *p <= '9'; ++p) {
      any = true;
      if (mant || *p != '0') {
        if (++ndigits > 19) { ok = false; break; }
        mant = mant * 10 + (uint64_t) (*p - '0');
      }
      --exp10;
    }
  }
  if (!ok || !any) return strtof(s, end);
  if (*p == 'e' || *p == 'E') {
    const char *q = p + 1;
    const bool eneg = *q == '-';
    if (*q == '-' || *q == '+') ++q;
    if (*q >= '0' && *q <= '9') {
      int e = 0;
      for (; *q >= '0' && *q <= '9'; ++q)
        if (e < 10000) e = e * 10 + (*q - '0');
      exp10 += eneg ? -e : e;
      p = q;
    }
  }
  // (a following 'x', 'n', 'i' ... cannot extend a number that started with a digit, except hex "0x": strtof's business)
  if ((*p == 'x' || *p == 'X') && ndigits == 0) return strtof(s, end);
  if (mant == 0) {
    if (end) *end = const_cast<char *>(p);
    return neg ? -0.0f : 0.0f;
  }
  if (mant >= (1ull << 53) || exp10 < -22 || exp10 > 22) return strtof(s, end);
  double d = (double) mant;
  d = exp10 < 0 ? d / p10[-exp10] : d * p10[exp10];
  if (!(d >= 1.1754943508222875e-38 && d <= 3.4028234663852886e+38)) return strtof(s, end);   // normal floats only
  uint64_t bits;
  std::memcpy(&bits, &d, sizeof(bits));
  const uint64_t low = bits & 0x1FFFFFFFull;            // what a float's significand drops
  if (low - 0x0FFFFFFFull <= 2ull) return strtof(s, end);   // d is (next to) a float rounding boundary
  if (end) *end = const_cast<char *>(p);
  const float f = (float) d;
  return neg ? -f : f;
}

}  // namespace host
}  // namespace quickrank
#endif
