// Test helper (no GPU needed): host::parse_float (host/include/qr_fast_float.h) against strtof on random and
// adversarial decimal strings — same value bit for bit, same end pointer.  usage: float_check [millions] [seed]
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "qr_fast_float.h"

static uint64_t failures = 0, checked = 0;

static void check(const char *s) {
  char *e1 = nullptr, *e2 = nullptr;
  const float a = quickrank::host::parse_float(s, &e1);
  const float b = strtof(s, &e2);
  uint32_t ua, ub;
  memcpy(&ua, &a, 4);
  memcpy(&ub, &b, 4);
  ++checked;
  if (ua != ub || e1 != e2) {
    if (failures++ < 20) fprintf(stderr, "MISMATCH '%s': %a (end +%td) vs strtof %a (end +%td)\n", s, a, e1 - s, b, e2 - s);
  }
}

int main(int argc, char **argv) {
  const uint64_t n = (uint64_t) (argc > 1 ? atof(argv[1]) * 1e6 : 5e6);
  std::mt19937_64 rng(argc > 2 ? strtoull(argv[2], nullptr, 10) : 1);
  char buf[128];
  // fixed cases: zeros, signs, exponents, trailing text, specials, hex, long digit strings, range limits
  const char *fixed[] = {"0", "-0", "+0.0", "0.0000", "1", "-1", "+1", ".5", "5.", "-.25", "1e0", "1E+2", "1e-2", "1e", "1e+", "12abc",
                         "3.5:7", "0x10", "0x1p3", "inf", "-inf", "nan", "infinity", "1e38", "3.4028235e38", "3.4028236e38", "1e39",
                         "1.17549435e-38", "1.17549e-38", "1e-45", "1e-46", "123456789012345678901234567890", "0.1", "0.2", "0.3",
                         "16777216", "16777217", "16777218", "16777219", "33554433", "0.000000000000000000000001", "1e22", "1e23",
                         "9007199254740993", "9007199254740992", "4.35", "2.675", "1.0000000596046448", "1.00000005960464477539"};
  for (const char *s : fixed) check(s);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t r = rng();
    switch (r % 7) {
      case 0:   // what the tests and most SVMLight writers produce: %.6g / %.9g of a float
        snprintf(buf, sizeof(buf), (r >> 8) & 1 ? "%.9g" : "%.6g", (double) (float) std::ldexp((double) (rng() >> 11) / 9007199254740992.0, (int) ((r >> 16) % 40) - 20));
        break;
      case 1:   // fixed-point decimals with 1..12 fractional digits
        snprintf(buf, sizeof(buf), "%s%" PRIu64 ".%0*" PRIu64, (r >> 9) & 1 ? "-" : "", (rng() % 100000), (int) ((r >> 12) % 12) + 1,
                 rng() % (uint64_t) std::pow(10.0, (double) ((r >> 12) % 12 + 1)));
        break;
      case 2: { // exactly a float rounding boundary (midpoint of two adjacent floats), printed exactly
        uint32_t u = (uint32_t) (rng() % 0x7f000000u) + 0x00800000u;
        float f0, f1;
        memcpy(&f0, &u, 4);
        ++u;
        memcpy(&f1, &u, 4);
        snprintf(buf, sizeof(buf), "%.60g", ((double) f0 + (double) f1) / 2);
        break;
      }
      case 3: { // a boundary nudged by one unit in its last printed digit (17 significant digits)
        uint32_t u = (uint32_t) (rng() % 0x7f000000u) + 0x00800000u;
        float f0, f1;
        memcpy(&f0, &u, 4);
        ++u;
        memcpy(&f1, &u, 4);
        const double m = ((double) f0 + (double) f1) / 2;
        snprintf(buf, sizeof(buf), "%.17g", std::nextafter(m, (r >> 20) & 1 ? 1e300 : -1e300));
        break;
      }
      case 4:   // integers up to 2^63
        snprintf(buf, sizeof(buf), "%" PRIu64, rng() >> ((r >> 8) % 60));
        break;
      case 5:   // scientific notation, 1..17 digits
        snprintf(buf, sizeof(buf), "%.*e", (int) ((r >> 8) % 17), std::ldexp((double) (rng() >> 11), (int) ((r >> 16) % 200) - 150));
        break;
      default:  // %.17g of a random double in the float range
        snprintf(buf, sizeof(buf), "%.17g", std::ldexp((double) (rng() >> 11) / 9007199254740992.0 + 0.5, (int) ((r >> 16) % 250) - 125));
    }
    check(buf);
  }
  printf("%" PRIu64 " strings, %" PRIu64 " mismatches\n", checked, failures);
  return failures ? 1 : 0;
}
