// Checks the restatement of glibc's exp() used by the lambda kernel (quickrank_b200/csrc/qr_kernels.cuh, exp_lambda)
// bit for bit against this machine's libm:  gcc -O2 -march=x86-64-v3 -ffp-contract=off scripts/exp_glibc_check.c -lm && ./a.out
// (variant 0 = the FMA build of libm, what an x86-64-v3 host runs; variant 1 = no contraction)
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "exp_glibc_table.h"
static inline uint64_t asu(double x){uint64_t u;memcpy(&u,&x,8);return u;}
static inline double asd(uint64_t u){double x;memcpy(&x,&u,8);return x;}
#define N 128
static const double InvLn2N = 0x1.71547652b82fep0 * N, NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47, Shift = 0x1.8p52;
static const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
double emu(double x, int variant){
  double z = InvLn2N * x;
  double kd = z + Shift; uint64_t ki = asu(kd); kd -= Shift;
  double r = fma(kd, NegLn2loN, fma(kd, NegLn2hiN, x));
  uint64_t idx = 2 * (ki % N), top = ki << (52 - 7);
  double tail = asd(T[idx]); uint64_t sbits = T[idx+1] + top;
  double r2 = r * r, tmp;
  if (variant == 0) { /* fully contracted, left to right */
    double t1 = tail + r; double t2 = fma(r, C3, C2); double t3 = fma(r2, t2, t1); double t4 = fma(r, C5, C4); tmp = fma(r2 * r2, t4, t3);
  } else if (variant == 1) { /* no contraction */
    tmp = tail + r + r2 * (C2 + r * C3) + r2 * r2 * (C4 + r * C5);
  } else { double t2 = fma(r, C3, C2); double t4 = fma(r, C5, C4); double t3 = fma(r2, t2, tail + r); tmp = fma(r2*r2, t4, t3); }
  double scale = asd(sbits);
  return variant == 1 ? scale + scale * tmp : fma(scale, tmp, scale);
}
int main(){
  srand48(7);
  for (int v = 0; v < 2; ++v) {
    long bad = 0, n = 20000000; double worst = 0;
    for (long i = 0; i < n; ++i) {
      double x = (drand48() - 0.5) * 2 * (i % 3 == 0 ? 60.0 : (i % 3 == 1 ? 5.0 : 0.01));
      double a = emu(x, v), b = exp(x);
      if (asu(a) != asu(b)) { ++bad; double e = fabs(a-b)/b; if (e > worst) worst = e; }
    }
    printf("variant %d: %ld of %ld differ from libm exp (worst rel %.3g)\n", v, bad, n, worst);
  }
  printf("exp(0)=%a emu=%a\n", exp(0.0), emu(0.0,0));
  return 0;
}
