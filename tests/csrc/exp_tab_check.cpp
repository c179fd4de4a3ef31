// Host-side accuracy check of magmaclust_b200/csrc/exp_tab.h against libm (tests/test_exp_table.py compiles and runs it).
// Prints: worst error in ulp, the argument where it occurred, then the special values.
#include <cmath>
#include <cstdio>
#include <random>
#include "../../magmaclust_b200/csrc/exp_tab.h"

static double ulp_err(double a, double b) {
  if (a == b) return 0.0;
  int e;
  frexp(b, &e);
  return fabs(a - b) / ldexp(1.0, e - 53);
}

int main() {
  std::mt19937_64 rng(3);
  double worst = 0, wx = 0;
  for (int rep = 0; rep < 4000000; ++rep) {
    const double u = (double)(rng() >> 11) * 0x1.0p-53;
    double x;
    switch (rep % 4) {
      case 0: x = -40.0 * u; break;           // the range the SE kernel lives in
      case 1: x = -708.0 * u; break;
      case 2: x = 709.0 * u; break;
      default: x = 4.0 * u - 2.0;
    }
    const double d = ulp_err(mcb::exp_tab(x, mcb::kExpTab), exp(x));
    if (d > worst) { worst = d; wx = x; }
  }
  printf("%.3f %.17g\n", worst, wx);
  const double xs[] = {0.0, -0.0, -708.5, -1e4, -INFINITY, 709.9, 1e4, INFINITY, NAN};
  for (double x : xs) printf("%.17g\n", mcb::exp_tab(x, mcb::kExpTab));
  return 0;
}
