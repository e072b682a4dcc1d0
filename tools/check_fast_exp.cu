// Host check of gpmdm_b200/csrc/fast_exp.cuh against libm:
//   nvcc -O2 -o /tmp/check_fast_exp tools/check_fast_exp.cu && /tmp/check_fast_exp
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../gpmdm_b200/csrc/fast_exp.cuh"
static const double TABLE[64] = {GPMDM_EXP_TABLE_VALUES};
int main() {
    double max_ulp = 0, worst = 0;
    srand48(1);
    for (long i = 0; i < 40000000; i++) {
        double x;
        int m = i % 4;
        if (m == 0) x = -drand48() * 1.0;
        else if (m == 1) x = -drand48() * 40.0;
        else if (m == 2) x = -drand48() * 745.0;
        else x = (drand48() - 0.5) * 1e-8;
        double a = gpmdm::fast_exp(x, TABLE), b = exp(x);
        if (b < 1e-300) continue;
        double ulp = fabs(a - b) / (nextafter(b, INFINITY) - b);
        if (ulp > max_ulp) { max_ulp = ulp; worst = x; }
    }
    printf("max ulp error %.3f at x = %.17g\n", max_ulp, worst);
    printf("exp(-800)=%g exp(-1e9)=%g exp(-1e300)=%g exp(0)=%.17g exp(-700)=%g vs %g\n", gpmdm::fast_exp(-800.0, TABLE),
           gpmdm::fast_exp(-1e9, TABLE), gpmdm::fast_exp(-1e300, TABLE), gpmdm::fast_exp(0.0, TABLE),
           gpmdm::fast_exp(-700.0, TABLE), exp(-700.0));
    return max_ulp < 2.5 ? 0 : 1;
}
