#include <stdio.h>
#include <math.h>
#include <stdlib.h>
// Closed-form minor-step count of line_cost (mppi_device.cuh: fp32 floor of (den / 2 + k numadd + 0.5) / den) against the serial
// Bresenham error term of nav2_util::LineIterator, exhaustive for den < 2048, numadd <= den, k <= den (2.9e9 cases, 13 s):
//   gcc -O2 -ffp-contract=off -o /tmp/check_line scripts/check_line_closed_form.c -lm && /tmp/check_line
int main(void) {
  long long checked = 0, bad = 0;
  for (int den = 1; den < 2048; ++den) {
    const float inv_den = 1.0f / (float)den;
    const int num0 = den / 2;
    const float a0 = (float)num0 + 0.5f;
    for (int numadd = 0; numadd <= den; ++numadd) {
      int num = num0, m = 0;
      const float numadd_f = (float)numadd;
      for (int k = 0; k <= den; ++k) {
        const float a = fmaf((float)k, numadd_f, a0);
        const int mk = (int)(a * inv_den);
        if (mk != m) {++bad; if (bad < 10) printf("den %d numadd %d k %d: %d vs %d\n", den, numadd, k, mk, m);}
        ++checked;
        num += numadd;
        if (num >= den) {num -= den; ++m;}
      }
    }
  }
  printf("checked %lld, mismatches %lld\n", checked, bad);
  return bad != 0;
}
