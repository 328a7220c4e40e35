/* CPU oracle: cascaded-biquad IIR recurrence used by zero-phase filtering.
 *
 * TEST INFRASTRUCTURE ONLY (checker for the CUDA LPF/BPF path; never linked
 * into opticomlib_b200).
 *
 * Restates the per-sample recurrence of scipy.signal.sosfilt (SciPy is the
 * third-party dependency that holds this arithmetic; it is not under
 * /root/reference -- requirements.txt pins scipy==1.12.0, the image has 1.18.1).
 * Reference call sites: opticomlib/devices.py:820-823 (BPF) and 1365-1368 (LPF),
 * both through scipy.signal.sosfiltfilt.
 *
 * Direct-form-II-transposed, section after section for every sample:
 *     y  = b0*x + z0
 *     z0 = b1*x - a1*y + z1
 *     z1 = b2*x - a2*y
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py).
 * -ffp-contract=off keeps the products and sums separately rounded, like the
 * SciPy loop compiled without FMA contraction.
 */
#include <stddef.h>

/* x: n samples with element stride `stride` (in doubles), filtered in place.
 * sos: S rows [b0 b1 b2 a0 a1 a2] (a0 == 1).  zi: S x 2 state, updated. */
void oracle_sosfilt_f64(const double *sos, int S, double *x, long n, long stride, double *zi)
{
    for (long i = 0; i < n; ++i) {
        double v = x[i * stride];
        for (int s = 0; s < S; ++s) {
            const double *c = sos + 6 * s;
            double *z = zi + 2 * s;
            double y = c[0] * v + z[0];
            z[0] = c[1] * v - c[4] * y + z[1];
            z[1] = c[2] * v - c[5] * y;
            v = y;
        }
        x[i * stride] = v;
    }
}
