"""Arithmetic identities the CUDA path relies on where it departs from the letter of the reference's
expressions: each is checked exhaustively against the reference's own expression on the CPU."""
import os

def test_svd_invdet_float_form(tmp_path):
    """k_leaves evaluates svd_invdet (qef.cl:107-109, tolerance 0.1f) without the reference's double
    division: (|x| < 0.1f || |x| >= 10.0f) ? 0 : 1.0f / x.  Exhaustive over all 2^32 floats against the
    reference's expression, bit for bit (NaN to NaN); a threshold off by one ulp must be caught."""
    import subprocess
    src = r"""
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline float ref(float x, float tol) { const double inv = 1.0 / (double)x; return (fabsf(x) < tol || fabs(inv) < (double)tol) ? 0.0f : (float)inv; }
static inline float alt(float x) { const float a = fabsf(x); return (a < 0.1f || a THRESH 10.0f) ? 0.0f : 1.0f / x; }
int main(void) {
    long bad = 0;
    #pragma omp parallel for reduction(+:bad) schedule(static)
    for (long long i = 0; i < (1LL << 32); i++) {
        uint32_t u = (uint32_t)i; float x; memcpy(&x, &u, 4);
        float r = ref(x, 0.1f), a = alt(x);
        uint32_t ur, ua; memcpy(&ur, &r, 4); memcpy(&ua, &a, 4);
        if (ur != ua && !(isnan(r) && isnan(a))) bad++;
    }
    printf("%ld\n", bad);
    return 0;
}
"""
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    for thresh, expect in ((">=", 0), (">", 2)):
        c = tmp_path / f"invdet_{expect}.c"
        c.write_text(src.replace("THRESH", thresh))
        exe = tmp_path / f"invdet_{expect}"
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", str(c), "-o", str(exe), "-lm"], env=env)
        assert int(subprocess.check_output([str(exe)], env=env).decode().strip()) == expect
