/*
 * kernel_model.c -- CPU model of the arithmetic the CUDA forward kernel performs (tests only).
 *
 * It is NOT the oracle and NOT product code: it mirrors the kernel's *reformulated* recurrence
 * (scaled insertion/deletion states, pad rows that carry M+I to the bottom of the strip, END
 * columns between back-to-back haplotypes, power-of-two initial condition) in the same operation
 * order and precision, so that tests can bound fp32-vs-oracle error on the build box, where no GPU
 * exists.  See DESIGN.md "Kernel recurrence".
 *
 *   I~[i][j] = I[i][j] / tMI_i          D~[i][j] = D[i][j] / tMD_i
 *   u        = M[i-1][j-1] + b'_i*I~[i-1][j-1] + c'_i*D~[i-1][j-1]         M[i][j] = (prior*tMM_i)*u
 *   D~[i][j] = M[i][j-1] + tDD_i*D~[i][j-1]
 *   I~[i][j] = M[i-1][j] + g_i*I~[i-1][j]
 *   b'=tIM_i*tMI_{i-1}/tMM_i  c'=tIM_i*tMD_{i-1}/tMM_i  g=tII_i*tMI_{i-1}/tMI_i   (tMI_0=tMD_0=1)
 * (5 FP instructions per cell; tMM = 0 is not representable in this form -- the kernels redo such pairs in fp64 --
 * and the model returns -2 for it.)
 * Pad rows below the read: b'=c'=tDD=0, prior=0, g=tMI_R for row R+1 and 1 after it, so the last
 * row of the strip holds (M+I)[R][j].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef MODEL_DOUBLE
typedef double real;
#define FMA fma
#define NAME(x) x##_f64
#else
typedef float real;
#define FMA fmaf
#define NAME(x) x##_f32
#endif

/* eps[q] = 10^(-q/10), m2m = triangular matchToMatch table (both double, supplied by the caller
 * exactly as the library computes them).  rows = total rows simulated (> R, multiple of K*32 in
 * the kernel; any value > R gives the same numbers).  stream: hap bases of nH haplotypes.
 * out[h] = raw last-row sum for haplotype h (before the log10 epilogue). c0 = 2^c0_exp. */
int NAME(model_task)(const double *eps, const double *m2m, const uint8_t *read, const uint8_t *bq,
                     const uint8_t *iq, const uint8_t *dq, const uint8_t *gq, int R, int rows,
                     const uint8_t *haps, const int32_t *hap_off, int nH, int c0_exp,
                     int tristate_off, real *out) {
    if (rows <= R) return -1;
    real *a = malloc(sizeof(real) * rows * 7), *b = a + rows, *c = b + rows, *g = c + rows, *dd = g + rows,
         *pm = dd + rows, *px = pm + rows;
    for (int i = 1; i <= rows; i++) {
        double A = 0, B = 0, C = 0, G = 1, DD = 0, PM = 0, PX = 0;
        if (i <= R) {
            const double ei = eps[iq[i - 1]], ed = eps[dq[i - 1]], ec = eps[gq[i - 1]];
            const double tmi_prev = i > 1 ? eps[iq[i - 2]] : 1.0, tmd_prev = i > 1 ? eps[dq[i - 2]] : 1.0;
            const int qi = iq[i - 1], qd = dq[i - 1];
            const int mn = qi < qd ? qi : qd, mx = qi < qd ? qd : qi;
            const double tIM = 1.0 - ec;
            A = m2m[((mx * (mx + 1)) >> 1) + mn];
            if (!(A > 0.0)) { free(a); return -2; }
            B = tIM * tmi_prev / A; C = tIM * tmd_prev / A; G = ec * tmi_prev / ei; DD = ec;
            const double e = eps[bq[i - 1]];
            PM = (1.0 - e) * A; PX = (tristate_off ? e : e / 3.0) * A;
        } else if (i == R + 1) {
            G = R >= 1 ? eps[iq[R - 1]] : 1.0;
        }
        a[i - 1] = (real)A; b[i - 1] = (real)B; c[i - 1] = (real)C; g[i - 1] = (real)G; dd[i - 1] = (real)DD;
        pm[i - 1] = (real)PM; px[i - 1] = (real)PX;
    }
    const real c0 = (real)ldexp(1.0, c0_exp);
    for (int h = 0; h < nH; h++) {
        const uint8_t *hap = haps + hap_off[h];
        const int H = hap_off[h + 1] - hap_off[h];
        /* column-major sweep: the kernel's wavefront computes exactly these values */
        real *Mp = calloc((size_t)(rows + 1) * 6, sizeof(real)), *Ip = Mp + rows + 1, *Dp = Ip + rows + 1,
             *Mc = Dp + rows + 1, *Ic = Mc + rows + 1, *Dc = Ic + rows + 1;
        /* index 0 = virtual row 0 */
        Dp[0] = c0; Dc[0] = c0;
        real sum = 0;
        for (int j = 1; j <= H; j++) {
            const uint8_t y = hap[j - 1];
            Mc[0] = 0; Ic[0] = 0; Dc[0] = c0;
            for (int i = 1; i <= rows; i++) {
                real prior = 0;
                if (i <= R) {
                    const uint8_t x = read[i - 1];
                    prior = (x == y || x == 'N' || y == 'N') ? pm[i - 1] : px[i - 1];
                }
                real u = FMA(c[i - 1], Dp[i - 1], Mp[i - 1]);
                u = FMA(b[i - 1], Ip[i - 1], u);
                Mc[i] = prior * u;
                Dc[i] = FMA(dd[i - 1], Dp[i], Mp[i]);
                Ic[i] = FMA(g[i - 1], Ic[i - 1], Mc[i - 1]);
            }
            sum += Ic[rows];
            real *s;
            s = Mp; Mp = Mc; Mc = s; s = Ip; Ip = Ic; Ic = s; s = Dp; Dp = Dc; Dc = s;
        }
        out[h] = sum;
        /* the six row pointers were carved from one allocation whose base is the min pointer */
        real *base = Mp < Mc ? Mp : Mc;
        free(base);
    }
    free(a);
    return 0;
}
