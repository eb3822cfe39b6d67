/* TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the PyWavelets 1.5.0 single-level analysis / synthesis along one axis (the same published
 * formulas as oracle/dwt_ref.py, which remains the checker).  It exists so that the CPU baseline legs of bench.py
 * (`cpu_baseline`, `--impl reference`) time a compiled transform, as the reference does (pywt.wavedec2 /
 * waverec2 at spiht/spiht_wrapper.py:163,276 are C loops), instead of numpy fancy indexing.
 * tests/test_oracle_dwt.py checks it against dwt_ref.py to 1e-12.
 *
 *   non-periodization: out[k] = sum_j f[j] x_ext[2k + 1 - j],  k < (N + F - 1) / 2
 *                      rec[n] = sum_t g[t] c[(n + F - 2 - t) / 2]  (even numerator, 0 <= index < m), n < 2m - F + 2
 *   periodization:     out[k] = sum_j f[j] x_per[(2k + F/2 - j) mod Np],  k < ceil(N / 2)   (odd N padded with its last)
 *                      rec[n] = sum_t g[t] c[((n + F/2 - 1 - t) / 2) mod m]  (even numerator), n < 2m
 * mode: 0 reflect (whole-sample), 1 symmetric (half-sample), 2 periodization.
 */
#include <stdlib.h>

static int ext_index(long g, long n, int mode)
{
    if (mode == 0) { /* ... x2 x1 | x0 x1 ... xN-1 | xN-2 ... */
        if (n == 1) return 0;
        long p = 2 * n - 2, m = g % p;
        if (m < 0) m += p;
        return (int)(m >= n ? p - m : m);
    }
    long p = 2 * n, m = g % p; /* ... x1 x0 | x0 x1 ... xN-1 | xN-1 ... */
    if (m < 0) m += p;
    return (int)(m >= n ? p - 1 - m : m);
}

/* source index of tap j for output k (analysis) */
static int *analysis_map(int n, int F, int mode, int m)
{
    int *map = (int *)malloc(sizeof(int) * (size_t)m * F);
    for (int k = 0; k < m; ++k)
        for (int j = 0; j < F; ++j) {
            int s;
            if (mode == 2) {
                int np = n + (n & 1);
                long g = (2L * k + F / 2 - j) % np;
                if (g < 0) g += np;
                s = (int)(g >= n ? n - 1 : g); /* the pad sample repeats the last one */
            } else {
                s = ext_index(2L * k + 1 - j, n, mode);
            }
            map[(size_t)k * F + j] = s;
        }
    return map;
}

int dwt_out_len(int n, int F, int mode) { return mode == 2 ? (n + 1) / 2 : (n + F - 1) / 2; }

/* analysis along the last axis of x[rows][n] -> lo, hi [rows][m] */
void dwt_last(const double *x, int rows, int n, int mode, const double *flo, const double *fhi, int F, double *lo,
              double *hi)
{
    const int m = dwt_out_len(n, F, mode);
    int *map = analysis_map(n, F, mode, m);
    for (int r = 0; r < rows; ++r) {
        const double *xr = x + (size_t)r * n;
        double *l = lo + (size_t)r * m, *h = hi + (size_t)r * m;
        for (int k = 0; k < m; ++k) {
            double a = 0.0, b = 0.0;
            const int *mk = map + (size_t)k * F;
            for (int j = 0; j < F; ++j) {
                const double v = xr[mk[j]];
                a += flo[j] * v;
                b += fhi[j] * v;
            }
            l[k] = a;
            h[k] = b;
        }
    }
    free(map);
}

/* analysis along the first axis of x[n][cols] -> lo, hi [m][cols] */
void dwt_first(const double *x, int n, int cols, int mode, const double *flo, const double *fhi, int F, double *lo,
               double *hi)
{
    const int m = dwt_out_len(n, F, mode);
    int *map = analysis_map(n, F, mode, m);
    for (int k = 0; k < m; ++k) {
        double *l = lo + (size_t)k * cols, *h = hi + (size_t)k * cols;
        for (int c = 0; c < cols; ++c) l[c] = h[c] = 0.0;
        for (int j = 0; j < F; ++j) {
            const double *xr = x + (size_t)map[(size_t)k * F + j] * cols;
            const double a = flo[j], b = fhi[j];
            for (int c = 0; c < cols; ++c) {
                l[c] += a * xr[c];
                h[c] += b * xr[c];
            }
        }
    }
    free(map);
}

int idwt_out_len(int m, int F, int mode) { return mode == 2 ? 2 * m : 2 * m - F + 2; }

/* for output n: the taps t with an even numerator and the coefficient index they read (-1: none) */
static int *synthesis_map(int m, int F, int mode, int nout)
{
    int *map = (int *)malloc(sizeof(int) * (size_t)nout * F);
    for (int n = 0; n < nout; ++n)
        for (int t = 0; t < F; ++t) {
            int idx = -1;
            if (mode == 2) {
                const long num = (long)n + F / 2 - 1 - t;
                if ((num & 1) == 0) {
                    long k = (num / 2) % m; /* num may be negative: floor division for even numbers is exact */
                    if (k < 0) k += m;
                    idx = (int)k;
                }
            } else {
                const long num = (long)n + F - 2 - t;
                if (num >= 0 && (num & 1) == 0 && num / 2 < m) idx = (int)(num / 2);
            }
            map[(size_t)n * F + t] = idx;
        }
    return map;
}

/* synthesis along the last axis: ca, cd [rows][m] -> out [rows][nout] */
void idwt_last(const double *ca, const double *cd, int rows, int m, int mode, const double *glo, const double *ghi,
               int F, double *out)
{
    const int nout = idwt_out_len(m, F, mode);
    int *map = synthesis_map(m, F, mode, nout);
    for (int r = 0; r < rows; ++r) {
        const double *a = ca + (size_t)r * m, *d = cd + (size_t)r * m;
        double *o = out + (size_t)r * nout;
        for (int n = 0; n < nout; ++n) {
            double s = 0.0;
            const int *mn = map + (size_t)n * F;
            for (int t = 0; t < F; ++t)
                if (mn[t] >= 0) s += glo[t] * a[mn[t]] + ghi[t] * d[mn[t]];
            o[n] = s;
        }
    }
    free(map);
}

/* synthesis along the first axis: ca, cd [m][cols] -> out [nout][cols] */
void idwt_first(const double *ca, const double *cd, int m, int cols, int mode, const double *glo, const double *ghi,
                int F, double *out)
{
    const int nout = idwt_out_len(m, F, mode);
    int *map = synthesis_map(m, F, mode, nout);
    for (int n = 0; n < nout; ++n) {
        double *o = out + (size_t)n * cols;
        for (int c = 0; c < cols; ++c) o[c] = 0.0;
        for (int t = 0; t < F; ++t) {
            const int k = map[(size_t)n * F + t];
            if (k < 0) continue;
            const double *a = ca + (size_t)k * cols, *d = cd + (size_t)k * cols;
            const double gl = glo[t], gh = ghi[t];
            for (int c = 0; c < cols; ++c) o[c] += gl * a[c] + gh * d[c];
        }
    }
    free(map);
}
