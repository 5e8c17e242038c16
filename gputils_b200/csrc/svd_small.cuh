// Sequential, LAPACK-faithful SVD of a small upper-triangular n x n matrix R:
//     R = Ur * diag(S) * Vt ,  S >= 0 descending.
// One GPU thread runs this per matrix (k_svd_bidiag in svd.cu). The reference's test
// `singularValuesMultipleMatrices` (testTensor.cu:1126-1171) pins the signs and the null-space basis
// that LAPACK's gesvd produces, i.e. the outcome of Householder bidiagonalisation (gebd2 + orgbr)
// followed by the implicit zero-shift / shifted QR iteration of bdsqr, so this file restates those
// published algorithms (LAPACK Users' Guide; Demmel & Kahan 1990 for the zero-shift sweep and the 2x2
// kernels lasv2 / las2) step for step: same sign rule, same deflation tests, same sweep direction
// choice, same final sign fix-up and ordering. (ref: tensor.cuh:1624-1676 calls cusolverDn?gesvd.)
//
// The code is __host__ __device__ and free of CUDA intrinsics so that tests/ can compile it with the host
// compiler and compare it with scipy's dgesvd without a GPU; the shipped library only instantiates it
// inside CUDA kernels. Arithmetic is written with explicit fma() where LAPACK's reference BLAS would
// fuse, and svd.cu is compiled with --fmad=false so host and device follow the same rounding sequence.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GPUB_HD __host__ __device__
#else
#define GPUB_HD
#endif

namespace gpub_svd {

template<typename T> struct Eps;
template<> struct Eps<double> {
    GPUB_HD static double eps() { return 1.1102230246251565e-16; }   // 2^-53 (dlamch 'E')
    GPUB_HD static double safmin() { return 2.2250738585072014e-308; }
};
template<> struct Eps<float> {
    GPUB_HD static float eps() { return 5.9604645e-08f; }             // 2^-24
    GPUB_HD static float safmin() { return 1.17549435e-38f; }
};

template<typename T> GPUB_HD inline T t_abs(T x) { return x < T(0) ? -x : x; }
template<typename T> GPUB_HD inline T t_max(T a, T b) { return a > b ? a : b; }
template<typename T> GPUB_HD inline T t_min(T a, T b) { return a < b ? a : b; }
// Fortran SIGN(a, b): |a| with the sign of b (b == +0 counts as positive)
template<typename T> GPUB_HD inline T t_sign(T a, T b) {
    T aa = t_abs(a);
    return (b < T(0) || (b == T(0) && signbit(b))) ? -aa : aa;
}
GPUB_HD inline double t_sqrt(double x) { return sqrt(x); }
GPUB_HD inline float t_sqrt(float x) { return sqrtf(x); }
GPUB_HD inline double t_fma(double a, double b, double c) { return fma(a, b, c); }
GPUB_HD inline float t_fma(float a, float b, float c) { return fmaf(a, b, c); }
GPUB_HD inline double t_pow(double a, double b) { return pow(a, b); }
GPUB_HD inline float t_pow(float a, float b) { return powf(a, b); }

// sqrt(x^2 + y^2) without unnecessary overflow (lapy2)
template<typename T> GPUB_HD inline T lapy2(T x, T y) {
    T xa = t_abs(x), ya = t_abs(y);
    T w = t_max(xa, ya), z = t_min(xa, ya);
    if (z == T(0)) return w;
    T q = z / w;
    return w * t_sqrt(T(1) + q * q);
}

// Householder generator (larfg): alpha <- beta, x <- v (scaled in place), returns tau
template<typename T>
GPUB_HD inline T larfg(int n, T *alpha, T *x, long incx) {
    if (n <= 1) return T(0);
    // nrm2 with scaling
    T scale = T(0), ssq = T(1);
    for (int i = 0; i < n - 1; i++) {
        T v = x[i * incx];
        if (v != T(0)) {
            T a = t_abs(v);
            if (scale < a) {
                T q = scale / a;
                ssq = T(1) + ssq * q * q;
                scale = a;
            } else {
                T q = a / scale;
                ssq += q * q;
            }
        }
    }
    T xnorm = scale * t_sqrt(ssq);
    if (xnorm == T(0)) return T(0);
    T beta = -t_sign(lapy2(*alpha, xnorm), *alpha);
    // dlarfg's guard: while |beta| is below safmin / eps, x, alpha and beta are scaled up by its reciprocal (the reciprocal of
    // alpha - beta would overflow otherwise: an all-ones matrix in fp32 gets there in a dozen columns, each being rounding noise
    // of the one before)
    const T safmin = sizeof(T) == 8 ? T(2.0041683600089728e-292) : T(1.9721522630525295e-31);
    const T rsafmn = T(1) / safmin;
    int knt = 0;
    T a = *alpha;
    while (t_abs(beta) < safmin && knt < 20) {
        knt++;
        for (int i = 0; i < n - 1; i++) x[i * incx] *= rsafmn;
        beta *= rsafmn;
        a *= rsafmn;
    }
    if (knt > 0) {
        // recompute the norm of the scaled vector (plain: it is in range now)
        T ss = T(0);
        for (int i = 0; i < n - 1; i++) ss += x[i * incx] * x[i * incx];
        xnorm = t_sqrt(ss);
        if (xnorm == T(0)) { *alpha = a; for (int j = 0; j < knt; j++) *alpha *= safmin; return T(0); }
        beta = -t_sign(lapy2(a, xnorm), a);
    }
    T tau = (beta - a) / beta;
    T s = T(1) / (a - beta);
    for (int i = 0; i < n - 1; i++) x[i * incx] *= s;
    for (int j = 0; j < knt; j++) beta *= safmin;
    *alpha = beta;
    return tau;
}

// plane rotation generator (lartg, LAPACK >= 3.10 convention: c >= 0, r carries the sign of f)
template<typename T>
GPUB_HD inline void lartg(T f, T g, T *c, T *s, T *r) {
    if (g == T(0)) {
        *c = T(1);
        *s = T(0);
        *r = f;
    } else if (f == T(0)) {
        *c = T(0);
        *s = t_sign(T(1), g);
        *r = t_abs(g);
    } else {
        T d = lapy2(f, g);
        *c = t_abs(f) / d;
        *r = t_sign(d, f);
        *s = g / *r;
    }
}

// singular values of the 2x2 upper triangular [f g; 0 h] (las2)
template<typename T>
GPUB_HD inline void las2(T f, T g, T h, T *ssmin, T *ssmax) {
    T fa = t_abs(f), ga = t_abs(g), ha = t_abs(h);
    T fhmn = t_min(fa, ha), fhmx = t_max(fa, ha);
    if (fhmn == T(0)) {
        *ssmin = T(0);
        if (fhmx == T(0)) {
            *ssmax = ga;
        } else {
            T mx = t_max(fhmx, ga), mn = t_min(fhmx, ga);
            *ssmax = mx * t_sqrt(T(1) + (mn / mx) * (mn / mx));
        }
    } else if (ga < fhmx) {
        T as = T(1) + fhmn / fhmx, at = (fhmx - fhmn) / fhmx, au = (ga / fhmx) * (ga / fhmx);
        T c = T(2) / (t_sqrt(as * as + au) + t_sqrt(at * at + au));
        *ssmin = fhmn * c;
        *ssmax = fhmx / c;
    } else {
        T au = fhmx / ga;
        if (au == T(0)) {
            *ssmin = (fhmn * fhmx) / ga;
            *ssmax = ga;
        } else {
            T as = T(1) + fhmn / fhmx, at = (fhmx - fhmn) / fhmx;
            T c = T(1) / (t_sqrt(T(1) + (as * au) * (as * au)) + t_sqrt(T(1) + (at * au) * (at * au)));
            *ssmin = (fhmn * c) * au;
            *ssmin = *ssmin + *ssmin;
            *ssmax = ga / (c + c);
        }
    }
}

// SVD of the 2x2 upper triangular [f g; 0 h] (lasv2):
// [csl snl; -snl csl] [f g; 0 h] [csr -snr; snr csr] = diag(ssmax, ssmin)
template<typename T>
GPUB_HD inline void lasv2(T f, T g, T h, T *ssmin, T *ssmax, T *snr, T *csr, T *snl, T *csl) {
    T ft = f, fa = t_abs(f), ht = h, ha = t_abs(h);
    int pmax = 1;
    bool swap = ha > fa;
    if (swap) {
        pmax = 3;
        T t = ft; ft = ht; ht = t;
        t = fa; fa = ha; ha = t;
    }
    T gt = g, ga = t_abs(g);
    T clt, crt, slt, srt;
    if (ga == T(0)) {
        *ssmin = ha;
        *ssmax = fa;
        clt = T(1); crt = T(1); slt = T(0); srt = T(0);
    } else {
        bool gasmal = true;
        if (ga > fa) {
            pmax = 2;
            if ((fa / ga) < Eps<T>::eps()) {
                gasmal = false;
                *ssmax = ga;
                if (ha > T(1)) *ssmin = fa / (ga / ha);
                else *ssmin = (fa / ga) * ha;
                clt = T(1);
                slt = ht / gt;
                srt = T(1);
                crt = ft / gt;
            }
        }
        if (gasmal) {
            T d = fa - ha;
            T l = (d == fa) ? T(1) : d / fa;
            T m = gt / ft;
            T t = T(2) - l;
            T mm = m * m, tt = t * t;
            T s = t_sqrt(tt + mm);
            T r = (l == T(0)) ? t_abs(m) : t_sqrt(l * l + mm);
            T a = T(0.5) * (s + r);
            *ssmin = ha / a;
            *ssmax = fa * a;
            if (mm == T(0)) {
                if (l == T(0)) t = t_sign(T(2), ft) * t_sign(T(1), gt);
                else t = gt / t_sign(d, ft) + m / t;
            } else {
                t = (m / (s + t) + m / (r + l)) * (T(1) + a);
            }
            l = t_sqrt(t * t + T(4));
            crt = T(2) / l;
            srt = t / l;
            clt = (crt + srt * m) / a;
            slt = (ht / ft) * srt / a;
        }
    }
    if (swap) {
        *csl = srt; *snl = crt; *csr = slt; *snr = clt;
    } else {
        *csl = clt; *snl = slt; *csr = crt; *snr = srt;
    }
    T tsign;
    if (pmax == 1) tsign = t_sign(T(1), *csr) * t_sign(T(1), *csl) * t_sign(T(1), f);
    else if (pmax == 2) tsign = t_sign(T(1), *snr) * t_sign(T(1), *csl) * t_sign(T(1), g);
    else tsign = t_sign(T(1), *snr) * t_sign(T(1), *snl) * t_sign(T(1), h);
    *ssmax = t_sign(*ssmax, tsign);
    *ssmin = t_sign(*ssmin, tsign * t_sign(T(1), f) * t_sign(T(1), h));
}

// column-major accessor
#define GPUB_AT(M, ld, i, j) (M)[(long) (i) + (long) (j) * (long) (ld)]

// apply H = I - tau v v^T (v[0] = 1 implicit, v[1:] = vtail with stride incv) from the left to C (rows x cols)
template<typename T>
GPUB_HD inline void apply_left(int rows, int cols, const T *vtail, long incv, T tau, T *C, long ldc) {
    if (tau == T(0)) return;
    for (int c = 0; c < cols; c++) {
        T *cc = C + (long) c * ldc;
        T w = cc[0];
        for (int r = 1; r < rows; r++) w = t_fma(vtail[(r - 1) * incv], cc[r], w);
        T tw = tau * w;
        cc[0] -= tw;
        for (int r = 1; r < rows; r++) cc[r] = t_fma(-tw, vtail[(r - 1) * incv], cc[r]);
    }
}

// apply H from the right to C (rows x cols): C <- C (I - tau v v^T), v over the columns
template<typename T>
GPUB_HD inline void apply_right(int rows, int cols, const T *vtail, long incv, T tau, T *C, long ldc) {
    if (tau == T(0)) return;
    for (int r = 0; r < rows; r++) {
        T w = C[r];
        for (int c = 1; c < cols; c++) w = t_fma(C[r + (long) c * ldc], vtail[(c - 1) * incv], w);
        T tw = tau * w;
        C[r] -= tw;
        for (int c = 1; c < cols; c++) C[r + (long) c * ldc] = t_fma(-tw, vtail[(c - 1) * incv], C[r + (long) c * ldc]);
    }
}

// y-rows rotation used by bdsqr on Vt: rows p and q of an (n x ncols) matrix
template<typename T>
GPUB_HD inline void rot_rows(int ncols, T *M, long ld, int p, int q, T c, T s) {
    for (int j = 0; j < ncols; j++) {
        T x = GPUB_AT(M, ld, p, j), y = GPUB_AT(M, ld, q, j);
        GPUB_AT(M, ld, p, j) = c * x + s * y;
        GPUB_AT(M, ld, q, j) = c * y - s * x;
    }
}
template<typename T>
GPUB_HD inline void rot_cols(int nrows, T *M, long ld, int p, int q, T c, T s) {
    for (int i = 0; i < nrows; i++) {
        T x = GPUB_AT(M, ld, i, p), y = GPUB_AT(M, ld, i, q);
        GPUB_AT(M, ld, i, p) = c * x + s * y;
        GPUB_AT(M, ld, i, q) = c * y - s * x;
    }
}

// lasr 'L','V',dir on rows lo..hi of Vt (ncols columns) and lasr 'R','V',dir on columns lo..hi of U
template<typename T>
GPUB_HD inline void lasr_rows(bool forward, int lo, int hi, int ncols, const T *c, const T *s, T *M, long ld) {
    const int cnt = hi - lo; // number of rotations
    for (int t = 0; t < cnt; t++) {
        const int j = forward ? t : cnt - 1 - t;
        const T ct = c[j], st = s[j];
        if (ct == T(1) && st == T(0)) continue;
        for (int col = 0; col < ncols; col++) {
            T temp = GPUB_AT(M, ld, lo + j + 1, col);
            GPUB_AT(M, ld, lo + j + 1, col) = ct * temp - st * GPUB_AT(M, ld, lo + j, col);
            GPUB_AT(M, ld, lo + j, col) = st * temp + ct * GPUB_AT(M, ld, lo + j, col);
        }
    }
}
template<typename T>
GPUB_HD inline void lasr_cols(bool forward, int lo, int hi, int nrows, const T *c, const T *s, T *M, long ld) {
    const int cnt = hi - lo;
    for (int t = 0; t < cnt; t++) {
        const int j = forward ? t : cnt - 1 - t;
        const T ct = c[j], st = s[j];
        if (ct == T(1) && st == T(0)) continue;
        for (int row = 0; row < nrows; row++) {
            T temp = GPUB_AT(M, ld, row, lo + j + 1);
            GPUB_AT(M, ld, row, lo + j + 1) = ct * temp - st * GPUB_AT(M, ld, row, lo + j);
            GPUB_AT(M, ld, row, lo + j) = st * temp + ct * GPUB_AT(M, ld, row, lo + j);
        }
    }
}

// bdsqr, upper bidiagonal (d[0..n-1], e[0..n-2]); Vt (n x ncvt) <- P^T Vt, U (nru x n) <- U Q.
// work: 4*n entries. Returns 0 on convergence, else the number of unconverged superdiagonals.
template<typename T>
GPUB_HD inline int bdsqr(int n, int ncvt, int nru, T *d, T *e, T *Vt, long ldvt, T *U, long ldu, T *work) {
    if (n == 0) return 0;
    const T eps = Eps<T>::eps(), unfl = Eps<T>::safmin();
    const int maxitr = 6;
    int info = 0;
    if (n > 1) {
        T *w_c1 = work, *w_s1 = work + n, *w_c2 = work + 2 * n, *w_s2 = work + 3 * n;
        const T tolmul = t_max(T(10), t_min(T(100), t_pow(eps, T(-0.125))));
        const T tol = tolmul * eps;
        T smax = T(0);
        for (int i = 0; i < n; i++) smax = t_max(smax, t_abs(d[i]));
        for (int i = 0; i < n - 1; i++) smax = t_max(smax, t_abs(e[i]));
        T sminl = T(0), sminoa = t_abs(d[0]);
        if (sminoa != T(0)) {
            T mu = sminoa;
            for (int i = 1; i < n; i++) {
                mu = t_abs(d[i]) * (mu / (mu + t_abs(e[i - 1])));
                sminoa = t_min(sminoa, mu);
                if (sminoa == T(0)) break;
            }
        }
        sminoa = sminoa / t_sqrt((T) n);
        const T thresh = t_max(tol * sminoa, (T) (maxitr * n) * ((T) n * unfl));
        const long maxit = (long) maxitr * n * n;
        long iter = 0;
        int oldll = -1, oldm = -1, idir = 0;
        int m = n - 1; // 0-based index of the bottom of the active block
        while (m > 0) {
            if (iter > maxit) {
                for (int i = 0; i < n - 1; i++)
                    if (e[i] != T(0)) info++;
                break;
            }
            // find a diagonal block to work on
            T smx = t_abs(d[m]);
            int ll = -1;
            bool split = false;
            for (int l = m - 1; l >= 0; l--) {
                T abss = t_abs(d[l]), abse = t_abs(e[l]);
                if (abse <= thresh) {
                    e[l] = T(0);
                    ll = l;
                    split = true;
                    break;
                }
                smx = t_max(smx, t_max(abss, abse));
            }
            if (split) {
                if (ll == m - 1) { // bottom singular value converged
                    m -= 1;
                    continue;
                }
                ll += 1;
            } else {
                ll = 0;
            }
            // active block d[ll..m], e[ll..m-1]
            if (ll == m - 1) {
                T sigmn, sigmx, sinr, cosr, sinl, cosl;
                lasv2(d[m - 1], e[m - 1], d[m], &sigmn, &sigmx, &sinr, &cosr, &sinl, &cosl);
                d[m - 1] = sigmx;
                e[m - 1] = T(0);
                d[m] = sigmn;
                if (ncvt > 0) rot_rows(ncvt, Vt, ldvt, m - 1, m, cosr, sinr);
                if (nru > 0) rot_cols(nru, U, ldu, m - 1, m, cosl, sinl);
                m -= 2;
                continue;
            }
            // direction: chase the bulge from the larger end to the smaller
            if (ll > oldm || m < oldll) idir = (t_abs(d[ll]) >= t_abs(d[m])) ? 1 : 2;
            // convergence tests (relative accuracy)
            bool restart = false;
            if (idir == 1) {
                if (t_abs(e[m - 1]) <= tol * t_abs(d[m])) {
                    e[m - 1] = T(0);
                    continue;
                }
                T mu = t_abs(d[ll]);
                sminl = mu;
                for (int l = ll; l < m; l++) {
                    if (t_abs(e[l]) <= tol * mu) {
                        e[l] = T(0);
                        restart = true;
                        break;
                    }
                    mu = t_abs(d[l + 1]) * (mu / (mu + t_abs(e[l])));
                    sminl = t_min(sminl, mu);
                }
            } else {
                if (t_abs(e[ll]) <= tol * t_abs(d[ll])) {
                    e[ll] = T(0);
                    continue;
                }
                T mu = t_abs(d[m]);
                sminl = mu;
                for (int l = m - 1; l >= ll; l--) {
                    if (t_abs(e[l]) <= tol * mu) {
                        e[l] = T(0);
                        restart = true;
                        break;
                    }
                    mu = t_abs(d[l]) * (mu / (mu + t_abs(e[l])));
                    sminl = t_min(sminl, mu);
                }
            }
            if (restart) continue;
            oldll = ll;
            oldm = m;
            // shift
            T shift = T(0), r;
            if ((T) n * tol * (sminl / smx) <= t_max(eps, T(0.01) * tol)) {
                shift = T(0);
            } else {
                T sll;
                if (idir == 1) {
                    sll = t_abs(d[ll]);
                    las2(d[m - 1], e[m - 1], d[m], &shift, &r);
                } else {
                    sll = t_abs(d[m]);
                    las2(d[ll], e[ll], d[ll + 1], &shift, &r);
                }
                if (sll > T(0) && (shift / sll) * (shift / sll) < eps) shift = T(0);
            }
            iter += m - ll;
            const int cnt = m - ll; // rotations in this sweep
            if (shift == T(0)) {
                if (idir == 1) {
                    T cs = T(1), oldcs = T(1), sn = T(0), oldsn = T(0);
                    for (int i = ll; i < m; i++) {
                        lartg(d[i] * cs, e[i], &cs, &sn, &r);
                        if (i > ll) e[i - 1] = oldsn * r;
                        lartg(oldcs * r, d[i + 1] * sn, &oldcs, &oldsn, &d[i]);
                        w_c1[i - ll] = cs; w_s1[i - ll] = sn; w_c2[i - ll] = oldcs; w_s2[i - ll] = oldsn;
                    }
                    T h = d[m] * cs;
                    d[m] = h * oldcs;
                    e[m - 1] = h * oldsn;
                    if (ncvt > 0) lasr_rows(true, ll, ll + cnt, ncvt, w_c1, w_s1, Vt, ldvt);
                    if (nru > 0) lasr_cols(true, ll, ll + cnt, nru, w_c2, w_s2, U, ldu);
                    if (t_abs(e[m - 1]) <= thresh) e[m - 1] = T(0);
                } else {
                    T cs = T(1), oldcs = T(1), sn = T(0), oldsn = T(0);
                    for (int i = m; i > ll; i--) {
                        lartg(d[i] * cs, e[i - 1], &cs, &sn, &r);
                        if (i < m) e[i] = oldsn * r;
                        lartg(oldcs * r, d[i - 1] * sn, &oldcs, &oldsn, &d[i]);
                        w_c1[i - ll - 1] = cs; w_s1[i - ll - 1] = -sn; w_c2[i - ll - 1] = oldcs; w_s2[i - ll - 1] = -oldsn;
                    }
                    T h = d[ll] * cs;
                    d[ll] = h * oldcs;
                    e[ll] = h * oldsn;
                    if (ncvt > 0) lasr_rows(false, ll, ll + cnt, ncvt, w_c2, w_s2, Vt, ldvt);
                    if (nru > 0) lasr_cols(false, ll, ll + cnt, nru, w_c1, w_s1, U, ldu);
                    if (t_abs(e[ll]) <= thresh) e[ll] = T(0);
                }
            } else {
                if (idir == 1) {
                    T f = (t_abs(d[ll]) - shift) * (t_sign(T(1), d[ll]) + shift / d[ll]);
                    T g = e[ll];
                    for (int i = ll; i < m; i++) {
                        T cosr, sinr, cosl, sinl;
                        lartg(f, g, &cosr, &sinr, &r);
                        if (i > ll) e[i - 1] = r;
                        f = cosr * d[i] + sinr * e[i];
                        e[i] = cosr * e[i] - sinr * d[i];
                        g = sinr * d[i + 1];
                        d[i + 1] = cosr * d[i + 1];
                        lartg(f, g, &cosl, &sinl, &r);
                        d[i] = r;
                        f = cosl * e[i] + sinl * d[i + 1];
                        d[i + 1] = cosl * d[i + 1] - sinl * e[i];
                        if (i < m - 1) {
                            g = sinl * e[i + 1];
                            e[i + 1] = cosl * e[i + 1];
                        }
                        w_c1[i - ll] = cosr; w_s1[i - ll] = sinr; w_c2[i - ll] = cosl; w_s2[i - ll] = sinl;
                    }
                    e[m - 1] = f;
                    if (ncvt > 0) lasr_rows(true, ll, ll + cnt, ncvt, w_c1, w_s1, Vt, ldvt);
                    if (nru > 0) lasr_cols(true, ll, ll + cnt, nru, w_c2, w_s2, U, ldu);
                    if (t_abs(e[m - 1]) <= thresh) e[m - 1] = T(0);
                } else {
                    T f = (t_abs(d[m]) - shift) * (t_sign(T(1), d[m]) + shift / d[m]);
                    T g = e[m - 1];
                    for (int i = m; i > ll; i--) {
                        T cosr, sinr, cosl, sinl;
                        lartg(f, g, &cosr, &sinr, &r);
                        if (i < m) e[i] = r;
                        f = cosr * d[i] + sinr * e[i - 1];
                        e[i - 1] = cosr * e[i - 1] - sinr * d[i];
                        g = sinr * d[i - 1];
                        d[i - 1] = cosr * d[i - 1];
                        lartg(f, g, &cosl, &sinl, &r);
                        d[i] = r;
                        f = cosl * e[i - 1] + sinl * d[i - 1];
                        d[i - 1] = cosl * d[i - 1] - sinl * e[i - 1];
                        if (i > ll + 1) {
                            g = sinl * e[i - 2];
                            e[i - 2] = cosl * e[i - 2];
                        }
                        w_c1[i - ll - 1] = cosr; w_s1[i - ll - 1] = -sinr; w_c2[i - ll - 1] = cosl; w_s2[i - ll - 1] = -sinl;
                    }
                    e[ll] = f;
                    if (t_abs(e[ll]) <= thresh) e[ll] = T(0);
                    if (ncvt > 0) lasr_rows(false, ll, ll + cnt, ncvt, w_c2, w_s2, Vt, ldvt);
                    if (nru > 0) lasr_cols(false, ll, ll + cnt, nru, w_c1, w_s1, U, ldu);
                }
            }
        }
    }
    // make singular values non-negative
    for (int i = 0; i < n; i++) {
        if (d[i] == T(0)) d[i] = T(0); // clears -0
        if (d[i] < T(0)) {
            d[i] = -d[i];
            for (int j = 0; j < ncvt; j++) GPUB_AT(Vt, ldvt, i, j) = -GPUB_AT(Vt, ldvt, i, j);
        }
    }
    // sort into decreasing order (selection of the smallest into the last free slot, as bdsqr does)
    for (int i = 0; i < n - 1; i++) {
        int isub = 0;
        T smin = d[0];
        for (int j = 1; j < n - i; j++) {
            if (d[j] <= smin) {
                isub = j;
                smin = d[j];
            }
        }
        const int last = n - 1 - i;
        if (isub != last) {
            d[isub] = d[last];
            d[last] = smin;
            for (int j = 0; j < ncvt; j++) {
                T t = GPUB_AT(Vt, ldvt, isub, j);
                GPUB_AT(Vt, ldvt, isub, j) = GPUB_AT(Vt, ldvt, last, j);
                GPUB_AT(Vt, ldvt, last, j) = t;
            }
            for (int r = 0; r < nru; r++) {
                T t = GPUB_AT(U, ldu, r, isub);
                GPUB_AT(U, ldu, r, isub) = GPUB_AT(U, ldu, r, last);
                GPUB_AT(U, ldu, r, last) = t;
            }
        }
    }
    return info;
}

// SVD of the n x n matrix G (upper triangular on entry, destroyed):
//   S[n], Vt (n x n, ldvt), Ur (n x n, ldur; only if want_u), scratch: 6*n entries.
template<typename T>
GPUB_HD inline int svd_upper_small(int n, T *G, long ldg, T *S, T *Vt, long ldvt, T *Ur, long ldur, bool want_u, T *scratch) {
    T *e = scratch, *tauq = scratch + n, *taup = scratch + 2 * n, *work = scratch; // work reuses all 6n after generation
    // gebd2: G <- Q^T G P, upper bidiagonal
    for (int i = 0; i < n; i++) {
        // H_i annihilates G(i+1:n-1, i)
        tauq[i] = larfg<T>(n - i, &GPUB_AT(G, ldg, i, i), &GPUB_AT(G, ldg, (i + 1 < n ? i + 1 : i), i), 1);
        S[i] = GPUB_AT(G, ldg, i, i);
        if (i + 1 < n) apply_left<T>(n - i, n - i - 1, &GPUB_AT(G, ldg, i + 1, i), 1, tauq[i], &GPUB_AT(G, ldg, i, i + 1), ldg);
        if (i < n - 1) {
            // G_i annihilates G(i, i+2:n-1)
            taup[i] = larfg<T>(n - i - 1, &GPUB_AT(G, ldg, i, i + 1), &GPUB_AT(G, ldg, i, (i + 2 < n ? i + 2 : i + 1)), ldg);
            e[i] = GPUB_AT(G, ldg, i, i + 1);
            if (i + 1 < n)
                apply_right<T>(n - i - 1, n - i - 1, &GPUB_AT(G, ldg, i, (i + 2 < n ? i + 2 : i + 1)), ldg, taup[i],
                               &GPUB_AT(G, ldg, i + 1, i + 1), ldg);
        } else {
            taup[i] = T(0);
        }
    }
    // orgbr 'P': Vt = P^T = G_{n-2} ... G_1 G_0, accumulated like orgl2: X <- X G_i for i = n-2 .. 0.
    // When G_i is applied only the trailing block X(i+1:, i+1:) differs from the identity.
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) GPUB_AT(Vt, ldvt, i, j) = (i == j) ? T(1) : T(0);
    for (int i = n - 2; i >= 0; i--) {
        if (taup[i] == T(0)) continue;
        const int len = n - i - 1;
        apply_right<T>(len, len, &GPUB_AT(G, ldg, i, (i + 2 < n ? i + 2 : i + 1)), ldg, taup[i], &GPUB_AT(Vt, ldvt, i + 1, i + 1), ldvt);
    }
    // orgbr 'Q': Ur = H_0 H_1 ... H_{n-1} applied to the identity, last first
    if (want_u) {
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) GPUB_AT(Ur, ldur, i, j) = (i == j) ? T(1) : T(0);
        for (int i = n - 1; i >= 0; i--) {
            if (tauq[i] == T(0) || i + 1 >= n) continue;
            apply_left<T>(n - i, n - i, &GPUB_AT(G, ldg, i + 1, i), 1, tauq[i], &GPUB_AT(Ur, ldur, i, i), ldur);
        }
    }
    // move e out of the scratch that bdsqr uses for its rotations: keep e in scratch[4n..5n)
    T *e2 = scratch + 4 * n;
    for (int i = 0; i < n - 1; i++) e2[i] = e[i];
    return bdsqr<T>(n, n, want_u ? n : 0, S, e2, Vt, ldvt, Ur, ldur, work);
}

// Full small gesvd, one matrix, sequential (m >= n): A (m x n, destroyed) = U diag(S) Vt.
// Follows gesvd's QR-first path: geqr2, org2r (full m x m Q when want_u), SVD of R, U(:, 0:n) <- Q(:, 0:n) Ur.
// scratch: 2*n*n + 7*n entries.
template<typename T>
GPUB_HD inline int gesvd_small(int m, int n, T *A, long lda, T *S, T *U, long ldu, T *Vt, long ldvt, bool want_u, T *scratch) {
    T *G = scratch, *Ur = scratch + (long) n * n, *tau = scratch + 2L * n * n, *rest = tau + n;
    for (int j = 0; j < n; j++) {
        tau[j] = larfg<T>(m - j, &GPUB_AT(A, lda, j, j), &GPUB_AT(A, lda, (j + 1 < m ? j + 1 : j), j), 1);
        if (j + 1 < n && j + 1 < m)
            apply_left<T>(m - j, n - j - 1, &GPUB_AT(A, lda, j + 1, j), 1, tau[j], &GPUB_AT(A, lda, j, j + 1), lda);
    }
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) GPUB_AT(G, n, i, j) = (i <= j) ? GPUB_AT(A, lda, i, j) : T(0);
    if (want_u) {
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) GPUB_AT(U, ldu, i, j) = (i == j) ? T(1) : T(0);
        for (int j = n - 1; j >= 0; j--) {
            if (j + 1 >= m) continue;
            apply_left<T>(m - j, m - j, &GPUB_AT(A, lda, j + 1, j), 1, tau[j], &GPUB_AT(U, ldu, j, j), ldu);
        }
    }
    int info = svd_upper_small<T>(n, G, n, S, Vt, ldvt, Ur, n, want_u, rest);
    if (want_u) {
        // U(:, 0:n) <- U(:, 0:n) * Ur, one row at a time through A's first row block as temporary
        for (int i = 0; i < m; i++) {
            for (int j = 0; j < n; j++) {
                T acc = T(0);
                for (int l = 0; l < n; l++) acc = t_fma(GPUB_AT(U, ldu, i, l), GPUB_AT(Ur, n, l, j), acc);
                G[j] = acc; // G is free after the bidiagonalisation
            }
            for (int j = 0; j < n; j++) GPUB_AT(U, ldu, i, j) = G[j];
        }
    }
    return info;
}

} // namespace gpub_svd
