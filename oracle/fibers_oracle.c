/*
 * CPU restatement (C + OpenMP) of the Fibers.jl voxel loops -- TEST INFRASTRUCTURE / CPU BASELINE.
 *
 * Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs load this library (oracle/_build/libfibers_oracle.so).
 * PARITY UNPINNED: the reference ships no tests or golden vectors and Julia is not installed,
 * so this port is validated against oracle/fibers_oracle.py (numpy restatement) only.
 *
 * It mirrors the STRUCTURE of the reference's hot loops so that timing it is a fair stand-in for
 * the multithreaded Julia path: `Threads.@threads for iz` (static z partition) with per-thread
 * scratch (src/gqi.jl:132-162, src/dti.jl:258-275, src/dsi.jl:197-261), a strided per-voxel gather
 * of the DWI series, one GEMV per voxel against the precomputed matrix, face-based local-maximum
 * suppression and a full stable descending sort (find_peaks!, src/gqi.jl:180-201), and the serial
 * odfmax post-pass (src/gqi.jl:164-168).  Matrices (pinv(A), sinc matrix, DSI tables) are built by
 * the numpy oracle and passed in.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NPEAK 3

static int nthreads_eff(int req) {
#ifdef _OPENMP
    return req > 0 ? req : omp_get_max_threads();
#else
    (void)req; return 1;
#endif
}

int oracle_max_threads(void) { return nthreads_eff(0); }

/* Dot product the way an optimised BLAS sgemv computes it (OpenBLAS behind Julia's mul!,
 * src/gqi.jl:144, src/dti.jl:296): 16 independent partial sums (vector lanes) with fused
 * multiply-add, reduced at the end.  Keeps the CPU baseline honest: a strictly sequential scalar
 * sum would not vectorise and would understate the reference's speed. */
static inline float dot_blas(const float* __restrict a, const float* __restrict b, int n) {
    float acc[16] = {0};
    int j = 0;
    for (; j + 16 <= n; j += 16)
        for (int l = 0; l < 16; ++l) acc[l] = __builtin_fmaf(a[j + l], b[j + l], acc[l]);
    float tail = 0.f;
    for (; j < n; ++j) tail = __builtin_fmaf(a[j], b[j], tail);
    for (int w = 8; w > 0; w >>= 1) for (int l = 0; l < w; ++l) acc[l] += acc[l + w];
    return acc[0] + tail;
}

/* ---------------------------------------------------------------- 3x3 symmetric eigen (StaticArrays closed form) */
static void cross3(const float a[3], const float b[3], float c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* values ascending in w, k-th eigenvector = column k of v (v[r][k]); src/dti.jl:311 */
static void eig3_sym(float a11, float a12, float a13, float a22, float a23, float a33, float w[3], float v[3][3]) {
    float p1 = a12 * a12 + a13 * a13 + a23 * a23;
    if (p1 == 0.f) {
        int o[3];
        if (a11 < a22) {
            if (a22 < a33) { o[0] = 0; o[1] = 1; o[2] = 2; }
            else if (a33 < a11) { o[0] = 2; o[1] = 0; o[2] = 1; }
            else { o[0] = 0; o[1] = 2; o[2] = 1; }
        } else {
            if (a11 < a33) { o[0] = 1; o[1] = 0; o[2] = 2; }
            else if (a33 < a22) { o[0] = 2; o[1] = 1; o[2] = 0; }
            else { o[0] = 1; o[1] = 2; o[2] = 0; }
        }
        float d[3] = {a11, a22, a33};
        for (int k = 0; k < 3; ++k) { w[k] = d[o[k]]; for (int r = 0; r < 3; ++r) v[r][k] = (r == o[k]) ? 1.f : 0.f; }
        return;
    }
    float q = (a11 + a22 + a33) / 3.f;
    float d11 = a11 - q, d22 = a22 - q, d33 = a33 - q;
    float p2 = d11 * d11 + d22 * d22 + d33 * d33 + 2.f * p1;
    float p = sqrtf(p2 / 6.f), invp = 1.f / p;
    float b11 = d11 * invp, b22 = d22 * invp, b33 = d33 * invp, b12 = a12 * invp, b13 = a13 * invp, b23 = a23 * invp;
    float x1[3] = {b12, b22, b23}, x2[3] = {b13, b23, b33}, c[3];
    cross3(x1, x2, c);
    float r = (b11 * c[0] + b12 * c[1] + b13 * c[2]) / 2.f;
    const float pif = 3.14159274f;
    float phi = r <= -1.f ? pif / 3.f : (r >= 1.f ? 0.f : acosf(r) / 3.f);
    float eig3 = q + 2.f * p * cosf(phi);
    float eig1 = q + 2.f * p * cosf(phi + (2.f * pif / 3.f));
    float eig2 = 3.f * q - eig1 - eig3;
    int swap = r > 0.f;
    float e1 = swap ? eig3 : eig1, e3 = swap ? eig1 : eig3;
    float r1[3] = {a11 - e1, a12, a13}, r2[3] = {a12, a22 - e1, a23}, r3[3] = {a13, a23, a33 - e1};
    float n1 = dot3(r1, r1), n2 = dot3(r2, r2), n3 = dot3(r3, r3);
    float r12[3], r23[3], r31[3];
    cross3(r1, r2, r12); cross3(r2, r3, r23); cross3(r3, r1, r31);
    float n12 = dot3(r12, r12), n23 = dot3(r23, r23), n31 = dot3(r31, r31);
    const float* best; float nb;
    if (n12 * n3 > n23 * n1) { if (n12 * n3 > n31 * n2) { best = r12; nb = n12; } else { best = r31; nb = n31; } }
    else { if (n23 * n1 > n31 * n2) { best = r23; nb = n23; } else { best = r31; nb = n31; } }
    float sn = sqrtf(nb);
    float v1[3] = {best[0] / sn, best[1] / sn, best[2] / sn}, o1[3], o2[3];
    if (fabsf(v1[0]) < fabsf(v1[1])) { float dn = sqrtf(v1[0] * v1[0] + v1[2] * v1[2]); o1[0] = -v1[2] / dn; o1[1] = 0.f; o1[2] = v1[0] / dn; }
    else { float dn = sqrtf(v1[1] * v1[1] + v1[2] * v1[2]); o1[0] = 0.f; o1[1] = v1[2] / dn; o1[2] = -v1[1] / dn; }
    cross3(v1, o1, o2);
    float ao1[3] = {a11 * o1[0] + a12 * o1[1] + a13 * o1[2], a12 * o1[0] + a22 * o1[1] + a23 * o1[2], a13 * o1[0] + a23 * o1[1] + a33 * o1[2]};
    float ao2[3] = {a11 * o2[0] + a12 * o2[1] + a13 * o2[2], a12 * o2[0] + a22 * o2[1] + a23 * o2[2], a13 * o2[0] + a23 * o2[1] + a33 * o2[2]};
    float c11 = dot3(o1, ao1) - eig2, c12 = dot3(o1, ao2), c22 = dot3(o2, ao2) - eig2;
    float s11 = c11 * c11, s12 = c12 * c12, s22 = c22 * c22, pp1 = 1.f, pp2 = 0.f;
    int degen = 0;
    if (s11 >= s22) {
        if (s11 > 0.f || s12 > 0.f) {
            if (s11 >= s12) { float t = c12 / c11; pp2 = 1.f / sqrtf(1.f + t * t); pp1 = t * pp2; }
            else { float t = c11 / c12; pp1 = 1.f / sqrtf(1.f + t * t); pp2 = t * pp1; }
        } else degen = 1;
    } else {
        if (s22 >= s12) { float t = c12 / c22; pp1 = 1.f / sqrtf(1.f + t * t); pp2 = t * pp1; }
        else { float t = c22 / c12; pp2 = 1.f / sqrtf(1.f + t * t); pp1 = t * pp2; }
    }
    float v2[3], v3[3];
    for (int k = 0; k < 3; ++k) v2[k] = degen ? o1[k] : pp1 * o1[k] - pp2 * o2[k];
    cross3(v1, v2, v3);
    w[0] = swap ? e3 : e1; w[1] = eig2; w[2] = swap ? e1 : e3;
    for (int k = 0; k < 3; ++k) { v[k][0] = swap ? v3[k] : v1[k]; v[k][1] = v2[k]; v[k][2] = swap ? v1[k] : v3[k]; }
}

/* ---------------------------------------------------------------- small pinv (double, Hestenes Jacobi) */
static void pinv_small(const double* A, int m, int n, double rtol, double* pA /*[n][m]*/) {
    double* U = (double*)malloc(sizeof(double) * m * n);
    double V[49];
    memcpy(U, A, sizeof(double) * m * n);
    for (int i = 0; i < n * n; ++i) V[i] = 0;
    for (int i = 0; i < n; ++i) V[i * n + i] = 1;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
            double al = 0, be = 0, ga = 0;
            for (int i = 0; i < m; ++i) { double up = U[i * n + p], uq = U[i * n + q]; al += up * up; be += uq * uq; ga += up * uq; }
            double lim = sqrt(al * be);
            if (ga == 0 || fabs(ga) <= 1e-17 * lim) continue;
            if (fabs(ga) / lim > off) off = fabs(ga) / lim;
            double zeta = (be - al) / (2 * ga);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
            double c = 1 / sqrt(1 + t * t), s = c * t;
            for (int i = 0; i < m; ++i) { double up = U[i * n + p], uq = U[i * n + q]; U[i * n + p] = c * up - s * uq; U[i * n + q] = s * up + c * uq; }
            for (int i = 0; i < n; ++i) { double vp = V[i * n + p], vq = V[i * n + q]; V[i * n + p] = c * vp - s * vq; V[i * n + q] = s * vp + c * vq; }
        }
        if (off < 1e-15) break;
    }
    double s2[7], smax = 0;
    for (int k = 0; k < n; ++k) { double s = 0; for (int i = 0; i < m; ++i) s += U[i * n + k] * U[i * n + k]; s2[k] = s; if (s > smax) smax = s; }
    memset(pA, 0, sizeof(double) * n * m);
    for (int k = 0; k < n; ++k) {
        if (!(s2[k] > rtol * rtol * smax) || s2[k] == 0) continue;
        for (int r = 0; r < n; ++r) { double f = V[r * n + k] / s2[k]; for (int i = 0; i < m; ++i) pA[r * m + i] += f * U[i * n + k]; }
    }
    free(U);
}

/* ---------------------------------------------------------------- DTI / ADC (src/dti.jl) */
/* nc = 7: outputs out[0..9] = s0,l1,l2,l3,v1,v2,v3,rd,md,fa ; nc = 2: out[0] = adc, out[1] = s0.
 * A [nvol][nc] row-major, pA [nc][nvol] row-major.  Outputs must be zero-filled by the caller. */
int oracle_linfit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, int nc,
                  const float* A, const float* pA, const uint8_t* ib0, float** out, uint8_t* valid, int nthreads) {
    const int64_t nxy = (int64_t)nx * ny, nvox = nxy * nz;
    nthreads = nthreads_eff(nthreads);
    int err = 0;
#pragma omp parallel num_threads(nthreads)
    {
        float* s = (float*)malloc(sizeof(float) * nvol);
        float* lg = (float*)malloc(sizeof(float) * nvol);
        double* Asub = (double*)malloc(sizeof(double) * nvol * nc);
        double* pAs = (double*)malloc(sizeof(double) * nvol * nc);
#pragma omp for schedule(static)
        for (int iz = 0; iz < nz; ++iz)
            for (int iy = 0; iy < ny; ++iy)
                for (int ix = 0; ix < nx; ++ix) {
                    int64_t vox = ix + (int64_t)nx * (iy + (int64_t)ny * iz);
                    if (mask[vox] == 0) continue;
                    for (int j = 0; j < nvol; ++j) s[j] = dwi[vox + (int64_t)j * nvox];       /* dwi.vol[ix,iy,iz,:] */
                    int npos = 0, b0pos = 0;
                    for (int j = 0; j < nvol; ++j) { int p = s[j] > 0; npos += p; b0pos |= p && ib0[j]; }
                    float d[7] = {0, 0, 0, 0, 0, 0, 0};
                    if (npos == nvol) {
                        for (int j = 0; j < nvol; ++j) lg[j] = logf(s[j]);
                        for (int k = 0; k < nc; ++k) d[k] = dot_blas(pA + (int64_t)k * nvol, lg, nvol);
                    } else if (npos > 6 && b0pos) {
                        int m = 0;
                        for (int j = 0; j < nvol; ++j) if (s[j] > 0) { for (int k = 0; k < nc; ++k) Asub[m * nc + k] = A[j * nc + k]; lg[m] = logf(s[j]); ++m; }
                        pinv_small(Asub, m, nc, 1.1920929e-7 * (m < nc ? m : nc), pAs);
                        for (int k = 0; k < nc; ++k) { double a = 0; for (int j = 0; j < m; ++j) a += (double)(float)pAs[k * m + j] * lg[j]; d[k] = (float)a; }
                    } else continue;
                    if (valid) valid[vox] = 1;
                    if (nc == 2) { out[0][vox] = d[0]; out[1][vox] = expf(d[1]); continue; }
                    float w[3], v[3][3];
                    eig3_sym(d[0], d[1], d[2], d[3], d[4], d[5], w, v);
                    float l1 = w[2], l2 = w[1], l3 = w[0];
                    float rd = l2 + l3, md = (l1 + rd) / 3.f; rd /= 2.f;
                    float fa = sqrtf(((l1 - md) * (l1 - md) + (l2 - md) * (l2 - md) + (l3 - md) * (l3 - md)) / (l1 * l1 + l2 * l2 + l3 * l3) * 1.5f);
                    out[0][vox] = expf(d[6]); out[1][vox] = l1; out[2][vox] = l2; out[3][vox] = l3;
                    for (int r = 0; r < 3; ++r) { out[4][vox + r * nvox] = v[r][2]; out[5][vox + r * nvox] = v[r][1]; out[6][vox + r * nvox] = v[r][0]; }
                    out[7][vox] = rd; out[8][vox] = md; out[9][vox] = fa;
                }
        free(s); free(lg); free(Asub); free(pAs);
    }
    return err;
}

/* ---------------------------------------------------------------- find_peaks! (src/gqi.jl:180-201) */
static void merge_sort_desc(int* idx, int* tmp, const float* key, int n) {   /* stable, descending */
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (key[idx[j]] > key[idx[i]]) ? idx[j++] : idx[i++];
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, sizeof(int) * n);
    }
}

/* faces: folded, 0-based, [nface][3] row-major.  Returns nvalid; isort filled. */
static int find_peaks(const float* o, float* odf_peak, int* isort, int* tmp, int M, const int32_t* faces, int nface) {
    memcpy(odf_peak, o, sizeof(float) * M);
    for (int f = 0; f < nface; ++f) {
        int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
        if (o[b] >= o[a] || o[c] >= o[a]) odf_peak[a] = 0;
        if (o[a] >= o[b] || o[c] >= o[b]) odf_peak[b] = 0;
        if (o[b] >= o[c] || o[a] >= o[c]) odf_peak[c] = 0;
    }
    for (int i = 0; i < M; ++i) isort[i] = i;
    merge_sort_desc(isort, tmp, odf_peak, M);
    int nvalid = 0;
    for (int i = 0; i < M; ++i) nvalid += odf_peak[i] > 0;
    return nvalid;
}

static void qa_postpass(const float* odf, int64_t nvox, int M, float** qa) {
    /* odfmax = maximum(mean(odf.vol, dims=4)) over all voxels; qa ./= odfmax  (src/gqi.jl:164-168) */
    float odfmax = -INFINITY;
    for (int64_t v = 0; v < nvox; ++v) {
        float acc = 0.f;
        for (int i = 0; i < M; ++i) acc += odf[v + (int64_t)i * nvox];
        acc /= (float)M;
        if (acc > odfmax) odfmax = acc;
    }
    for (int k = 0; k < NPEAK; ++k) for (int64_t v = 0; v < nvox; ++v) qa[k][v] /= odfmax;
}

/* ---------------------------------------------------------------- GQI (src/gqi.jl:109-171) */
/* A [M][nvol] row-major; vertices [M][3] row-major (first half); outputs zero-filled by caller.
 * peak[k]: [nvox*3] frame-major; peak_idx (optional) int16 [nvox*3] frame-major. */
int oracle_gqi_rec(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, const float* A, int M,
                   const int32_t* faces, int nface, const float* vertices, float* odf, float** peak, float** qa,
                   int16_t* peak_idx, int nthreads) {
    const int64_t nvox = (int64_t)nx * ny * nz;
    nthreads = nthreads_eff(nthreads);
    if (peak_idx) for (int64_t i = 0; i < 3 * nvox; ++i) peak_idx[i] = -1;
#pragma omp parallel num_threads(nthreads)
    {
        float* s = (float*)malloc(sizeof(float) * nvol);
        float* o = (float*)malloc(sizeof(float) * M);
        float* op = (float*)malloc(sizeof(float) * M);
        int* isort = (int*)malloc(sizeof(int) * M);
        int* tmp = (int*)malloc(sizeof(int) * M);
#pragma omp for schedule(static)
        for (int iz = 0; iz < nz; ++iz)
            for (int iy = 0; iy < ny; ++iy)
                for (int ix = 0; ix < nx; ++ix) {
                    int64_t vox = ix + (int64_t)nx * (iy + (int64_t)ny * iz);
                    if (mask[vox] == 0) continue;
                    float mx = 0.f;
                    for (int j = 0; j < nvol; ++j) { float v = dwi[vox + (int64_t)j * nvox]; v = v < 0 ? 0 : v; s[j] = v; if (v > mx) mx = v; }
                    if (mx == 0.f) continue;
                    for (int i = 0; i < M; ++i) o[i] = dot_blas(A + (int64_t)i * nvol, s, nvol);
                    float mn = o[0];
                    for (int i = 0; i < M; ++i) { odf[vox + (int64_t)i * nvox] = o[i]; if (o[i] < mn) mn = o[i]; }
                    int nvalid = find_peaks(o, op, isort, tmp, M, faces, nface);
                    int n = nvalid < NPEAK ? nvalid : NPEAK;
                    for (int k = 0; k < n; ++k) {
                        int id = isort[k];
                        for (int c = 0; c < 3; ++c) peak[k][vox + c * nvox] = vertices[id * 3 + c];
                        qa[k][vox] = o[id] - mn;
                        if (peak_idx) peak_idx[vox + k * nvox] = (int16_t)id;
                    }
                }
        free(s); free(o); free(op); free(isort); free(tmp);
    }
    qa_postpass(odf, nvox, M, qa);
    return 0;
}

/* ---------------------------------------------------------------- DSI (src/dsi.jl:171-270), FFT form */
typedef struct { float re, im; } cpx;

static void fft1d(cpx* x, int n, int stride, const cpx* tw) {   /* in-place radix-2 DIT, forward */
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cpx t = x[i * stride]; x[i * stride] = x[j * stride]; x[j * stride] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                cpx w = tw[k * step];
                cpx u = x[(i + k) * stride], v = x[(i + k + len / 2) * stride];
                cpx t = {v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re};
                x[(i + k) * stride].re = u.re + t.re; x[(i + k) * stride].im = u.im + t.im;
                x[(i + k + len / 2) * stride].re = u.re - t.re; x[(i + k + len / 2) * stride].im = u.im - t.im;
            }
    }
}

/* iq_lin [nvol]: 0-based linear index (x fastest) of each volume in the nfft^3 grid; H [nfft^3];
 * coords [M][nrad][3]: 1-based continuous subscripts; qr2 [nrad]. */
int oracle_dsi_rec(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, int nfft,
                   const int32_t* iq_lin, const float* H, const float* coords, const float* qr2, int nrad, float dqr,
                   int M, const int32_t* faces, int nface, const float* vertices, float* pdf, float* odf,
                   float** peak, float** qa, int16_t* peak_idx, int nthreads) {
    const int64_t nvox = (int64_t)nx * ny * nz;
    const int n3 = nfft * nfft * nfft, sh = nfft / 2;
    nthreads = nthreads_eff(nthreads);
    if (peak_idx) for (int64_t i = 0; i < 3 * nvox; ++i) peak_idx[i] = -1;
    cpx* tw = (cpx*)malloc(sizeof(cpx) * nfft);
    for (int k = 0; k < nfft; ++k) { tw[k].re = (float)cos(-2 * M_PI * k / nfft); tw[k].im = (float)sin(-2 * M_PI * k / nfft); }
#pragma omp parallel num_threads(nthreads)
    {
        float* X = (float*)calloc(n3, sizeof(float));
        cpx* x = (cpx*)malloc(sizeof(cpx) * n3);
        float* p = (float*)malloc(sizeof(float) * n3);
        float* o = (float*)malloc(sizeof(float) * M);
        float* op = (float*)malloc(sizeof(float) * M);
        int* isort = (int*)malloc(sizeof(int) * M);
        int* tmp = (int*)malloc(sizeof(int) * M);
#pragma omp for schedule(static)
        for (int iz = 0; iz < nz; ++iz)
            for (int iy = 0; iy < ny; ++iy)
                for (int ix = 0; ix < nx; ++ix) {
                    int64_t vox = ix + (int64_t)nx * (iy + (int64_t)ny * iz);
                    if (mask[vox] == 0) continue;
                    for (int j = 0; j < nvol; ++j) X[iq_lin[j]] = dwi[vox + (int64_t)j * nvox];   /* last write wins */
                    float mx = X[0];
                    for (int i = 1; i < n3; ++i) if (X[i] > mx) mx = X[i];
                    if (mx == 0.f) continue;
                    /* clamp, window, fftshift into x */
                    for (int c = 0; c < nfft; ++c) for (int b = 0; b < nfft; ++b) for (int a = 0; a < nfft; ++a) {
                        int src = a + nfft * (b + nfft * c);
                        float v = X[src]; v = v > 0 ? v : 0; v *= H[src]; X[src] = v;
                        int dst = ((a + sh) % nfft) + nfft * (((b + sh) % nfft) + nfft * ((c + sh) % nfft));
                        x[dst].re = v; x[dst].im = 0.f;
                    }
                    for (int c = 0; c < nfft; ++c) for (int b = 0; b < nfft; ++b) fft1d(x + nfft * (b + nfft * c), nfft, 1, tw);
                    for (int c = 0; c < nfft; ++c) for (int a = 0; a < nfft; ++a) fft1d(x + a + nfft * nfft * c, nfft, nfft, tw);
                    for (int b = 0; b < nfft; ++b) for (int a = 0; a < nfft; ++a) fft1d(x + a + nfft * b, nfft, nfft * nfft, tw);
                    float sum = 0.f;
                    for (int c = 0; c < nfft; ++c) for (int b = 0; b < nfft; ++b) for (int a = 0; a < nfft; ++a) {
                        int src = ((a + sh) % nfft) + nfft * (((b + sh) % nfft) + nfft * ((c + sh) % nfft));
                        float v = x[src].re; p[a + nfft * (b + nfft * c)] = v; sum += v;
                    }
                    for (int i = 0; i < n3; ++i) p[i] /= sum;
                    for (int j = 0; j < nvol; ++j) pdf[vox + (int64_t)j * nvox] = p[iq_lin[j]];
                    for (int v = 0; v < M; ++v) {
                        float acc = 0.f;
                        for (int r = 0; r < nrad; ++r) {
                            const float* cc = coords + ((int64_t)v * nrad + r) * 3;
                            int i0 = (int)floorf(cc[0]), j0 = (int)floorf(cc[1]), k0 = (int)floorf(cc[2]);
                            float fx = cc[0] - i0, fy = cc[1] - j0, fz = cc[2] - k0;
                            i0 -= 1; j0 -= 1; k0 -= 1;
                            int i1 = i0 + 1 < nfft ? i0 + 1 : i0, j1 = j0 + 1 < nfft ? j0 + 1 : j0, k1 = k0 + 1 < nfft ? k0 + 1 : k0;
#define PV(a, b, c) p[(a) + nfft * ((b) + nfft * (c))]
                            float c00 = PV(i0, j0, k0) * (1.f - fx) + PV(i1, j0, k0) * fx;
                            float c10 = PV(i0, j1, k0) * (1.f - fx) + PV(i1, j1, k0) * fx;
                            float c01 = PV(i0, j0, k1) * (1.f - fx) + PV(i1, j0, k1) * fx;
                            float c11 = PV(i0, j1, k1) * (1.f - fx) + PV(i1, j1, k1) * fx;
#undef PV
                            float c0 = c00 * (1.f - fy) + c10 * fy, c1 = c01 * (1.f - fy) + c11 * fy;
                            acc += (c0 * (1.f - fz) + c1 * fz) * qr2[r];
                        }
                        o[v] = acc * dqr;
                    }
                    float mn = o[0];
                    for (int i = 0; i < M; ++i) { odf[vox + (int64_t)i * nvox] = o[i]; if (o[i] < mn) mn = o[i]; }
                    int nvalid = find_peaks(o, op, isort, tmp, M, faces, nface);
                    int n = nvalid < NPEAK ? nvalid : NPEAK;
                    for (int k = 0; k < n; ++k) {
                        int id = isort[k];
                        for (int c = 0; c < 3; ++c) peak[k][vox + c * nvox] = vertices[id * 3 + c];
                        qa[k][vox] = o[id] - mn;
                        if (peak_idx) peak_idx[vox + k * nvox] = (int16_t)id;
                    }
                }
        free(X); free(x); free(p); free(o); free(op); free(isort); free(tmp);
    }
    free(tw);
    qa_postpass(odf, nvox, M, qa);
    return 0;
}
