// Host-side construction of the per-protocol constants (the GPU analogue of the reference's
// DTIwork / ADCwork / GQIwork / DSIwork constructors).  Runs once per plan; not on the hot path.
#include <cmath>
#include <cstring>
#include <map>
#include <algorithm>
#include "common.cuh"

namespace fibers {

// ---- pseudo-inverse: one-sided (Hestenes) Jacobi SVD in float64 ----------------------------
// A is row-major [m][n], m >= n.  pA is row-major [n][m].  Singular values below
// rtol * sigma_max are dropped, as LinearAlgebra.pinv does (reference: src/dti.jl:72,143).
void pinv_rowmajor(const double* A, int m, int n, double rtol, double* pA) {
    std::vector<double> U(A, A + (size_t)m * n);        // columns get orthogonalised in place
    std::vector<double> V((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < m; ++i) {
                    double up = U[(size_t)i * n + p], uq = U[(size_t)i * n + q];
                    alpha += up * up; beta += uq * uq; gamma += up * uq;
                }
                if (gamma == 0.0) continue;
                double lim = std::sqrt(alpha * beta);
                if (std::fabs(gamma) <= 1e-17 * lim) continue;
                off = std::max(off, std::fabs(gamma) / (lim > 0 ? lim : 1.0));
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; ++i) {
                    double up = U[(size_t)i * n + p], uq = U[(size_t)i * n + q];
                    U[(size_t)i * n + p] = c * up - s * uq;
                    U[(size_t)i * n + q] = s * up + c * uq;
                }
                for (int i = 0; i < n; ++i) {
                    double vp = V[(size_t)i * n + p], vq = V[(size_t)i * n + q];
                    V[(size_t)i * n + p] = c * vp - s * vq;
                    V[(size_t)i * n + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    std::vector<double> sig2(n);
    double smax2 = 0;
    for (int k = 0; k < n; ++k) {
        double s2 = 0;
        for (int i = 0; i < m; ++i) s2 += U[(size_t)i * n + k] * U[(size_t)i * n + k];
        sig2[k] = s2; smax2 = std::max(smax2, s2);
    }
    std::fill(pA, pA + (size_t)n * m, 0.0);
    for (int k = 0; k < n; ++k) {
        if (!(sig2[k] > rtol * rtol * smax2) || sig2[k] == 0) continue;
        for (int r = 0; r < n; ++r) {
            double f = V[(size_t)r * n + k] / sig2[k];
            for (int i = 0; i < m; ++i) pA[(size_t)r * m + i] += f * U[(size_t)i * n + k];
        }
    }
}

static std::string design_common(int nvol, const float* bval, std::vector<uint8_t>& ib0) {
    if (nvol <= 0 || !bval) return "Missing b-value table from input DWI structure";
    float bmin = bval[0];
    for (int j = 1; j < nvol; ++j) bmin = std::min(bmin, bval[j]);
    ib0.resize(nvol);
    for (int j = 0; j < nvol; ++j) ib0[j] = (bval[j] == bmin);      // src/dti.jl:117
    return "";
}

static void pinv_f32(const std::vector<float>& A, int m, int n, std::vector<float>& pA) {
    std::vector<double> Ad(A.begin(), A.end()), pAd((size_t)n * m);
    // default rtol of pinv: eps(real(float(one(T)))) * min(size(A)...) with T = Float32
    double rtol = 1.1920929e-7 * std::min(m, n);
    pinv_rowmajor(Ad.data(), m, n, rtol, pAd.data());
    pA.resize((size_t)n * m);
    for (size_t i = 0; i < pA.size(); ++i) pA[i] = (float)pAd[i];
}

// DTIwork: src/dti.jl:110-155.  A [nvol][7] in the reference's fp32 arithmetic; pA = pinv(A).
std::string build_dti_design(int nvol, const float* bval, const float* bvec,
                             std::vector<float>& A, std::vector<float>& pA, std::vector<uint8_t>& ib0) {
    std::string e = design_common(nvol, bval, ib0);
    if (!e.empty()) return e;
    if (!bvec) return "Missing gradient table from input DWI structure";
    A.resize((size_t)nvol * 7);
    for (int j = 0; j < nvol; ++j) {
        volatile float gx = bvec[j], gy = bvec[nvol + j], gz = bvec[2 * nvol + j];   // column-major [nvol,3]
        float nb = -bval[j];
        float c[6];
        c[0] = gx * gx;                       // :133
        c[1] = (2.0f * gx) * gy;              // :134  (2*bvec[:,1]) .* bvec[:,2]
        c[2] = (2.0f * gx) * gz;
        c[3] = gy * gy;
        c[4] = (2.0f * gy) * gz;
        c[5] = gz * gz;
        for (int k = 0; k < 6; ++k) A[(size_t)j * 7 + k] = c[k] * nb;                 // :140
        A[(size_t)j * 7 + 6] = 1.0f;                                                   // :142
    }
    pinv_f32(A, nvol, 7, pA);
    return "";
}

// ADCwork: src/dti.jl:49-83.
std::string build_adc_design(int nvol, const float* bval, std::vector<float>& A, std::vector<float>& pA,
                             std::vector<uint8_t>& ib0) {
    std::string e = design_common(nvol, bval, ib0);
    if (!e.empty()) return e;
    A.resize((size_t)nvol * 2);
    for (int j = 0; j < nvol; ++j) { A[(size_t)j * 2] = -bval[j]; A[(size_t)j * 2 + 1] = 1.0f; }
    pinv_f32(A, nvol, 2, pA);
    return "";
}

// GQIwork system matrix: src/gqi.jl:66-69, fp32 constant chain.
std::string build_gqi_matrix(int nvol, const float* bval, const float* bvec, const float* vertices,
                             int nvert2, float sigma, std::vector<float>& A) {
    if (nvol <= 0 || !bval) return "Missing b-value table from input DWI structure";
    if (!bvec) return "Missing gradient table from input DWI structure";
    if (!vertices || nvert2 < 2 || (nvert2 & 1)) return "ODF vertex table must have an even, positive row count";
    const int M = nvert2 / 2;
    const float pif = (float)M_PI;
    volatile float sig = sigma / pif;                       // T(σ/π) with σ::Float32
    std::vector<float> bq((size_t)nvol * 3);
    for (int j = 0; j < nvol; ++j) {
        volatile float t = bval[j] * 0.01506f;              // bval * T(0.01506)
        volatile float s = std::sqrt((float)t);
        volatile float sc = s * sig;
        for (int c = 0; c < 3; ++c) bq[(size_t)j * 3 + c] = bvec[(size_t)c * nvol + j] * sc;
    }
    A.resize((size_t)M * nvol);
    for (int i = 0; i < M; ++i) {
        // second (antipodal) half of the vertex table, as the reference uses (vertices[nvert+1:end,:])
        float vx = vertices[M + i], vy = vertices[(size_t)nvert2 + M + i], vz = vertices[(size_t)2 * nvert2 + M + i];
        for (int j = 0; j < nvol; ++j) {
            volatile float p0 = vx * bq[(size_t)j * 3], p1 = vy * bq[(size_t)j * 3 + 1], p2 = vz * bq[(size_t)j * 3 + 2];
            volatile float x = (p0 + p1);
            x = x + p2;
            float xv = x;
            double px = M_PI * (double)xv;
            A[(size_t)i * nvol + j] = (xv == 0.0f) ? 1.0f : (float)(std::sin(px) / px);   // Base.sinc
        }
    }
    return "";
}

// DSIwork + the per-voxel pipeline of dsi_rec folded into one linear map (src/dsi.jl:59-143,
// :205-242): odf = (Mo s+)/den, pdf = (Mp s+)/den, den = nfft^3 H_c s+_c.
std::string build_dsi_matrix(int nvol, const float* bval, const float* bvec, const float* vertices,
                             int nvert2, int hann_width, std::vector<float>& MoMp, int& cvol, float& dscale) {
    if (nvol <= 0 || !bval) return "Missing b-value table from input DWI structure";
    if (!bvec) return "Missing gradient table from input DWI structure";
    if (!vertices || nvert2 < 2 || (nvert2 & 1)) return "ODF vertex table must have an even, positive row count";
    if (hann_width < 0) return "hann_width must be >= 0";
    const int M = nvert2 / 2, N = nvol;
    float bmin = bval[0];
    for (int j = 1; j < N; ++j) bmin = std::min(bmin, bval[j]);
    float b1 = INFINITY;
    for (int j = 0; j < N; ++j) if (bval[j] > bmin) b1 = std::min(b1, bval[j]);
    if (!std::isfinite(b1)) return "DSI needs at least two distinct b-values";
    const float dq = std::sqrt(b1);                                           // :66
    std::vector<int> iq((size_t)N * 3);
    int lo = INT32_MAX, hi = INT32_MIN;
    for (int j = 0; j < N; ++j) {
        float sb = std::sqrt(bval[j]);
        for (int c = 0; c < 3; ++c) {
            volatile float q = bvec[(size_t)c * N + j] * sb;                  // :62
            volatile float qs = q / dq;
            int v = (int)std::nearbyint((float)qs);                          // round half to even (:67)
            iq[(size_t)j * 3 + c] = v; lo = std::min(lo, v); hi = std::max(hi, v);
        }
    }
    int nfft = hi - lo + 1;                                                   // :70
    int p2 = 1; while (p2 < nfft) p2 <<= 1; nfft = p2;                        // :71
    if (nfft > 64) return "DSI q-space grid too large (nfft > 64)";
    const int shift = nfft / 2 + 1;                                           // :73 (1-based)
    for (int j = 0; j < N; ++j)
        for (int c = 0; c < 3; ++c) {
            int s = iq[(size_t)j * 3 + c] + shift;
            if (s < 1 || s > nfft) return "DSI q-space point falls outside the FFT grid";
        }
    // last write wins for duplicate grid cells (:205)
    std::map<int, int> last;
    std::vector<int> cell(N);
    std::vector<uint8_t> live(N, 1);
    for (int j = 0; j < N; ++j) {
        int c = ((iq[(size_t)j * 3] + shift - 1) * nfft + (iq[(size_t)j * 3 + 1] + shift - 1)) * nfft +
                (iq[(size_t)j * 3 + 2] + shift - 1);
        cell[j] = c;
        auto it = last.find(c);
        if (it != last.end()) live[it->second] = 0;
        last[c] = j;
    }
    // Hann window on sampled points, Float64 then rounded to Float32 (:80-85)
    std::vector<double> H(N);
    for (int j = 0; j < N; ++j) {
        if (hann_width == 0) H[j] = 1.0;
        else {
            double r2 = 0;
            for (int c = 0; c < 3; ++c) r2 += (double)iq[(size_t)j * 3 + c] * iq[(size_t)j * 3 + c];
            H[j] = (double)(float)((1.0 + std::cos(std::sqrt(r2) * (2.0 * M_PI / hann_width))) * 0.5);
        }
        if (!live[j]) H[j] = 0.0;
    }
    std::vector<double> ctab(nfft);
    for (int k = 0; k < nfft; ++k) ctab[k] = std::cos(2.0 * M_PI * k / nfft);
    ctab[0] = 1.0;
    if (nfft % 4 == 0) { ctab[nfft / 4] = 0.0; ctab[3 * nfft / 4] = 0.0; ctab[nfft / 2] = -1.0; }
    auto cosdot = [&](int gx, int gy, int gz, int j) {   // g relative to the grid centre
        long d = (long)gx * iq[(size_t)j * 3] + (long)gy * iq[(size_t)j * 3 + 1] + (long)gz * iq[(size_t)j * 3 + 2];
        int k = (int)(((d % nfft) + nfft) % nfft);
        return ctab[k];
    };
    // radial sampling (:104-109), fp32 chain
    const int nrad = 21;
    float qr[nrad], qr2[nrad];
    const float rscale = (float)(nfft / 2 - 1);
    for (int r = 0; r < nrad; ++r) {
        volatile float t = rscale * (float)(0.3 + 0.03 * r);
        qr[r] = t;
        volatile float t2 = qr[r] * qr[r];
        qr2[r] = t2;
    }
    volatile float dqrv = qr[1] - qr[0];
    const double dqr = (double)(float)dqrv;
    MoMp.assign((size_t)(M + N) * N, 0.f);
    std::vector<double> row(N);
    for (int v = 0; v < M; ++v) {
        std::fill(row.begin(), row.end(), 0.0);
        float vv[3] = {vertices[M + v], vertices[(size_t)nvert2 + M + v], vertices[(size_t)2 * nvert2 + M + v]};
        for (int r = 0; r < nrad; ++r) {
            int c0[3]; double fr[3];
            for (int c = 0; c < 3; ++c) {
                volatile float prod = vv[c] * qr[r];
                volatile float co = prod + (float)shift;                      // 1-based continuous subscript
                float cf = co;
                float fl = std::floor(cf);
                c0[c] = (int)fl; fr[c] = (double)(cf - fl);
                if (c0[c] < 1 || c0[c] + 1 > nfft) {
                    if (!(c0[c] == nfft && fr[c] == 0.0)) return "DSI interpolation point outside the grid";
                }
            }
            for (int dx = 0; dx < 2; ++dx) for (int dy = 0; dy < 2; ++dy) for (int dz = 0; dz < 2; ++dz) {
                double w = (dx ? fr[0] : 1 - fr[0]) * (dy ? fr[1] : 1 - fr[1]) * (dz ? fr[2] : 1 - fr[2]);
                if (w == 0.0) continue;
                w *= (double)qr2[r];
                int gx = c0[0] + dx - shift, gy = c0[1] + dy - shift, gz = c0[2] + dz - shift;
                for (int j = 0; j < N; ++j)
                    if (H[j] != 0.0) row[j] += w * H[j] * cosdot(gx, gy, gz, j);
            }
        }
        for (int j = 0; j < N; ++j) MoMp[(size_t)v * N + j] = (float)(row[j] * dqr);
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            MoMp[(size_t)(M + i) * N + j] =
                (float)(H[j] * cosdot(iq[(size_t)i * 3], iq[(size_t)i * 3 + 1], iq[(size_t)i * 3 + 2], j));
    int centre = ((shift - 1) * nfft + (shift - 1)) * nfft + (shift - 1);
    auto it = last.find(centre);
    if (it == last.end()) { cvol = -1; dscale = 0.f; }
    else { cvol = it->second; dscale = (float)((double)nfft * nfft * nfft * H[cvol]); }
    return "";
}

// Folded-mesh neighbour table (src/gqi.jl:63-64 folding; :185-196 suppression rule).
std::string build_neighbours(const int32_t* faces, int nface, int nvert, std::vector<uint16_t>& nbr) {
    if (!faces || nface <= 0) return "Missing ODF face table";
    if (nvert >= 0xFFFF) return "too many ODF vertices";
    nbr.assign((size_t)nvert * NBR_W, NBR_NONE);
    std::vector<int> deg(nvert, 0);
    auto add = [&](int a, int b) -> bool {
        for (int k = 0; k < deg[a]; ++k) if (nbr[(size_t)a * NBR_W + k] == b) return true;
        if (deg[a] >= NBR_W) return false;
        nbr[(size_t)a * NBR_W + deg[a]++] = (uint16_t)b;
        return true;
    };
    for (int f = 0; f < nface; ++f) {
        int v[3];
        for (int c = 0; c < 3; ++c) {
            int x = faces[(size_t)c * nface + f];                 // column-major [nface,3], 1-based
            if (x < 1 || x > 2 * nvert) return "ODF face index out of range";
            if (x > nvert) x -= nvert;
            v[c] = x - 1;
        }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                if (a != b && !add(v[a], v[b])) return "ODF mesh vertex degree exceeds the supported maximum (8)";
    }
    for (int a = 0; a < nvert; ++a) std::sort(&nbr[(size_t)a * NBR_W], &nbr[(size_t)a * NBR_W] + deg[a]);
    return "";
}

}  // namespace fibers
