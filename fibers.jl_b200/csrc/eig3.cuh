// Closed-form eigen-decomposition of a real symmetric 3x3 matrix (StaticArrays.jl `eigen(Symmetric(::SMatrix{3,3}))`), shared by
// the DTI fit (src/dti.jl:311) and the structure-tensor eigen-decomposition (src/structens.jl:26).  Include only from translation
// units compiled with -fmad=false: the solver follows the reference's unfused fp32 operation order.
#pragma once
#include <math_constants.h>

namespace fibers {
namespace {

struct Cross { float x, y, z; };
__device__ __forceinline__ Cross cross3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return {a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0};
}

// Closed-form eigen-decomposition of a real symmetric 3x3 matrix; values ascending in w[],
// k-th eigenvector in (v[0][k], v[1][k], v[2][k]).  Restates StaticArrays.jl
// `_eig(::Size{(3,3)}, ::RealHermSymComplexHerm)` (mirrored by oracle/fibers_oracle.py:eig3_sym).
__device__ void eig3_sym(float a11, float a12, float a13, float a22, float a23, float a33,
                         float w[3], float v[3][3]) {
    float p1 = a12 * a12 + a13 * a13 + a23 * a23;
    if (p1 == 0.f) {   // diagonal matrix: sorted diagonal, unit vectors
        int o0, o1, o2;
        if (a11 < a22) {
            if (a22 < a33) { o0 = 0; o1 = 1; o2 = 2; }
            else if (a33 < a11) { o0 = 2; o1 = 0; o2 = 1; }
            else { o0 = 0; o1 = 2; o2 = 1; }
        } else {
            if (a11 < a33) { o0 = 1; o1 = 0; o2 = 2; }
            else if (a33 < a22) { o0 = 2; o1 = 1; o2 = 0; }
            else { o0 = 1; o1 = 2; o2 = 0; }
        }
        float d[3] = {a11, a22, a33};
        int o[3] = {o0, o1, o2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            w[k] = d[o[k]];
#pragma unroll
            for (int r = 0; r < 3; ++r) v[r][k] = (r == o[k]) ? 1.f : 0.f;
        }
        return;
    }
    float q = (a11 + a22 + a33) / 3.f;
    float d11 = a11 - q, d22 = a22 - q, d33 = a33 - q;
    float p2 = d11 * d11 + d22 * d22 + d33 * d33 + 2.f * p1;
    float p = sqrtf(p2 / 6.f);
    float invp = 1.f / p;
    float b11 = d11 * invp, b22 = d22 * invp, b33 = d33 * invp;
    float b12 = a12 * invp, b13 = a13 * invp, b23 = a23 * invp;
    Cross c = cross3(b12, b22, b23, b13, b23, b33);
    float r = (b11 * c.x + b12 * c.y + b13 * c.z) / 2.f;
    const float pif = 3.14159274f;
    float phi;
    if (r <= -1.f) phi = pif / 3.f;
    else if (r >= 1.f) phi = 0.f;
    else phi = acosf(r) / 3.f;
    float eig3 = q + 2.f * p * cosf(phi);
    float eig1 = q + 2.f * p * cosf(phi + (2.f * pif / 3.f));
    float eig2 = 3.f * q - eig1 - eig3;
    const bool swap = r > 0.f;
    float e1 = swap ? eig3 : eig1;
    float e3 = swap ? eig1 : eig3;
    // first eigenvector: best-conditioned cross product of two rows of A - e1 I
    float r1x = a11 - e1, r1y = a12, r1z = a13;
    float r2x = a12, r2y = a22 - e1, r2z = a23;
    float r3x = a13, r3y = a23, r3z = a33 - e1;
    float n1 = r1x * r1x + r1y * r1y + r1z * r1z;
    float n2 = r2x * r2x + r2y * r2y + r2z * r2z;
    float n3 = r3x * r3x + r3y * r3y + r3z * r3z;
    Cross r12 = cross3(r1x, r1y, r1z, r2x, r2y, r2z);
    Cross r23 = cross3(r2x, r2y, r2z, r3x, r3y, r3z);
    Cross r31 = cross3(r3x, r3y, r3z, r1x, r1y, r1z);
    float n12 = r12.x * r12.x + r12.y * r12.y + r12.z * r12.z;
    float n23 = r23.x * r23.x + r23.y * r23.y + r23.z * r23.z;
    float n31 = r31.x * r31.x + r31.y * r31.y + r31.z * r31.z;
    Cross best; float nb;
    if (n12 * n3 > n23 * n1) {
        if (n12 * n3 > n31 * n2) { best = r12; nb = n12; } else { best = r31; nb = n31; }
    } else {
        if (n23 * n1 > n31 * n2) { best = r23; nb = n23; } else { best = r31; nb = n31; }
    }
    float sn = sqrtf(nb);
    float v1x = best.x / sn, v1y = best.y / sn, v1z = best.z / sn;
    // orthonormal complement of v1
    float o1x, o1y, o1z;
    if (fabsf(v1x) < fabsf(v1y)) {
        float dn = sqrtf(v1x * v1x + v1z * v1z);
        o1x = -v1z / dn; o1y = 0.f; o1z = v1x / dn;
    } else {
        float dn = sqrtf(v1y * v1y + v1z * v1z);
        o1x = 0.f; o1y = v1z / dn; o1z = -v1y / dn;
    }
    Cross o2 = cross3(v1x, v1y, v1z, o1x, o1y, o1z);
    // projected 2x2 problem of A - eig2 I on {o1, o2}
    float ao1x = a11 * o1x + a12 * o1y + a13 * o1z;
    float ao1y = a12 * o1x + a22 * o1y + a23 * o1z;
    float ao1z = a13 * o1x + a23 * o1y + a33 * o1z;
    float ao2x = a11 * o2.x + a12 * o2.y + a13 * o2.z;
    float ao2y = a12 * o2.x + a22 * o2.y + a23 * o2.z;
    float ao2z = a13 * o2.x + a23 * o2.y + a33 * o2.z;
    float c11 = o1x * ao1x + o1y * ao1y + o1z * ao1z - eig2;
    float c12 = o1x * ao2x + o1y * ao2y + o1z * ao2z;
    float c22 = o2.x * ao2x + o2.y * ao2y + o2.z * ao2z - eig2;
    float s11 = c11 * c11, s12 = c12 * c12, s22 = c22 * c22;
    float v2x, v2y, v2z;
    float pp1, pp2;
    bool degen = false;
    if (s11 >= s22) {
        if (s11 > 0.f || s12 > 0.f) {
            if (s11 >= s12) { float t = c12 / c11; pp2 = 1.f / sqrtf(1.f + t * t); pp1 = t * pp2; }
            else            { float t = c11 / c12; pp1 = 1.f / sqrtf(1.f + t * t); pp2 = t * pp1; }
        } else { degen = true; pp1 = 1.f; pp2 = 0.f; }
    } else {
        if (s22 >= s12) { float t = c12 / c22; pp1 = 1.f / sqrtf(1.f + t * t); pp2 = t * pp1; }
        else            { float t = c22 / c12; pp2 = 1.f / sqrtf(1.f + t * t); pp1 = t * pp2; }
    }
    if (degen) { v2x = o1x; v2y = o1y; v2z = o1z; }
    else {
        v2x = pp1 * o1x - pp2 * o2.x; v2y = pp1 * o1y - pp2 * o2.y; v2z = pp1 * o1z - pp2 * o2.z;
    }
    Cross v3 = cross3(v1x, v1y, v1z, v2x, v2y, v2z);
    if (swap) {
        w[0] = e3; w[1] = eig2; w[2] = e1;
        v[0][0] = v3.x; v[1][0] = v3.y; v[2][0] = v3.z;
        v[0][2] = v1x;  v[1][2] = v1y;  v[2][2] = v1z;
    } else {
        w[0] = e1; w[1] = eig2; w[2] = e3;
        v[0][0] = v1x;  v[1][0] = v1y;  v[2][0] = v1z;
        v[0][2] = v3.x; v[1][2] = v3.y; v[2][2] = v3.z;
    }
    v[0][1] = v2x; v[1][1] = v2y; v[2][1] = v2z;
}

}  // namespace
}  // namespace fibers
