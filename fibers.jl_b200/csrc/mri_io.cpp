// Volume I/O either side of the reconstruction path: NIfTI-1 (.nii / .nii.gz) and MGH
// (.mgh / .mgz) readers and writers with the reference's header semantics, behind the C ABI.  Host code only.
//
//   fibers_mri_read_info   load_nifti_hdr (src/mri.jl:1394-1551) / the header part of load_mgh (:1217-1283), plus the
//                          geometry mri_read derives (:611-700): dims beyond the 4th folded into frames, vox2ras0 = sform if
//                          sform_code != 0, else qform if qform_code != 0, else diag(pixdim); units converted to mm / ms
//   fibers_mri_read_data   load_nifti (:1576-1672) / load_mgh (:1284-1372): byte order, scl_slope / scl_inter rule (:1665-1669),
//                          big-endian MGH payload, trailing mr_parms
//   fibers_mri_write       mri_write + save_nifti / save_mgh (:1695-1937, :2059-2176, :1939-2036): same header fields
//                          ("FreeSurfer julia", intent_name "huh?", qform from vox2ras_to_qform :391-463, sform = vox2ras0,
//                          vox_offset 352, cal_min / cal_max), same byte layout
// Differences by design: compressed files go through zlib in-process, streaming straight into / out of the caller's buffer
// (the reference shells out to zcat / gzip through a temporary file: two more passes over the disk, src/mri.jl:1581-1592,
// :2160-2163); the destination may be pinned memory obtained from fibers_cuda_host_register, so that the reconstruction
// entry points DMA from it directly.  Bruker directories (load_bruker) are not read.
#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>
#include "common.cuh"

namespace fibers {
namespace {

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i) if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}
// mri_filename (src/mri.jl:520-560): format from the extension (no search for a stem on disk here: the wrapper does that)
int format_of(const std::string& path, int& gz) {
    gz = 0;
    if (ends_with(path, ".nii")) return FIBERS_FMT_NIFTI;
    if (ends_with(path, ".nii.gz")) { gz = 1; return FIBERS_FMT_NIFTI; }
    if (ends_with(path, ".mgh")) return FIBERS_FMT_MGH;
    if (ends_with(path, ".mgz")) { gz = 1; return FIBERS_FMT_MGH; }
    return 0;
}
size_t dtype_size(int dt) {
    switch (dt) { case FIBERS_F32: case FIBERS_I32: case FIBERS_U32: return 4; case FIBERS_F64: case FIBERS_I64: return 8;
                  case FIBERS_I16: case FIBERS_U16: return 2; case FIBERS_U8: case FIBERS_I8: return 1; default: return 0; }
}
void swap_bytes(void* p, size_t elem, size_t n) {
    uint8_t* b = (uint8_t*)p;
    if (elem == 2) for (size_t i = 0; i < n; ++i) std::swap(b[2 * i], b[2 * i + 1]);
    else if (elem == 4) for (size_t i = 0; i < n; ++i) { std::swap(b[4 * i], b[4 * i + 3]); std::swap(b[4 * i + 1], b[4 * i + 2]); }
    else if (elem == 8) for (size_t i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) std::swap(b[8 * i + k], b[8 * i + 7 - k]);
}
bool host_is_little() { const uint16_t x = 1; return *(const uint8_t*)&x == 1; }

// gzopen reads plain files transparently, so one code path serves .nii / .nii.gz / .mgh / .mgz
struct GzIn {
    gzFile f = nullptr;
    ~GzIn() { if (f) gzclose(f); }
    bool open(const char* path) { f = gzopen(path, "rb"); if (f) gzbuffer(f, 1 << 20); return f != nullptr; }
    bool read(void* dst, size_t n) {
        uint8_t* d = (uint8_t*)dst;
        while (n) { const unsigned c = (unsigned)std::min<size_t>(n, 1u << 30); const int r = gzread(f, d, c); if (r <= 0) return false; d += r; n -= (size_t)r; }
        return true;
    }
    bool skip(size_t n) { return gzseek(f, (z_off_t)n, SEEK_CUR) >= 0; }
    bool at_eof() { uint8_t b; return gzread(f, &b, 1) <= 0; }
};
struct Out {
    gzFile g = nullptr; FILE* p = nullptr; size_t nb = 0;
    ~Out() { close(); }
    bool open(const char* path, bool gz) { if (gz) { g = gzopen(path, "wb1"); if (g) gzbuffer(g, 1 << 20); return g != nullptr; } p = fopen(path, "wb"); return p != nullptr; }
    bool write(const void* src, size_t n) {
        nb += n;
        if (p) return fwrite(src, 1, n, p) == n;
        const uint8_t* s = (const uint8_t*)src;
        while (n) { const unsigned c = (unsigned)std::min<size_t>(n, 1u << 30); if (gzwrite(g, s, c) != (int)c) return false; s += c; n -= c; }
        return true;
    }
    bool close() { bool ok = true; if (g) { ok = gzclose(g) == Z_OK; g = nullptr; } if (p) { ok = fclose(p) == 0; p = nullptr; } return ok; }
};

template <class T> T rd(const uint8_t* b, size_t off, bool sw) { T v; memcpy(&v, b + off, sizeof(T)); if (sw) swap_bytes(&v, sizeof(T), 1); return v; }

int nifti_dtype(int code) {                            // (src/mri.jl:1604-1629)
    switch (code) { case 2: return FIBERS_U8; case 4: return FIBERS_I16; case 8: return FIBERS_I32; case 16: return FIBERS_F32; case 64: return FIBERS_F64;
                    case 256: return FIBERS_I8; case 512: return FIBERS_U16; case 768: return FIBERS_U32; default: return -1; }
}
int nifti_code(int dt, int& bitpix) {                  // (src/mri.jl:1767-1793)
    switch (dt) { case FIBERS_U8: bitpix = 8; return 2; case FIBERS_I16: bitpix = 16; return 4; case FIBERS_I32: bitpix = 32; return 8;
                  case FIBERS_F32: bitpix = 32; return 16; case FIBERS_F64: bitpix = 64; return 64; case FIBERS_I8: bitpix = 8; return 256;
                  case FIBERS_U16: bitpix = 16; return 512; case FIBERS_U32: bitpix = 32; return 768; default: return -1; }
}

int read_nifti_info(GzIn& in, fibers_mri_info* h) {
    uint8_t b[348];
    if (!in.read(b, 348)) return fail(FIBERS_ERR_ARG, "NIfTI header is shorter than 348 bytes");
    const int32_t hs = rd<int32_t>(b, 0, false);
    bool sw;
    if (hs == 348) sw = false;
    else if (rd<int32_t>(b, 0, true) == 348) sw = true;
    else return fail(FIBERS_ERR_ARG, "Invalid header size " + std::to_string(hs) + " found in NIfTI header");     // (:1416)
    h->bswap = sw;
    int32_t dim[8];
    for (int i = 0; i < 8; ++i) dim[i] = rd<int16_t>(b, 40 + 2 * i, sw);
    const int32_t glmin = rd<int32_t>(b, 144, sw);
    if (dim[1] < 0) dim[1] = glmin;                                                  // > 32k columns (FreeSurfer; :1429-1433)
    if ((int64_t)dim[1] * dim[2] * dim[3] == 163842) { dim[1] = 163842; dim[2] = 1; dim[3] = 1; }      // ico7 (:1435-1438)
    const int code = rd<int16_t>(b, 70, sw);
    h->dtype = nifti_dtype(code);
    if (h->dtype < 0) return fail(FIBERS_ERR_ARG, "Data type " + std::to_string(code) + " not supported");          // (:1636)
    float pixdim[8];
    for (int i = 0; i < 8; ++i) pixdim[i] = rd<float>(b, 76 + 4 * i, sw);
    h->data_offset = (int64_t)std::llround((double)rd<float>(b, 108, sw));          // round(vox_offset) (:1653)
    h->scl_slope = rd<float>(b, 112, sw); h->scl_inter = rd<float>(b, 116, sw);
    const int8_t units = rd<int8_t>(b, 123, false);
    const int xyzu = units & 7, tu = units & 56;
    const float xyzscale = xyzu == 1 ? 1000.f : xyzu == 3 ? 0.001f : 1.f;           // m, (mm), um; unknown -> mm (:1444-1455)
    const float tscale = tu == 8 ? 1000.f : tu == 16 ? 1.f : tu == 32 ? 0.001f : 0.f;   // (:1457-1466)
    for (int i = 1; i <= 3; ++i) pixdim[i] *= xyzscale;
    pixdim[4] *= tscale;
    memcpy(h->pixdim, pixdim, sizeof(pixdim));
    h->qform_code = rd<int16_t>(b, 252, sw); h->sform_code = rd<int16_t>(b, 254, sw);
    float S[16] = {0}, Q[16] = {0};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) S[4 * r + c] = rd<float>(b, 280 + 16 * r + 4 * c, sw) * xyzscale;
    S[15] = 1.f;
    {                                                                               // qform (:1493-1530), fp32 throughout
        float qb = rd<float>(b, 256, sw), qc = rd<float>(b, 260, sw), qd = rd<float>(b, 264, sw);
        const float qx = rd<float>(b, 268, sw), qy = rd<float>(b, 272, sw), qz = rd<float>(b, 276, sw);
        float a = 1.f - (qb * qb + qc * qc + qd * qd);
        if (std::fabs(a) < 1.0e-7f) { a = 1.f / std::sqrt(qb * qb + qc * qc + qd * qd); qb *= a; qc *= a; qd *= a; a = 0.f; }
        else a = std::sqrt(a);
        float r[9] = {a * a + qb * qb - qc * qc - qd * qd, 2 * qb * qc - 2 * a * qd, 2 * qb * qd + 2 * a * qc,
                      2 * qb * qc + 2 * a * qd, a * a + qc * qc - qb * qb - qd * qd, 2 * qc * qd - 2 * a * qb,
                      2 * qb * qd - 2 * a * qc, 2 * qc * qd + 2 * a * qb, a * a + qd * qd - qc * qc - qb * qb};
        if (pixdim[0] < 0.f) { r[2] = -r[2]; r[5] = -r[5]; r[8] = -r[8]; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Q[4 * i + j] = r[3 * i + j] * pixdim[1 + j];
        Q[3] = qx; Q[7] = qy; Q[11] = qz; Q[15] = 1.f;
    }
    memcpy(h->sform, S, sizeof(S)); memcpy(h->qform, Q, sizeof(Q));
    if (h->sform_code != 0) memcpy(h->vox2ras0, S, sizeof(S));                      // (:1537-1550)
    else if (h->qform_code != 0) memcpy(h->vox2ras0, Q, sizeof(Q));
    else { float D[16] = {0}; D[0] = pixdim[1]; D[5] = pixdim[2]; D[10] = pixdim[3]; D[15] = 1.f; memcpy(h->vox2ras0, D, sizeof(D)); }
    // mri_read (:646-673): dims are hdr.dim[2:end] that are > 0; more than 4 -> everything beyond x, y, z goes into dim 4
    std::vector<int64_t> vs;
    for (int i = 1; i < 8; ++i) if (dim[i] > 0) vs.push_back(dim[i]);
    while (vs.size() < 3) vs.push_back(1);
    int64_t nfr = 1;
    for (size_t i = 3; i < vs.size(); ++i) nfr *= vs[i];
    h->ndim = vs.size() > 3 ? 4 : 3;
    h->dim[0] = (int32_t)vs[0]; h->dim[1] = (int32_t)vs[1]; h->dim[2] = (int32_t)vs[2]; h->dim[3] = (int32_t)nfr;
    h->tr = pixdim[4]; h->flip_angle = h->te = h->ti = 0.f;                         // (:664-666)
    return 0;
}

int read_mgh_info(GzIn& in, fibers_mri_info* h) {
    uint8_t b[30];
    if (!in.read(b, 30)) return fail(FIBERS_ERR_ARG, "MGH header is shorter than 30 bytes");
    const bool sw = host_is_little();                                               // everything is big-endian (ntoh)
    const int32_t nd1 = rd<int32_t>(b, 4, sw), nd2 = rd<int32_t>(b, 8, sw), nd3 = rd<int32_t>(b, 12, sw), nfr = rd<int32_t>(b, 16, sw);
    const int32_t type = rd<int32_t>(b, 20, sw);
    const int16_t ras_good = rd<int16_t>(b, 28, sw);
    int unused = 256 - 2;
    float M[16] = {0};
    if (ras_good > 0) {                                                             // (:1263-1279)
        uint8_t r[60];
        if (!in.read(r, 60)) return fail(FIBERS_ERR_ARG, "MGH header is truncated");
        float delta[3], Mdc[9], Pc[3];
        for (int i = 0; i < 3; ++i) delta[i] = rd<float>(r, 4 * i, sw);
        for (int i = 0; i < 9; ++i) Mdc[i] = rd<float>(r, 12 + 4 * i, sw);           // column-major 3 x 3 (reshape(Mdc, (3,3)))
        for (int i = 0; i < 3; ++i) Pc[i] = rd<float>(r, 48 + 4 * i, sw);
        const float crs[3] = {(float)nd1 / 2.f, (float)nd2 / 2.f, (float)nd3 / 2.f};
        for (int i = 0; i < 3; ++i) {
            float acc = 0.f;
            for (int j = 0; j < 3; ++j) { const float md = Mdc[3 * j + i] * delta[j]; M[4 * i + j] = md; acc += md * crs[j]; }
            M[4 * i + 3] = Pc[i] - acc;                                             // Pxyz_0 = Pxyz_c - Mdc*D*Pcrs_c
        }
        M[15] = 1.f;
        unused -= 60;
    } else return fail(FIBERS_ERR_ARG, "MGH file has no RAS transform (ras_good_flag = 0)");       // mri_read errors on an empty M (:629-631)
    if (!in.skip((size_t)unused)) return fail(FIBERS_ERR_ARG, "MGH header is truncated");
    switch (type) { case 3: h->dtype = FIBERS_F32; break; case 0: h->dtype = FIBERS_U8; break; case 4: h->dtype = FIBERS_I16; break;
                    case 10: h->dtype = FIBERS_U16; break; case 1: h->dtype = FIBERS_I32; break;
                    default: return fail(FIBERS_ERR_ARG, "MGH data type " + std::to_string(type) + " not supported"); }
    h->bswap = sw;
    h->dim[0] = nd1; h->dim[1] = nd2; h->dim[2] = nd3; h->dim[3] = nfr; h->ndim = nfr > 1 ? 4 : 3;
    h->data_offset = 284;
    memcpy(h->vox2ras0, M, sizeof(M));
    h->scl_slope = 0.f; h->scl_inter = 0.f;
    return 0;
}

void volres_of(const float* M, float* res) {           // sqrt of the column sums of squares of vox2ras0[1:3,1:3] (mri_set_geometry!)
    for (int j = 0; j < 3; ++j) res[j] = std::sqrt(M[j] * M[j] + M[4 + j] * M[4 + j] + M[8 + j] * M[8 + j]);
}

template <class T> bool rescale(T* v, size_t n, float slope, float inter) {        // vol .= dtype.(vol .* slope .+ inter) (:1668)
    for (size_t i = 0; i < n; ++i) {
        const float y = (float)v[i] * slope + inter;
        const T t = (T)y;
        if ((float)t != y) return false;                // Julia: InexactError for an integer element type
        v[i] = t;
    }
    return true;
}

}  // namespace
}  // namespace fibers

using namespace fibers;

extern "C" int fibers_mri_read_info(const char* path, fibers_mri_info* info) {
    if (!path || !info) return fail(FIBERS_ERR_ARG, "NULL pointer");
    memset(info, 0, sizeof(*info));
    int gz = 0;
    const int fmt = format_of(path, gz);
    if (!fmt) return fail(FIBERS_ERR_ARG, std::string("Cannot determine format of ") + path);          // (:618)
    GzIn in;
    if (!in.open(path)) return fail(FIBERS_ERR_ARG, std::string("Could not open ") + path);
    info->format = fmt; info->gz = gz;
    if (int rc = fmt == FIBERS_FMT_NIFTI ? read_nifti_info(in, info) : read_mgh_info(in, info)) return rc;
    if (fmt == FIBERS_FMT_MGH) {                       // trailing mr_parms, if present (:1296-1299, :1361-1363)
        const size_t nv = (size_t)info->dim[0] * info->dim[1] * info->dim[2] * info->dim[3] * dtype_size(info->dtype);
        uint8_t t[16];
        if (in.skip(nv) && in.read(t, 16)) {
            const bool sw = host_is_little();
            info->tr = rd<float>(t, 0, sw); info->flip_angle = rd<float>(t, 4, sw); info->te = rd<float>(t, 8, sw); info->ti = rd<float>(t, 12, sw);
        }
    }
    volres_of(info->vox2ras0, info->volres);
    return 0;
}

extern "C" int fibers_mri_read_data(const char* path, const fibers_mri_info* info, void* dst, int64_t dst_bytes) {
    if (!path || !info || !dst) return fail(FIBERS_ERR_ARG, "NULL pointer");
    const size_t es = dtype_size(info->dtype);
    const size_t n = (size_t)info->dim[0] * info->dim[1] * info->dim[2] * info->dim[3];
    if (es == 0 || (int64_t)(n * es) > dst_bytes) return fail(FIBERS_ERR_ARG, "destination buffer too small");
    GzIn in;
    if (!in.open(path)) return fail(FIBERS_ERR_ARG, std::string("Could not open ") + path);
    if (!in.skip((size_t)info->data_offset)) return fail(FIBERS_ERR_ARG, std::string("Could not seek in ") + path);
    if (!in.read(dst, n * es)) return fail(FIBERS_ERR_ARG, std::string(path) + ": fewer bytes than the header announces");
    if (info->format == FIBERS_FMT_NIFTI && !in.at_eof())                           // (:1657-1660)
        return fail(FIBERS_ERR_ARG, std::string(path) + ", read a volume but did not reach end of file");
    if (info->bswap) swap_bytes(dst, es, n);
    const float sl = info->scl_slope, ic = info->scl_inter;
    if (info->format == FIBERS_FMT_NIFTI && sl != 0.f && !(ic == 0.f && sl == 1.f)) {   // (:1665-1669)
        bool ok = true;
        switch (info->dtype) {
            case FIBERS_F32: ok = rescale((float*)dst, n, sl, ic); break;
            case FIBERS_F64: { double* v = (double*)dst; for (size_t i = 0; i < n; ++i) v[i] = v[i] * (double)sl + (double)ic; } break;
            case FIBERS_I16: ok = rescale((int16_t*)dst, n, sl, ic); break;
            case FIBERS_U16: ok = rescale((uint16_t*)dst, n, sl, ic); break;
            case FIBERS_I32: ok = rescale((int32_t*)dst, n, sl, ic); break;
            case FIBERS_U32: ok = rescale((uint32_t*)dst, n, sl, ic); break;
            case FIBERS_U8: ok = rescale((uint8_t*)dst, n, sl, ic); break;
            case FIBERS_I8: ok = rescale((int8_t*)dst, n, sl, ic); break;
        }
        if (!ok) return fail(FIBERS_ERR_ARG, "InexactError: scl_slope / scl_inter do not map the stored integers to integers");
    }
    return 0;
}

namespace {
// vox2ras_to_qform (src/mri.jl:391-463), Float64 like the reference (vox2ras0 is promoted by the 1.0 literals)
int qform_of(const float* Mf, double q[7]) {
    double M[16]; for (int i = 0; i < 16; ++i) M[i] = Mf[i];
    double d[3], R[9];
    for (int j = 0; j < 3; ++j) d[j] = std::sqrt(M[j] * M[j] + M[4 + j] * M[4 + j] + M[8 + j] * M[8 + j] + M[12 + j] * M[12 + j]);   // sum over ALL four rows (:401)
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = M[4 * i + j] / d[j];
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det == 0) return 1;
    double r11 = R[0], r12 = R[1], r13 = R[2], r21 = R[3], r22 = R[4], r23 = R[5], r31 = R[6], r32 = R[7], r33 = R[8], qfac = 1.0;
    if (!(det > 0)) { r13 = -r13; r23 = -r23; r33 = -r33; qfac = -1.0; }
    double a = r11 + r22 + r33 + 1.0, b, c, dd;
    if (a > 0.5) { a = 0.5 * std::sqrt(a); b = 0.25 * (r32 - r23) / a; c = 0.25 * (r13 - r31) / a; dd = 0.25 * (r21 - r12) / a; }
    else {
        const double xd = 1.0 + r11 - (r22 + r33), yd = 1.0 + r22 - (r11 + r33), zd = 1.0 + r33 - (r11 + r22);
        if (xd > 1) { b = 0.5 * std::sqrt(xd); c = 0.25 * (r12 + r21) / b; dd = 0.25 * (r13 + r31) / b; a = 0.25 * (r32 - r23) / b; }
        else if (yd > 1) { c = 0.5 * std::sqrt(yd); b = 0.25 * (r12 + r21) / c; dd = 0.25 * (r23 + r32) / c; a = 0.25 * (r13 - r31) / c; }
        else { dd = 0.5 * std::sqrt(zd); b = 0.25 * (r13 + r31) / dd; c = 0.25 * (r23 + r32) / dd; a = 0.25 * (r21 - r12) / dd; }
        if (a < 0) { b = -b; c = -c; dd = -dd; }
    }
    q[0] = b; q[1] = c; q[2] = dd; q[3] = M[3]; q[4] = M[7]; q[5] = M[11]; q[6] = qfac;
    return 0;
}

// dtype.(vol) (src/mri.jl:2138): conversion to an integer type must be exact (Julia raises InexactError), to a float type it rounds
template <class S, class D> bool convert_block(const S* s, D* d, size_t n) {
    bool ok = true;
    for (size_t i = 0; i < n; ++i) { d[i] = (D)s[i]; if (!std::is_floating_point<D>::value && (S)d[i] != s[i]) ok = false; }
    return ok;
}
template <class S> bool convert_to(const S* s, void* d, int out, size_t n) {
    switch (out) {
        case FIBERS_F32: return convert_block(s, (float*)d, n);   case FIBERS_F64: return convert_block(s, (double*)d, n);
        case FIBERS_I16: return convert_block(s, (int16_t*)d, n); case FIBERS_U16: return convert_block(s, (uint16_t*)d, n);
        case FIBERS_I32: return convert_block(s, (int32_t*)d, n); case FIBERS_U32: return convert_block(s, (uint32_t*)d, n);
        case FIBERS_U8: return convert_block(s, (uint8_t*)d, n);  case FIBERS_I8: return convert_block(s, (int8_t*)d, n);
        default: return false;
    }
}
bool convert_any(const void* s, int in, void* d, int out, size_t n) {
    switch (in) {
        case FIBERS_F32: return convert_to((const float*)s, d, out, n);   case FIBERS_F64: return convert_to((const double*)s, d, out, n);
        case FIBERS_I16: return convert_to((const int16_t*)s, d, out, n); case FIBERS_U16: return convert_to((const uint16_t*)s, d, out, n);
        case FIBERS_I32: return convert_to((const int32_t*)s, d, out, n); case FIBERS_U32: return convert_to((const uint32_t*)s, d, out, n);
        case FIBERS_U8: return convert_to((const uint8_t*)s, d, out, n);  case FIBERS_I8: return convert_to((const int8_t*)s, d, out, n);
        default: return false;
    }
}
template <class T> void minmax(const T* v, size_t n, double& lo, double& hi) { lo = hi = (double)v[0]; for (size_t i = 1; i < n; ++i) { lo = std::min(lo, (double)v[i]); hi = std::max(hi, (double)v[i]); } }
}  // namespace

extern "C" int fibers_mri_write(const char* path, const void* vol, int dtype, const int32_t* dim, const float* vox2ras0, const float* volres,
                                float tr, float flip_angle, float te, float ti, float scl_slope, float scl_inter, int out_dtype) {
    if (!path || !vol || !dim || !vox2ras0) return fail(FIBERS_ERR_ARG, "NULL pointer");
    int gz = 0;
    const int fmt = format_of(path, gz);
    if (!fmt) return fail(FIBERS_ERR_ARG, std::string("Cannot determine format of ") + path);          // (:1727)
    const size_t es = dtype_size(dtype);
    if (!es || dim[0] <= 0 || dim[1] <= 0 || dim[2] <= 0 || dim[3] <= 0) return fail(FIBERS_ERR_ARG, "Input structure has empty vol field");   // (:1699)
    const size_t n = (size_t)dim[0] * dim[1] * dim[2] * dim[3];
    float res[3];
    if (volres) memcpy(res, volres, sizeof(res)); else volres_of(vox2ras0, res);
    Out out;
    if (!out.open(path, gz != 0)) return fail(FIBERS_ERR_ARG, std::string("Could not open ") + path + " for writing");
    const size_t CH = (size_t)1 << 22;                  // elements per conversion / byte-swap chunk
    std::vector<uint8_t> tmp;
    if (fmt == FIBERS_FMT_MGH) {                        // save_mgh (:1939-2036): everything big-endian, payload in the array's own type
        int type;
        switch (dtype) { case FIBERS_F32: type = 3; break; case FIBERS_U8: type = 0; break; case FIBERS_I32: type = 1; break; case FIBERS_I64: type = 2; break;
                         case FIBERS_I16: type = 4; break; case FIBERS_U16: type = 10; break;
                         default: return fail(FIBERS_ERR_ARG, "MGH: element type not supported"); }
        const bool sw = host_is_little();
        uint8_t h[284]; memset(h, 0, sizeof(h));
        auto put32 = [&](size_t off, int32_t v) { if (sw) swap_bytes(&v, 4, 1); memcpy(h + off, &v, 4); };
        auto putf = [&](size_t off, float v) { if (sw) swap_bytes(&v, 4, 1); memcpy(h + off, &v, 4); };
        put32(0, 1); put32(4, dim[0]); put32(8, dim[1]); put32(12, dim[2]); put32(16, dim[3]); put32(20, type); put32(24, 1);
        int16_t good = 1; if (sw) swap_bytes(&good, 2, 1); memcpy(h + 28, &good, 2);
        double M[16]; for (int i = 0; i < 16; ++i) M[i] = vox2ras0[i];
        double delta[3];
        for (int j = 0; j < 3; ++j) delta[j] = std::sqrt(M[j] * M[j] + M[4 + j] * M[4 + j] + M[8 + j] * M[8 + j]);
        for (int j = 0; j < 3; ++j) putf(30 + 4 * j, (float)delta[j]);
        for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) putf(42 + 4 * (3 * j + i), (float)(M[4 * i + j] / delta[j]));      // Mdc column-major
        const double crs[4] = {dim[0] / 2.0, dim[1] / 2.0, dim[2] / 2.0, 1.0};
        for (int i = 0; i < 3; ++i) { double acc = 0; for (int j = 0; j < 4; ++j) acc += M[4 * i + j] * crs[j]; putf(78 + 4 * i, (float)acc); }
        if (!out.write(h, 284)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
        if (sw && es > 1) {
            tmp.resize(std::min(n, CH) * es);
            for (size_t o = 0; o < n; o += CH) {
                const size_t c = std::min(CH, n - o);
                memcpy(tmp.data(), (const uint8_t*)vol + o * es, c * es); swap_bytes(tmp.data(), es, c);
                if (!out.write(tmp.data(), c * es)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
            }
        } else if (!out.write(vol, n * es)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
        float parms[4] = {tr, flip_angle, te, ti};
        if (sw) swap_bytes(parms, 4, 4);
        if (!out.write(parms, 16) || !out.close()) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
        return 0;
    }
    // ---- NIfTI-1: the header mri_write builds (:1733-1885), in the host's byte order ----
    if (out_dtype < 0) out_dtype = dtype;
    int bitpix = 0;
    const int code = nifti_code(out_dtype, bitpix);
    if (code < 0) return fail(FIBERS_ERR_ARG, "Data type not supported");                                 // (:1792)
    uint8_t h[352]; memset(h, 0, sizeof(h));
    auto put = [&](size_t off, auto v) { memcpy(h + off, &v, sizeof(v)); };
    put(0, (int32_t)348);
    int16_t d8[8] = {(int16_t)(dim[3] > 1 ? 4 : 3), (int16_t)dim[0], (int16_t)dim[1], (int16_t)dim[2], (int16_t)dim[3], 1, 1, 1};
    int32_t glmin = 0;
    if (dim[0] > 32768) { glmin = dim[0]; d8[1] = -1; }                                                      // (:1755-1758)
    memcpy(h + 40, d8, 16);
    put(70, (int16_t)code); put(72, (int16_t)bitpix);
    double q[7];
    if (qform_of(vox2ras0, q)) return fail(FIBERS_ERR_ARG, "vox2ras determinant is 0");                      // (:404)
    const float pixdim[8] = {(float)q[6], res[0], res[1], res[2], tr, 0.f, 0.f, 0.f};
    memcpy(h + 76, pixdim, 32);
    put(108, 352.f); put(112, scl_slope); put(116, scl_inter);
    h[123] = 2 | 16;                                                                                          // mm, msec
    {
        double lo = 0, hi = 0;                                                                               // cal_max / cal_min in the array's own type (:1822-1823)
        switch (dtype) { case FIBERS_F32: minmax((const float*)vol, n, lo, hi); break; case FIBERS_F64: minmax((const double*)vol, n, lo, hi); break;
                         case FIBERS_I16: minmax((const int16_t*)vol, n, lo, hi); break; case FIBERS_U16: minmax((const uint16_t*)vol, n, lo, hi); break;
                         case FIBERS_I32: minmax((const int32_t*)vol, n, lo, hi); break; case FIBERS_U32: minmax((const uint32_t*)vol, n, lo, hi); break;
                         case FIBERS_U8: minmax((const uint8_t*)vol, n, lo, hi); break; case FIBERS_I8: minmax((const int8_t*)vol, n, lo, hi); break;
                         default: return fail(FIBERS_ERR_ARG, "element type not supported"); }
        put(124, (float)hi); put(128, (float)lo);
    }
    put(144, glmin);
    { char desc[81]; snprintf(desc, sizeof(desc), "%-80s", "FreeSurfer julia"); memcpy(h + 148, desc, 80); }
    put(252, (int16_t)1); put(254, (int16_t)1);
    for (int i = 0; i < 6; ++i) put(256 + 4 * i, (float)q[i]);
    memcpy(h + 280, vox2ras0, 48);                                                                           // srow_x, srow_y, srow_z
    memcpy(h + 328, "huh?", 4);
    memcpy(h + 344, "n+1\0", 4);
    if (!out.write(h, 352)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
    if (out_dtype == dtype) { if (!out.write(vol, n * es)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path); }
    else {
        const size_t os = dtype_size(out_dtype);
        tmp.resize(std::min(n, CH) * os);
        for (size_t o = 0; o < n; o += CH) {
            const size_t c = std::min(CH, n - o);
            if (!convert_any((const uint8_t*)vol + o * es, dtype, tmp.data(), out_dtype, c)) return fail(FIBERS_ERR_ARG, "InexactError: the volume cannot be stored exactly in the requested integer type (or the type is not supported)");
            if (!out.write(tmp.data(), c * os)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
        }
    }
    if (!out.close()) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
    return 0;
}

// ---- TrackVis .trk (version 2) writer: trk_write (src/trk.jl:433-495) with the header Tract{T}(ref::MRI) builds (:88-145) ----
extern "C" int fibers_trk_write(const char* path, const int32_t* volsize, const float* volres, const float* vox2ras,
                                int64_t nstr, const int32_t* npts, const float* xyz) {
    return fibers_trk_write_ex(path, volsize, volres, vox2ras, nstr, npts, xyz, 0, nullptr, 0, nullptr);
}

extern "C" int fibers_trk_write_ex(const char* path, const int32_t* volsize, const float* volres, const float* vox2ras, int64_t nstr, const int32_t* npts,
                                   const float* xyz, int ns, const float* scalars, int np, const float* properties) {
    if (!path || !volsize || !volres || !vox2ras || (nstr > 0 && (!npts || !xyz))) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (ns < 0 || ns > 10 || np < 0 || np > 10 || (ns > 0 && nstr > 0 && !scalars) || (np > 0 && nstr > 0 && !properties))
        return fail(FIBERS_ERR_ARG, "between 0 and 10 scalars / properties, with their arrays");
    uint8_t h[1000]; memset(h, 0, sizeof(h));
    memcpy(h, "TRACK", 6);
    for (int i = 0; i < 3; ++i) { const int16_t d = (int16_t)volsize[i]; memcpy(h + 6 + 2 * i, &d, 2); }
    memcpy(h + 12, volres, 12);                                   // voxel_size; origin (24..35) = 0; names stay empty
    { const int16_t a = (int16_t)ns, b = (int16_t)np; memcpy(h + 36, &a, 2); memcpy(h + 238, &b, 2); }
    memcpy(h + 440, vox2ras, 64);                                 // vox_to_ras, row by row (permutedims before the write, :452)
    // voxel_order from vox2ras_to_orient (src/mri.jl:471-500): the dominant axis of every column and its sign
    for (int c = 0; c < 3; ++c) {
        int im = 0; float am = std::fabs(vox2ras[c]);
        for (int r = 1; r < 3; ++r) if (std::fabs(vox2ras[4 * r + c]) > am) { am = std::fabs(vox2ras[4 * r + c]); im = r; }
        const bool pos = vox2ras[4 * im + c] > 0;
        const char o = im == 0 ? (pos ? 'R' : 'L') : im == 1 ? (pos ? 'A' : 'P') : (pos ? 'S' : 'I');
        h[948 + c] = (uint8_t)o; h[952 + c] = (uint8_t)o;         // voxel_order, voxel_order_original (4 bytes each, NUL-terminated)
    }
    // image_orientation_patient = ([-1 0 0; 0 -1 0; 0 0 1] * vox2ras[1:3, 1:2] * Diagonal(1 ./ volres[1:2]))[:]   (:105-110)
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 3; ++r) {
            const float v = (float)((r < 2 ? -1.0 : 1.0) * (double)vox2ras[4 * r + c] * (1.0 / (double)volres[c]));
            memcpy(h + 956 + 4 * (3 * c + r), &v, 4);
        }
    const int32_t n32 = (int32_t)nstr, ver = 2, hs = 1000;
    memcpy(h + 988, &n32, 4); memcpy(h + 992, &ver, 4); memcpy(h + 996, &hs, 4);
    Out out;
    if (!out.open(path, false)) return fail(FIBERS_ERR_ARG, std::string("Could not open ") + path + " for writing");
    if (!out.write(h, 1000)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
    std::vector<float> buf;
    int64_t p0 = 0;
    for (int64_t i = 0; i < nstr; ++i) {
        const int32_t n = npts[i];
        buf.resize((size_t)(3 + ns) * n + 1 + np);
        memcpy(buf.data(), &n, 4);
        for (int32_t k = 0; k < n; ++k) {
            for (int c = 0; c < 3; ++c)                            // T.((xyz .+ .5) .* voxel_size): Float64 arithmetic, rounded once (:477)
                buf[1 + (size_t)(3 + ns) * k + c] = (float)(((double)xyz[3 * (p0 + k) + c] + 0.5) * (double)volres[c]);
            for (int c = 0; c < ns; ++c) buf[1 + (size_t)(3 + ns) * k + 3 + c] = scalars[(size_t)ns * (p0 + k) + c];
        }
        for (int c = 0; c < np; ++c) buf[1 + (size_t)(3 + ns) * n + c] = properties[(size_t)np * i + c];
        if (!out.write(buf.data(), buf.size() * 4)) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
        p0 += n;
    }
    if (!out.close()) return fail(FIBERS_ERR_ARG, std::string("Problem saving ") + path);
    return 0;
}

// trk_read (src/trk.jl:358-425): header fields + streamlines; points come back as xyz ./ voxel_size .- .5 (0-based voxel
// coordinates, Float32 arithmetic like the reference's broadcast over Vector{Float32}), scalars and properties as stored.
namespace {
bool trk_header(FILE* f, fibers_trk_info* info) {
    uint8_t h[1000];
    if (fread(h, 1, 1000, f) != 1000) return false;
    memset(info, 0, sizeof(*info));
    int16_t d[3]; memcpy(d, h + 6, 6);
    for (int i = 0; i < 3; ++i) info->dim[i] = d[i];
    memcpy(info->voxel_size, h + 12, 12); memcpy(info->origin, h + 24, 12);
    int16_t ns, np; memcpy(&ns, h + 36, 2); memcpy(&np, h + 238, 2);
    info->n_scalars = ns; info->n_properties = np;
    memcpy(info->vox_to_ras, h + 440, 64);                        // row by row (the reference transposes what it reads, :385)
    memcpy(info->voxel_order, h + 948, 4); memcpy(info->voxel_order_original, h + 952, 4);
    memcpy(info->image_orientation_patient, h + 956, 24);
    memcpy(&info->n_count, h + 988, 4); memcpy(&info->version, h + 992, 4); memcpy(&info->hdr_size, h + 996, 4);
    return memcmp(h, "TRACK", 5) == 0 && info->n_scalars >= 0 && info->n_properties >= 0 && info->n_count >= 0;
}
}  // namespace

extern "C" int fibers_trk_read_info(const char* path, fibers_trk_info* info) {
    if (!path || !info) return fail(FIBERS_ERR_ARG, "NULL pointer");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(FIBERS_ERR_ARG, std::string("Could not open ") + path + " for reading");
    bool ok = trk_header(f, info);
    int64_t total = 0;
    for (int32_t i = 0; ok && i < info->n_count; ++i) {          // one pass over the point counts
        int32_t n;
        ok = fread(&n, 4, 1, f) == 1 && n >= 0 &&
             fseek(f, (long)(((int64_t)n * (3 + info->n_scalars) + info->n_properties) * 4), SEEK_CUR) == 0;
        total += ok ? n : 0;
    }
    fclose(f);
    if (!ok) return fail(FIBERS_ERR_ARG, std::string("Not a readable .trk file: ") + path);
    info->total_points = total;
    return 0;
}

extern "C" int fibers_trk_read_data(const char* path, const fibers_trk_info* info, int32_t* npts, float* xyz, float* scalars, float* properties) {
    if (!path || !info || (info->n_count > 0 && (!npts || !xyz))) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if ((info->n_scalars > 0 && !scalars) || (info->n_properties > 0 && !properties)) return fail(FIBERS_ERR_ARG, "NULL scalars / properties");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(FIBERS_ERR_ARG, std::string("Could not open ") + path + " for reading");
    fibers_trk_info h2;
    bool ok = trk_header(f, &h2) && h2.n_count == info->n_count && h2.n_scalars == info->n_scalars && h2.n_properties == info->n_properties;
    const int ns = info->n_scalars, np = info->n_properties;
    std::vector<float> buf;
    int64_t p0 = 0;
    for (int32_t i = 0; ok && i < info->n_count; ++i) {
        int32_t n;
        ok = fread(&n, 4, 1, f) == 1 && n >= 0 && p0 + n <= info->total_points;
        if (!ok) break;
        npts[i] = n;
        buf.resize((size_t)n * (3 + ns) + np);
        ok = buf.empty() || fread(buf.data(), 4, buf.size(), f) == buf.size();
        if (!ok) break;
        for (int32_t k = 0; k < n; ++k) {
            const float* src = buf.data() + (size_t)k * (3 + ns);
            for (int c = 0; c < 3; ++c) xyz[3 * (p0 + k) + c] = src[c] / info->voxel_size[c] - 0.5f;
            for (int c = 0; c < ns; ++c) scalars[(size_t)ns * (p0 + k) + c] = src[3 + c];
        }
        for (int c = 0; c < np; ++c) properties[(size_t)np * i + c] = buf[(size_t)n * (3 + ns) + c];
        p0 += n;
    }
    fclose(f);
    if (!ok || p0 != info->total_points) return fail(FIBERS_ERR_ARG, std::string("Problem reading ") + path);
    return 0;
}
