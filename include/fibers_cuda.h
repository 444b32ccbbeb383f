/*
 * libfibers_cuda.so -- C ABI of the B200-native voxel-wise diffusion reconstruction path
 * (drop-in for the hot loops of lincbrain/Fibers.jl v1.0.0).
 *
 * The reference has no FFI layer: its boundary is the exported Julia API
 *   adc_fit(dwi::MRI, mask::MRI)                       src/dti.jl:164
 *   dti_fit(dwi::MRI, mask::MRI)                       src/dti.jl:221  (-> dti_fit_ls :243)
 *   gqi_rec(dwi, mask, odf_dirs=sphere_642, σ=1.25f0)  src/gqi.jl:109
 *   dsi_rec(dwi, mask, odf_dirs=sphere_642, hann=32)   src/dsi.jl:171
 * Each host entry point below replaces the `Threads.@threads for iz` voxel nest (and, for
 * GQI/DSI, the serial odfmax post-pass) of one of those functions with one blocking call;
 * julia/FibersCUDA.jl shows the `ccall` a maintainer adds.
 *
 * Conventions
 *  - Every array is Julia column-major [nx, ny, nz, nframes]: the voxel index is contiguous,
 *    the frame (volume / vertex / xyz component) has stride nx*ny*nz  (src/mri.jl:249-255).
 *  - Host entry points take HOST pointers; the caller owns every buffer; the library only
 *    borrows them for the duration of the call.  Outputs are fully overwritten (voxels that
 *    the reference leaves untouched are written as 0, matching its zero-filled MRI ctor).
 *  - `mask` is uint8, non-zero = inside (Julia side: UInt8.(mask.vol .!= 0)).
 *  - `faces` is int32, 1-based, [nface, 3] column-major, exactly `odf_dirs.faces`;
 *    `vertices` is float32 [nvert2, 3] column-major, exactly `odf_dirs.vertices`.
 *  - `bvec` is float32 [nvol, 3] column-major, `bval` float32 [nvol].
 *  - Return value: 0 = OK, otherwise a FIBERS_ERR_* code; fibers_cuda_last_error() gives the
 *    message for the calling thread (Julia wrapper re-raises it with error(msg), matching the
 *    reference's exception convention, e.g. src/gqi.jl:111-117).
 *  - There is NO CPU fallback: without a usable CUDA device every compute call fails with
 *    FIBERS_ERR_NODEV.
 *  - ngpu: number of GPUs to shard z-slabs over (>=1; clipped to the configured device list).
 *    No inter-GPU collective is used; the host gathers slabs and reduces the one scalar (odfmax).
 *  - Host memory: arrays in PINNED / registered memory (cudaHostAlloc, cudaHostRegister,
 *    fibers_cuda_host_register) are copied by DMA straight from / into the caller's buffers.  Arrays in
 *    ordinary PAGEABLE memory (what a Julia `Array` is, src/mri.jl:249-255) are accepted as they are and go
 *    through an internal pinned bounce ring filled / emptied by a few host copy threads; the choice is made per
 *    call from cudaPointerGetAttributes (override: env FIBERS_CUDA_HOST_PATH=direct|bounce;
 *    FIBERS_CUDA_COPY_THREADS, default 16 per GPU, never more than the CPUs the process may use).  The worker and copy threads of a GPU are bound to the CPUs
 *    local to it (FIBERS_CUDA_AFFINITY=0 disables).
 *  - gqi_rec / dsi_rec / dti_gqi_fit: `odf` may be NULL when the caller only needs peaks and QA (the only
 *    downstream consumer, stream(), reads nothing else: src/stream.jl:76-173); the ODF is then formed on the
 *    device for the peak search but never crosses PCIe (4.7 of the 4.9 GB a HCP-shaped subject returns).
 */
#ifndef FIBERS_CUDA_H
#define FIBERS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIBERS_OK          0
#define FIBERS_ERR_ARG     1   /* bad argument (null pointer, non-positive size, unsupported mesh) */
#define FIBERS_ERR_TABLE   2   /* missing b-value / gradient table (reference: error(...) src/dti.jl:166-168,223-229) */
#define FIBERS_ERR_NODEV   3   /* no CUDA device / driver: there is no CPU fallback */
#define FIBERS_ERR_CUDA    4   /* CUDA runtime error, message has the detail */
#define FIBERS_ERR_NOMEM   5   /* device or pinned-host allocation failed */

/* dwi element types accepted by gqi_rec / dsi_rec (the reference converts with `.=`,
 * src/gqi.jl:139; dti_fit / adc_fit are Float32-only like the reference's method signature
 * src/dti.jl:286). */
#define FIBERS_F32 0
#define FIBERS_F64 1
#define FIBERS_I16 2
#define FIBERS_U16 3
#define FIBERS_I32 4
#define FIBERS_U8  5
/* further element types of the volume I/O entry points only (NIfTI Int8 / UInt32, MGH long) */
#define FIBERS_I8  6
#define FIBERS_U32 7
#define FIBERS_I64 8

/* Reconstruction kernel selection for GQI/DSI (all are CUDA paths). */
#define FIBERS_KERNEL_AUTO  0   /* tensor-core path when the shape allows it, else SIMT */
#define FIBERS_KERNEL_SIMT  1   /* fp32 CUDA-core tiled contraction + fused peak epilogue */
#define FIBERS_KERNEL_TC    2   /* tcgen05 split-fp16 (fp32-accurate) contraction + fused epilogue */

/* ---- library / device management ------------------------------------------------------- */
int         fibers_cuda_version(void);                 /* 100*major + minor */
int         fibers_cuda_device_count(void);            /* number of visible CUDA devices (0 if none) */
const char* fibers_cuda_last_error(void);              /* thread-local, never NULL */
/* Device ordinals used by the host entry points for shards 0..n-1 (default 0,1,2,...; the env
 * var FIBERS_CUDA_DEVICES="3,1" is read at first use). */
int         fibers_cuda_set_devices(const int* devices, int n);
/* The host entry points keep per-device streams, slab buffers and the last plan between calls (a batch
 * of subjects with one protocol pays for set-up once); this frees them (optional, e.g. before exit). */
void        fibers_cuda_release_cache(void);
/* Force a reconstruction kernel for subsequently created plans / host calls (FIBERS_KERNEL_*). */
int         fibers_cuda_set_kernel(int kernel);

/* ---- host-pointer entry points (what the Julia wrapper ccalls) ------------------------- */

/* dti_fit / dti_fit_ls: src/dti.jl:243-278 voxel nest + :286-316 voxel fit + :325-335 maps.
 * Outputs s0, eval*, rd, md, fa: [nx,ny,nz]; evec*: [nx,ny,nz,3].  valid (optional, may be
 * NULL): uint8 [nx,ny,nz], 1 where the voxel took the full or partial fit branch. */
int fibers_dti_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec,
                   float* s0, float* eval1, float* eval2, float* eval3,
                   float* evec1, float* evec2, float* evec3,
                   float* rd, float* md, float* fa, uint8_t* valid, int ngpu);

/* adc_fit: src/dti.jl:164-213. */
int fibers_adc_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                   const float* bval, float* adc, float* s0, int ngpu);

/* gqi_rec: src/gqi.jl:109-171 (work set-up :42-82, peaks :180-201, odfmax post-pass :164-168).
 * odf: [nx,ny,nz,nvert2/2]; peak1..3: [nx,ny,nz,3]; qa1..3: [nx,ny,nz] (already divided by odfmax).
 * peak_idx (optional, test aid; may be NULL): int16 [nx,ny,nz,3], 0-based vertex index or -1. */
int fibers_gqi_rec(const void* dwi, int dwi_dtype, const uint8_t* mask,
                   int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec,
                   const float* vertices, int nvert2, const int32_t* faces, int nface, float sigma,
                   float* odf, float* peak1, float* peak2, float* peak3,
                   float* qa1, float* qa2, float* qa3, int16_t* peak_idx, int ngpu);

/* dsi_rec: src/dsi.jl:171-270 (work set-up :59-143).  pdf: [nx,ny,nz,nvol]. */
int fibers_dsi_rec(const void* dwi, int dwi_dtype, const uint8_t* mask,
                   int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec,
                   const float* vertices, int nvert2, const int32_t* faces, int nface, int hann_width,
                   float* pdf, float* odf, float* peak1, float* peak2, float* peak3,
                   float* qa1, float* qa2, float* qa3, int16_t* peak_idx, int ngpu);

/* dti_fit + gqi_rec on the SAME dwi volume in one pass: the result of
 * calling fibers_dti_fit (src/dti.jl:221-316) and fibers_gqi_rec (src/gqi.jl:109-171) one after the other,
 * bit for bit, but every z-slab chunk of the DWI crosses PCIe once and feeds both kernels while it is
 * resident.  dwi must be Float32 (the reference's dti_fit accepts nothing else, src/dti.jl:286). */
int fibers_dti_gqi_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                       const float* bval, const float* bvec,
                       float* s0, float* eval1, float* eval2, float* eval3,
                       float* evec1, float* evec2, float* evec3,
                       float* rd, float* md, float* fa,
                       const float* vertices, int nvert2, const int32_t* faces, int nface, float sigma,
                       float* odf, float* peak1, float* peak2, float* peak3,
                       float* qa1, float* qa2, float* qa3, int ngpu);

/* cfg4-style batch: dti_fit + gqi_rec (src/dti.jl:221, src/gqi.jl:109) of `nsub` subjects that share one
 * protocol and one volume shape, in one call.  dwi[i], mask[i]: subject i's arrays; dti_out: nsub x 10 pointers
 * (s0, eval1-3, evec1-3, rd, md, fa per subject; the whole table may be NULL = GQI only); gqi_out: nsub x 7
 * pointers (odf, peak1-3, qa1-3 per subject; an odf entry may be NULL).  Subjects are handed to the `ngpu`
 * devices from an ordered queue; each device keeps its stream ring rolling across subject boundaries, so the H2D
 * of subject i+1 overlaps the kernels and the D2H of subject i, and the plans are built once per device.  With
 * fewer subjects than GPUs every subject is also split into z-slabs.  Results are bit-identical to nsub calls
 * of fibers_dti_gqi_fit. */
int fibers_dti_gqi_fit_batch(int nsub, const float* const* dwi, const uint8_t* const* mask,
                             int nx, int ny, int nz, int nvol, const float* bval, const float* bvec,
                             float* const* dti_out,
                             const float* vertices, int nvert2, const int32_t* faces, int nface, float sigma,
                             float* const* gqi_out, int ngpu);

/* rumba_rec: src/rusd.jl:419-636 (RUMBA-SD: Richardson-Lucy deconvolution with a Rician / noncentral-chi likelihood and
 * total-variation regularisation).  dwi must be Float32 (the reference's method signature).
 * mask_pos: uint8, 1 where mask.vol > 0 (voxels that are deconvolved, src/rusd.jl:441); mask_any: 1 where mask.vol != 0 (voxels
 * whose peaks are extracted, :614; NULL = mask_pos).  ang_neig: angular neighbourhood of the peak search in degrees (12.5
 * for sphere_724 / sphere_642, 16 for sphere_362, :479-483).  coil_combine: 0 = "SMF-SENSE", 1 = "SoS-GRAPPA".
 * Outputs: fodf [nx,ny,nz,nvert2/2]; fgm, fcsf, gfa, var [nx,ny,nz]; peak1..5 [nx,ny,nz,3]; *snr_mean, *snr_std;
 * peak_idx (optional test aid): int16 [nx,ny,nz,5], 0-based vertex or -1.
 * The total-variation term couples neighbouring voxels in every iteration: the volume is processed on ONE device
 * (`device` ordinal); a batch of subjects uses one device per subject.  The three matrix products of an iteration are plain
 * GEMMs and go through cuBLAS, loaded at run time; everything else is fused into two kernels per iteration. */
int fibers_rumba_rec(const float* dwi, const uint8_t* mask_pos, const uint8_t* mask_any,
                     int nx, int ny, int nz, int nvol, const float* bval, const float* bvec,
                     const float* vertices, int nvert2, float ang_neig, int niter,
                     float lambda_para, float lambda_perp, float lambda_csf, float lambda_gm,
                     int ncoils, int coil_combine, int ipat_factor, int use_tv,
                     float* fodf, float* fgm, float* fcsf,
                     float* peak1, float* peak2, float* peak3, float* peak4, float* peak5,
                     float* gfa, float* var, float* snr_mean, float* snr_std, int16_t* peak_idx, int device);

/* st_eigen / st_recon: src/structens.jl:13-34, :40-88 (structure tensor).  Float32 volumes [nx,ny,nz].
 * eigvec: [nx,ny,nz,3,3] with eigvec[x,y,z,:,k] the k-th eigenvector, eigval: [nx,ny,nz,3] ascending -- what
 * StaticArrays' eigen(Symmetric(S, :L)) returns, i.e. the reference's output arrays verbatim.  st_recon applies
 * ImageFiltering's separable Gaussian (sigma: pre-smoothing, rho: tensor smoothing; <= 0 skips) and Scharr factors with
 * "reflect" borders, forms the six products and calls st_eigen.  The _device variant works on device pointers
 * (frame f of eigvec / eigval starts at f * nvox). */
int fibers_st_eigen(const float* sxx, const float* sxy, const float* sxz, const float* syy, const float* syz, const float* szz,
                    int nx, int ny, int nz, float* eigvec, float* eigval, int device);
int fibers_st_recon(const float* vol, int nx, int ny, int nz, float sigma, float rho, float* eigvec, float* eigval, int device);
int fibers_st_eigen_device(const float* d_sxx, const float* d_sxy, const float* d_sxz, const float* d_syy, const float* d_syz,
                           const float* d_szz, int64_t nvox, float* d_eigvec, float* d_eigval, void* stream);

/* stream: src/stream.jl:730-790 (streamline tractography), the deterministic regime: orientation
 * VECTORS (ovec[i] = [nx,ny,nz,3] float32, i < nvec <= 8, e.g. the peak volumes of gqi_rec / dsi_rec or eigvec1 of dti_fit),
 * no local connection matrices (those: fibers_stream_lcm below); macroscopic voxels and the microscopy regime.  Replaces the StreamWork constructor's masking (:72-147), the
 * `Threads.@threads` seed loop (:759-781) and stream_new_line / stream_new_point! / stream_pick_by_angle! (:621-690,
 * :497-541, :355-387).
 *   f        nvec volumes [nx,ny,nz] (vector amplitudes, e.g. qa1..3) or NULL; vectors with f < f_thresh are dropped (:136-138)
 *   fa       [nx,ny,nz] or NULL; voxels with fa < fa_thresh leave the mask (:128)
 *   mask     uint8 [nx,ny,nz] (Julia side: UInt8.(mask.vol[:,:,:,1] .> 0)) or NULL = "any vector component non-zero" (:107-112)
 *   seed     uint8 [nx,ny,nz] (seed.vol .> 0) or NULL = the final mask (:744-754)
 *   sublist  [nsub][3] sub-voxel offsets.  The reference draws them from rand(Uniform(-.5+eps(), .5-eps()), 3) (:177-183, or a
 *            single zero offset when nsub == 0): the WRAPPER draws them (same call, Julia's RNG) and passes them in.
 *   cosang_thresh = cosd(T(ang_thresh)), step_size, smooth_coeff, len_min, len_max: the StreamWork fields of the same name.
 *   micro_search_dist  NULL = macroscopic regime; [3] half widths of the search box = the microscopy regime (StreamWork sets it when
 *            min(volres) <= 0.05: fill(search_dist, 3), 0 along the through-plane axis of 2-D data, :84-90, :153), in which the next
 *            position is searched around the tentative step (stream_micro_new_point! :547-617); micro_search_cosang = cosd(T(search_ang)).
 * Output order = the reference's: seed voxels in column-major order, sub-voxel samples innermost, lines shorter than len_min
 * dropped; coordinates are the reference's 1-based voxel coordinates.  The number of points is not known beforehand, so the call
 * returns an opaque handle plus the counts; the caller allocates npts[nstr] (Int32) and xyz[3, npts_total] (Float32: the
 * concatenated [3, npts] matrices of Tract.xyz), calls fibers_stream_fetch and releases the handle with fibers_stream_free.
 * The _device variant takes DEVICE volume pointers (peaks / qa planes a reconstruction left on the current device). */
int fibers_stream(const float* const* ovec, int nvec, int nx, int ny, int nz, const float* const* f, float f_thresh,
                  const float* fa, float fa_thresh, const uint8_t* mask, const uint8_t* seed, const float* sublist, int nsub,
                  int len_min, int len_max, float cosang_thresh, float step_size, float smooth_coeff,
                  const int32_t* micro_search_dist /*[3] or NULL*/, float micro_search_cosang, int device,
                  void** result, int64_t* nstr, int64_t* npts_total);
int fibers_stream_device(const float* const* d_ovec, int nvec, int nx, int ny, int nz, const float* const* d_f, float f_thresh,
                         const float* d_fa, float fa_thresh, const uint8_t* d_mask, const uint8_t* d_seed,
                         const float* sublist, int nsub, int len_min, int len_max, float cosang_thresh,
                         float step_size, float smooth_coeff, const int32_t* micro_search_dist, float micro_search_cosang,
                         void** result, int64_t* nstr, int64_t* npts_total);
int fibers_stream_fetch(void* result, int32_t* npts, float* xyz);
void fibers_stream_free(void* result);

/* stream(...; lcms, lcm_thresh): the branch that follows local connection matrices (src/stream.jl:208-236 set-up, :523-538
 * stream_new_point!, :380-494 stream_pick_by_lcm!), macroscopic regime.  Every step first makes the conventional pick (for the
 * method-difference flag), then, on entering a new voxel, zeroes the LCM elements that do not touch the entry edge, normalises
 * them and draws one connection: rand(Categorical(lcm)) in the reference, on Julia's task-local generator (the answer depends on
 * the thread schedule).  Here draw k of streamline l (reference order, before the len_min filter) is the counter-based uniform
 * number  u = top 24 bits of splitmix64_finalise(lcm_seed + (l + 1) * 0x9E3779B97F4A7C15 + (k + 1) * 0xD1B54A32D192ED03) * 2^-24,
 * and the connection is the first i with p[1] + ... + p[i] > u (Distributions' DiscreteNonParametric sampler).  No angle
 * threshold is applied (:676).
 *   lcms      [nx,ny,nz,10] float32 like lcms.vol; elements < lcm_thresh (compared in Float64, :220) count as 0
 *   strdim1/2 the in-plane dimensions, 0-based: setdiff(1:3, thrudim) .- 1 with thrudim = the component of the FIRST orientation
 *             volume that is zero everywhere (:224-226; the wrapper evaluates it, like cosd(ang_thresh) for the other branch)
 * fibers_stream_fetch_scalars: one Float32 per point, in the order of fibers_stream_fetch's xyz: 1 where the LCM pick differed
 * from the conventional one (the scalars str_add! receives, :783). */
int fibers_stream_lcm(const float* const* ovec, int nvec, int nx, int ny, int nz, const float* const* f, float f_thresh,
                      const float* fa, float fa_thresh, const uint8_t* mask, const uint8_t* seed, const float* sublist, int nsub,
                      int len_min, int len_max, float step_size, float smooth_coeff, const float* lcms, double lcm_thresh,
                      int strdim1, int strdim2, uint64_t lcm_seed, int device, void** result, int64_t* nstr, int64_t* npts_total);
int fibers_stream_fetch_scalars(void* result, float* scalars);

/* ---- volume I/O either side of the path: NIfTI-1 (.nii, .nii.gz) and MGH (.mgh, .mgz) ----------
 * fibers_mri_read_info   = load_nifti_hdr (src/mri.jl:1394-1551) / the header of load_mgh (:1217-1283) + what mri_read derives
 *                          (:611-700): dims beyond the 4th folded into frames, units converted to mm / ms, vox2ras0 = sform if
 *                          sform_code != 0, else qform if qform_code != 0, else diag(pixdim); volres from vox2ras0.
 * fibers_mri_read_data   = the payload of load_nifti (:1640-1672) / load_mgh (:1305-1310): stored element type, byte order
 *                          fixed, scl_slope / scl_inter applied by the reference's rule.  `dst` is column-major
 *                          [dim0, dim1, dim2, dim3] of info->dtype; it may be pinned memory (fibers_cuda_host_register), so the
 *                          reconstruction calls DMA straight from it.
 * fibers_mri_write       = mri_write + save_nifti / save_mgh (:1695-1937, :2059-2176, :1939-2036): the format follows the
 *                          extension; `out_dtype` (NIfTI only, < 0 = keep) is mri_write's `datatype` argument; `volres` may be NULL
 *                          (derived from vox2ras0 as mri_write does :1719-1721).
 * Compressed files are inflated / deflated in-process (zlib) straight into / from the caller's buffer; the reference pipes them
 * through zcat / gzip and a temporary file (:1581-1592, :2160-2163).  Matrices are ROW-major 4 x 4 here. */
#define FIBERS_FMT_NIFTI 1
#define FIBERS_FMT_MGH   2
typedef struct fibers_mri_info {
    int32_t format;          /* FIBERS_FMT_* */
    int32_t gz;              /* .nii.gz / .mgz */
    int32_t ndim;            /* 3 or 4 */
    int32_t dim[4];          /* volsize[3], nframes */
    int32_t dtype;           /* FIBERS_* element type as stored */
    int32_t bswap;           /* stored byte order differs from the host's (the reader fixes it) */
    int32_t sform_code, qform_code;
    int64_t data_offset;     /* round(vox_offset) (NIfTI) or 284 (MGH) */
    float vox2ras0[16], sform[16], qform[16];
    float pixdim[8];         /* NIfTI, in mm / ms */
    float volres[3];
    float tr, flip_angle, te, ti;
    float scl_slope, scl_inter;
} fibers_mri_info;
int fibers_mri_read_info(const char* path, fibers_mri_info* info);
int fibers_mri_read_data(const char* path, const fibers_mri_info* info, void* dst, int64_t dst_bytes);
int fibers_mri_write(const char* path, const void* vol, int dtype, const int32_t* dim /*[4]*/, const float* vox2ras0 /*[16]*/,
                     const float* volres /*[3] or NULL*/, float tr, float flip_angle, float te, float ti,
                     float scl_slope, float scl_inter, int out_dtype);

/* trk_write (src/trk.jl:433-495) for the Tract that stream() returns: TrackVis .trk version 2 with the header Tract{T}(ref::MRI)
 * builds (:88-145) from the reference volume (volsize, volres, vox2ras row-major; voxel order from vox2ras_to_orient,
 * src/mri.jl:471-500), no scalars / properties; points are stored as (xyz + .5) * voxel_size like the reference. */
int fibers_trk_write(const char* path, const int32_t* volsize /*[3]*/, const float* volres /*[3]*/, const float* vox2ras /*[16]*/,
                     int64_t nstr, const int32_t* npts, const float* xyz /*[3, sum(npts)]*/);
/* the same with n_scalars values per point (scalars [n_scalars, sum(npts)]) and n_properties per streamline (properties
 * [n_properties, nstr]), unnamed like the Tract str_add! fills (src/trk.jl:166-260, write loop :470-490) */
int fibers_trk_write_ex(const char* path, const int32_t* volsize, const float* volres, const float* vox2ras, int64_t nstr, const int32_t* npts,
                        const float* xyz, int n_scalars, const float* scalars, int n_properties, const float* properties);

/* trk_read (src/trk.jl:358-425): two calls, like the volume reader.  `info` = the header fields a caller of the reference's
 * Tract reads back + the total number of points (so that the caller can allocate); `data` fills npts [n_count], xyz
 * [3, total_points] as xyz ./ voxel_size .- .5 (0-based voxel coordinates, :412-413), scalars [n_scalars, total_points] and
 * properties [n_properties, n_count] (NULL allowed when the counts are 0). */
typedef struct {
    int32_t dim[3];
    float voxel_size[3], origin[3];
    int32_t n_scalars, n_properties;
    float vox_to_ras[16];        /* row-major */
    char voxel_order[4], voxel_order_original[4];
    float image_orientation_patient[6];
    int32_t n_count, version, hdr_size;
    int64_t total_points;
} fibers_trk_info;
int fibers_trk_read_info(const char* path, fibers_trk_info* info);
int fibers_trk_read_data(const char* path, const fibers_trk_info* info, int32_t* npts, float* xyz, float* scalars, float* properties);

/* Optional: page-lock a caller-owned host array (and release it) so that later calls take the direct DMA path.
 * Worth it for arrays that are used more than once (registration itself costs about as much as one copy). */
int fibers_cuda_host_register(void* ptr, size_t bytes);
int fibers_cuda_host_unregister(void* ptr);

/* ---- device-resident entry points (kernel-only timing, slab pipelines, batch drivers) ---
 * A plan holds the per-protocol constants on ONE device (reconstruction matrix, neighbour
 * table, vertex table, pinv of the design matrix): the GPU analogue of GQIwork / DSIwork /
 * DTIwork (src/gqi.jl:32-82, src/dsi.jl:41-143, src/dti.jl:101-155).
 * All d_* pointers are device pointers on the plan's device.  Frame f of an array starts at
 * base + f*pitch (pitch in ELEMENTS, >= nvox): a z-slab of a larger volume is addressed by
 * offsetting base and keeping the full-volume pitch.  Calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream). */
typedef struct fibers_plan fibers_plan;

int  fibers_dti_plan_create(fibers_plan** plan, int device, int nvol, const float* bval, const float* bvec);
int  fibers_adc_plan_create(fibers_plan** plan, int device, int nvol, const float* bval);
int  fibers_gqi_plan_create(fibers_plan** plan, int device, int nvol, const float* bval, const float* bvec,
                            const float* vertices, int nvert2, const int32_t* faces, int nface, float sigma);
int  fibers_dsi_plan_create(fibers_plan** plan, int device, int nvol, const float* bval, const float* bvec,
                            const float* vertices, int nvert2, const int32_t* faces, int nface, int hann_width);
void fibers_plan_destroy(fibers_plan* plan);
/* Introspection for tests: copies the plan's reconstruction matrix (row-major [rows, nvol],
 * float32) to host.  GQI: rows = nvert (A of src/gqi.jl:69); DSI: rows = nvert + nvol (Mo;Mp);
 * DTI/ADC: rows = 7 / 2 (pinv(A), src/dti.jl:143,72).  Returns rows, or <0 on error. */
int  fibers_plan_matrix(const fibers_plan* plan, float* out, int64_t capacity);
/* Which kernel the plan's recon calls will launch (FIBERS_KERNEL_SIMT or FIBERS_KERNEL_TC). */
int  fibers_plan_kernel(const fibers_plan* plan);

int fibers_dti_fit_device(fibers_plan* plan, const float* d_dwi, int64_t dwi_pitch,
                          const uint8_t* d_mask, int64_t nvox, int64_t out_pitch,
                          float* d_s0, float* d_eval1, float* d_eval2, float* d_eval3,
                          float* d_evec1, float* d_evec2, float* d_evec3,
                          float* d_rd, float* d_md, float* d_fa, uint8_t* d_valid, void* stream);
int fibers_adc_fit_device(fibers_plan* plan, const float* d_dwi, int64_t dwi_pitch,
                          const uint8_t* d_mask, int64_t nvox, float* d_adc, float* d_s0, void* stream);

/* GQI / DSI on a device-resident slab.  d_pdf is used by DSI plans only (NULL for GQI).
 * d_peak_idx may be NULL.  d_stats: 2 x int32 device scratch owned by the caller:
 *   [0] = max over the slab's voxels of mean_v(odf), as an order-preserving int encoding
 *   [1] = reserved.
 * If `finalize` != 0 the call zero-initialises d_stats, runs the slab and divides the three QA
 * planes by the slab's own odfmax on the same stream (single-slab use, src/gqi.jl:164-168).
 * If `finalize` == 0 the caller initialises d_stats once with fibers_stats_init_device, may
 * run several slabs into it, and then calls fibers_qa_scale_device. */
int fibers_recon_device(fibers_plan* plan, const float* d_dwi, int64_t dwi_pitch,
                        const uint8_t* d_mask, int64_t nvox, int64_t out_pitch,
                        float* d_pdf, float* d_odf,
                        float* d_peak1, float* d_peak2, float* d_peak3,
                        float* d_qa1, float* d_qa2, float* d_qa3,
                        int16_t* d_peak_idx, int32_t* d_stats, int finalize, void* stream);
int fibers_stats_init_device(int32_t* d_stats, void* stream);
/* qa_k[v] /= odfmax for v < nvox.  If d_stats != NULL the divisor is decoded from d_stats[0]
 * on the device (no host sync) and `odfmax` is ignored. */
int fibers_qa_scale_device(float* d_qa1, float* d_qa2, float* d_qa3, int64_t nvox,
                           const int32_t* d_stats, float odfmax, void* stream);
/* Decode helper for hosts that combine several slabs: order-preserving int -> float. */
float fibers_stats_decode_max(int32_t encoded);
/* Number of kernel launches the library has issued in this process (all devices). */
int64_t fibers_cuda_launch_count(void);

/* ---- host-only set-up helpers (no device needed; exercised by the CPU test-suite) -------
 * fibers_host_build_matrix: kind 1 = DTI pinv [7,nvol], 2 = ADC pinv [2,nvol], 3 = GQI A [M,nvol]
 * (src/gqi.jl:66-69), 4 = DSI [Mo;Mp] [M+nvol,nvol] with *cvol / *dscale (den = dscale*s+[cvol]).
 * Writes row-major float32 into out (capacity in elements), returns the row count or <0. */
int fibers_host_build_matrix(int kind, int nvol, const float* bval, const float* bvec,
                             const float* vertices, int nvert2, float sigma, int hann_width,
                             float* out, int64_t capacity, int* cvol, float* dscale);
/* RUMBA-SD set-up (src/rusd.jl:444-520, :478-492): kernel [ndir, nvert2/2 + 2] column-major (capacity in elements, may be NULL),
 * vol_row int32 [nvol] (0 = minimum-b volume, else the 1-based kernel row), nbr uint16 [nvert2/2, 16] row-major, 0xFFFF
 * terminated angular neighbourhoods.  Returns ndir or < 0. */
int fibers_host_build_rumba(int nvol, const float* bval, const float* bvec, const float* vertices, int nvert2, float ang_neig,
                            float lambda_para, float lambda_perp, float lambda_csf, float lambda_gm,
                            float* kernel, int64_t capacity, int32_t* vol_row, uint16_t* nbr);
/* Folded-mesh neighbour table: out is uint16 [nvert, 8] row-major, 0xFFFF = none. */
int fibers_host_build_neighbours(const int32_t* faces, int nface, int nvert, uint16_t* out);
/* z-slab partition used by the host entry points: out[2g], out[2g+1] = voxel range of shard g. */
int fibers_host_partition_slabs(const uint8_t* mask, int64_t nxny, int nz, int ngpu, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* FIBERS_CUDA_H */
