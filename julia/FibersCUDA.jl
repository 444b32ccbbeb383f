#=
  FibersCUDA.jl -- drop-in GPU back end for the voxel-wise reconstruction functions of Fibers.jl

  Same signatures, same MRI / ODF inputs, same DTI / GQI / DSI outputs as the reference
  (src/dti.jl:164,221  src/gqi.jl:109  src/dsi.jl:171); the `Threads.@threads for iz` voxel nest
  and the serial odfmax post-pass of each function are replaced by ONE blocking `ccall` into
  libfibers_cuda.so (include/fibers_cuda.h).  No CUDA.jl, no CPU fallback: if the library or a
  CUDA device is missing the call raises `error(msg)`.

  NOTE: Julia is not installed in the build image, so this file has never been executed; the
  identical C ABI is exercised from Python/ctypes (fibers.jl_b200/recon.py) with Fortran-ordered
  arrays, which is byte-for-byte what the ccalls below pass.  See INTEGRATION.md.

  Usage inside Fibers.jl (after `include("gqi.jl")` etc.):
      include("FibersCUDA.jl"); using .FibersCUDA
      gqi = FibersCUDA.gqi_rec(dwi, mask)              # or: Fibers.gqi_rec = FibersCUDA.gqi_rec
=#
module FibersCUDA

using ..Fibers: MRI, ODF, DTI, GQI, DSI, sphere_642

export adc_fit, dti_fit, gqi_rec, dsi_rec, dti_gqi_fit, device_count

const libfibers = get(ENV, "FIBERS_CUDA_LIB", "libfibers_cuda.so")
const NGPU = Ref{Cint}(parse(Cint, get(ENV, "FIBERS_CUDA_NGPU", "1")))

device_count() = Int(ccall((:fibers_cuda_device_count, libfibers), Cint, ()))

function check(rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:fibers_cuda_last_error, libfibers), Cstring, ()))
  error(msg)                       # same exception convention as the reference (src/gqi.jl:111-117)
end

# dwi element-type codes of include/fibers_cuda.h
dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float64}) = Cint(1)
dtype_code(::Type{Int16})   = Cint(2)
dtype_code(::Type{UInt16})  = Cint(3)
dtype_code(::Type{Int32})   = Cint(4)
dtype_code(::Type{UInt8})   = Cint(5)

mask_u8(mask::MRI) = UInt8.(reshape(mask.vol, size(mask.vol)[1:3]) .!= 0)    # masks may be [nx,ny,nz,1] label maps

"""
    adc_fit(dwi::MRI, mask::MRI)

GPU version of `Fibers.adc_fit` (src/dti.jl:164-213).
"""
function adc_fit(dwi::MRI, mask::MRI)
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  adc = MRI(mask, 1, Float32)
  s0  = MRI(mask, 1, Float32)
  vol = dwi.vol::Array{Float32,4}                      # reference method signature is Float32-only
  nx, ny, nz, nvol = size(vol)
  m = mask_u8(mask)
  check(ccall((:fibers_adc_fit, libfibers), Cint,
              (Ptr{Float32}, Ptr{UInt8}, Cint, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cint),
              vol, m, nx, ny, nz, nvol, dwi.bval, adc.vol, s0.vol, NGPU[]))
  return adc, s0
end

"""
    dti_fit(dwi::MRI, mask::MRI)

GPU version of `Fibers.dti_fit` / `dti_fit_ls` (src/dti.jl:221-316, maps :325-335).
"""
function dti_fit(dwi::MRI, mask::MRI)
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  isempty(dwi.bvec) && error("Missing gradient table from input DWI structure")
  S0    = MRI(mask, 1, Float32); Eval1 = MRI(mask, 1, Float32)
  Eval2 = MRI(mask, 1, Float32); Eval3 = MRI(mask, 1, Float32)
  Evec1 = MRI(mask, 3, Float32); Evec2 = MRI(mask, 3, Float32); Evec3 = MRI(mask, 3, Float32)
  RD    = MRI(mask, 1, Float32); MD    = MRI(mask, 1, Float32); FA    = MRI(mask, 1, Float32)
  vol = dwi.vol::Array{Float32,4}
  nx, ny, nz, nvol = size(vol)
  m = mask_u8(mask)
  bvec = Matrix{Float32}(dwi.bvec)                     # [nvol, 3] column-major
  check(ccall((:fibers_dti_fit, libfibers), Cint,
              (Ptr{Float32}, Ptr{UInt8}, Cint, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}, Cint),
              vol, m, nx, ny, nz, nvol, dwi.bval, bvec,
              S0.vol, Eval1.vol, Eval2.vol, Eval3.vol, Evec1.vol, Evec2.vol, Evec3.vol,
              RD.vol, MD.vol, FA.vol, C_NULL, NGPU[]))
  return DTI(S0, Eval1, Eval2, Eval3, Evec1, Evec2, Evec3, RD, MD, FA)
end

"""
    gqi_rec(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25))

GPU version of `Fibers.gqi_rec` (src/gqi.jl:109-171, peaks :180-201).
"""
function gqi_rec(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25))
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  isempty(dwi.bvec) && error("Missing gradient table from input DWI structure")
  npeak = 3
  nvert = div(size(odf_dirs.vertices, 1), 2)
  odf  = MRI(mask, nvert, Float32)
  peak = [MRI(mask, 3, Float32) for _ in 1:npeak]
  qa   = [MRI(mask, 1, Float32) for _ in 1:npeak]
  nx, ny, nz, nvol = size(dwi.vol)
  m = mask_u8(mask)
  bvec  = Matrix{Float32}(dwi.bvec)
  faces = Matrix{Int32}(odf_dirs.faces)                # Matrix{Integer} cannot cross ccall
  check(ccall((:fibers_gqi_rec, libfibers), Cint,
              (Ptr{Cvoid}, Cint, Ptr{UInt8}, Cint, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Cint, Ptr{Int32}, Cint, Cfloat,
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
               Ptr{Int16}, Cint),
              dwi.vol, dtype_code(eltype(dwi.vol)), m, nx, ny, nz, nvol, dwi.bval, bvec,
              odf_dirs.vertices, size(odf_dirs.vertices, 1), faces, size(faces, 1), σ,
              odf.vol, peak[1].vol, peak[2].vol, peak[3].vol, qa[1].vol, qa[2].vol, qa[3].vol,
              C_NULL, NGPU[]))
  return GQI(odf, peak, qa)
end

"""
    dti_gqi_fit(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25))

`dti_fit(dwi, mask)` and `gqi_rec(dwi, mask, odf_dirs, σ)` in ONE pass over `dwi.vol` (every z-slab chunk is
copied to the GPU once and feeds both kernels); returns `(DTI, GQI)`, bit-identical to the two separate calls.
"""
function dti_gqi_fit(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25))
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  isempty(dwi.bvec) && error("Missing gradient table from input DWI structure")
  S0    = MRI(mask, 1, Float32); Eval1 = MRI(mask, 1, Float32)
  Eval2 = MRI(mask, 1, Float32); Eval3 = MRI(mask, 1, Float32)
  Evec1 = MRI(mask, 3, Float32); Evec2 = MRI(mask, 3, Float32); Evec3 = MRI(mask, 3, Float32)
  RD    = MRI(mask, 1, Float32); MD    = MRI(mask, 1, Float32); FA    = MRI(mask, 1, Float32)
  nvert = div(size(odf_dirs.vertices, 1), 2)
  odf  = MRI(mask, nvert, Float32)
  peak = [MRI(mask, 3, Float32) for _ in 1:3]
  qa   = [MRI(mask, 1, Float32) for _ in 1:3]
  vol = dwi.vol::Array{Float32,4}                      # dti_fit's method signature is Float32-only
  nx, ny, nz, nvol = size(vol)
  m = mask_u8(mask)
  bvec  = Matrix{Float32}(dwi.bvec)
  faces = Matrix{Int32}(odf_dirs.faces)
  check(ccall((:fibers_dti_gqi_fit, libfibers), Cint,
              (Ptr{Float32}, Ptr{UInt8}, Cint, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Cint, Ptr{Int32}, Cint, Cfloat,
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cint),
              vol, m, nx, ny, nz, nvol, dwi.bval, bvec,
              S0.vol, Eval1.vol, Eval2.vol, Eval3.vol, Evec1.vol, Evec2.vol, Evec3.vol, RD.vol, MD.vol, FA.vol,
              odf_dirs.vertices, size(odf_dirs.vertices, 1), faces, size(faces, 1), σ,
              odf.vol, peak[1].vol, peak[2].vol, peak[3].vol, qa[1].vol, qa[2].vol, qa[3].vol, NGPU[]))
  return DTI(S0, Eval1, Eval2, Eval3, Evec1, Evec2, Evec3, RD, MD, FA), GQI(odf, peak, qa)
end

"""
    dsi_rec(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, hann_width::Int=32)

GPU version of `Fibers.dsi_rec` (src/dsi.jl:171-270).
"""
function dsi_rec(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, hann_width::Int=32)
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  isempty(dwi.bvec) && error("Missing gradient table from input DWI structure")
  npeak = 3
  nvert = div(size(odf_dirs.vertices, 1), 2)
  nx, ny, nz, nvol = size(dwi.vol)
  pdf  = MRI(mask, nvol, Float32)
  odf  = MRI(mask, nvert, Float32)
  peak = [MRI(mask, 3, Float32) for _ in 1:npeak]
  qa   = [MRI(mask, 1, Float32) for _ in 1:npeak]
  m = mask_u8(mask)
  bvec  = Matrix{Float32}(dwi.bvec)
  faces = Matrix{Int32}(odf_dirs.faces)
  check(ccall((:fibers_dsi_rec, libfibers), Cint,
              (Ptr{Cvoid}, Cint, Ptr{UInt8}, Cint, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32},
               Ptr{Float32}, Cint, Ptr{Int32}, Cint, Cint,
               Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
               Ptr{Int16}, Cint),
              dwi.vol, dtype_code(eltype(dwi.vol)), m, nx, ny, nz, nvol, dwi.bval, bvec,
              odf_dirs.vertices, size(odf_dirs.vertices, 1), faces, size(faces, 1), Cint(hann_width),
              pdf.vol, odf.vol, peak[1].vol, peak[2].vol, peak[3].vol, qa[1].vol, qa[2].vol, qa[3].vol,
              C_NULL, NGPU[]))
  return DSI(pdf, odf, peak, qa)
end

end # module
