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

using ..Fibers: MRI, ODF, DTI, GQI, DSI, sphere_642, Tract, str_add!
using Distributions: Uniform

export adc_fit, dti_fit, gqi_rec, dsi_rec, dti_gqi_fit, dti_gqi_fit_batch, device_count

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
The C ABI receives bare pointers and takes every length from the dwi volume: a mask of another size or a short
b-table would be read / written out of bounds.  The reference throws BoundsError / DimensionMismatch in these
cases (its loops index the arrays); the wrapper checks explicitly and raises the same kind of exception.
"""
function check_dims(dwi::MRI, mask::MRI; need_bvec::Bool=true, odf_dirs::Union{ODF,Nothing}=nothing)
  ndims(dwi.vol) == 4 || throw(DimensionMismatch("dwi.vol must be [nx, ny, nz, nvol], got size $(size(dwi.vol))"))
  nx, ny, nz, nvol = size(dwi.vol)
  msz = size(mask.vol)
  (length(msz) >= 3 && msz[1:3] == (nx, ny, nz) && prod(msz) == nx * ny * nz) ||
    throw(DimensionMismatch("mask size $(msz) does not match dwi size $((nx, ny, nz))"))
  length(dwi.bval) == nvol ||
    throw(DimensionMismatch("b-value table has $(length(dwi.bval)) entries for $nvol volumes"))
  eltype(dwi.bval) == Float32 || throw(ArgumentError("dwi.bval must be Vector{Float32} (as mri_read returns it)"))
  if need_bvec
    size(dwi.bvec) == (nvol, 3) ||
      throw(DimensionMismatch("gradient table has size $(size(dwi.bvec)), expected ($nvol, 3)"))
  end
  if odf_dirs !== nothing
    nv2 = size(odf_dirs.vertices, 1)
    (size(odf_dirs.vertices, 2) == 3 && iseven(nv2) && nv2 > 0) ||
      throw(DimensionMismatch("odf_dirs.vertices must be [2M, 3], got $(size(odf_dirs.vertices))"))
    eltype(odf_dirs.vertices) == Float32 || throw(ArgumentError("odf_dirs.vertices must be Matrix{Float32}"))
    size(odf_dirs.faces, 2) == 3 || throw(DimensionMismatch("odf_dirs.faces must be [F, 3]"))
    all(f -> 1 <= f <= nv2, odf_dirs.faces) || throw(BoundsError(odf_dirs.vertices, maximum(odf_dirs.faces)))
  end
  return nx, ny, nz, nvol
end

"""
    adc_fit(dwi::MRI, mask::MRI)

GPU version of `Fibers.adc_fit` (src/dti.jl:164-213).
"""
function adc_fit(dwi::MRI, mask::MRI)
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  check_dims(dwi, mask; need_bvec=false)
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
  check_dims(dwi, mask)
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
function gqi_rec(dwi::MRI, mask::MRI, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25); want_odf::Bool=true)
  isempty(dwi.bval) && error("Missing b-value table from input DWI structure")
  isempty(dwi.bvec) && error("Missing gradient table from input DWI structure")
  check_dims(dwi, mask; odf_dirs=odf_dirs)
  npeak = 3
  nvert = div(size(odf_dirs.vertices, 1), 2)
  # want_odf=false (extension): the ODF is formed on the GPU for the peak search but not copied back; `.odf` is then
  # a 1-frame placeholder.  `stream` reads peaks and QA only (src/stream.jl:76-173).
  odf  = MRI(mask, want_odf ? nvert : 1, Float32)
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
              want_odf ? pointer(odf.vol) : Ptr{Float32}(C_NULL),
              peak[1].vol, peak[2].vol, peak[3].vol, qa[1].vol, qa[2].vol, qa[3].vol,
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
  check_dims(dwi, mask; odf_dirs=odf_dirs)
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
  check_dims(dwi, mask; odf_dirs=odf_dirs)
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

"""
    dti_gqi_fit_batch(dwis::Vector{MRI}, masks::Vector{MRI}, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25))

`dti_gqi_fit` of a batch of subjects that share one protocol and one volume shape in ONE library call: the
subjects are queued over `FIBERS_CUDA_NGPU` devices and every device overlaps the transfers of the next subject
with the kernels of the current one.  Returns a vector of `(DTI, GQI)`.
"""
function dti_gqi_fit_batch(dwis::Vector{MRI}, masks::Vector{MRI}, odf_dirs::ODF=sphere_642, σ::Float32=Float32(1.25))
  length(dwis) == length(masks) || throw(DimensionMismatch("dwis and masks must have the same length"))
  isempty(dwis) && return Tuple{DTI,GQI}[]
  d1 = dwis[1]
  isempty(d1.bval) && error("Missing b-value table from input DWI structure")
  isempty(d1.bvec) && error("Missing gradient table from input DWI structure")
  for (d, m) in zip(dwis, masks)
    check_dims(d, m; odf_dirs=odf_dirs)
    (size(d.vol) == size(d1.vol) && d.bval == d1.bval && d.bvec == d1.bvec) ||
      throw(DimensionMismatch("all subjects of a batch must share the volume shape and the b-table"))
  end
  nx, ny, nz, nvol = size(d1.vol)
  nvert = div(size(odf_dirs.vertices, 1), 2)
  vols = [d.vol::Array{Float32,4} for d in dwis]
  ms   = [mask_u8(m) for m in masks]
  res  = [(DTI(MRI(m, 1, Float32), MRI(m, 1, Float32), MRI(m, 1, Float32), MRI(m, 1, Float32), MRI(m, 3, Float32),
               MRI(m, 3, Float32), MRI(m, 3, Float32), MRI(m, 1, Float32), MRI(m, 1, Float32), MRI(m, 1, Float32)),
           GQI(MRI(m, nvert, Float32), [MRI(m, 3, Float32) for _ in 1:3], [MRI(m, 1, Float32) for _ in 1:3])) for m in masks]
  dti_tab = Ptr{Float32}[]; gqi_tab = Ptr{Float32}[]
  for (d, g) in res
    append!(dti_tab, pointer.([d.s0.vol, d.eigval1.vol, d.eigval2.vol, d.eigval3.vol, d.eigvec1.vol, d.eigvec2.vol,
                               d.eigvec3.vol, d.rd.vol, d.md.vol, d.fa.vol]))
    append!(gqi_tab, pointer.([g.odf.vol, g.peak[1].vol, g.peak[2].vol, g.peak[3].vol, g.qa[1].vol, g.qa[2].vol, g.qa[3].vol]))
  end
  bvec  = Matrix{Float32}(d1.bvec)
  faces = Matrix{Int32}(odf_dirs.faces)
  GC.@preserve vols ms res begin                        # the pointer tables do not root the arrays they point to
    check(ccall((:fibers_dti_gqi_fit_batch, libfibers), Cint,
                (Cint, Ptr{Ptr{Float32}}, Ptr{Ptr{UInt8}}, Cint, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32},
                 Ptr{Ptr{Float32}}, Ptr{Float32}, Cint, Ptr{Int32}, Cint, Cfloat, Ptr{Ptr{Float32}}, Cint),
                length(dwis), pointer.(vols), pointer.(ms), nx, ny, nz, nvol, d1.bval, bvec,
                dti_tab, odf_dirs.vertices, size(odf_dirs.vertices, 1), faces, size(faces, 1), σ, gqi_tab, NGPU[]))
  end
  return res
end

"""
    stream(ovec; f, f_thresh, fa, fa_thresh, mask, seed, nsub, len_min, len_max, ang_thresh, step_size, smooth_coeff)

GPU version of `Fibers.stream` (src/stream.jl:730-790): macroscopic and microscopy regime, and (macroscopic) local
connection matrices `lcms` (`fibers_stream_lcm`; the connection draws come from a counter-based generator seeded with
`lcm_seed`, which this wrapper draws from Julia's RNG unless given).  The StreamWork constructor's masking, the seed loop and the propagation run in
`fibers_stream`; this wrapper keeps the reference's defaults, draws the sub-voxel offsets exactly like StreamWork
does (src/stream.jl:177-183) and assembles the `Tract`.
"""
function stream(ovec::Union{MRI,Vector{MRI}}; f::Union{MRI,Vector{MRI},Nothing}=nothing, f_thresh::Real=.03,
                fa::Union{MRI,Nothing}=nothing, fa_thresh::Real=.1, mask::Union{MRI,Nothing}=nothing,
                seed::Union{MRI,Nothing}=nothing, nsub::Union{Integer,Nothing}=3, len_min::Integer=3,
                len_max::Integer=(isa(ovec,MRI) ? maximum(ovec.volsize) : maximum(ovec[1].volsize)),
                ang_thresh::Union{Real,Nothing}=45, step_size::Union{Real,Nothing}=.5,
                smooth_coeff::Union{Real,Nothing}=.2, search_dist::Integer=15, search_ang::Real=10,
                lcms::Union{MRI,Nothing}=nothing, lcm_thresh::Real=.099, lcm_seed::UInt64=rand(UInt64))
  ovecs = isa(ovec, MRI) ? MRI[ovec] : ovec
  fs    = isa(f, MRI) ? MRI[f] : f
  nx, ny, nz = size(ovecs[1].vol)[1:3]
  domicro = (minimum(ovecs[1].volres) <= 0.05)                    # src/stream.jl:84
  micro_search_dist = domicro ? fill(Int32(search_dist), 3) : Int32[]
  # 2-D orientation angles -> in-plane unit vectors, exactly as the StreamWork constructor does (src/stream.jl:145-172)
  function as_vectors(o::MRI)
    size(o.vol, 4) == 3 && return Array{Float32,4}(o.vol)
    size(o.vol, 4) == 1 || error("Input orientations should be 3D vectors or angles ∊ [-90, 90]")
    thrudim = argmax(o.volres); strdims = setdiff(1:3, thrudim)
    domicro && (micro_search_dist[thrudim] = 0)
    a = Float32.(o.vol[:,:,:,1]); v = zeros(Float32, nx, ny, nz, 3)
    if -π/2-eps(Float32) <= minimum(a) && maximum(a) <= π/2+eps(Float32)
      v[:,:,:,strdims[1]] .= cos.(a);  v[:,:,:,strdims[2]] .= sin.(a)
    elseif -90 <= minimum(a) && maximum(a) <= 90
      v[:,:,:,strdims[1]] .= cosd.(a); v[:,:,:,strdims[2]] .= sind.(a)
    else
      error("Input orientations should be 3D vectors or angles ∊ [-90, 90]")
    end
    isnothing(mask) && (v .*= (a .!= 0))                          # the mask would be any(x -> x != 0, vol) (:107-112)
    return v
  end
  isnothing(nsub) && (nsub = domicro ? 0 : 3); isnothing(ang_thresh) && (ang_thresh = domicro ? 20 : 45)
  isnothing(step_size) && (step_size = domicro ? 1 : .5); isnothing(smooth_coeff) && (smooth_coeff = domicro ? 0 : .2)
  if !isnothing(seed) && !isnothing(mask) && size(seed.vol) != size(mask.vol)
    error("Dimension mismatch between seed mask " * string(size(seed.vol)) * " and brain mask " * string(size(mask.vol)))
  end
  sublist = nsub > 0 ? [Float32.(rand(Uniform(-.5+eps(), .5-eps()), 3)) for isub in 1:nsub] : [zeros(Float32, 3)]
  sub  = reduce(hcat, sublist)                                   # [3, nsub] column-major = [nsub][3] for the C side
  vols = [as_vectors(o) for o in ovecs]
  fvol = isnothing(fs) ? nothing : [Array{Float32,3}(x.vol[:,:,:,1]) for x in fs]
  favol = isnothing(fa) ? nothing : Array{Float32,3}(fa.vol[:,:,:,1])
  m  = isnothing(mask) ? nothing : UInt8.(mask.vol[:,:,:,1] .> 0)
  sd = isnothing(seed) ? nothing : UInt8.(seed.vol[:,:,:,1] .> 0)
  handle = Ref{Ptr{Cvoid}}(C_NULL); nstr = Ref{Int64}(0); ntot = Ref{Int64}(0)
  if !isnothing(lcms)
    domicro && error("stream: local connection matrices are only defined for the macroscopic regime")
    size(lcms.vol) == (nx, ny, nz, 10) || error("stream: lcms must be [nx ny nz 10]")
    thrudim = findall(vec(all(x -> x==0, ovecs[1].vol, dims=(1,2,3))))   # src/stream.jl:224-226
    strdims = setdiff(1:3, thrudim)
    lv = Array{Float32,4}(lcms.vol)
    GC.@preserve vols fvol begin
      check(ccall((:fibers_stream_lcm, libfibers), Cint,
                  (Ptr{Ptr{Float32}}, Cint, Cint, Cint, Cint, Ptr{Ptr{Float32}}, Cfloat, Ptr{Float32}, Cfloat, Ptr{UInt8}, Ptr{UInt8},
                   Ptr{Float32}, Cint, Cint, Cint, Cfloat, Cfloat, Ptr{Float32}, Cdouble, Cint, Cint, UInt64, Cint,
                   Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}),
                  pointer.(vols), length(vols), nx, ny, nz, isnothing(fvol) ? C_NULL : pointer.(fvol), Float32(f_thresh),
                  isnothing(favol) ? C_NULL : favol, Float32(fa_thresh), isnothing(m) ? C_NULL : m, isnothing(sd) ? C_NULL : sd,
                  sub, size(sub, 2), len_min, len_max, Float32(step_size), Float32(smooth_coeff), lv, Float64(lcm_thresh),
                  strdims[1]-1, strdims[2]-1, lcm_seed, 0, handle, nstr, ntot))
    end
  else
  GC.@preserve vols fvol begin
    check(ccall((:fibers_stream, libfibers), Cint,
                (Ptr{Ptr{Float32}}, Cint, Cint, Cint, Cint, Ptr{Ptr{Float32}}, Cfloat, Ptr{Float32}, Cfloat, Ptr{UInt8}, Ptr{UInt8},
                 Ptr{Float32}, Cint, Cint, Cint, Cfloat, Cfloat, Cfloat, Ptr{Int32}, Cfloat, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}),
                pointer.(vols), length(vols), nx, ny, nz, isnothing(fvol) ? C_NULL : pointer.(fvol), Float32(f_thresh),
                isnothing(favol) ? C_NULL : favol, Float32(fa_thresh), isnothing(m) ? C_NULL : m, isnothing(sd) ? C_NULL : sd,
                sub, size(sub, 2), len_min, len_max, cosd(Float32(ang_thresh)), Float32(step_size), Float32(smooth_coeff),
                domicro ? micro_search_dist : C_NULL, domicro ? cosd(Float32(search_ang)) : 0f0, 0, handle, nstr, ntot))
  end
  end
  npts = Vector{Int32}(undef, nstr[]); xyz = Matrix{Float32}(undef, 3, ntot[]); flags = Vector{Float32}(undef, ntot[])
  try
    check(ccall((:fibers_stream_fetch, libfibers), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float32}), handle[], npts, xyz))
    isnothing(lcms) || check(ccall((:fibers_stream_fetch_scalars, libfibers), Cint, (Ptr{Cvoid}, Ptr{Float32}), handle[], flags))
  finally
    ccall((:fibers_stream_free, libfibers), Cvoid, (Ptr{Cvoid},), handle[])
  end
  ends = cumsum(Int.(npts))
  str = [xyz[:, (ends[i]-npts[i]+1):ends[i]] for i in eachindex(npts)]
  tr = Tract{Float32}(isnothing(mask) ? ovecs[1] : mask)
  if isnothing(lcms)
    str_add!(tr, str)                                             # src/stream.jl:785-787
  else
    str_add!(tr, str, [flags[(ends[i]-npts[i]+1):ends[i]] for i in eachindex(npts)])   # the method-difference flags (:783)
  end
  return tr
end

# fibers_mri_info (include/fibers_cuda.h); matrices are row-major on the C side
struct MriInfo
  format::Int32; gz::Int32; ndim::Int32; dim::NTuple{4,Int32}; dtype::Int32; bswap::Int32; sform_code::Int32; qform_code::Int32
  data_offset::Int64
  vox2ras0::NTuple{16,Float32}; sform::NTuple{16,Float32}; qform::NTuple{16,Float32}; pixdim::NTuple{8,Float32}
  volres::NTuple{3,Float32}; tr::Float32; flip_angle::Float32; te::Float32; ti::Float32; scl_slope::Float32; scl_inter::Float32
end
const IO_TYPES = (Float32, Float64, Int16, UInt16, Int32, UInt8, Int8, UInt32, Int64)      # FIBERS_F32 ... FIBERS_I64

"""
    load_volume(fname) -> (info::MriInfo, vol)

Header + payload of a NIfTI-1 / MGH file through the library (replaces `load_nifti` / `load_mgh`, src/mri.jl:1576, :1217,
inside `mri_read`): no temporary file for .nii.gz / .mgz, byte order and scl_slope / scl_inter handled as in the reference.
"""
function load_volume(fname::String)
  info = Ref{MriInfo}()
  check(ccall((:fibers_mri_read_info, libfibers), Cint, (Cstring, Ptr{MriInfo}), fname, info))
  h = info[]
  T = IO_TYPES[h.dtype + 1]
  vol = Array{T}(undef, Int.(h.ndim >= 4 ? h.dim : h.dim[1:3])...)
  check(ccall((:fibers_mri_read_data, libfibers), Cint, (Cstring, Ptr{MriInfo}, Ptr{Cvoid}, Int64), fname, info, vol, sizeof(vol)))
  return h, vol
end

"""
    save_volume(fname, vol, vox2ras0, volres, mr_parms, scl_slope, scl_inter, datatype)

`save_nifti` / `save_mgh` (src/mri.jl:2059, :1939) with the header `mri_write` builds (:1733-1885), format by extension.
"""
function save_volume(fname::String, vol::Array{T}, vox2ras0::Matrix, volres::Vector, mr_parms::Vector=zeros(4),
                     scl_slope::Real=0, scl_inter::Real=0, datatype::DataType=T) where T<:Number
  dim = Int32[size(vol, 1), size(vol, 2), size(vol, 3), size(vol, 4)]
  M = Matrix{Float32}(permutedims(vox2ras0))                      # row-major for the C side
  code(t) = Cint(findfirst(==(t), IO_TYPES) - 1)
  check(ccall((:fibers_mri_write, libfibers), Cint,
              (Cstring, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Cint),
              fname, vol, code(T), dim, M, Float32.(volres), mr_parms[1], mr_parms[2], mr_parms[3], mr_parms[4],
              scl_slope, scl_inter, code(datatype)))
  return false
end

end # module
