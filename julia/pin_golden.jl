#=
  pin_golden.jl -- pinning kit, step 2 of 3 (see tools/pin/export_inputs.py): runs the UNMODIFIED reference (Fibers.jl
  dti_fit / adc_fit / gqi_rec / dsi_rec with their default arguments, rumba_rec, stream, st_recon) on the inputs of this repository's golden
  fixtures and writes its volumes with the reference's own *_write functions.  NOT RUN where this repository was built
  (no Julia there): it exists so that anyone with Julia can pin the oracle against the real thing.

      julia --project=<Fibers.jl checkout> julia/pin_golden.jl <inputs dir> <outputs dir>
=#
using Fibers, DelimitedFiles

indir, outdir = ARGS[1], ARGS[2]
mkpath(outdir)

function load(name)
  dwi  = mri_read(joinpath(indir, name * "_dwi.nii.gz"))
  mask = mri_read(joinpath(indir, name * "_mask.nii.gz"))
  dwi.bval = vec(readdlm(joinpath(indir, name * "_bval.txt"), Float32))     # assigned as they are: no re-normalisation
  dwi.bvec = readdlm(joinpath(indir, name * "_bvec.txt"), Float32)
  return dwi, mask
end

dwi, mask = load("dti_small")
dti_write(dti_fit(dwi, mask), joinpath(outdir, "dti_small"))                # src/dti.jl:221, :344
adc, s0 = adc_fit(dwi, mask)                                                # src/dti.jl:164
mri_write(adc, joinpath(outdir, "dti_small_adc.nii.gz")); mri_write(s0, joinpath(outdir, "dti_small_adc_s0.nii.gz"))

dwi, mask = load("gqi_small")
gqi_write(gqi_rec(dwi, mask), joinpath(outdir, "gqi_small"))                # src/gqi.jl:109 (sphere_642, sigma = 1.25), :210

dwi, mask = load("dsi_small")
dsi_write(dsi_rec(dwi, mask), joinpath(outdir, "dsi_small"))                # src/dsi.jl:171 (sphere_642, hann_width = 32), :279

dwi, mask = load("rumba_small")
rumba_write(rumba_rec(dwi, mask, sphere_362, 40), joinpath(outdir, "rumba_small"))   # src/rusd.jl:419 (40 iterations, other arguments default), :645

# stream (src/stream.jl:730) without random sub-voxel offsets (nsub = 0): deterministic; the Tract goes out as .trk (src/trk.jl:433)
ovec = [mri_read(joinpath(indir, "stream_small_ovec$i.nii.gz")) for i in 1:2]
f    = [mri_read(joinpath(indir, "stream_small_f$i.nii.gz")) for i in 1:2]
tr = stream(ovec; f=f, f_thresh=0.05, mask=mri_read(joinpath(indir, "stream_small_mask.nii.gz")), nsub=0)
trk_write(tr, joinpath(outdir, "stream_small.trk"))

# structure tensor (src/structens.jl:40-88), sigma = 1, rho = 2: eigenvalues as 3 frames, eigenvectors as 9 frames ([:, :, :, i, j] -> frame i + 3 (j - 1))
v = mri_read(joinpath(indir, "structens_small_vol.nii.gz"))
eigvec, eigval = st_recon(Float32.(v.vol[:, :, :, 1]), 1.0, 2.0)
o = MRI(v, 3); o.vol = eigval;                                   mri_write(o, joinpath(outdir, "structens_small_eigval.nii.gz"))
o = MRI(v, 9); o.vol = reshape(eigvec, size(eigvec)[1:3]..., 9); mri_write(o, joinpath(outdir, "structens_small_eigvec.nii.gz"))

println("reference outputs written to ", outdir)
