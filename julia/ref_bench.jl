#=
  ref_bench.jl -- times the REAL reference (Fibers.jl `gqi_rec`, src/gqi.jl:109) on the synthetic cfg2-shaped
  workload of bench.py, for anyone who has Julia and the package.  NOT RUN in this repository's environment
  (no Julia in the build image or on the GPU boxes): bench.py's `--impl reference` arm times the C/OpenMP
  restatement of the same loop instead and says so in its JSON line.

      JULIA_NUM_THREADS=auto julia --project=<Fibers.jl checkout> julia/ref_bench.jl [nx ny nz]

  Prints one JSON line: voxels/s of gqi_rec + peaks on all Julia threads.
=#
using Fibers, Random, LinearAlgebra, Printf

nx, ny, nz = length(ARGS) >= 3 ? parse.(Int, ARGS[1:3]) : (145, 174, 16)     # default: a z-sub-slab of 145x174x145
nb0, ndir, shells = 18, 90, (1000f0, 2000f0, 3000f0)

# gradient table: Fibonacci-sphere directions per shell (same construction as fibers.jl_b200/phantom.py)
function fib_dirs(n, rot)
  v = zeros(Float32, n, 3)
  for i in 1:n
    z = 1 - (2i - 1) / n
    r = sqrt(max(0, 1 - z^2)); ϕ = Float32(π) * (3 - sqrt(5f0)) * (i - 1) + rot
    v[i, :] .= (r * cos(ϕ), r * sin(ϕ), z)
  end
  return v
end
bval = vcat(zeros(Float32, nb0), [fill(b, ndir) for b in shells]...)
bvec = vcat(zeros(Float32, nb0, 3), [fib_dirs(ndir, 0.7f0 * k) for k in 1:length(shells)]...)
nvol = length(bval)

# two-fibre multi-tensor signal + isotropic compartment, Rician noise (SNR 30), ~0.1 % negatives
rng = MersenneTwister(2)
nvox = nx * ny * nz
randdir() = (v = randn(rng, Float32, 3); v ./ norm(v))
S = zeros(Float32, nvox, nvol)
for i in 1:nvox
  e1, e2 = randdir(), randdir()
  f1 = (0.3f0 + 0.4f0 * rand(rng, Float32)) * 0.9f0; f2 = 0.9f0 - f1
  s0 = 500f0 + 1000f0 * rand(rng, Float32); σn = s0 / 30
  for k in 1:nvol
    g = @view bvec[k, :]
    s = s0 * (f1 * exp(-bval[k] * (2f-4 + 1.5f-3 * dot(g, e1)^2)) + f2 * exp(-bval[k] * (2f-4 + 1.5f-3 * dot(g, e2)^2)) +
              0.1f0 * exp(-bval[k] * 3f-3))
    s = sqrt((s + σn * randn(rng, Float32))^2 + (σn * randn(rng, Float32))^2)
    S[i, k] = rand(rng) < 1e-3 ? -0.1f0 * s : s
  end
end
dwi = MRI(reshape(S, nx, ny, nz, nvol))
dwi.bval = bval; dwi.bvec = bvec
mask = MRI(ones(Float32, nx, ny, nz))

gqi_rec(dwi, mask)                                  # warm-up (compilation)
t = @elapsed gqi_rec(dwi, mask)
@printf("{\"impl\": \"Fibers.jl gqi_rec\", \"metric\": \"voxels/sec (GQI recon+peaks)\", \"value\": %.1f, \"unit\": \"voxels/s\", \"threads\": %d, \"shape\": [%d, %d, %d, %d], \"seconds\": %.3f}\n",
        nvox / t, Threads.nthreads(), nx, ny, nz, nvol, t)
