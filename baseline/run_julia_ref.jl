# run_julia_ref.jl — times the UNMODIFIED reference tracer (RayTraceHeatTransfer.jl, method = :exchange) on the
# bench workload, for `bench.py --impl reference` on machines that have a `julia` binary and the package installed
# (e.g. `julia --project=baseline/_ref -e 'using Pkg; Pkg.develop(path="/root/reference")'`).  This image has no Julia,
# so bench.py falls back to the C restatement in oracle/ and says so (cpu_baseline.kind = "port").
#
#   julia -t auto baseline/run_julia_ref.jl <Ndim> <kappa> <sigma_s> <rays_total> <steps>
# prints one line:  RTHX_JULIA_REF rays_per_s=<value> threads=<n> rays=<traced per step>
using RayTraceHeatTransfer
using StaticArrays, GeometryBasics

Ndim   = parse(Int, ARGS[1]); kappa = parse(Float64, ARGS[2]); sigma_s = parse(Float64, ARGS[3])
rays   = parse(Int, ARGS[4]); steps = parse(Int, ARGS[5])

verts = SVector(Point2(0.0, 0.0), Point2(1.0, 0.0), Point2(1.0, 1.0), Point2(0.0, 1.0))
face = PolyVolume2D{Float64}(verts, SVector(true, true, true, true), 1, kappa, sigma_s)
face.T_in_w = [1000.0, 0.0, 0.0, 0.0]; face.epsilon = [1.0, 1.0, 1.0, 1.0]; face.T_in_g = -1.0; face.q_in_g = 0.0
mesh = RayTracingDomain2D([face], [(Ndim, Ndim)]; verbose = false)
nudge = 10_000 * eps(Float64)

# time the traced part only (parallelRayTracing), not the smoothing that follows it in mesh(...)
RayTraceHeatTransfer.parallelRayTracing(mesh, max(rays ÷ 100, 1000), nudge, false)     # compile + warm-up
t = @elapsed for _ in 1:steps
    RayTraceHeatTransfer.parallelRayTracing(mesh, rays, nudge, false)
end
n_el = length(mesh.surface_mapping) + length(mesh.volume_mapping)
traced = div(rays, n_el) * n_el
println("RTHX_JULIA_REF rays_per_s=$(traced * steps / t) threads=$(Threads.nthreads()) rays=$traced")
