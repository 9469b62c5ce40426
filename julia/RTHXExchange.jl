# RTHXExchange.jl — the reference-side binding of librthx.so (include/rthx.h).
#
# `include("RTHXExchange.jl")` after `using RayTraceHeatTransfer` replaces the emitter loop of
#     RayTraceHeatTransfer.parallelRayTracing(rtm, rays_total, nudge, verbose; rec)
#         (src/RayTracing/RayTracing2D/ExchangeFactors2D/parallelRayTracing.jl:1-62)
# by one `ccall` into the CUDA library; `exchangeRayTracing!`, `smooth_F` and `solveEquilibrium!` stay untouched, so
# `mesh(N_rays; method=:exchange, rec)` keeps its signature, its F_raw / F_smooth outputs and its RayRecorder
# semantics.  Julia is not available in the build image of this repository: this file is exercised only through its
# 1:1 Python twin (raytraceheattransfer.jl_b200/{flatten,tracing,_lib}.py), which drives the same C ABI.
module RTHXExchange

using RayTraceHeatTransfer
using SparseArrays
using GeometryBasics: Point2
import RayTraceHeatTransfer: RayTracingDomain2D, group_uniform_bins

const LIB = get(ENV, "RTHX_LIB", "librthx.so")
const SEED = Ref{UInt64}(rand(UInt64))      # set RTHXExchange.SEED[] for reproducible runs
const DEVICES = Ref{Vector{Cint}}(Cint[0])  # devices of the trace (rthx_create_multi / rthx_trace_exchange_multi); e.g. Cint.(0:7)

# ---- struct twins of include/rthx.h (field order and types must match) -------------------------------------------
struct RthxMesh
    n_coarse::Int32; n_cells::Int32; n_bands::Int32; n_surfaces::Int32
    coarse_nv::Ptr{Int32}; coarse_vx::Ptr{Float64}; coarse_vy::Ptr{Float64}; coarse_solid::Ptr{UInt8}
    fine_off::Ptr{Int32}
    cell_nv::Ptr{Int32}; cell_vx::Ptr{Float64}; cell_vy::Ptr{Float64}; cell_mid::Ptr{Float64}
    cell_volume::Ptr{Float64}; cell_surf_id::Ptr{Int32}
    kappa::Ptr{Float64}; sigma_s::Ptr{Float64}; epsilon::Ptr{Float64}; uniform_beta::Ptr{Float64}
end

struct RthxTraceArgs
    rays_per_emitter::Int64; ray_id_offset::Int64; seed::UInt64; nudge::Float64
    n_bins::Int32; bins::Ptr{Int32}
    mode::Int32; locator::Int32
    emitter_rank::Int32; emitter_world::Int32
    n_rec_ids::Int32; rec_ids::Ptr{Int32}; rec_bin::Int32
    block_threads::Int32; row_chunks::Int32
end

mutable struct RthxRecOut
    capacity::Int64; origins::Ptr{Float64}; endpoints::Ptr{Float64}; n_recorded::Int64
end

mutable struct RthxStats
    rays_traced::Int64; rays_lost::Int64; kernel_ms::Float64; total_ms::Float64
    n_blocks::Int32; block_threads::Int32; row_chunks::Int32; smem_bytes::Int32; hist_in_smem::Int32; n_launches::Int32
    RthxStats() = new(0, 0, 0.0, 0.0, 0, 0, 0, 0, 0, 0)
end

bandvalue(x::AbstractVector, b) = x[b]
bandvalue(x::Number, b) = x

# ---- flatten RayTracingDomain2D (DomainStructs.jl:89-130) into the SoA of rthx_mesh; indices become 0-based -------
function flatten(rtm::RayTracingDomain2D)
    nc = length(rtm.coarse_mesh); nb = rtm.n_spectral_bins
    ncell = length(rtm.volume_mapping); ns = length(rtm.surface_mapping)
    coarse_nv = zeros(Int32, nc); coarse_vx = zeros(4, nc); coarse_vy = zeros(4, nc); coarse_solid = zeros(UInt8, 4, nc)
    fine_off = zeros(Int32, nc + 1)
    cell_nv = zeros(Int32, ncell); cell_vx = zeros(4, ncell); cell_vy = zeros(4, ncell); cell_mid = zeros(2, ncell)
    cell_volume = zeros(ncell); cell_surf_id = fill(Int32(-1), 4, ncell)
    kappa = zeros(ncell, nb); sigma_s = zeros(ncell, nb); epsilon = zeros(max(ns, 1), nb)
    g = 0
    for (c, face) in enumerate(rtm.coarse_mesh)
        coarse_nv[c] = length(face.vertices)
        for (i, v) in enumerate(face.vertices)
            coarse_vx[i, c] = v[1]; coarse_vy[i, c] = v[2]; coarse_solid[i, c] = face.solidWalls[i]
        end
        fine_off[c] = g
        for (f, cell) in enumerate(rtm.fine_mesh[c])
            g += 1
            cell_nv[g] = length(cell.vertices)
            for (i, v) in enumerate(cell.vertices)
                cell_vx[i, g] = v[1]; cell_vy[i, g] = v[2]
            end
            cell_mid[1, g] = cell.midPoint[1]; cell_mid[2, g] = cell.midPoint[2]
            cell_volume[g] = cell.volume
            for w in 1:length(cell.vertices)
                s = get(rtm.surface_mapping, (c, f, w), 0)
                if s > 0
                    cell_surf_id[w, g] = s - 1
                    for b in 1:nb
                        epsilon[s, b] = bandvalue(cell.epsilon[w], b)
                    end
                end
            end
            for b in 1:nb
                kappa[g, b] = bandvalue(cell.kappa_g, b); sigma_s[g, b] = bandvalue(cell.sigma_s_g, b)
            end
        end
    end
    fine_off[nc + 1] = g
    ub = Float64.(rtm.uniform_across_bin)
    arrays = (coarse_nv, coarse_vx, coarse_vy, coarse_solid, fine_off, cell_nv, cell_vx, cell_vy, cell_mid,
              cell_volume, cell_surf_id, kappa, sigma_s, epsilon, ub)
    mesh = RthxMesh(nc, ncell, nb, ns, map(pointer, arrays)...)
    return mesh, arrays      # keep `arrays` alive (GC.@preserve) while `mesh` is in use
end

check(rc, h) = rc == 0 || error("rthx: " * unsafe_string(ccall((:rthx_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# ---- page-locked result arrays: the device writes them by DMA, Julia wraps them without a copy -----------------------
# (wrapped foreign memory cannot be resized: code that inserts NEW stored entries into the returned F_raw — `F[i, j] = x` on a
#  structural zero, `dropzeros!` — must `copy(F)` first; `smooth_F` and `equilibriumGrey2D!` only read it)
function pinned_vector(::Type{T}, n::Integer) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rthx_host_alloc, LIB), Cint, (Ref{Ptr{Cvoid}}, UInt64), p, max(n, 1) * sizeof(T)), C_NULL)
    v = unsafe_wrap(Array, Ptr{T}(p[]), n; own = false)
    finalizer(_ -> ccall((:rthx_host_free, LIB), Cint, (Ptr{Cvoid},), p[]), v)
    return v
end

"""
    trace_bins(rtm, rays_per_emitter, nudge, bins, rec) -> Vector{SparseMatrixCSC{Float64,Int}}

Batched replacement of `computeExchangeFactorsBin` (parallelRayTracing.jl:64-159): all `bins` in one launch per device.
The UInt64 count matrix never reaches the host: with several devices their rows are gathered on the first one over NVLink
inside the trace kernels (`counts_out = C_NULL`), and `rthx_counts_csc` emits the three arrays of the
`SparseMatrixCSC{Float64,Int}` that `sparse(I, J, V)` + `row_normalize!` (:144-169) would have built — 1-based, rows
ascending within each column — straight into page-locked memory.
"""
function trace_bins(rtm::RayTracingDomain2D, rays_per_emitter::Integer, nudge::Float64, bins::Vector{Int}, rec)
    mesh, arrays = flatten(rtm)
    N = Int(mesh.n_surfaces + mesh.n_cells)
    nbins = length(bins)
    lost = Array{UInt64}(undef, N, nbins)
    bins0 = Int32.(bins .- 1)
    rec_ids = rec === nothing ? Int32[] : Int32.(rec.ids .- 1)
    cap = length(rec_ids) * rays_per_emitter
    origins = zeros(2, max(cap, 1)); endpoints = zeros(2, max(cap, 1))
    recout = RthxRecOut(cap, pointer(origins), pointer(endpoints), 0)
    stats = RthxStats()
    devs = DEVICES[]
    handles = fill(Ptr{Cvoid}(C_NULL), length(devs))
    Fs = Vector{SparseMatrixCSC{Float64,Int}}(undef, nbins)
    GC.@preserve arrays lost bins0 rec_ids origins endpoints handles begin
        check(ccall((:rthx_create_multi, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ref{RthxMesh}, Ptr{Cint}, Cint),
                    handles, mesh, devs, length(devs)), C_NULL)
        args = RthxTraceArgs(rays_per_emitter, 0, SEED[], nudge, nbins, pointer(bins0), 0, 0, 0, 1,
                             length(rec_ids), isempty(rec_ids) ? C_NULL : pointer(rec_ids),
                             rec === nothing ? 0 : rec.bin - 1, 0, 0)
        rc = ccall((:rthx_trace_exchange_multi, LIB), Cint,
                   (Ptr{Ptr{Cvoid}}, Cint, Ref{RthxTraceArgs}, Ptr{UInt64}, Ptr{UInt64}, Ref{RthxRecOut}, Ref{RthxStats}),
                   handles, length(handles), args, C_NULL, lost, recout, stats)
        check(rc, handles[1])
        length(handles) > 1 &&                                                   # the read-outs leave through every device's PCIe link
            check(ccall((:rthx_set_copy_helpers, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint),
                        handles[1], pointer(handles, 2), length(handles) - 1), handles[1])
        for b in 1:nbins
            nnz = Ref{Int64}(0)
            check(ccall((:rthx_counts_nnz, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), handles[1], b - 1, nnz), handles[1])
            colptr = Vector{Int}(undef, N + 1)
            rowval = pinned_vector(Int, nnz[]); nzval = pinned_vector(Float64, nnz[])
            check(ccall((:rthx_counts_csc, LIB), Cint,
                        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Ptr{UInt64}, Ptr{Float64}),
                        handles[1], b - 1, 1, 1, colptr, rowval, C_NULL, nzval), handles[1])
            maxloss = Int(maximum(@view lost[:, b]))
            println("Maximum ray tracing ray loss per emitter: $maxloss/$rays_per_emitter")   # unconditional, :163
            Fs[b] = SparseMatrixCSC{Float64,Int}(N, N, colptr, rowval, nzval)                # no sort, no copy
        end
        foreach(h -> ccall((:rthx_destroy, LIB), Cint, (Ptr{Cvoid},), h), handles)
    end
    SEED[] += 0x9E3779B97F4A7C15                          # a fresh stream for the next call, like unseeded rand()
    if rec !== nothing
        for k in 1:recout.n_recorded                      # parallelRayTracing.jl:120-123; any slot works for collect_rays
            push!(rec.origins[1], Point2{Float64}(origins[1, k], origins[2, k]))
            push!(rec.endpoints[1], Point2{Float64}(endpoints[1, k], endpoints[2, k]))
        end
    end
    return Fs
end

# ---- method replacement: same signature and return value as parallelRayTracing.jl:1-62 ---------------------------
function RayTraceHeatTransfer.parallelRayTracing(rtm::RayTracingDomain2D, rays_total::P, nudge::G, verbose::Bool;
                                                 rec=nothing) where {P<:Integer, G}
    num_emitters = length(rtm.surface_mapping) + length(rtm.volume_mapping)
    rays_per_emitter = div(rays_total, num_emitters)
    n_bins = rtm.n_spectral_bins
    if rtm.spectral_mode == :spectral_variable
        verbose && println("Computing $n_bins separate F matrices for variable spectral extinction")
        groups, reps, nonuniform = group_uniform_bins(rtm.uniform_across_bin)
        to_trace = vcat(nonuniform, [first(g) for g in groups])
        Fs = trace_bins(rtm, rays_per_emitter, Float64(nudge), to_trace, rec)
        F_raw_vector = Vector{AbstractMatrix}(undef, n_bins)
        for (k, bin) in enumerate(nonuniform)
            F_raw_vector[bin] = Fs[k]
        end
        for (k, g) in enumerate(groups), j in g
            F_raw_vector[j] = Fs[length(nonuniform) + k]        # grouped bins alias one matrix (:38-41)
        end
        return F_raw_vector, rays_per_emitter
    end
    verbose && println(rtm.spectral_mode != :grey ?
                       "Computing single F matrix for uniform spectral extinction ($n_bins bins)" :
                       "Computing single F matrix for grey extinction")
    return trace_bins(rtm, rays_per_emitter, Float64(nudge), [1], rec)[1], rays_per_emitter
end

# ---- optional: the linear solve of equilibriumGrey2D! on the device (rthx_solve_grey, include/rthx.h) ----------------
struct RthxSolveArgs
    n::Int32; source::Int32; layout::Int32; memory::Int32; max_iters::Int32; measure_pass::Int32
    F_dense::Ptr{Float64}; colptr::Ptr{Int64}; rowval::Ptr{Int32}; nzval::Ptr{Float64}
    coeff::Ptr{Float64}; rhs::Ptr{Float64}; rtol::Float64; atol::Float64
end

mutable struct RthxSolveStats
    iterations::Int32; restarts::Int32; launches::Int32; converged::Int32; matvecs::Int32; pad_::Int32
    residual::Float64; rhs_norm::Float64; total_ms::Float64; matvec_ms::Float64; matvec_gbs::Float64; matvec_bytes::Int64
    RthxSolveStats() = new(0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0)
end

"""
    solve_grey(rtm, F, coeff, h; memory = 50, rtol = 1e-12) -> (j, g, stats)

Solves `(I - Diagonal(coeff) * F') * j = h` on the GPU and returns the incident power `g = F' * j` with it: the two
O(N^2) steps of `equilibriumGrey2D!` (src/HeatTransfer/equilibrium/equilibriumGrey2D.jl:148-158 and :168-194).
A `Matrix{Float64}` is handed over as it lies in memory (column-major), a `SparseMatrixCSC` by its three fields.
In `equilibriumGrey2D!` the maintainer replaces
    M = I - Diagonal(coeff) * permutedims(F); j = F isa SparseMatrixCSC ? gmres(...) : M \ h      (:149-158)
    the two receiver loops                                                                          (:168-194)
by
    j, g, _ = RTHXExchange.solve_grey(mesh, F, coeff, h);  r .= b .* g;  Abs .= (1 .- b) .* g
"""
function solve_grey(rtm::RayTracingDomain2D, F::AbstractMatrix, coeff::Vector{Float64}, h::Vector{Float64};
                    memory::Integer = 50, rtol::Float64 = 1e-12)
    n = length(h)
    size(F) == (n, n) && length(coeff) == n || error("solve_grey: size mismatch")
    mesh, arrays = flatten(rtm)
    j = zeros(n); g = zeros(n); stats = RthxSolveStats()
    hd = Ref{Ptr{Cvoid}}(C_NULL)
    if F isa SparseMatrixCSC
        colptr = Int64.(F.colptr .- 1); rowval = Int32.(F.rowval .- 1); nzval = Vector{Float64}(F.nzval)
        Fd = Float64[]
    else
        colptr = Int64[]; rowval = Int32[]; nzval = Float64[]
        Fd = Matrix{Float64}(F)
    end
    GC.@preserve arrays colptr rowval nzval Fd coeff h j g begin
        check(ccall((:rthx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{RthxMesh}, Cint), hd, mesh, DEVICES[][1]), C_NULL)
        args = F isa SparseMatrixCSC ?
            RthxSolveArgs(n, 2, 0, memory, 0, 0, C_NULL, pointer(colptr), pointer(rowval), pointer(nzval),
                          pointer(coeff), pointer(h), rtol, -1.0) :
            RthxSolveArgs(n, 1, 1, memory, 0, 0, pointer(Fd), C_NULL, C_NULL, C_NULL,        # RTHX_COL_MAJOR
                          pointer(coeff), pointer(h), rtol, -1.0)
        rc = ccall((:rthx_solve_grey, LIB), Cint, (Ptr{Cvoid}, Ref{RthxSolveArgs}, Ptr{Float64}, Ptr{Float64}, Ref{RthxSolveStats}),
                   hd[], args, j, g, stats)
        check(rc, hd[])
        ccall((:rthx_destroy, LIB), Cint, (Ptr{Cvoid},), hd[])
    end
    stats.converged == 1 || @warn "rthx_solve_grey: residual $(stats.residual) after $(stats.iterations) iterations"
    return j, g, stats
end

end # module
