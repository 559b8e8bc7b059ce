# BarBayB200.jl -- Julia glue for the B200 ADVI backend (ccall into libbarbay_b200.so).
#
# NOT EXERCISED IN THIS REPOSITORY'S CI: the build image has no Julia.  The file shows the
# binding a BarBay.jl maintainer would add; the C ABI it targets (include/barbay_b200.h) is
# exercised end to end by the Python/ctypes mirror (barbay.jl_b200/vi.py) and tests/.
#
# Drop-in point: src/vi.jl:201   q = Turing.vi(bayes_model, advi; optimizer=opt)
# becomes                         q = BarBayB200.vi(data_arrays, model, model_kwargs, advi, opt)
# Everything before (argument checks, data_to_arrays, vi.jl:102-169) and after
# (advi_to_df, CSV.write, vi.jl:203-234) stays byte for byte.
module BarBayB200

import Distributions
import Bijectors
import DistributionsAD
import Turing

const LIB = get(ENV, "BARBAY_B200_LIB", "libbarbay_b200.so")
const BB_ABI_VERSION = Int32(2)

struct BBPrior
    data::Ptr{Cdouble}
    n::Int64
    is_matrix::Int32
end

struct BBOpt
    kind::Int32
    eta::Cdouble
    tau::Cdouble
    post::Cdouble
    n::Int32
end

# field order and types mirror `struct bb_desc` in include/barbay_b200.h
struct BBDesc
    abi_version::Int32
    model::Int32
    dtype::Int32
    n_rep::Int32
    n_time::Ptr{Int32}
    n_neutral::Int32
    n_bc::Int32
    bc_count::Ptr{Int64}
    n_env::Int32
    env_idx::Ptr{Int32}
    n_geno::Int32
    geno_idx::Ptr{Int32}
    s_pop_prior::BBPrior
    logsig_pop_prior::BBPrior
    s_bc_prior::BBPrior
    logsig_bc_prior::BBPrior
    loglam_prior::BBPrior
    logtau_prior::BBPrior
    ragged_as_written::Int32
    n_samples::Int32
    seed::UInt64
    device::Int32
    rank::Int32
    world::Int32
    n_devices::Int32     # ABI 2: N > 1 = one handle drives N GPUs of this process (single blocking advi call)
    env_per_rep::Int32   # 1: env_idx holds one environment list per replicate
end

"""
    naive_prior(da; device=-1) -> Dict

`BarBay.stats.naive_prior` (src/stats.jl:1175-1359) on the GPU from the packed counts `da = data_to_arrays(data; ...)`
of the frame WITH the pseudocount added (the reference adds it at stats.jl:1185 before packing, :1189): the
replacement for the arithmetic of stats.jl:1199-1352, same keys and vector orders.
"""
function naive_prior(da; device::Integer=-1)
    counts = da.bc_count isa Vector ? reduce(vcat, vec.(da.bc_count)) : vec(da.bc_count)
    n_time = Int32.(da.bc_count isa Vector ? size.(da.bc_count, 1) :
                    fill(size(da.bc_count, 1), ndims(da.bc_count) == 3 ? size(da.bc_count, 3) : 1))
    n_pop = sum(n_time) - length(n_time)
    s_pop, logσ_pop = Vector{Float64}(undef, n_pop), Vector{Float64}(undef, n_pop)
    logλ = Vector{Float64}(undef, length(counts))
    rc = ccall((:bb_naive_prior, LIB), Cint,
               (Ptr{Int64}, Int32, Ptr{Int32}, Int32, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
               counts, length(n_time), n_time, da.n_neutral, da.n_bc, device, s_pop, logσ_pop, logλ)
    rc == 0 || error(unsafe_string(ccall((:bb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    return Dict(:s_pop_prior => s_pop, :logσ_pop_prior => logσ_pop, :logλ_prior => logλ)
end

# model function name -> bb_model (the reference dispatches on the name too, src/vi.jl:111-169)
function model_id(model::Function)
    name = "$(model)"
    occursin("multienv_replicate", name) && return Int32(4)
    occursin("replicate", name) && return Int32(1)
    occursin("multienv", name) && return Int32(2)
    occursin("genotype", name) && return Int32(3)
    return Int32(0)
end

check(h, rc) = rc == 0 || error(unsafe_string(ccall((:bb_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# `[mean, std]` or an n x 2 Matrix{Float64} (column-major: means then stds), model_fitness_normal.jl:125-129
function prior(p::VecOrMat{Float64}, keep::Vector{Any})
    buf = p isa Vector ? copy(p) : vec(copy(p))
    push!(keep, buf)
    return BBPrior(pointer(buf), p isa Vector ? 2 : size(p, 1), p isa Vector ? 0 : 1)
end

"""
    vi(data_arrays, model, model_kwargs, advi, opt; seed, dtype) -> TransformedDistribution

Replacement for `Turing.vi(bayes_model, advi; optimizer=opt)` (src/vi.jl:201).  Returns a real
`Bijectors.transformed(TuringDiagMvNormal(m, σ), Stacked(identity...))` so that the unchanged
`utils.advi_to_df` (which reads `dist.dist.m`, `dist.dist.σ`, `dist.transform.ranges_out`,
src/utils.jl:1049-1060, and is typed `::Distributions.Sampleable`, :1411) accepts it.
"""
function vi(da, model::Function, model_kwargs::Dict, advi, opt; seed::Integer=0, dtype::Symbol=:f64,
            device::Integer=-1, n_devices::Integer=1, corrected_ragged::Bool=false)
    keep = Any[]
    kw = Dict{Symbol,Any}(model_kwargs)
    counts = da.bc_count isa Vector ? reduce(vcat, vec.(da.bc_count)) : vec(da.bc_count)
    n_time = Int32.(da.bc_count isa Vector ? size.(da.bc_count, 1) :
                    fill(size(da.bc_count, 1), ndims(da.bc_count) == 3 ? size(da.bc_count, 3) : 1))
    # one environment list shared by all replicates, or one list per replicate (the Vector{Matrix} method of the
    # multienv x replicate model, …replicates.jl:465-472): indexin.(envs, Ref(unique(vcat(envs...)))) back to back
    env_per_rep = haskey(kw, :envs) && eltype(kw[:envs]) <: AbstractVector
    env_flat = haskey(kw, :envs) ? (env_per_rep ? reduce(vcat, kw[:envs]) : kw[:envs]) : Any[]
    env_idx = isempty(env_flat) ? Int32[] : Int32.(indexin(env_flat, unique(env_flat)))
    geno_idx = haskey(kw, :genotypes) ? Int32.(indexin(kw[:genotypes], unique(kw[:genotypes]))) : Int32[]
    push!(keep, counts, n_time, env_idx, geno_idx)
    getp(k, d) = prior(get(kw, k, d), keep)
    desc = BBDesc(
        BB_ABI_VERSION, model_id(model), dtype == :f32 ? 0 : 1, length(n_time), pointer(n_time),
        da.n_neutral, da.n_bc, pointer(counts),
        isempty(env_idx) ? 1 : length(unique(env_idx)), isempty(env_idx) ? C_NULL : pointer(env_idx),
        isempty(geno_idx) ? 0 : length(unique(geno_idx)), isempty(geno_idx) ? C_NULL : pointer(geno_idx),
        getp(:s_pop_prior, [0.0, 2.0]), getp(:logσ_pop_prior, [0.0, 1.0]), getp(:s_bc_prior, [0.0, 2.0]),
        getp(:logσ_bc_prior, [0.0, 1.0]), getp(:logλ_prior, [3.0, 3.0]), getp(:logτ_prior, [-2.0, 1.0]),
        Int32(corrected_ragged ? 0 : 1), advi.samples_per_step, UInt64(seed), Int32(device), 0, 1, Int32(n_devices),
        Int32(env_per_rep))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin
        rc = ccall((:bb_create, LIB), Cint, (Ref{BBDesc}, Ref{Ptr{Cvoid}}), desc, h)
        rc == 0 || error(unsafe_string(ccall((:bb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    try
        check(h[], ccall((:bb_init_params, LIB), Cint, (Ptr{Cvoid}, UInt64), h[], UInt64(seed)))
        bo = opt isa Turing.Variational.DecayedADAGrad ?
             BBOpt(1, opt.eta, opt.pre, opt.post, 0) : BBOpt(0, opt.eta, opt.tau, 0.0, opt.n)
        check(h[], ccall((:bb_set_optimizer, LIB), Cint, (Ptr{Cvoid}, Ref{BBOpt}), h[], bo))
        check(h[], ccall((:bb_step, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}), h[], advi.max_iters, C_NULL))
        D = ccall((:bb_n_latent, LIB), Int64, (Ptr{Cvoid},), h[])
        m, σ = Vector{Float64}(undef, D), Vector{Float64}(undef, D)
        check(h[], ccall((:bb_get_posterior, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h[], m, σ))
        ranges = var_ranges(da, model, kw)
        base = DistributionsAD.TuringDiagMvNormal(m, σ)
        return Bijectors.transformed(base, Bijectors.Stacked(fill(identity, length(ranges)), ranges))
    finally
        ccall((:bb_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

# one contiguous 1-based range per variable group, VarInfo order (SURVEY.md §8a rows M1-M5)
function var_ranges(da, model, kw)
    name = "$(model)"
    nts = da.bc_count isa Vector ? size.(da.bc_count, 1) :
          fill(size(da.bc_count, 1), ndims(da.bc_count) == 3 ? size(da.bc_count, 3) : 1)
    R, B, M = length(nts), da.n_neutral + da.n_bc, da.n_bc
    E = haskey(kw, :envs) ? length(unique(kw[:envs])) : 1
    nst, nlam = sum(nts .- 1), sum(nts .* B)
    hier = occursin("replicate", name) || occursin("genotype", name)
    lens = hier ?
        [nst, nst, occursin("genotype", name) ? length(unique(kw[:genotypes])) : E * M, E * M * R, E * M * R, E * M * R, nlam] :
        [nst, nst, E * M, E * M, nlam]
    stops = cumsum(lens)
    return [(s - l + 1):s for (s, l) in zip(stops, lens)]
end

# "<group>[i]" strings in VarInfo order, exactly what src/vi.jl:184-198 derives from the Turing VarInfo
function var_names(da, model, kw)
    name = "$(model)"
    hier = occursin("replicate", name) || occursin("genotype", name)
    groups = hier ? ["s̲ₜ", "logσ̲ₜ", "θ̲⁽ᵐ⁾", "θ̲̃⁽ᵐ⁾", "logτ̲⁽ᵐ⁾", "logσ̲⁽ᵐ⁾", "logΛ̲̲"] :
                    ["s̲ₜ", "logσ̲ₜ", "s̲⁽ᵐ⁾", "logσ̲⁽ᵐ⁾", "logΛ̲̲"]
    names = Any[]
    for (g, r) in zip(groups, var_ranges(da, model, Dict{Symbol,Any}(kw)))
        append!(names, ["$g[$i]" for i in 1:length(r)])
    end
    return names
end

end # module
