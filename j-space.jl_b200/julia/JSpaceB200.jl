### JSpaceB200.jl — thin Julia wrapper over libjspace_b200.so (include/jspace_b200.h).
###
### Drop-in for the spatial clonal-dynamics path of J_Space.jl: same exported names, positional and
### keyword arguments and 7-tuple return as
###     spatial_graph        (src/J_Space.jl:43-71)
###     simulate_evolution   (src/J_Space.jl:638-842 and :848-1104)
### Host code stays Julia; all simulation work is done by `ccall`s — no CUDA.jl kernels, no CPU
### fallback (a missing GPU raises an error).
###
### NOT EXERCISED IN THIS REPOSITORY'S CI: the build image has no Julia.  The Python package
### j-space.jl_b200/jspace_b200 binds the same symbols and is what the tests drive; this file is the
### binding a J_Space.jl maintainer would add (INTEGRATION.md).  Struct layouts are checked against
### the header by tests/test_abi_cpu.py for the ctypes twin; keep the two in sync.

module JSpaceB200

using DataFrames, DelimitedFiles, UUIDs

export spatial_graph, simulate_evolution, PhiloxSeed, B200Graph

const LIB = get(ENV, "JSPACE_B200_LIB", joinpath(@__DIR__, "..", "lib", "libjspace_b200.so"))

# ---- records (must equal include/jspace_b200.h) -------------------------------------------------
struct JsbEvent
    time::Float64
    cell::UInt32
    parent::UInt32
    site::UInt32
    site2::UInt32
    clone::UInt16
    label::UInt16
    type::UInt8
    flags::UInt8
    pad::UInt16
end

struct JsbClone
    alpha::Float64
    count::UInt32
    parent::UInt16
    label::UInt16
end

struct JsbSummary
    n_iterations::UInt64
    n_events::UInt64
    t_final::Float64
    n_effective::UInt32
    n_births::UInt32
    n_deaths::UInt32
    n_migrations::UInt32
    n_displaced::UInt32
    n_excised::UInt32
    n_alive::UInt32
    n_clones::UInt32
    next_cell_id::UInt32
    status::Int32
end

struct JsbSnapshot
    time::Float64
    iteration::UInt64
    n_events_before::UInt64
    n_alive::UInt32
    index::UInt32
end

Base.@kwdef mutable struct JsbParams
    size::UInt32 = 0
    method::UInt32 = 1
    Tf::Float64 = 0.0
    rate_birth::Float64 = 0.0
    rate_death::Float64 = 0.0
    rate_migration::Float64 = 0.0
    mu_driver::Float64 = 0.0
    adv_mean::Float64 = 0.0
    adv_std::Float64 = 0.0
    rule::Int32 = 0
    n_start_cells::Int32 = 1
    n_bottleneck::Int32 = 0
    n_snapshots::Int32 = 0
    t_bottleneck::Ptr{Float64} = C_NULL
    ratio_bottleneck::Ptr{Float64} = C_NULL
    t_snapshot::Ptr{Float64} = C_NULL
    n_tree_nodes::Int32 = 0
    n_tree_edges::Int32 = 0
    edge_parent::Ptr{Int32} = C_NULL
    edge_child::Ptr{Int32} = C_NULL
    node_rate::Ptr{Float64} = C_NULL
    root_node::Int32 = 0
    quirk_flags::UInt32 = 0
    root_rate::Float64 = 0.0
    max_iterations::UInt64 = 0
end

Base.@kwdef mutable struct JsbRunOpts
    size::UInt32 = 0
    flags::UInt32 = 0
    seed::UInt64 = 0
    g_begin::Int64 = 0
    g_end::Int64 = 0
    device::Int32 = 0
    warps_per_sm::Int32 = 0
    log_pool_bytes::UInt64 = 0
end

const OUT_EVENTS, OUT_LATTICE, OUT_LISTS = 0x1, 0x2, 0x4
const EVENT_NAME = ("Duplicate", "Mutation", "Death", "Migration")

last_error() = unsafe_string(ccall((:jsb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 || error("libjspace_b200: $(last_error()) (code $rc)")

# ---- seed: stands in for MersenneTwister(seed) --------------------------------------------------
"""Philox key plus stream counter.  Every `simulate_evolution` call on the same object uses the
next stream, as successive calls on one MersenneTwister see fresh numbers."""
mutable struct PhiloxSeed
    seed::UInt64
    stream::Int64
end
PhiloxSeed(seed::Integer) = PhiloxSeed(UInt64(seed), 0)

# ---- graph: stands in for the MetaGraph ----------------------------------------------------------
mutable struct B200Graph
    handle::Ptr{Cvoid}
    n_cell::Int
    nv::Int
    clone::Vector{UInt16}      # :Subpop per vertex (0 = empty)
    cell_id::Vector{UInt32}    # :id per vertex
    set_mut::Vector{Any}
    alpha::Vector{Float64}
    function B200Graph(h, n_cell)
        g = new(h, n_cell, Int(ccall((:jsb_graph_nv, LIB), Int64, (Ptr{Cvoid},), h)), UInt16[], UInt32[], Any[], Float64[])
        finalizer(x -> ccall((:jsb_graph_free, LIB), Cvoid, (Ptr{Cvoid},), x.handle), g)
        return g
    end
end

"spatial_graph(row, col, seed; dim, n_cell) — src/J_Space.jl:43-64"
function spatial_graph(row::Int, col::Int, seed; dim::Int = 2, n_cell::Int = 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jsb_graph_lattice, LIB), Cint, (Cint, Int64, Int64, Ref{Ptr{Cvoid}}), dim, row, col, h))
    return B200Graph(h[], n_cell)
end

"spatial_graph(path, seed; n_cell) — src/J_Space.jl:67-71"
function spatial_graph(path::String, seed; n_cell::Int = 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jsb_graph_dense_file, LIB), Cint, (Cstring, Ref{Ptr{Cvoid}}), path, h))
    return B200Graph(h[], n_cell)
end

rule_code(model::String) = model == "contact" ? 0 : model == "voter" ? 1 :
                           (model == "h_voter" || model == "hvoter") ? 2 : 3   # :767-786

cell_uuid(id::Integer) = UUID(UInt128(id))

function run_one(G::B200Graph, p::JsbParams, seed::PhiloxSeed, keep)
    stream = seed.stream
    seed.stream += 1
    p.size = sizeof(JsbParams)
    opts = JsbRunOpts(size = sizeof(JsbRunOpts), flags = OUT_EVENTS | OUT_LATTICE, seed = seed.seed,
                      g_begin = stream, g_end = stream + 1)
    res = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin
        check(ccall((:jsb_run, LIB), Cint,
                    (Ptr{Cvoid}, Ref{JsbParams}, Int64, Int64, Ref{JsbRunOpts}, Ref{Ptr{Cvoid}}),
                    G.handle, p, 1, stream + 1, opts, res))
    end
    return res[]
end

function materialise(G::B200Graph, res::Ptr{Cvoid}, names)
    try
        ps = Ref{Ptr{JsbSummary}}(C_NULL)
        check(ccall((:jsb_result_summaries, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{JsbSummary}}), res, ps))
        s = unsafe_load(ps[])
        s.status == 2 && error("all 999 driver labels used (the reference throws at src/J_Space.jl:371)")
        pe = Ref{Ptr{JsbEvent}}(C_NULL); ne = Ref{Int64}(0)
        check(ccall((:jsb_result_events, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{JsbEvent}}, Ref{Int64}), res, 0, pe, ne))
        ev = copy(unsafe_wrap(Array, pe[], ne[]))
        pc = Ref{Ptr{JsbClone}}(C_NULL); nc = Ref{Int32}(0)
        check(ccall((:jsb_result_clones, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{JsbClone}}, Ref{Int32}), res, 0, pc, nc))
        cl = copy(unsafe_wrap(Array, pc[], nc[]))
        pl = Ref{Ptr{UInt16}}(C_NULL); pi = Ref{Ptr{UInt32}}(C_NULL)
        check(ccall((:jsb_result_lattice, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{UInt16}}, Ref{Ptr{UInt32}}), res, 0, pl, pi))
        clone = copy(unsafe_wrap(Array, pl[], G.nv)); ids = copy(unsafe_wrap(Array, pi[], G.nv))
        # set_mut_pop and α (src/J_Space.jl:841)
        set_mut = Any[]
        for c in cl
            item = names === nothing ? Int(c.label) : names[c.label]
            if c.parent == 0
                push!(set_mut, item)
            else
                par = set_mut[c.parent]
                push!(set_mut, par isa Vector ? vcat(par, [item]) : Any[par, item])
            end
        end
        α = [c.alpha for c in cl]
        # df (src/J_Space.jl:284-290)
        df = DataFrame(Event = String[], Time = Float64[], Cell = UUID[], Notes = Any[])
        for e in ev
            notes = e.type == 0 ? [cell_uuid(e.parent)] :
                    e.type == 1 ? Any[cell_uuid(e.parent), names === nothing ? Int(e.label) : set_mut[e.clone]] : undef
            push!(df, [EVENT_NAME[e.type + 1], e.time, cell_uuid(e.cell), notes])
        end
        # list_len_node_occ / CA_Alive_TOT: replayed from the group flags exactly as
        # jspace_b200/api.py:replay_series does (one entry per JSB_EVF_GROUP_START group)
        n = count(!iszero, clone) - (Int(s.n_births) - Int(s.n_deaths) - Int(s.n_excised) - Int(s.n_displaced))
        series = Any[n]; ca_tot = Any[]; counts = n > 0 ? [n] : Int[]
        i = 1
        while i <= length(ev)
            j = i
            while j < length(ev) && (ev[j + 1].flags & 0x1) == 0; j += 1; end
            for k in i:j
                if ev[k].type == 2; counts[ev[k].clone] -= 1; n -= 1; end
            end
            if j - i >= 1 && ev[j].type == 0 && ev[j - 1].type <= 1
                ev[j - 1].type == 1 ? push!(counts, 1) : (counts[ev[j - 1].clone] += 1)
                n += 1
            end
            push!(series, n); push!(ca_tot, copy(counts))
            i = j + 1
        end
        out = B200Graph(G.handle, G.n_cell)   # shares the immutable graph handle
        out.clone, out.cell_id, out.set_mut, out.alpha = clone, ids, set_mut, α
        return df, out, series, set_mut, Any[], ca_tot, α
    finally
        ccall((:jsb_result_free, LIB), Cvoid, (Ptr{Cvoid},), res)
    end
end

"simulate_evolution, random drivers — src/J_Space.jl:638-650"
function simulate_evolution(G::B200Graph, Tf::AbstractFloat, rate_birth::AbstractFloat, rate_death::AbstractFloat,
                            rate_migration::AbstractFloat, μ_dri::AbstractFloat, driv_average_advantage::AbstractFloat,
                            driv_std_advantage::AbstractFloat, model::String, seed::PhiloxSeed;
                            Time_of_sampling = [], t_bottleneck = [], ratio_bottleneck = [])
    tb = Float64.(t_bottleneck); rb = Float64.(ratio_bottleneck); ts = Float64.(Time_of_sampling)
    p = JsbParams(method = 1, Tf = Tf, rate_birth = rate_birth, rate_death = rate_death, rate_migration = rate_migration,
                  mu_driver = μ_dri, adv_mean = driv_average_advantage, adv_std = driv_std_advantage,
                  rule = rule_code(model), n_start_cells = G.n_cell, n_bottleneck = length(tb), n_snapshots = length(ts),
                  t_bottleneck = pointer(tb), ratio_bottleneck = pointer(rb), t_snapshot = pointer(ts))
    return materialise(G, run_one(G, p, seed, (tb, rb, ts)), nothing)
end

"simulate_evolution, driver tree — src/J_Space.jl:848-859"
function simulate_evolution(G::B200Graph, Tf::AbstractFloat, rate_death::AbstractFloat, rate_migration::AbstractFloat,
                            μ_dri::AbstractFloat, model::String, edge_list_path::String, driv_adv_path::String,
                            seed::PhiloxSeed; Time_of_sampling = [], t_bottleneck = [], ratio_bottleneck = [])
    edge_list = readdlm(edge_list_path, String)          # :871
    driv_adv = readdlm(driv_adv_path, String)            # :872
    names = unique(driv_adv[:, 1])
    idx = Dict(n => Int32(i - 1) for (i, n) in enumerate(names))
    ep = Int32[idx[x] for x in edge_list[:, 1]]; ec = Int32[idx[x] for x in edge_list[:, 2]]
    rate = [parse(Float64, driv_adv[findfirst(==(n), driv_adv[:, 1]), 2]) for n in names]   # :503-504
    root = setdiff(edge_list[:, 1], edge_list[:, 2])[1]  # :873
    tb = Float64.(t_bottleneck); rb = Float64.(ratio_bottleneck); ts = Float64.(Time_of_sampling)
    p = JsbParams(method = 2, Tf = Tf, rate_death = rate_death, rate_migration = rate_migration, mu_driver = μ_dri,
                  rule = rule_code(model), n_start_cells = G.n_cell, n_bottleneck = length(tb), n_snapshots = length(ts),
                  t_bottleneck = pointer(tb), ratio_bottleneck = pointer(rb), t_snapshot = pointer(ts),
                  n_tree_nodes = length(names), n_tree_edges = length(ep), edge_parent = pointer(ep),
                  edge_child = pointer(ec), node_rate = pointer(rate), root_node = idx[root],
                  root_rate = parse(Float64, driv_adv[1, 2]))   # :881
    return materialise(G, run_one(G, p, seed, (tb, rb, ts, ep, ec, rate)), names)
end

end # module
