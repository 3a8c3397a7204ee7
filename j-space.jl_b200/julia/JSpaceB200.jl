### JSpaceB200.jl — thin Julia wrapper over libjspace_b200.so (include/jspace_b200.h).
###
### Drop-in for the spatial clonal-dynamics path of J_Space.jl: same exported names, positional and
### keyword arguments and 7-tuple return as
###     spatial_graph        (src/J_Space.jl:43-71)
###     simulate_evolution   (src/J_Space.jl:638-842 and :848-1104)
### and the same OBJECTS: the graphs are `MetaGraphs.MetaGraph`s carrying the reference's vertex
### properties (`:name`, `:id`, `:mutation`, `:Subpop`, `:Fit`), `df` is the reference's DataFrame,
### `Gs_plot` holds one MetaGraph per `Time_of_sampling` entry, and `seed` is the caller's
### `MersenneTwister` — so `Start_J_Space` (src/Start.jl:12-13,34,40,71,88), `plot_lattice`
### (src/Start.jl:119-147) and `sampling_phylogentic_relation`
### (src/sampling_phylogentic_relation_and_genotype.jl:15-136) run unchanged on what comes back.
### Host code stays Julia; all simulation work is done by `ccall`s — no CUDA.jl kernels, no CPU
### fallback (a missing GPU raises an error).
###
### NOT EXERCISED IN THIS REPOSITORY'S CI: the build image has no Julia.  The C call sequence of this
### file (handles, argument types, order of frees) is what tests/c_abi_driver.c performs from C;
### the Python package j-space.jl_b200/jspace_b200 binds the same symbols and
### is what the parity tests drive.  Struct layouts are checked against the header by
### tests/test_abi_cpu.py for the ctypes twin; keep the two in sync.
###
### What cannot be hidden (DESIGN.md §6): random numbers come from the counter-based draw contract
### keyed by the MersenneTwister's seed (the caller's generator is not advanced), and cell ids are
### `UUID(n)` with n the sequential id in `uuid1` call order instead of time-based UUIDs.

module JSpaceB200

using DataFrames, DelimitedFiles, UUIDs, Random
using Graphs, MetaGraphs

export spatial_graph, simulate_evolution

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
    n_devices::Int32 = 0
    pad::Int32 = 0
    devices::Ptr{Int32} = C_NULL
end

const OUT_EVENTS, OUT_LATTICE = 0x1, 0x2
const EVENT_NAME = ("Duplicate", "Mutation", "Death", "Migration")

last_error() = unsafe_string(ccall((:jsb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 || error("libjspace_b200: $(last_error()) (code $rc)")

# ---- the native graph handle: owned ONCE --------------------------------------------------------
# Every MetaGraph this module returns carries the same `GraphHandle` object in its graph-level
# property `:jsb_handle`; the one finalizer runs when the last of them is collected.
mutable struct GraphHandle
    ptr::Ptr{Cvoid}
    n_cell::Int
    function GraphHandle(ptr::Ptr{Cvoid}, n_cell::Int)
        h = new(ptr, n_cell)
        finalizer(x -> (x.ptr == C_NULL || ccall((:jsb_graph_free, LIB), Cvoid, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
        return h
    end
end

# ---- the random stream: MersenneTwister -> Philox key + stream counter --------------------------
# `seed::MersenneTwister` is what src/Start.jl:12-13 constructs and passes to every stage.  The key is
# the generator's seed (rng.seed, the UInt32 words of Config.toml's `seed`); every call on the same
# generator object uses the next stream, as successive calls on one MersenneTwister see fresh numbers.
const STREAMS = WeakKeyDict{Any,Int64}()
philox_key(rng::MersenneTwister) = reduce((a, w) -> (a << 32) | UInt64(w), reverse(rng.seed); init = UInt64(0))
philox_key(seed::Integer) = UInt64(seed)
function next_stream!(rng)
    s = get(STREAMS, rng, Int64(0))
    STREAMS[rng] = s + 1
    return s
end
next_stream!(::Integer) = Int64(0)

cell_uuid(id::Integer) = UUID(UInt128(id))

# ---- spatial_graph ---------------------------------------------------------------------------------
function attach!(G::MetaGraph, h::Ptr{Cvoid}, n_cell::Int)
    set_prop!(G, :jsb_handle, GraphHandle(h, n_cell))
    for i in 1:nv(G)                                  # src/J_Space.jl:137-146: every vertex is named
        set_props!(G, i, Dict(:name => "node$i"))
    end
    return G
end

"spatial_graph(row, col, seed; dim, n_cell) — src/J_Space.jl:43-64.  Seeding happens inside the run (the i-th
replicate seeds itself); the returned MetaGraph has the reference's topology and `:name` properties."
function spatial_graph(row::Int, col::Int, seed; dim::Int = 2, n_cell::Int = 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jsb_graph_lattice, LIB), Cint, (Cint, Int64, Int64, Ref{Ptr{Cvoid}}), dim, row, col, h))
    if dim == 2
        G = Graphs.grid((row, col, 1))                # :54
    else
        d = convert(Int, trunc((row * col)^(1 / 3))) + 1   # :50 / :58
        G = Graphs.grid((d, d, d))
    end
    return attach!(MetaGraph(G), h[], n_cell)
end

"spatial_graph(path, seed; n_cell) — src/J_Space.jl:67-71 (the evident intent: SimpleGraph(adjacency))."
function spatial_graph(path::String, seed; n_cell::Int = 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jsb_graph_dense_file, LIB), Cint, (Cstring, Ref{Ptr{Cvoid}}), path, h))
    return attach!(MetaGraph(SimpleGraph(readdlm(path, Int))), h[], n_cell)
end

# A MetaGraph that did not come from this module (e.g. built by the reference's own spatial_graph): its
# topology goes through the CSR entry point, its seeded vertices (those with an :id) become the start cells.
function handle_of(G::MetaGraph)
    has_prop(G, :jsb_handle) && return get_prop(G, :jsb_handle)::GraphHandle
    n = nv(G)
    rowptr = Int64[0]; colidx = Int32[]
    for v in 1:n
        append!(colidx, Int32.(neighbors(G.graph, v)))   # SimpleGraph adjacency lists are ascending
        push!(rowptr, length(colidx))
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jsb_graph_csr, LIB), Cint, (Int64, Ptr{Int64}, Ptr{Int32}, Ref{Ptr{Cvoid}}), n, rowptr, colidx, h))
    seeded = Int64[v for v in 1:n if has_prop(G, v, :id)]
    length(seeded) == 1 && check(ccall((:jsb_graph_set_seed_candidates, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), h[], seeded, 1))
    gh = GraphHandle(h[], max(length(seeded), 0))
    set_prop!(G, :jsb_handle, gh)
    return gh
end

rule_code(model::String) = model == "contact" ? 0 : model == "voter" ? 1 :
                           (model == "h_voter" || model == "hvoter") ? 2 : 3   # :767-786

function run_one(gh::GraphHandle, p::JsbParams, key::UInt64, stream::Int64, keep; flags = OUT_EVENTS | OUT_LATTICE)
    p.size = sizeof(JsbParams)
    opts = JsbRunOpts(size = sizeof(JsbRunOpts), flags = flags, seed = key, g_begin = stream, g_end = stream + 1)
    res = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep gh begin
        check(ccall((:jsb_run, LIB), Cint,
                    (Ptr{Cvoid}, Ref{JsbParams}, Int64, Int64, Ref{JsbRunOpts}, Ref{Ptr{Cvoid}}),
                    gh.ptr, p, 1, stream + 1, opts, res))
    end
    return res[]
end

# initialize_graph's lattice (src/J_Space.jl:79-163): the same replicate with Tf = 0, whose loop is never entered
function seeded_lattice(gh::GraphHandle, p::JsbParams, key::UInt64, stream::Int64, keep, n::Int)
    Tf = p.Tf
    p.Tf = 0.0
    res = run_one(gh, p, key, stream, keep; flags = OUT_LATTICE)
    p.Tf = Tf
    try
        pl = Ref{Ptr{UInt16}}(C_NULL); pi = Ref{Ptr{UInt32}}(C_NULL)
        check(ccall((:jsb_result_lattice, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{UInt16}}, Ref{Ptr{UInt32}}), res, 0, pl, pi))
        return copy(unsafe_wrap(Array, pl[], n)), copy(unsafe_wrap(Array, pi[], n))
    finally
        ccall((:jsb_result_free, LIB), Cvoid, (Ptr{Cvoid},), res)
    end
end

# vertex properties of one lattice state, as the reference keeps them (:355-362, :675-680)
function fill_props!(G::MetaGraph, clone, ids, set_mut, α)
    for v in 1:nv(G)
        c = Int(clone[v])
        if c == 0
            for k in (:id, :mutation, :Subpop, :Fit)
                has_prop(G, v, k) && rem_prop!(G, v, k)
            end
        else
            set_props!(G, v, Dict(:mutation => set_mut[c], :id => cell_uuid(ids[v]), :Subpop => c, :Fit => α[c]))
        end
    end
    return G
end

function materialise(G::MetaGraph, gh::GraphHandle, res::Ptr{Cvoid}, names, seeded)
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
        n = nv(G)
        pl = Ref{Ptr{UInt16}}(C_NULL); pi = Ref{Ptr{UInt32}}(C_NULL)
        check(ccall((:jsb_result_lattice, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{UInt16}}, Ref{Ptr{UInt32}}), res, 0, pl, pi))
        clone = copy(unsafe_wrap(Array, pl[], n)); ids = copy(unsafe_wrap(Array, pi[], n))
        psn = Ref{Ptr{JsbSnapshot}}(C_NULL); nsn = Ref{Int32}(0)
        check(ccall((:jsb_result_snapshots, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{JsbSnapshot}}, Ref{Int32}), res, 0, psn, nsn))
        snaps = nsn[] > 0 ? copy(unsafe_wrap(Array, psn[], nsn[])) : JsbSnapshot[]

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

        # df (src/J_Space.jl:284-290), rows in the reference's push! order
        df = DataFrame(Event = String[], Time = Float64[], Cell = UUID[], Notes = Any[])
        for e in ev
            notes = e.type == 0 ? [cell_uuid(e.parent)] :
                    e.type == 1 ? Any[cell_uuid(e.parent), names === nothing ? Int(e.label) : set_mut[e.clone]] : undef
            push!(df, [EVENT_NAME[e.type + 1], e.time, cell_uuid(e.cell), notes])
        end

        # Replay of the log from the seeded lattice: list_len_node_occ / CA_Alive_TOT (one entry per group of
        # rows flagged JSB_EVF_GROUP_START, :740-741, :758-759, :807-808, :835-836), and the lattice at every
        # snapshot (`copy(G)` when t_curr > Time_of_sampling[k], :708-714: the state after the first
        # n_events_before rows; CA_Alive_TOT gets one extra row there, :712).
        st_clone, st_id = seeded                       # initialize_graph's lattice (run_one with Tf = 0)
        nalive = count(!iszero, st_clone)
        counts = nalive > 0 ? [nalive] : Int[]
        series = Any[nalive]; ca_tot = Any[]; Gs_plot = Any[]
        snap_k = 1
        take_snapshot!() = begin
            Gk = copy(G)
            fill_props!(Gk, st_clone, st_id, set_mut, α)
            push!(Gs_plot, Gk); push!(ca_tot, copy(counts))
        end
        i = 1
        while i <= length(ev)
            while snap_k <= length(snaps) && snaps[snap_k].n_events_before <= i - 1
                take_snapshot!(); snap_k += 1
            end
            j = i
            while j < length(ev) && (ev[j + 1].flags & 0x1) == 0; j += 1; end
            for k in i:j
                e = ev[k]
                if e.type == 2                         # Death (also displaced / excised cells)
                    counts[e.clone] -= 1; nalive -= 1
                    st_clone[e.site] = 0; st_id[e.site] = 0
                elseif e.type == 3                     # Migration: site2 -> site
                    st_clone[e.site2] = 0; st_id[e.site2] = 0
                    st_clone[e.site] = e.clone; st_id[e.site] = e.cell
                else                                   # a daughter takes the vertex the row names
                    st_clone[e.site] = e.clone; st_id[e.site] = e.cell
                end
            end
            if j - i >= 1 && ev[j].type == 0 && ev[j - 1].type <= 1   # a birth pair closes the group
                ev[j - 1].type == 1 ? push!(counts, 1) : (counts[ev[j - 1].clone] += 1)
                nalive += 1
            end
            push!(series, nalive); push!(ca_tot, copy(counts))
            i = j + 1
        end
        while snap_k <= length(snaps)
            take_snapshot!(); snap_k += 1
        end

        Gout = copy(G)                                 # shares the GraphHandle object (graph-level property)
        fill_props!(Gout, clone, ids, set_mut, α)
        return df, Gout, series, set_mut, Gs_plot, ca_tot, α
    finally
        ccall((:jsb_result_free, LIB), Cvoid, (Ptr{Cvoid},), res)
    end
end

"simulate_evolution, random drivers — src/J_Space.jl:638-650"
function simulate_evolution(G::AbstractGraph, Tf::AbstractFloat, rate_birth::AbstractFloat, rate_death::AbstractFloat,
                            rate_migration::AbstractFloat, μ_dri::AbstractFloat, driv_average_advantage::AbstractFloat,
                            driv_std_advantage::AbstractFloat, model::String, seed::Union{MersenneTwister,Integer};
                            Time_of_sampling = [], t_bottleneck = [], ratio_bottleneck = [])
    gh = handle_of(G)
    tb = Float64.(t_bottleneck); rb = Float64.(ratio_bottleneck); ts = Float64.(Time_of_sampling)
    p = JsbParams(method = 1, Tf = Tf, rate_birth = rate_birth, rate_death = rate_death, rate_migration = rate_migration,
                  mu_driver = μ_dri, adv_mean = driv_average_advantage, adv_std = driv_std_advantage,
                  rule = rule_code(model), n_start_cells = gh.n_cell, n_bottleneck = length(tb), n_snapshots = length(ts),
                  t_bottleneck = pointer(tb), ratio_bottleneck = pointer(rb), t_snapshot = pointer(ts))
    key = philox_key(seed); stream = next_stream!(seed); keep = (tb, rb, ts)
    return materialise(G, gh, run_one(gh, p, key, stream, keep), nothing, seeded_lattice(gh, p, key, stream, keep, nv(G)))
end

"simulate_evolution, driver tree — src/J_Space.jl:848-859"
function simulate_evolution(G::AbstractGraph, Tf::AbstractFloat, rate_death::AbstractFloat, rate_migration::AbstractFloat,
                            μ_dri::AbstractFloat, model::String, edge_list_path::String, driv_adv_path::String,
                            seed::Union{MersenneTwister,Integer}; Time_of_sampling = [], t_bottleneck = [], ratio_bottleneck = [])
    gh = handle_of(G)
    edge_list = readdlm(edge_list_path, String)          # :871
    driv_adv = readdlm(driv_adv_path, String)            # :872
    names = unique(driv_adv[:, 1])
    idx = Dict(n => Int32(i - 1) for (i, n) in enumerate(names))
    ep = Int32[idx[x] for x in edge_list[:, 1]]; ec = Int32[idx[x] for x in edge_list[:, 2]]
    rate = [parse(Float64, driv_adv[findfirst(==(n), driv_adv[:, 1]), 2]) for n in names]   # :503-504
    root = setdiff(edge_list[:, 1], edge_list[:, 2])[1]  # :873
    tb = Float64.(t_bottleneck); rb = Float64.(ratio_bottleneck); ts = Float64.(Time_of_sampling)
    p = JsbParams(method = 2, Tf = Tf, rate_death = rate_death, rate_migration = rate_migration, mu_driver = μ_dri,
                  rule = rule_code(model), n_start_cells = gh.n_cell, n_bottleneck = length(tb), n_snapshots = length(ts),
                  t_bottleneck = pointer(tb), ratio_bottleneck = pointer(rb), t_snapshot = pointer(ts),
                  n_tree_nodes = length(names), n_tree_edges = length(ep), edge_parent = pointer(ep),
                  edge_child = pointer(ec), node_rate = pointer(rate), root_node = idx[root],
                  root_rate = parse(Float64, driv_adv[1, 2]))   # :881
    key = philox_key(seed); stream = next_stream!(seed); keep = (tb, rb, ts, ep, ec, rate)
    return materialise(G, gh, run_one(gh, p, key, stream, keep), names, seeded_lattice(gh, p, key, stream, keep, nv(G)))
end

end # module
