### PhiloxShim.jl — TEST HARNESS (not product): an AbstractRNG that serves the draw contract of
### DESIGN.md §3 to the UNMODIFIED reference loop, so that on a box with Julia the real
### `simulate_evolution` (src/J_Space.jl:638-1104) can be replayed event for event against the GPU
### log.  NOT RUN IN THIS REPOSITORY (no Julia in the image); it restates, in Julia, exactly what
### oracle/jspace_oracle.cpp and the CUDA kernel implement.
###
### Harness patches the reference needs (SURVEY.md §8c): relax `seed::MersenneTwister` to
### `seed::AbstractRNG` in the signatures at src/J_Space.jl:45,67,79,126,332,431,562,647,856, and map
### uuid1 values to sequential ids by order of first appearance in the log.
###
### The shim recognises each draw site by the order of calls inside one loop iteration:
###   randexp            -> starts iteration it+1; D1 = -log(((a >> 12) + 0.5)·2^-52), block (it, site 0)
###   rand() #1          -> D2 = (b >> 11)·2^-53 of the same block
###   rand(1:n) #1, #2   -> D3 / D4: Lemire bounded draw on words a / b of block (it, site 1)
###   rand(UInt128) ×2   -> uuid1 draws: ignored (ids are sequential)
###   rand() #2          -> D7 = word a of block (it, site 4);  rand() #3 -> coin = word b
###   rand(::Set{Int})   -> D8: j-th smallest element, j = bounded draw on block (it, site 5)
###   randn              -> D9: polar method on blocks (it, site 6, seq = attempt)
###   rand(i:n) #3...    -> excision draws: block (it, site 7, seq = draw number)

module PhiloxShimModule
using Random
export PhiloxShim

const M0 = 0xD2511F53; const M1 = 0xCD9E8D57; const W0 = 0x9E3779B9; const W1 = 0xBB67AE85

function philox4x32_10(c::NTuple{4,UInt32}, k::NTuple{2,UInt32})
    c0, c1, c2, c3 = c; k0, k1 = k
    for r in 0:9
        if r > 0
            k0 += W0 % UInt32; k1 += W1 % UInt32
        end
        p0 = UInt64(M0) * c0; p1 = UInt64(M1) * c2
        c0, c1, c2, c3 = (UInt32(p1 >> 32) ⊻ c1 ⊻ k0, UInt32(p1 & 0xffffffff), UInt32(p0 >> 32) ⊻ c3 ⊻ k1, UInt32(p0 & 0xffffffff))
    end
    return (c0, c1, c2, c3)
end

mutable struct PhiloxShim <: AbstractRNG
    seed::UInt64
    stream::UInt32
    it::UInt64
    n_float::Int      # rand() calls in this iteration
    n_index::Int      # rand(a:b) calls in this iteration
end
PhiloxShim(seed, stream) = PhiloxShim(UInt64(seed), UInt32(stream), 0, 0, 0)

function block(r::PhiloxShim, it, attempt, site, seq)
    o = philox4x32_10((UInt32(it & 0xffffffff), UInt32((it >> 32) & 0xffff) | (UInt32(attempt) << 16), r.stream,
                       (UInt32(site) << 28) | (UInt32(seq) & 0x0fffffff)),
                      (UInt32(r.seed & 0xffffffff), UInt32(r.seed >> 32)))
    return UInt64(o[1]) | (UInt64(o[2]) << 32), UInt64(o[3]) | (UInt64(o[4]) << 32)
end

u53(w) = Float64(w >> 11) * 2.0^-53

function bounded(r::PhiloxShim, it, site, seq, half, s::UInt64)
    attempt = 0
    w = block(r, it, attempt, site, seq)[half + 1]
    m = UInt128(w) * s; l = UInt64(m & typemax(UInt64))
    if l < s
        t = (-s) % s
        while l < t
            attempt += 1
            w = block(r, it, attempt, site, seq)[half + 1]
            m = UInt128(w) * s; l = UInt64(m & typemax(UInt64))
        end
    end
    return UInt64(m >> 64)
end

# contract_log: same operation order as oracle/jspace_oracle.cpp (no fused multiply-add)
function contract_log(x::Float64)
    ln2_hi = 0x1.62e42fee00000p-1; ln2_lo = 0x1.a39ef35793c76p-33
    Lg = (0x1.5555555555593p-1, 0x1.999999997fa04p-2, 0x1.2492494229359p-2, 0x1.c71c51d8e78afp-3,
          0x1.7466496cb03dep-3, 0x1.39a09d078c69fp-3, 0x1.2f112df3e5244p-3)
    bits = reinterpret(UInt64, x); hx = UInt32(bits >> 32); lx = UInt32(bits & 0xffffffff)
    hx += 0x3ff00000 - 0x3fe6a09e
    k = Int(hx >> 20) - 0x3ff
    hx = (hx & 0x000fffff) + 0x3fe6a09e
    x = reinterpret(Float64, (UInt64(hx) << 32) | lx)
    f = x - 1.0; hfsq = (0.5 * f) * f; s = f / (2.0 + f); z = s * s; w = z * z
    t1 = w * (Lg[2] + w * (Lg[4] + w * Lg[6])); t2 = z * (Lg[1] + w * (Lg[3] + w * (Lg[5] + w * Lg[7])))
    R = t2 + t1; dk = Float64(k)
    return (((s * (hfsq + R) + dk * ln2_lo) - hfsq) + f) + dk * ln2_hi
end

function Random.randexp(r::PhiloxShim)
    r.it += 1; r.n_float = 0; r.n_index = 0
    a, _ = block(r, r.it, 0, 0, 0)
    return -contract_log((Float64(a >> 12) + 0.5) * 2.0^-52)
end
Random.randexp(r::PhiloxShim, ::Type{Float64}) = randexp(r)

function Random.rand(r::PhiloxShim, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}})
    r.n_float += 1
    if r.n_float == 1
        return u53(block(r, r.it, 0, 0, 0)[2])          # D2
    elseif r.n_float == 2
        return u53(block(r, r.it, 0, 4, 0)[1])          # D7
    else
        return u53(block(r, r.it, 0, 4, 0)[2])          # coin
    end
end

function Random.rand(r::PhiloxShim, sp::Random.SamplerRangeNDL{UInt64,Int})
    r.n_index += 1
    s = UInt64(sp.s)
    j = r.n_index <= 2 ? bounded(r, r.it, 1, 0, r.n_index - 1, s) : bounded(r, r.it, 7, r.n_index - 3, 0, s)
    return Int(sp.a + j)
end

Random.rand(r::PhiloxShim, ::Random.SamplerType{UInt128}) = UInt128(0)   # uuid1: ids are sequential

function Random.rand(r::PhiloxShim, sp::Random.SamplerSimple{<:Set{Int}})
    v = sort!(collect(sp[]))
    return v[bounded(r, r.it, 5, 0, 0, UInt64(length(v))) + 1]            # D8: j-th smallest unused label
end

function Random.randn(r::PhiloxShim)
    attempt = 0
    while true
        a, b = block(r, r.it, 0, 6, attempt)
        v1 = 2.0 * u53(a) - 1.0; v2 = 2.0 * u53(b) - 1.0
        s = v1 * v1 + v2 * v2
        if s >= 1.0 || s == 0.0
            attempt += 1
            continue
        end
        return v1 * sqrt((-2.0 * contract_log(s)) / s)
    end
end
Random.randn(r::PhiloxShim, ::Type{Float64}) = randn(r)

end # module
