# Integrals.jl backend over the B200 path (unexecuted here: no Julia in the build image — written against
# Integrals.jl v4's extension protocol, `Integrals.__solvebp_call(prob, alg, sensealg, domain, p; reltol, abstol, maxiters)`).
#
#   using Integrals, FastTanhSinhQuadratureB200
#   f   = Builtin(:gauss)                                   # exp(-p[1] x^2) on the device
#   sol = solve(IntegralProblem((x, p) -> exp(-p[1] * x^2), (-1.0, 1.0), [2.0]), TanhSinhB200(f); reltol=1e-10)
#   sol.u ≈ sqrt(pi / 2) * erf(sqrt(2))
#
# A matrix `p` with one COLUMN per integral, or vectors of bounds, make the solve a batch (one device launch).
module FastTanhSinhQuadratureB200IntegralsExt

using Integrals, SciMLBase
using FastTanhSinhQuadratureB200
const FTS = FastTanhSinhQuadratureB200

# the algorithm type joins Integrals.jl's hierarchy through a wrapper, so the shim itself does not depend on SciMLBase
struct B200Alg{A<:FTS.TanhSinhB200} <: SciMLBase.AbstractIntegralAlgorithm
    alg::A
end
SciMLBase.solve(prob::IntegralProblem, alg::FTS.TanhSinhB200; kw...) = SciMLBase.solve(prob, B200Alg(alg); kw...)

function Integrals.__solvebp_call(prob::IntegralProblem, w::B200Alg, sensealg, domain, p;
    reltol=nothing, abstol=nothing, maxiters=nothing)
    lb, ub = domain
    alg = w.alg
    params = p isa SciMLBase.NullParameters ? nothing : (p isa AbstractMatrix ? permutedims(p) : p)   # rows = integrals in the C ABI
    atol = abstol === nothing ? 0 : abstol
    kw = alg.max_levels === nothing ? (;) : (; max_levels=alg.max_levels)
    val = alg.complement ? FTS.quad_cmpl(alg.integrand, lb, ub; rtol=reltol, atol=atol, params=params, kw...) :
          FTS.quad(alg.integrand, lb, ub; rtol=reltol, atol=atol, params=params, kw...)
    # quad's warn path has already reported an integral that stopped at max_levels (adaptive.jl:71-73)
    return SciMLBase.build_solution(prob, w, val, nothing; retcode=ReturnCode.Success)
end

end # module
