# AltroB200TO.jl -- the reference-facing constructor: ALTROSolver(prob::TO.Problem, opts::SolverOptions; kw...).
#
# Drop-in for the path the reference's scripts drive (random_linear_problem.jl:87-139, mpc.jl:11-47,
# simple_rocket.jl:59-82,128-129, altro_solver.jl:35-72, grasp_mpc_helpers.jl:46-55, flexible_sat_mpc.jl:162-166,271-272):
#   * reads size(prob), prob.model, prob.obj, prob.x0 and walks prob.constraints, lowering every constraint type the
#     benchmarks use to the affine conic blocks of the C ABI -- by the SAME table the Python mirror uses
#     (altro_mpc_icra2021_b200/lowering.tbl, checked by tests/test_lowering_table.py);
#   * the Problem is shared by reference, as in Altro: the scripts mutate prob.x0, prob.model.A[i], cons[i].A, the
#     objective (TO.update_trajectory!) and prob.Z (RD.shift_fill!) in place BETWEEN solves, so solve! re-reads all of
#     them (uploading only what changed, by hash) and writes the solution back into prob.Z;
#   * a Vector of structurally identical Problems is one batch (B = length): the GPU solves them together.
#
# NOT EXECUTED in the build environment (no julia, and the pinned TrajectoryOptimization / RobotDynamics / Altro
# branches are not vendored): derived mechanically from problem.py / solver.py, which the test-suite exercises.
# Unknown constraint or model types are an error, never a CPU fallback.
module AltroB200TO

using LinearAlgebra
import TrajectoryOptimization
import RobotDynamics
const TO = TrajectoryOptimization
const RD = RobotDynamics

include("AltroB200.jl")
import .AltroB200: EQUALITY, INEQUALITY, SECOND_ORDER_CONE, STATE, CONTROL
const LL = AltroB200

# Altro.SolverOptions as the scripts construct it (run_random_linear.jl:41-49, run_simple_rocket.jl:121-129,
# ALTROParams.jl:86-95): every field of the C struct (LL.SolverOptions, same names, types and defaults) plus the
# keywords the scripts pass that have no counterpart on this path.  projected_newton = true is refused at upload (the
# polish step is not built; every benchmark sets it false); static_bp, verbose and show_summary are accepted and ignored.
@eval Base.@kwdef mutable struct SolverOptions
    $([:($f::$(fieldtype(LL.SolverOptions, f)) = $(getfield(LL.SolverOptions(), f))) for f in fieldnames(LL.SolverOptions)]...)
    projected_newton::Bool = false
    static_bp::Bool = true
    verbose::Int = 0
    show_summary::Bool = false
end
function ll_options(o::SolverOptions)
    o.projected_newton && error("AltroB200: projected_newton = true is not supported by this solve path")
    return LL.SolverOptions(; (f => getfield(o, f) for f in fieldnames(LL.SolverOptions))...)
end

export ALTROSolver, SolverOptions, solve!, set_options!, iterations, status, states, controls, cost, max_violation,
       benchmark_solve!, get_constraints, get_trajectory, shift_fill!

# ---------------------------------------------------------------------------------------------- lowering table
struct Rule
    sense::Symbol; side::Symbol; G::Symbol; h::Symbol; layout::Symbol
end

function read_lowering_table(path=joinpath(@__DIR__, "..", "altro_mpc_icra2021_b200", "lowering.tbl"))
    rules = Dict{Symbol,Rule}()
    for line in eachline(path)
        (isempty(strip(line)) || startswith(line, "#")) && continue
        f = strip.(split(line, "|"))
        rules[Symbol(f[1])] = Rule(Symbol(f[2]), Symbol(f[3]), Symbol(f[4]), Symbol(f[5]), Symbol(f[6]))
    end
    return rules
end
const LOWERING = read_lowering_table()

sense_code(::TO.Equality) = EQUALITY
sense_code(::TO.Inequality) = INEQUALITY
sense_code(::TO.SecondOrderCone) = SECOND_ORDER_CONE
sense_code(r::Rule, con) = r.sense == :from_con ? sense_code(TO.sense(con)) :
                           r.sense == :Equality ? EQUALITY : r.sense == :Inequality ? INEQUALITY : SECOND_ORDER_CONE

"One affine block: side, 1-based indices on that side, G (p x w) or one per knot, h, sense."
struct Block
    sense::Cint; side::Cint; inds::Vector{Int}; G::Vector{Matrix{Float64}}; h::Vector{Vector{Float64}}; per_knot::Bool
end

# side / index rules ---------------------------------------------------------------------------------------
function side_inds(::Val{:state}, con, n, m) ; return [(STATE, collect(1:n))] end
function side_inds(::Val{:control}, con, n, m) ; return [(CONTROL, collect(1:m))] end
function side_inds(::Val{:con_inds}, con, n, m)
    inds = hasproperty(con, :inds) ? collect(Int, con.inds) : collect(n .+ (1:m))
    if all(inds .<= n)
        return [(STATE, inds)]
    elseif all(inds .> n)
        return [(CONTROL, inds .- n)]
    elseif hasproperty(con, :m) && !hasproperty(con, :n) && maximum(inds) <= m   # TO.ControlConstraint: inds over u
        return [(CONTROL, inds)]
    end
    error("AltroB200: a constraint block must live on one side (state or control): $(typeof(con))")
end
side_inds(::Val{:foot}, con, n, m) = [(CONTROL, collect(3 * (con.i - 1) .+ (1:3)))]
function side_inds(::Val{:finite}, con::TO.BoundConstraint, n, m)
    zmax, zmin = Vector(con.z_max), Vector(con.z_min)
    out = Tuple{Cint,Vector{Int}}[]
    for (side, rng, off) in ((STATE, 1:n, 0), (CONTROL, n .+ (1:m), n))
        idx = [j - off for j in rng if isfinite(zmax[j]) || isfinite(zmin[j])]
        isempty(idx) || push!(out, (side, idx))
    end
    return out
end

# G / h rules (data of knot i of the constraint's range; shared constraints ignore i) -------------------------
knot_data(x::AbstractVector{<:AbstractArray}, i) = x[i]
knot_data(x, i) = x
G_rule(::Val{:identity}, con, side, inds, n, i) = Matrix{Float64}(I, length(inds), length(inds))
G_rule(::Val{:identity_0}, con, side, inds, n, i) = [Matrix{Float64}(I, length(inds), length(inds)); zeros(1, length(inds))]
G_rule(::Val{:stack_A_c}, con, side, inds, n, i) = [Matrix{Float64}(knot_data(con.A, i)); Vector{Float64}(knot_data(con.c, i))']
G_rule(::Val{:A}, con, side, inds, n, i) = Matrix{Float64}(knot_data(con.A, i))
G_rule(::Val{:friction_pyramid}, con, side, inds, n, i) = [1.0 0 -con.μ; -1.0 0 -con.μ; 0 1.0 -con.μ; 0 -1.0 -con.μ]
function G_rule(::Val{:bound_rows}, con, side, inds, n, i)
    off = side == STATE ? 0 : n
    zmax, zmin = Vector(con.z_max), Vector(con.z_min)
    up = [k for (k, j) in enumerate(inds) if isfinite(zmax[j + off])]
    dn = [k for (k, j) in enumerate(inds) if isfinite(zmin[j + off])]
    G = zeros(length(up) + length(dn), length(inds))
    for (r, k) in enumerate(up); G[r, k] = 1.0; end
    for (r, k) in enumerate(dn); G[length(up) + r, k] = -1.0; end
    return G
end
h_rule(::Val{:zero}, con, side, inds, n, p, i) = zeros(p)
h_rule(::Val{:minus_b}, con, side, inds, n, p, i) = -Vector{Float64}(knot_data(con.b, i))
h_rule(::Val{:minus_xf}, con, side, inds, n, p, i) = -Vector{Float64}(con.xf)[inds]
h_rule(::Val{:val_last}, con, side, inds, n, p, i) = [zeros(p - 1); Float64(con.val)]
function h_rule(::Val{:bound_rhs}, con, side, inds, n, p, i)
    off = side == STATE ? 0 : n
    zmax, zmin = Vector(con.z_max), Vector(con.z_min)
    return [[-zmax[j + off] for j in inds if isfinite(zmax[j + off])]; [zmin[j + off] for j in inds if isfinite(zmin[j + off])]]
end

"Lowers one TO constraint over the 1-based knot range `knots` of an N-knot problem by the shared table."
function lower(con, knots::UnitRange, n, m, N)
    name = Symbol(nameof(typeof(con)))
    haskey(LOWERING, name) || error("AltroB200: constraint type $name is not in lowering.tbl (no CPU fallback)")
    r = LOWERING[name]
    blocks = Tuple{Block,UnitRange{Int}}[]
    for (side, inds) in side_inds(Val(r.side), con, n, m)
        kn = side == CONTROL ? (first(knots):min(last(knots), N - 1)) : knots   # u_N is not a decision variable
        isempty(kn) && continue
        per_knot = r.layout == :per_knot
        idx = per_knot ? collect(1:length(kn)) : [1]
        Gs = [G_rule(Val(r.G), con, side, inds, n, i) for i in idx]
        hs = [h_rule(Val(r.h), con, side, inds, n, size(Gs[1], 1), i) for i in idx]
        push!(blocks, (Block(sense_code(r, con), side, inds, Gs, hs, per_knot), kn))
    end
    return blocks
end

# ---------------------------------------------------------------------------------------------- model / objective
"(A, B, d) of one knot of a discrete affine model: RD.LinearModel (random_linear_problem.jl:8, ALTROParams.jl:61), or any
model whose discrete dynamics are affine (rocket, grasp: rocket_landing_problem.jl:33-39, grasp_model.jl:74-92), probed
exactly with n + m + 1 evaluations."
function affine_knot(prob, k)
    mdl = prob.model
    n, m, N = size(prob)
    if mdl isa RD.LinearModel
        i = length(mdl.A) == 1 ? 1 : k
        d = isempty(mdl.d) ? zeros(n) : Vector{Float64}(mdl.d[min(i, length(mdl.d))])
        return Matrix{Float64}(mdl.A[i]), Matrix{Float64}(mdl.B[i]), d
    end
    z = prob.Z[k]
    f(x, u) = Vector{Float64}(RD.discrete_dynamics(TO.integration(prob), mdl, RD.StaticKnotPoint(z, [x; u])))
    d = f(zeros(n), zeros(m))
    A = hcat([f(Float64.(1:n .== j), zeros(m)) - d for j in 1:n]...)
    B = hcat([f(zeros(n), Float64.(1:m .== j)) - d for j in 1:m]...)
    return A, B, d
end
time_varying(prob) = prob.model isa RD.LinearModel && length(prob.model.A) > 1

"Diagonal weights and tracking reference of a TO.Objective of DiagonalCost / QuadraticCost terms: q = -Q xref, r = -R uref."
function diag_objective(obj, n, m, N)
    c1, cN = obj[1], obj[N]
    Q, R, Qf = diag(Matrix(c1.Q)), diag(Matrix(c1.R)), diag(Matrix(cN.Q))
    (isdiag(Matrix(c1.Q)) && isdiag(Matrix(c1.R)) && isdiag(Matrix(cN.Q))) || error("AltroB200: only diagonal weights")
    Xref = zeros(n, N); Uref = zeros(m, N - 1)
    for k in 1:N
        Qk = k == N ? Qf : Q
        Xref[:, k] = -Vector(obj[k].q) ./ Qk
        k < N && (Uref[:, k] = -Vector(obj[k].r) ./ R)
    end
    return Q, R, Qf, Xref, Uref
end

# ---------------------------------------------------------------------------------------------- the solver
mutable struct Stats
    tsolve::Float64; iterations::Int; iterations_outer::Int; status::LL.TerminationStatus
    cost::Vector{Float64}; c_max::Vector{Float64}
end

mutable struct ALTROSolver
    probs::Vector                     # the shared Problems (batch of B), re-read at every solve!
    ll::LL.ALTROSolver
    opts::SolverOptions
    blocks::Vector{Tuple{Any,UnitRange{Int},Block,Int}}   # (constraint object, knots, lowered block, ABI id)
    hashes::Dict{Symbol,UInt}
    stats::Stats
    solver_al::Any                    # `altro.solver_al` (simple_rocket.jl:81): the object shift_fill! accepts
end
struct ConstraintHandle; s::ALTROSolver; end

ALTROSolver(prob::TO.Problem, opts::SolverOptions=SolverOptions(); kw...) = ALTROSolver([prob], opts; kw...)
function ALTROSolver(probs::Vector{<:TO.Problem}, opts::SolverOptions=SolverOptions(); device=0, kw...)
    opts = deepcopy(opts)
    for (k, v) in kw   # ALTROSolver(prob, opts; show_summary=true, verbose=1) (run_simple_rocket.jl:66)
        k in fieldnames(SolverOptions) && setfield!(opts, k, convert(fieldtype(SolverOptions, k), v))
    end
    p1 = probs[1]
    n, m, N = size(p1)
    B = length(probs)
    dt = p1.Z[1].dt
    A, Bm, d = gather_dynamics(probs)
    Q, R, Qf, _, _ = diag_objective(p1.obj, n, m, N)
    Xref = zeros(n, N, B); Uref = zeros(m, N - 1, B); x0 = zeros(n, B); U0 = zeros(m, N - 1, B)
    for (b, p) in enumerate(probs)
        _, _, _, Xref[:, :, b], Uref[:, :, b] = diag_objective(p.obj, n, m, N)
        x0[:, b] = p.x0
        for k in 1:N-1; U0[:, k, b] = RD.control(p.Z[k]); end
    end
    ll = LL.ALTROSolver(n, m, N, B, dt; A=A, Bm=Bm, d=d, Q=Q, R=R, Qf=Qf, Xref=Xref, Uref=Uref, x0=x0, U0=U0, opts=ll_options(opts), device=device)
    blocks = Tuple{Any,UnitRange{Int},Block,Int}[]
    for (inds, con) in zip(p1.constraints)
        for (blk, kn) in lower(con, inds, n, m, N)
            G, h = pack_block(probs, con, inds, blk, n, m, N)
            id = LL.add_constraint!(ll, blk.sense, blk.side, kn, blk.inds, G, h; per_knot=blk.per_knot, per_instance=B > 1)
            push!(blocks, (con, inds, blk, id))
        end
    end
    s = ALTROSolver(probs, ll, opts, blocks, Dict{Symbol,UInt}(),
                    Stats(0.0, 0, 0, LL.UNSOLVED, zeros(B), zeros(B)), nothing)
    s.solver_al = ConstraintHandle(s)
    return s
end

"Per-instance / per-knot dynamics of the batch in the C layout (Julia dims reversed): (n, n[, N-1][, B])."
function gather_dynamics(probs)
    n, m, N = size(probs[1])
    B = length(probs)
    tv = time_varying(probs[1])
    K = tv ? N - 1 : 1
    A = zeros(n, n, K, B); Bm = zeros(m, n, K, B); d = zeros(n, K, B)
    for (b, p) in enumerate(probs), k in 1:K
        Ak, Bk, dk = affine_knot(p, k)
        A[:, :, k, b] = Ak'; Bm[:, :, k, b] = Bk'; d[:, k, b] = dk      # row-major for C
    end
    shared = B == 1 || all(A[:, :, :, b] == A[:, :, :, 1] && Bm[:, :, :, b] == Bm[:, :, :, 1] for b in 2:B)
    if shared && !tv
        return A[:, :, 1, 1], Bm[:, :, 1, 1], d[:, 1, 1]
    elseif shared
        return A[:, :, :, 1], Bm[:, :, :, 1], d[:, :, 1]
    end
    return tv ? (A, Bm, d) : (A[:, :, 1, :], Bm[:, :, 1, :], d[:, 1, :])
end

"G, h of one block for the whole batch in the C layout: G is (w, p[, nk][, B])."
function pack_block(probs, con1, inds, blk, n, m, N)
    B = length(probs)
    nk = blk.per_knot ? length(blk.G) : 1
    p, w = size(blk.G[1])
    G = zeros(w, p, nk, B); h = zeros(p, nk, B)
    for (b, pr) in enumerate(probs)
        conb = b == 1 ? con1 : nth_constraint(pr, con1, inds)
        lb = first(x for x in lower(conb, inds, n, m, N) if x[1].side == blk.side)[1]
        for k in 1:nk
            G[:, :, k, b] = lb.G[k]'; h[:, k, b] = lb.h[k]
        end
    end
    blk.per_knot || (G = G[:, :, 1, :]; h = h[:, 1, :])
    B == 1 && (G = dropdims(G, dims=ndims(G)); h = dropdims(h, dims=ndims(h)))
    return G, h
end
nth_constraint(pr, con1, inds) = first(c for (i, c) in zip(pr.constraints) if i == inds && typeof(c) == typeof(con1))

# what the scripts mutate in place between two solve! calls --------------------------------------------------
function refresh!(s::ALTROSolver)
    p1 = s.probs[1]
    n, m, N = size(p1)
    B = length(s.probs)
    changed(key, val) = (hv = hash(val); old = get(s.hashes, key, UInt(0)); s.hashes[key] = hv; hv != old)
    x0 = hcat([Vector{Float64}(p.x0) for p in s.probs]...)
    changed(:x0, x0) && LL.set_initial_state!(s.ll, x0)                      # problem.x0 .= x0_new / TO.set_initial_state!
    Xref = zeros(n, N, B); Uref = zeros(m, N - 1, B)
    for (b, p) in enumerate(s.probs)
        _, _, _, Xref[:, :, b], Uref[:, :, b] = diag_objective(p.obj, n, m, N)
    end
    changed(:ref, (Xref, Uref)) && LL.update_trajectory!(s.ll, Xref, Uref)   # TO.update_trajectory!(prob.obj, Z_track, k)
    if p1.model isa RD.LinearModel                                           # opt.model.A[i] = ... (altro_solver.jl:35-37)
        A, Bm, d = gather_dynamics(s.probs)
        if changed(:dyn, (A, Bm, d))
            pk = time_varying(p1)
            LL.check(s.ll.h, ccall((:altro_set_dynamics, LL.lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                                   s.ll.h, pk, ndims(A) > (pk ? 3 : 2), A, Bm, d))
        end
    end
    for (con, inds, blk, id) in s.blocks                                     # cons[1].A[i] = ... (grasp_mpc_helpers.jl:46-55)
        blk.per_knot || continue
        G, h = pack_block(s.probs, con, inds, blk, n, m, N)
        changed(Symbol(:con, id), (G, h)) &&
            LL.check(s.ll.h, ccall((:altro_update_constraint_data, LL.lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), s.ll.h, id, G, h))
    end
    U0 = zeros(m, N - 1, B)                                                  # RD.shift_fill!(prob.Z) / initial_controls!
    for (b, p) in enumerate(s.probs), k in 1:N-1; U0[:, k, b] = RD.control(p.Z[k]); end
    changed(:U0, U0) && LL.check(s.ll.h, ccall((:altro_set_trajectory, LL.lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.ll.h, C_NULL, U0))
end

"solve!(solver): re-read the shared Problems, batched AL-iLQR solve on the GPU, write the solution back into prob.Z."
function solve!(s::ALTROSolver)
    s.ll.opts = ll_options(s.opts)
    refresh!(s)
    LL.solve!(s.ll)
    n, m, N = size(s.probs[1])
    for (b, p) in enumerate(s.probs), k in 1:N
        RD.set_state!(p.Z[k], s.ll.X[:, k, b])
        k < N && RD.set_control!(p.Z[k], s.ll.U[:, k, b])
    end
    s.hashes[:U0] = hash(s.ll.U)   # the warm start now on the device IS the solution
    s.stats = Stats(s.ll.tsolve, Int(maximum(s.ll.iters)), Int(maximum(s.ll.iters_outer)),
                    LL.TerminationStatus(all(s.ll.stat .== 1) ? 1 : Int(first(x for x in s.ll.stat if x != 1))),
                    copy(s.ll.J), copy(s.ll.cmax))
    return s
end

set_options!(s::ALTROSolver; kw...) = (for (k, v) in kw; k in fieldnames(SolverOptions) && setfield!(s.opts, k, convert(fieldtype(SolverOptions, k), v)); end; s)
iterations(s::ALTROSolver) = length(s.probs) == 1 ? Int(s.ll.iters[1]) : Int.(s.ll.iters)
status(s::ALTROSolver) = s.stats.status
states(s::ALTROSolver) = TO.states(s.probs[1])
controls(s::ALTROSolver) = TO.controls(s.probs[1])
get_trajectory(s::ALTROSolver) = s.probs[1].Z
cost(s::ALTROSolver) = length(s.probs) == 1 ? s.ll.J[1] : copy(s.ll.J)
max_violation(s::ALTROSolver) = length(s.probs) == 1 ? s.ll.cmax[1] : copy(s.ll.cmax)
get_constraints(s::ALTROSolver) = ConstraintHandle(s)
"Altro.shift_fill!(TO.get_constraints(altro)) / Altro.shift_fill!(altro.solver_al): dual warm start, on the device."
shift_fill!(c::ConstraintHandle) = LL.shift_fill!(c.s.ll; primal=false, dual=true)
function benchmark_solve!(s::ALTROSolver; samples=10, evals=10)
    refresh!(s)
    t = LL.benchmark_solve!(s.ll; samples=samples, evals=evals)
    return (times=t .* 1e6, median_ns=sort(t)[cld(length(t), 2)] * 1e6)   # median(b).time is in ns (random_linear_problem.jl:173)
end

end # module
