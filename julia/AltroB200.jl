# AltroB200.jl -- thin `ccall` shim over libaltro_b200.so (include/altro_b200.h) that keeps the names the
# reference's benchmark scripts use: ALTROSolver, SolverOptions, solve!, set_options!, iterations, status,
# states, controls, cost, max_violation, shift_fill!, benchmark_solve!.
#
# NOT EXECUTED in the build environment (no julia there); derived mechanically from the C header and from the
# Python mirror altro_mpc_icra2021_b200/solver.py, which is exercised by the test-suite.  A batch of B
# structurally identical TrajectoryOptimization problems is described by plain arrays (instance-major,
# row-major as C sees them, i.e. Julia arrays with the *reversed* dimension order).
module AltroB200

export ALTROSolver, SolverOptions, solve!, set_options!, iterations, status, states, controls, cost,
       max_violation, shift_fill!, benchmark_solve!, add_constraint!, set_initial_state!, update_trajectory!,
       mpc_run!, EQUALITY, INEQUALITY, SECOND_ORDER_CONE, STATE, CONTROL

const lib = get(ENV, "ALTRO_B200_LIB", joinpath(@__DIR__, "..", "altro_mpc_icra2021_b200", "libaltro_b200.so"))

const EQUALITY, INEQUALITY, SECOND_ORDER_CONE = Cint(0), Cint(1), Cint(2)
const STATE, CONTROL = Cint(0), Cint(1)
@enum TerminationStatus UNSOLVED SOLVE_SUCCEEDED MAX_ITERATIONS MAX_ITERATIONS_OUTER MAXIMUM_COST STATE_LIMIT CONTROL_LIMIT NO_PROGRESS COST_INCREASE NOT_PD

# altro_opts_t, field for field (Altro.SolverOptions)
Base.@kwdef mutable struct SolverOptions
    constraint_tolerance::Cdouble = 1e-6
    cost_tolerance::Cdouble = 1e-4
    cost_tolerance_intermediate::Cdouble = 1e-4
    gradient_tolerance::Cdouble = 10.0
    gradient_tolerance_intermediate::Cdouble = 1.0
    penalty_initial::Cdouble = 1.0
    penalty_scaling::Cdouble = 10.0
    penalty_max::Cdouble = 1e8
    dual_max::Cdouble = 1e8
    line_search_lower_bound::Cdouble = 1e-8
    line_search_upper_bound::Cdouble = 10.0
    max_cost_value::Cdouble = 1e8
    max_state_value::Cdouble = 1e8
    bp_reg_initial::Cdouble = 0.0
    bp_reg_increase_factor::Cdouble = 1.6
    bp_reg_max::Cdouble = 1e8
    bp_reg_min::Cdouble = 1e-8
    bp_reg_fp::Cdouble = 10.0
    iterations::Cint = 1000
    iterations_inner::Cint = 300
    iterations_outer::Cint = 30
    iterations_linesearch::Cint = 20
    dJ_counter_limit::Cint = 10
    reset_duals::Cint = 1
    reset_penalties::Cint = 1
    kickout_max_penalty::Cint = 0
    dj_zero_converges::Cint = 1
    soc_hess_exact::Cint = 1
    soc_viol_proj::Cint = 1
    first_step_unconditional::Cint = 1
end

check(h, rc) = rc == 0 || error("altro_b200: " * unsafe_string(ccall((:altro_last_error, lib), Cstring, (Ptr{Cvoid},), h)))

mutable struct ALTROSolver
    h::Ptr{Cvoid}
    n::Int; m::Int; N::Int; B::Int; P::Int
    opts::SolverOptions
    X::Array{Float64,3}   # (n, N, B)   == C layout [B][N][n]
    U::Array{Float64,3}   # (m, N-1, B)
    iters::Vector{Cint}; iters_outer::Vector{Cint}; stat::Vector{Cint}; ls::Vector{Cint}
    J::Vector{Float64}; Jal::Vector{Float64}; cmax::Vector{Float64}; pmax::Vector{Float64}
    tsolve::Float64       # ms, device time of the last batched solve (solver.stats.tsolve)
end

"""
    ALTROSolver(n, m, N, B, dt; A, Bm, d, Q, R, Qf, Xref, Uref, x0, U0, opts, device=0)

`A`,`Bm`,`d`: `(n,n)` shared LTI, or with trailing `(N-1)` and/or `(B)` dimensions (per knot / per instance).
"""
function ALTROSolver(n, m, N, B, dt; A, Bm, d=nothing, Q, R, Qf, Xref, Uref, x0, U0=zeros(m, N - 1, B),
                     opts=SolverOptions(), device=0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:altro_create, lib), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint, Cdouble), href, device, n, m, N, B, dt)
    rc == 0 || error("altro_create: " * unsafe_string(ccall((:altro_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = href[]
    per_knot = ndims(A) >= 3 && size(A, 3) == N - 1
    per_inst = size(A, ndims(A)) == B && ndims(A) >= 3 && !(ndims(A) == 3 && per_knot)
    check(h, ccall((:altro_set_dynamics, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                   h, per_knot, per_inst, A, Bm, d === nothing ? C_NULL : d))
    check(h, ccall((:altro_set_cost_diag, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h, Q, R, Qf))
    check(h, ccall((:altro_set_reference, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h, Xref, Uref))
    check(h, ccall((:altro_set_x0, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h, x0))
    check(h, ccall((:altro_set_trajectory, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h, C_NULL, U0))
    s = ALTROSolver(h, n, m, N, B, 0, opts, zeros(n, N, B), copy(U0), zeros(Cint, B), zeros(Cint, B), zeros(Cint, B),
                    zeros(Cint, B), zeros(B), zeros(B), zeros(B), zeros(B), 0.0)
    finalizer(x -> ccall((:altro_destroy, lib), Cint, (Ptr{Cvoid},), x.h), s)
    return s
end

"TO.add_constraint!(cons, con, inds) for one affine conic block c = G z[inds] + h (1-based knots `k0:k1`)."
function add_constraint!(s::ALTROSolver, sense, side, knots::UnitRange, inds::Vector{<:Integer}, G, h;
                         per_knot=false, per_instance=false)
    p, w = size(G, 2), size(G, 1)               # G is (w, p, ...) in Julia == [..][p][w] in C
    id = Ref{Cint}(0)
    check(s.h, ccall((:altro_add_constraint, lib), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cint}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cint}),
                     s.h, sense, side, first(knots) - 1, last(knots), p, w, Cint.(inds .- 1), per_knot, per_instance, G, h, id))
    return id[]
end

function push_options!(s::ALTROSolver)
    o = s.opts
    vals = Cdouble[getfield(o, f) for f in fieldnames(SolverOptions)[1:18]]
    ints = Cint[getfield(o, f) for f in fieldnames(SolverOptions)[19:end]]
    buf = vcat(reinterpret(UInt8, vals), reinterpret(UInt8, ints))
    check(s.h, ccall((:altro_set_options, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), s.h, buf))
end

set_options!(s::ALTROSolver; kw...) = (for (k, v) in kw; setfield!(s.opts, k, convert(fieldtype(SolverOptions, k), v)); end; s)
set_initial_state!(s::ALTROSolver, x0) = check(s.h, ccall((:altro_set_x0, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.h, x0))
update_trajectory!(s::ALTROSolver, Xref, Uref) =
    check(s.h, ccall((:altro_set_reference, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, Xref, Uref))

"solve!(solver): batched AL-iLQR solve of all B instances on the GPU; fills states/controls/stats."
function solve!(s::ALTROSolver)
    push_options!(s)
    check(s.h, ccall((:altro_solve, lib), Cint, (Ptr{Cvoid},), s.h))
    fetch!(s)
end

function fetch!(s::ALTROSolver)
    check(s.h, ccall((:altro_get_trajectory, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, s.X, s.U))
    check(s.h, ccall((:altro_get_stats, lib), Cint,
                     (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                     s.h, s.iters, s.iters_outer, s.stat, s.ls, s.J, s.Jal, s.cmax, s.pmax))
    ms = Ref{Cdouble}(0.0)
    check(s.h, ccall((:altro_get_timing, lib), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ptr{Clonglong}), s.h, ms, C_NULL))
    s.tsolve = ms[]
    return s
end

"RD.shift_fill!(Z) and Altro.shift_fill!(get_constraints(solver)) on the device."
shift_fill!(s::ALTROSolver; primal=true, dual=true) =
    check(s.h, ccall((:altro_shift_fill, lib), Cint, (Ptr{Cvoid}, Cint, Cint), s.h, primal, dual))

"Closed-loop MPC run on the device: steps x {transition; solve!} per instance in one launch."
function mpc_run!(s::ALTROSolver, steps::Integer; shift=true)
    push_options!(s)
    check(s.h, ccall((:altro_mpc_run, lib), Cint, (Ptr{Cvoid}, Cint, Cint), s.h, steps, shift))
    fetch!(s)
end

"Quadruped: per-instance bank of linearisations + gait schedule (altro_solver.jl:40-62 without the per-tick upload)."
function set_dynamics_slots!(s::ALTROSolver, A, Bm, d, sched::Array{Cint})
    nslots, sched_len = size(A, 3), size(sched, 1)   # A is (n, n, nslots, B), sched is (sched_len, B) in Julia order
    check(s.h, ccall((:altro_set_dynamics_slots, lib), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
                     s.h, nslots, A, Bm, d, sched, sched_len))
end

"Grasp: constraint data along the whole object trajectory, read at each instance's position (grasp_mpc_helpers.jl:46-55)."
function add_track_constraint!(s::ALTROSolver, sense, side, knots::UnitRange, inds::Vector{<:Integer}, G, h)
    id = Ref{Cint}(0)
    p, w, Nt = size(G, 2), size(G, 1), size(G, 3)    # G is (w, p, Nt) in Julia order
    check(s.h, ccall((:altro_add_track_constraint, lib), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ref{Cint}),
                     s.h, sense, side, first(knots) - 1, last(knots), p, w, Cint.(inds .- 1), G, h, Nt, id))
    return id[]
end
set_track_index!(s::ALTROSolver, kidx::Vector{Cint}) =
    check(s.h, ccall((:altro_set_track_index, lib), Cint, (Ptr{Cvoid}, Ptr{Cint}), s.h, kidx))

"Launch tuning; neither call changes a bit of the results."
set_launch_config!(s::ALTROSolver, threads::Integer) =
    check(s.h, ccall((:altro_set_launch_config, lib), Cint, (Ptr{Cvoid}, Cint), s.h, threads))
set_line_search_mode!(s::ALTROSolver, speculative::Integer) =
    check(s.h, ccall((:altro_set_line_search_mode, lib), Cint, (Ptr{Cvoid}, Cint), s.h, speculative))
reserve_steps!(s::ALTROSolver, steps::Integer) =
    check(s.h, ccall((:altro_reserve_steps, lib), Cint, (Ptr{Cvoid}, Cint), s.h, steps))

"benchmark_solve!(solver; samples, evals): restore-and-resolve; returns device times in ms."
function benchmark_solve!(s::ALTROSolver; samples=10, evals=10)
    push_options!(s)
    check(s.h, ccall((:altro_snapshot, lib), Cint, (Ptr{Cvoid},), s.h))
    t = Float64[]
    for _ in 1:samples*evals
        check(s.h, ccall((:altro_restore, lib), Cint, (Ptr{Cvoid},), s.h))
        check(s.h, ccall((:altro_solve, lib), Cint, (Ptr{Cvoid},), s.h))
        ms = Ref{Cdouble}(0.0)
        check(s.h, ccall((:altro_get_timing, lib), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ptr{Clonglong}), s.h, ms, C_NULL))
        push!(t, ms[])
    end
    fetch!(s)
    return t
end

iterations(s::ALTROSolver) = s.iters
status(s::ALTROSolver) = TerminationStatus.(s.stat)
states(s::ALTROSolver) = s.X
controls(s::ALTROSolver) = s.U
cost(s::ALTROSolver) = s.J
max_violation(s::ALTROSolver) = s.cmax

end # module
