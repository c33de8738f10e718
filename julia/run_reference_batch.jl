# run_reference_batch.jl -- runs the REAL Altro.jl (the versions pinned by benchmarks/Manifest.toml of the reference) on
# a batch serialised by scripts/dump_case.py, with Threads.@threads over the instances, and compares iteration counts,
# costs, constraint violations and trajectories with the results recorded by this repository.
#
#   julia --project=<reference>/benchmarks -t auto julia/run_reference_batch.jl <dump_dir>
#
# NOT EXECUTED in the build environment (no julia; the pinned packages are not vendored).  It exists so that a
# maintainer with the reference's environment can close the parity loop that SURVEY.md 8c leaves open, and -- being the
# "Threads.@threads over the batch" baseline BASELINE.json names -- time the reference CPU path on the box's cores.
# Only the rocket / random-linear / flexible-satellite families are wired here (built-in TO constraint types); the grasp
# and quadruped families need the custom constraint structs of the reference's scripts (new_constraints.jl,
# FrictionConstraint.jl), which a maintainer includes from the reference tree.
using LinearAlgebra, StaticArrays, Statistics
using TrajectoryOptimization, RobotDynamics, Altro
const TO = TrajectoryOptimization
const RD = RobotDynamics

# --- minimal JSON reader (numbers, strings, arrays, objects, true/false/null): no package needed ---------------
mutable struct JP; s::String; i::Int; end
ws!(p) = (while p.i <= lastindex(p.s) && isspace(p.s[p.i]); p.i += 1; end)
function jval(p::JP)
    ws!(p); c = p.s[p.i]
    if c == '{'
        p.i += 1; d = Dict{String,Any}(); ws!(p)
        p.s[p.i] == '}' && (p.i += 1; return d)
        while true
            ws!(p); k = jval(p); ws!(p); p.i += 1; d[k] = jval(p); ws!(p)
            p.s[p.i] == ',' ? (p.i += 1) : (p.i += 1; return d)
        end
    elseif c == '['
        p.i += 1; a = Any[]; ws!(p)
        p.s[p.i] == ']' && (p.i += 1; return a)
        while true
            push!(a, jval(p)); ws!(p)
            p.s[p.i] == ',' ? (p.i += 1) : (p.i += 1; return a)
        end
    elseif c == '"'
        j = findnext('"', p.s, p.i + 1); v = p.s[p.i+1:j-1]; p.i = j + 1; return v
    elseif startswith(SubString(p.s, p.i), "true");  p.i += 4; return true
    elseif startswith(SubString(p.s, p.i), "false"); p.i += 5; return false
    elseif startswith(SubString(p.s, p.i), "null");  p.i += 4; return nothing
    else
        j = p.i
        while j <= lastindex(p.s) && (isdigit(p.s[j]) || p.s[j] in "+-.eE"); j += 1; end
        v = parse(Float64, p.s[p.i:j-1]); p.i = j; return v
    end
end
readjson(path) = jval(JP(read(path, String), 1))
mat(a) = a isa Vector && !isempty(a) && a[1] isa Vector ? permutedims(hcat([mat(x) for x in a]...)) : Float64.(a)

function read_step(path, layout)
    raw = reinterpret(Float64, read(path)); out = Dict{String,Any}(); o = 0
    for (name, dims) in layout
        d = Int.(dims); cnt = prod(d)
        out[name] = permutedims(reshape(raw[o+1:o+cnt], reverse(d)...), reverse(1:length(d)))   # back to [B][...] order
        o += cnt
    end
    return out
end

function build_problem(meta, b, x0, Xref, Uref, U0)
    n, m, N, dt = Int(meta["n"]), Int(meta["m"]), Int(meta["N"]), meta["dt"]
    dyn = meta["dynamics"]
    (dyn["per_knot"] || dyn["per_instance"]) && error("only shared LTI dynamics are wired in this driver")
    A, Bm, d = mat(dyn["A"]), mat(dyn["B"]), Float64.(dyn["d"])
    model = RD.LinearModel(SMatrix{n,n}(A), SMatrix{n,m}(Bm), SVector{n}(d); dt=dt)
    Q, R, Qf = Diagonal(SVector{n}(Float64.(meta["Q"]))), Diagonal(SVector{m}(Float64.(meta["R"]))), Diagonal(SVector{n}(Float64.(meta["Qf"])))
    Z = Traj([SVector{n}(Xref[b, k, :]) for k in 1:N], [SVector{m}(Uref[b, min(k, N - 1), :]) for k in 1:N-1], fill(dt, N))
    obj = TO.TrackingObjective(Q, R, Z, Qf=Qf)
    cons = ConstraintList(n, m, N)
    for c in meta["constraints"]
        G, h = mat(c["G"]), Float64.(c["h"]); inds = Int.(c["inds"]) .+ 1; kn = Int(c["k0"])+1:Int(c["k1"])
        side = Int(c["side"]) == 0 ? :state : :control
        zin = side == :state ? inds : n .+ inds
        if Int(c["sense"]) == 2 && size(G, 1) == length(inds) + 1 && G[1:end-1, :] == I && all(G[end, :] .== 0)
            TO.add_constraint!(cons, NormConstraint(n, m, h[end], TO.SecondOrderCone(), side), kn)      # |z| <= val
        elseif Int(c["sense"]) == 1 && all(sum(G .!= 0, dims=2) .<= 1)                                    # bounds
            zmax, zmin = fill(Inf, n + m), fill(-Inf, n + m)
            for r in 1:size(G, 1)
                j = zin[findfirst(!=(0), G[r, :])]
                G[r, findfirst(!=(0), G[r, :])] > 0 ? (zmax[j] = -h[r]) : (zmin[j] = h[r])
            end
            TO.add_constraint!(cons, BoundConstraint(n, m, x_min=zmin[1:n], x_max=zmax[1:n], u_min=zmin[n+1:end], u_max=zmax[n+1:end]), kn)
        else
            error("constraint $(c["name"]): include the reference's custom constraint types (NormConstraint2, ...) to run it")
        end
    end
    prob = TO.Problem(model, obj, SVector{n}(Xref[b, N, :]), (N - 1) * dt; x0=SVector{n}(x0[b, :]), constraints=cons, integration=RD.PassThrough)
    initial_controls!(prob, [SVector{m}(U0[b, k, :]) for k in 1:N-1])
    return prob
end

function main(dir)
    meta = readjson(joinpath(dir, "problem.json"))
    o = meta["options"]
    opts = SolverOptions(cost_tolerance=o["cost_tolerance"], cost_tolerance_intermediate=o["cost_tolerance_intermediate"],
                         constraint_tolerance=o["constraint_tolerance"], penalty_initial=o["penalty_initial"],
                         penalty_scaling=o["penalty_scaling"], reset_duals=o["reset_duals"] != 0, projected_newton=false)
    B, steps = Int(meta["B"]), Int(meta["steps"])
    t_total = 0.0
    for s in 0:steps-1
        st = read_step(joinpath(dir, "step_" * lpad(s, 4, '0') * ".bin"), meta["step_layout"])
        iters = zeros(Int, B); dX = zeros(B); dU = zeros(B); dJ = zeros(B)
        t_total += @elapsed Threads.@threads for b in 1:B
            prob = build_problem(meta, b, st["x0"], st["Xref"], st["Uref"], st["U0"])
            solver = ALTROSolver(prob, opts)
            # duals in: reset_duals = false benchmarks warm-start them (Altro.shift_fill!); written through get_duals-style access
            solve!(solver)
            iters[b] = iterations(solver)
            X = hcat(Vector.(states(solver))...)'; U = hcat(Vector.(controls(solver))...)'
            dX[b] = maximum(abs.(X .- st["X"][b, :, :])) / max(1.0, maximum(abs.(st["X"][b, :, :])))
            dU[b] = maximum(abs.(U .- st["U"][b, :, :])) / max(1.0, maximum(abs.(st["U"][b, :, :])))
            dJ[b] = abs(cost(solver) - st["cost"][b]) / max(1.0, abs(st["cost"][b]))
        end
        same = count(iters .== Int.(st["iterations"]))
        println("step $s: iterations equal in $same / $B instances;  max rel |dX| = $(maximum(dX))  |dU| = $(maximum(dU))  |dJ| = $(maximum(dJ))")
    end
    println("reference CPU throughput: $(B * steps / t_total) solves/s on $(Threads.nthreads()) threads (includes problem construction)")
end

main(ARGS[1])
