# SqpQpB200.jl -- the Julia side of the drop-in boundary (NOT executable in the build image:
# no Julia toolchain there; kept thin so that review can substitute for execution).
#
# Two ways into libsqpqp.so (include/sqpqp.h), both through `ccall`:
#
#  B2 (fast lane)  QpDevice <: SqpSolver.AbstractSubOptimizer
#      same method set as QpJuMP (src/algorithms/subproblem_JuMP.jl:23-24, 36-183, 352-393):
#      create_model!, sub_optimize!, sub_optimize_FR!, sub_optimize_lp -- the raw COO value
#      arrays of the evaluator (sqp.dE, sqp.h_val, sqp.df, sqp.E; sqp.jl:86-117) go to the GPU,
#      JuMP is skipped.  Needs the one-line dispatch edit shown in INTEGRATION.md.
#
#  B1 (mandatory)  SqpQpB200.Optimizer <: MOI.AbstractOptimizer
#      usable as `"external_optimizer" => SqpQpB200.Optimizer` with an UNMODIFIED SqpSolver:
#      JuMP copies the QP it builds (ScalarQuadratic objective, ScalarAffine rows in
#      EqualTo/GreaterThan/LessThan/Interval, variable bounds) through `MOI.copy_to`, the shim
#      flattens it to triplets and calls the generic lane (sqpqp_qp_setup / sqpqp_qp_solve).
#      `MOI.supports_incremental_interface` is false, so JuMP's CachingOptimizer re-copies the
#      (edited) model at every `optimize!` -- exactly what SqpSolver's per-iteration
#      `set_normalized_coefficient` / `set_normalized_rhs` / `@objective` edits need.
module SqpQpB200

import MathOptInterface
const MOI = MathOptInterface

const libsqpqp = get(ENV, "SQPQP_LIB", joinpath(@__DIR__, "..", "csrc", "libsqpqp.so"))

struct Info
    moi_status::Int32; admm_iters::Int32; cg_iters::Int32; polish_tries::Int32
    polish_cg_iters::Int32; polished::Int32; rho_updates::Int32; checks::Int32
    ipm_iters::Int32; chol_factorizations::Int32
    rho::Float64; rho_box_floor::Float64; res_prim::Float64; res_dual::Float64; objective::Float64
end

check(h, rc) = rc == 0 || error("sqpqp error $rc: " * unsafe_string(ccall((:sqpqp_last_error, libsqpqp), Cstring, (Ptr{Cvoid},), h)))

function create_handle(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:sqpqp_create, libsqpqp), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
    rc == 0 || error("sqpqp_create failed ($rc): no CUDA device $device?  There is no CPU fallback.")
    return h[]
end
destroy_handle(h) = ccall((:sqpqp_destroy, libsqpqp), Cint, (Ptr{Cvoid},), h)

# ------------------------------------------------------------------------------------------
# B2: QpDevice  (drop-in for QpJuMP inside SqpSolver)
# ------------------------------------------------------------------------------------------
mutable struct QpDevice
    h::Ptr{Cvoid}
    n::Int; m::Int; num_linear::Int; S::Int
    slack_rows::Vector{Int}     # row (1-based) of every slack column, in the library's column order
    function QpDevice(n, m, num_linear, j_row::Vector{Int}, j_col::Vector{Int}, h_row::Vector{Int}, h_col::Vector{Int},
                      x_L, x_U, g_L, g_U; device = 0)
        h = create_handle(device)
        check(h, ccall((:sqpqp_setup_nlp, libsqpqp), Cint,
            (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
            h, 1, n, m, num_linear, length(j_row), j_row, j_col, length(h_row), h_row, h_col, x_L, x_U, g_L, g_U, 0))
        S = Ref{Int32}(0)
        check(h, ccall((:sqpqp_num_slacks, libsqpqp), Cint, (Ptr{Cvoid}, Ref{Int32}), h, S))
        # slack columns of create_model! (subproblem_JuMP.jl:59-65): rows i > num_linear get u_i, and v_i as well when
        # both bounds are finite -- the column order of sqpqp_num_slacks
        slack_rows = Int[]
        for i in (num_linear + 1):m
            push!(slack_rows, i)
            (g_L[i] > -Inf && g_U[i] < Inf) && push!(slack_rows, i)
        end
        @assert length(slack_rows) == S[]
        qp = new(h, n, m, num_linear, S[], slack_rows)
        finalizer(q -> destroy_handle(q.h), qp)
        return qp
    end
end

"eval_functions!/eval_Jacobian! scatter (sqp.jl:86-117) + QpData refresh (sqp.jl:66-79)"
# Page-lock the host's persistent vectors once (sqp.dE, sqp.h_val, sqp.df, sqp.E, sqp.x, sqp.lambda, ...: sqp.jl:16-59): every
# later call copies them straight over the link instead of through the library's staging area.
register!(qp::QpDevice, v::Vector) = check(qp.h, ccall((:sqpqp_host_register, libsqpqp), Cint,
    (Ptr{Cvoid}, Ptr{Cvoid}, Int64), qp.h, pointer(v), sizeof(v)))
update!(qp::QpDevice, dE, h_val, df, E) = check(qp.h, ccall((:sqpqp_update_nlp, libsqpqp), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), qp.h, dE, h_val, df, E))

function _solve(qp::QpDevice, phase::Integer, x_k::Vector{Float64}, Δ::Float64, E_override = C_NULL)
    p = zeros(qp.n); λ = zeros(qp.m); mxL = zeros(qp.n); mxU = zeros(qp.n); slack = zeros(max(qp.S, 1))
    st = Ref{Int32}(0); info = Ref{Info}()
    check(qp.h, ccall((:sqpqp_solve_tr, libsqpqp), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ref{Int32}, Ref{Info}),
        qp.h, phase, x_k, Δ, E_override, C_NULL, p, λ, mxL, mxU, slack, st, info))
    # same 6-tuple as collect_solution! (subproblem_JuMP.jl:182, 514-563), p_slack in the reference's shape
    # Dict{Int,Vector{Float64}} keyed by row (field type sqp.jl:22; filled at subproblem_JuMP.jl:519-529: one or two
    # slack values per nonlinear row, absent for the linear rows)
    p_slack = Dict{Int,Vector{Float64}}()
    for (c, i) in enumerate(qp.slack_rows)
        push!(get!(p_slack, i, Float64[]), slack[c])
    end
    return p, λ, mxU, mxL, p_slack, MOI.TerminationStatusCode(Int(st[]))
end
create_model!(qp::QpDevice, Δ) = nothing                                    # pattern was built in the constructor
sub_optimize!(qp::QpDevice, x_k, Δ) = _solve(qp, 0, x_k, Δ)                  # subproblem_JuMP.jl:127-183
sub_optimize_FR!(qp::QpDevice, x_k, Δ) = _solve(qp, 1, x_k, Δ)               # :352-393
sub_optimize_soc!(qp::QpDevice, x_k, Δ, E_soc) = _solve(qp, 2, x_k, Δ, E_soc) # sqp_trust_region.jl:341-360
function sub_optimize_lp(qp::QpDevice, x_k)                                   # subproblem_JuMP.jl:185-244
    x, λ, mxU, mxL, _, st = _solve(qp, 3, x_k, Inf)
    λ[(qp.num_linear + 1):end] .= 0.0
    return x, λ, mxU, mxL, st
end

# The arithmetic around each QP solve, on the device matrices of the last update!
"norm_violations / compute_phi / compute_qmodel (common.jl:54-77, sqp.jl:170-183, sqp_trust_region.jl:487-508)"
function merit(qp::QpDevice, x, p, E_trial, f_trial::Float64, μ::Float64, fr::Bool)
    o = [Ref(0.0) for _ in 1:5]
    check(qp.h, ccall((:sqpqp_merit, libsqpqp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Float64}, Ref{Int32},
         Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}),
        qp.h, x, p, E_trial, f_trial, μ, Int32(fr), o[1], o[2], o[3], o[4], o[5]))
    return (viol0 = o[1][], viol_trial = o[2][], phi_trial = o[3][], q0 = o[4][], qk = o[5][])
end
"KT_residuals (common.jl:14-23), as coded"
function kt_residuals(qp::QpDevice, λ, mult_x_U, mult_x_L)
    kt = Ref(0.0)
    check(qp.h, ccall((:sqpqp_kt_residuals, libsqpqp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}), qp.h, λ, mult_x_U, mult_x_L, kt))
    return kt[]
end
"which = 0: Jacobian * x, 1: Jacobian' * x, 2: Hessian * x (sqp_trust_region.jl:343,490,492, common.jl:17)"
function spmv(qp::QpDevice, which::Integer, x::Vector{Float64})
    y = zeros(which == 0 ? qp.m : qp.n)
    check(qp.h, ccall((:sqpqp_spmv, libsqpqp), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), qp.h, which, x, y))
    return y
end
"line-search merit primitives (sqp_line_search.jl:271-334, sqp.jl:190-213, merit.jl:13-17, common.jl:30-47)"
function linesearch_terms(qp::QpDevice, x, p, α::Float64, E_trial, μ_rows, λ)
    out = zeros(8)
    check(qp.h, ccall((:sqpqp_linesearch_terms, libsqpqp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        qp.h, x, p, α, E_trial, μ_rows, λ, out))
    return (dfp = out[1], pHp = out[2], viol1 = out[3], violinf = out[4], wviol0 = out[5], wviol_trial = out[6],
            viol1_trial = out[7], compl = out[8])
end

# ------------------------------------------------------------------------------------------
# B1: generic MOI optimizer (copy_to based)
# ------------------------------------------------------------------------------------------
mutable struct Optimizer <: MOI.AbstractOptimizer
    h::Ptr{Cvoid}
    silent::Bool
    options::Dict{String,Any}          # unknown raw attributes are accepted and ignored
    nv::Int; nc::Int
    x::Vector{Float64}; row_dual::Vector{Float64}; col_dual::Vector{Float64}
    rows::Vector{Tuple{MOI.ConstraintIndex,Int}}   # affine constraint -> row
    status::Int32; objective::Float64; solved::Bool
    pattern_key::UInt
    Optimizer(; kwargs...) = (o = new(C_NULL, false, Dict{String,Any}(string(k) => v for (k, v) in kwargs), 0, 0,
                                      Float64[], Float64[], Float64[], Tuple{MOI.ConstraintIndex,Int}[], 0, NaN, false, UInt(0));
                              finalizer(x -> x.h == C_NULL || destroy_handle(x.h), o); o)
end
MOI.get(::Optimizer, ::MOI.SolverName) = "sqpqp-b200"
MOI.is_empty(o::Optimizer) = o.nv == 0
function MOI.empty!(o::Optimizer); o.nv = 0; o.nc = 0; o.solved = false; empty!(o.rows); return; end
MOI.supports(::Optimizer, ::MOI.Silent) = true
MOI.set(o::Optimizer, ::MOI.Silent, v::Bool) = (o.silent = v)
MOI.supports(::Optimizer, ::MOI.RawOptimizerAttribute) = true     # print_level, mu_strategy, linear_solver, ... (ignored)
MOI.set(o::Optimizer, a::MOI.RawOptimizerAttribute, v) = (o.options[a.name] = v)
MOI.get(o::Optimizer, a::MOI.RawOptimizerAttribute) = o.options[a.name]
MOI.supports_incremental_interface(::Optimizer) = false
const _SAF = MOI.ScalarAffineFunction{Float64}
const _SQF = MOI.ScalarQuadraticFunction{Float64}
const _SETS = Union{MOI.EqualTo{Float64},MOI.GreaterThan{Float64},MOI.LessThan{Float64},MOI.Interval{Float64}}
MOI.supports_constraint(::Optimizer, ::Type{_SAF}, ::Type{<:_SETS}) = true
MOI.supports_constraint(::Optimizer, ::Type{MOI.VariableIndex}, ::Type{<:_SETS}) = true
MOI.supports(::Optimizer, ::MOI.ObjectiveSense) = true
MOI.supports(::Optimizer, ::MOI.ObjectiveFunction{<:Union{_SAF,_SQF}}) = true

_bounds(s::MOI.EqualTo) = (s.value, s.value)
_bounds(s::MOI.GreaterThan) = (s.lower, Inf)
_bounds(s::MOI.LessThan) = (-Inf, s.upper)
_bounds(s::MOI.Interval) = (s.lower, s.upper)

# The whole (small) QP is flattened here; the solve is deferred to optimize!.
mutable struct _Flat
    p_row::Vector{Int64}; p_col::Vector{Int64}; p_val::Vector{Float64}; q::Vector{Float64}
    a_row::Vector{Int64}; a_col::Vector{Int64}; a_val::Vector{Float64}
    rl::Vector{Float64}; ru::Vector{Float64}; cl::Vector{Float64}; cu::Vector{Float64}; sense::Float64
end
const _FLAT = IdDict{Optimizer,_Flat}()

function MOI.copy_to(dest::Optimizer, src::MOI.ModelLike)
    MOI.empty!(dest)
    vis = MOI.get(src, MOI.ListOfVariableIndices())
    idx = MOI.Utilities.IndexMap()
    for (k, vi) in enumerate(vis); idx[vi] = MOI.VariableIndex(k); end
    nv = length(vis)
    F = _Flat(Int64[], Int64[], Float64[], zeros(nv), Int64[], Int64[], Float64[], Float64[], Float64[], fill(-Inf, nv), fill(Inf, nv), 1.0)
    sense = MOI.get(src, MOI.ObjectiveSense())
    F.sense = sense == MOI.MAX_SENSE ? -1.0 : 1.0
    if sense != MOI.FEASIBILITY_SENSE
        T = MOI.get(src, MOI.ObjectiveFunctionType())
        f = MOI.get(src, MOI.ObjectiveFunction{T}())
        if f isa _SQF
            for t in f.quadratic_terms  # MOI: 1/2 x'Qx, off-diagonal term c <=> Q_ij = Q_ji = c (sqpqp.h generic lane)
                push!(F.p_row, idx[t.variable_1].value); push!(F.p_col, idx[t.variable_2].value); push!(F.p_val, F.sense * t.coefficient)
            end
            for t in f.affine_terms; F.q[idx[t.variable].value] += F.sense * t.coefficient; end
        elseif f isa _SAF
            for t in f.terms; F.q[idx[t.variable].value] += F.sense * t.coefficient; end
        end
    end
    for S in (MOI.EqualTo{Float64}, MOI.GreaterThan{Float64}, MOI.LessThan{Float64}, MOI.Interval{Float64})
        for ci in MOI.get(src, MOI.ListOfConstraintIndices{MOI.VariableIndex,S}())
            j = idx[MOI.get(src, MOI.ConstraintFunction(), ci)].value
            lo, hi = _bounds(MOI.get(src, MOI.ConstraintSet(), ci))
            F.cl[j] = max(F.cl[j], lo); F.cu[j] = min(F.cu[j], hi)
            idx[ci] = MOI.ConstraintIndex{MOI.VariableIndex,S}(j)
        end
        for ci in MOI.get(src, MOI.ListOfConstraintIndices{_SAF,S}())
            f = MOI.get(src, MOI.ConstraintFunction(), ci)
            lo, hi = _bounds(MOI.get(src, MOI.ConstraintSet(), ci))
            push!(F.rl, lo - f.constant); push!(F.ru, hi - f.constant)
            r = length(F.rl)
            for t in f.terms; push!(F.a_row, r); push!(F.a_col, idx[t.variable].value); push!(F.a_val, t.coefficient); end
            idx[ci] = MOI.ConstraintIndex{_SAF,S}(r)
            push!(dest.rows, (idx[ci], r))
        end
    end
    dest.nv = nv; dest.nc = length(F.rl)
    _FLAT[dest] = F
    return idx
end

function MOI.optimize!(o::Optimizer)
    F = _FLAT[o]
    o.h == C_NULL && (o.h = create_handle(get(o.options, "device", 0)))
    key = hash((F.p_row, F.p_col, F.a_row, F.a_col, o.nv, o.nc))
    if key != o.pattern_key  # same pattern across SQP iterations -> device structure is kept
        check(o.h, ccall((:sqpqp_qp_setup, libsqpqp), Cint,
            (Ptr{Cvoid}, Int32, Int32, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}),
            o.h, o.nv, o.nc, length(F.p_row), F.p_row, F.p_col, length(F.a_row), F.a_row, F.a_col))
        o.pattern_key = key
    end
    o.x = zeros(o.nv); o.row_dual = zeros(max(o.nc, 1)); o.col_dual = zeros(o.nv)
    st = Ref{Int32}(0); info = Ref{Info}()
    check(o.h, ccall((:sqpqp_qp_solve, libsqpqp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}, Ref{Info}),
        o.h, F.p_val, F.q, F.a_val, F.rl, F.ru, F.cl, F.cu, o.x, o.row_dual, o.col_dual, st, info))
    o.status = st[]; o.objective = F.sense * info[].objective; o.solved = true
    return
end

_ok(o) = o.status in (1, 4, 7, 10)
MOI.get(o::Optimizer, ::MOI.TerminationStatus) = o.solved ? MOI.TerminationStatusCode(Int(o.status)) : MOI.OPTIMIZE_NOT_CALLED
MOI.get(o::Optimizer, ::MOI.RawStatusString) = string(MOI.get(o, MOI.TerminationStatus()))
MOI.get(o::Optimizer, ::MOI.ResultCount) = (o.solved && _ok(o)) ? 1 : 0
MOI.get(o::Optimizer, ::MOI.PrimalStatus) = _ok(o) ? MOI.FEASIBLE_POINT : MOI.NO_SOLUTION
MOI.get(o::Optimizer, ::MOI.DualStatus) = _ok(o) ? MOI.FEASIBLE_POINT : MOI.NO_SOLUTION
MOI.get(o::Optimizer, ::MOI.ObjectiveValue) = o.objective
MOI.get(o::Optimizer, ::MOI.VariablePrimal, vi::MOI.VariableIndex) = o.x[vi.value]
# engine duals are already in MOI sign for a minimisation (grad = A'lambda + r); flip for MAX_SENSE
MOI.get(o::Optimizer, ::MOI.ConstraintDual, ci::MOI.ConstraintIndex{_SAF,<:_SETS}) = _FLAT[o].sense * o.row_dual[ci.value]
function MOI.get(o::Optimizer, ::MOI.ConstraintDual, ci::MOI.ConstraintIndex{MOI.VariableIndex,S}) where {S<:_SETS}
    r = _FLAT[o].sense * o.col_dual[ci.value]   # JuMP.reduced_cost sums the bound duals of a variable
    S <: MOI.GreaterThan && return max(r, 0.0)
    S <: MOI.LessThan && return min(r, 0.0)
    return r
end

end # module
