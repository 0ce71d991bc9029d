# EasySIMPB200.jl -- the reference-side binding: how EasySIMP.jl reaches libeasysimp_b200.so through `ccall`.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain (probed: `which julia` -> none), so this
# file is the binding a maintainer adds to EasySIMP.jl, written against include/easysimp_b200.h and mirrored 1:1 by the
# ctypes binding that IS exercised on the GPU (easysimp_b200/lib.py + easysimp_b200/host.py, tests/test_gpu_*.py).
# tests/golden/make_julia_golden.jl runs the UNMODIFIED reference and this binding side by side on a machine that has
# Julia and a B200 and writes the vectors tests/test_julia_golden.py consumes.
#
# Drop-in scope (BASELINE.json north_star): `simp_optimize`, `OptimizationParameters`, `FilterCache` / `create_filter_cache`
# and the `assemble_stiffness_matrix_simp!` entry points keep their names and argument meaning; `MeshImport` and
# `PostProcessing` are untouched.  The matrix-free operator handle `B200Operator` takes the place of the
# `SparseMatrixCSC` K everywhere the user scripts touch it (README.md:110-153):
#
#     include("EasySIMPB200.jl"); using .EasySIMPB200
#     dh, cellvalues, K, f = setup_problem_b200(grid)                       # K::B200Operator, never allocate_matrix(dh)
#     assemble_stiffness_matrix_simp!(K, f, dh, cellvalues, material_model, fill(0.4, getncells(grid)))
#     ch = apply_fixed_boundary!(K, f, dh, fixed_nodes)                     # same ConstraintHandler, no apply!(K, f, ch)
#     results = simp_optimize(grid, dh, cellvalues, [PointLoad(dh, nodes, F)], [ch], params; operator = K)
#     results_data = create_results_data(grid, dh, cellvalues, results)     # unchanged PostProcessing
#
# or, to swap the backend globally, define in src/Optimization/Optimization.jl
#     simp_optimize(grid, dh, cv, loads, bcs, params::OptimizationParameters, acc = nothing) =
#         EasySIMPB200.simp_optimize(grid, dh, cv, loads, bcs, params, acc)
module EasySIMPB200

using Ferrite
using Tensors
using Printf
using LinearAlgebra
using ..FiniteElementAnalysis: apply_load_condition!, calculate_element_volumes, create_volume_quadrature,
                               create_simp_material_model
using ..Optimization: OptimizationParameters, OptimizationResult, FilterCache, create_filter_cache,
                      OptimizationLogger, log_iteration!, write_summary, close_logger
using ..PostProcessing: create_results_data, export_results_vtu
import ..PostProcessing

const LIB = get(ENV, "ESIMP_LIB", joinpath(@__DIR__, "..", "libeasysimp_b200.so"))

# esimp_params (include/easysimp_b200.h:69-85) -- field order and types must match the C struct
struct EsimpParams
    E0::Cdouble; Emin::Cdouble; nu::Cdouble; p::Cdouble
    volume_fraction::Cdouble; move_limit::Cdouble; damping::Cdouble; filter_radius::Cdouble
    filter_type::Int32; pcg_maxit::Int32
    pcg_rtol::Cdouble
    pcg_check_every::Int32; matvec_variant::Int32; warm_start::Int32; precond::Int32
end

# esimp_iter_stats (include/easysimp_b200.h:89-102)
struct EsimpIterStats
    energy::Cdouble; volume::Cdouble; change::Cdouble; lambda::Cdouble; grayness::Cdouble; max_u::Cdouble
    pcg_relres::Cdouble
    pcg_iters::Int32; oc_steps::Int32; status::Int32; reserved::Int32
    ms_filter::Cdouble; ms_solve::Cdouble; ms_sens::Cdouble; ms_oc::Cdouble; ms_total::Cdouble
end

mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(device::Integer = -1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:esimp_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device)
        st < 0 && error("esimp_create: " * unsafe_string(ccall((:esimp_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(r[])
        finalizer(c -> (c.ptr != C_NULL && ccall((:esimp_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL), ctx)
        return ctx
    end
end

function check(ctx::Context, st::Integer, what)
    st < 0 && error("$what failed ($st): " * unsafe_string(ccall((:esimp_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.ptr)))
    return st
end

"""
    B200Operator

Matrix-free stand-in for the `SparseMatrixCSC` K of the reference (`setup_problem`, FiniteElementAnalysis.jl:130-157):
holds the library context plus what the entry points that used to mutate K have to remember (the prescribed dofs of every
`apply_*_boundary!` call).  `op * u` applies P K(rho) P + m (I - P).
"""
mutable struct B200Operator
    ctx::Context
    grid::Grid
    dh::DofHandler
    structured::Union{Nothing,NTuple{6,Float64}}   # (nx, ny, nz, hx, hy, hz) of a generate_grid layout
    prescribed::Vector{Int}                        # union of ch.prescribed_dofs (1-based), in call order
    mesh_set::Bool
end

"Detect Ferrite's generate_grid(Hexahedron, (nx,ny,nz), left, right) layout (SURVEY Appendix A.1); nothing otherwise."
function structured_dims(grid::Grid)
    getcells(grid, 1) isa Ferrite.Hexahedron || return nothing
    x = [n.x for n in grid.nodes]
    x0 = x[1]
    nx = findfirst(i -> abs(x[i+1][2] - x0[2]) > 0 || abs(x[i+1][3] - x0[3]) > 0, 1:length(x)-1)
    nx === nothing && return nothing
    nxn = nx                                   # nodes per x row
    nyn = findfirst(j -> abs(x[j*nxn+1][3] - x0[3]) > 0, 1:(length(x) ÷ nxn - 1))
    nyn === nothing && return nothing
    nzn = length(x) ÷ (nxn * nyn)
    (nxn * nyn * nzn == length(x) && getncells(grid) == (nxn - 1) * (nyn - 1) * (nzn - 1)) || return nothing
    h = (x[2][1] - x0[1], x[nxn+1][2] - x0[2], x[nxn*nyn+1][3] - x0[3])
    # verify node positions and the first cell's VTK vertex order
    for k in 0:nzn-1, j in 0:nyn-1, i in 0:nxn-1
        p = x[1+i+nxn*(j+nyn*k)]
        (abs(p[1] - x0[1] - i * h[1]) < 1e-9 * abs(h[1]) + 1e-12 && abs(p[2] - x0[2] - j * h[2]) < 1e-9 * abs(h[2]) + 1e-12 &&
         abs(p[3] - x0[3] - k * h[3]) < 1e-9 * abs(h[3]) + 1e-12) || return nothing
    end
    collect(getcells(grid, 1).nodes) == [1, 2, nxn + 2, nxn + 1, nxn * nyn + 1, nxn * nyn + 2, nxn * nyn + nxn + 2, nxn * nyn + nxn + 1] || return nothing
    return (Float64(nxn - 1), Float64(nyn - 1), Float64(nzn - 1), h[1], h[2], h[3])
end

"dof_of_node (3 x n_nodes, 0-based) from the DofHandler: what get_node_dofs (FiniteElementAnalysis.jl:257-279) builds."
function dof_of_node(dh::DofHandler)
    grid = dh.grid
    out = fill(Int64(-1), 3, getnnodes(grid))
    for cell in CellIterator(dh)
        dofs = celldofs(cell)
        for (a, n) in enumerate(getnodes(cell))
            out[1, n] < 0 && (out[:, n] .= dofs[3a-2:3a] .- 1)
        end
    end
    return out
end

"Upload the mesh once (idempotent): structured grids by (nx, ny, nz, h) + Ferrite's dof numbering, imported meshes as they are."
function ensure_mesh!(op::B200Operator)
    op.mesh_set && return op
    ctx, grid, dh = op.ctx, op.grid, op.dh
    sd = op.structured
    if sd !== nothing
        don = dof_of_node(dh)
        check(ctx, ccall((:esimp_set_grid_structured, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cdouble, Cdouble, Ptr{Int64}),
                         ctx.ptr, Int(sd[1]), Int(sd[2]), Int(sd[3]), sd[4], sd[5], sd[6], don), "esimp_set_grid_structured")
    else
        npe = length(getcells(grid, 1).nodes)
        xyz = reduce(hcat, [collect(n.x) for n in grid.nodes])                         # 3 x n_nodes
        conn = Int64.(reduce(hcat, [collect(c.nodes) for c in getcells(grid)])) .- 1   # npe x n_cells
        cd = Int64.(reduce(hcat, [celldofs(dh, i) for i in 1:getncells(grid)])) .- 1   # 3npe x n_cells
        check(ctx, ccall((:esimp_set_mesh_unstructured, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}),
                         ctx.ptr, npe, getncells(grid), getnnodes(grid), xyz, conn, cd), "esimp_set_mesh_unstructured")
    end
    op.mesh_set = true
    return op
end

"""
    setup_problem_b200(grid, interpolation_order = 1; device = -1) -> dh, cellvalues, K::B200Operator, f

`setup_problem` (FiniteElementAnalysis.jl:130-157) without `allocate_matrix(dh)`: at 100 M dofs K cannot exist (~100 GB).
"""
function setup_problem_b200(grid::Grid, interpolation_order::Int = 1; device::Integer = -1)
    dim = 3
    if getcells(grid, 1) isa Ferrite.Hexahedron
        ip = Lagrange{RefHexahedron,interpolation_order}()^dim
        qr = QuadratureRule{RefHexahedron}(2)
    else
        ip = Lagrange{RefTetrahedron,interpolation_order}()^dim
        qr = QuadratureRule{RefTetrahedron}(2)
    end
    cellvalues = CellValues(qr, ip)
    dh = DofHandler(grid)
    add!(dh, :u, ip)
    close!(dh)
    op = B200Operator(Context(device), grid, dh, structured_dims(grid), Int[], false)
    return dh, cellvalues, op, zeros(ndofs(dh))
end

"Wrap an existing (grid, dh) pair -- for scripts that already called the reference's `setup_problem`."
B200Operator(grid::Grid, dh::DofHandler; device::Integer = -1) = B200Operator(Context(device), grid, dh, structured_dims(grid), Int[], false)

# ---- the boundary-condition entry points on the operator handle (FiniteElementAnalysis.jl:293-340).  The ConstraintHandler
# is built exactly as in the reference; instead of `apply!(K, f, ch)` the prescribed dofs are remembered (the library applies
# P K P + m (I - P) with m = mean |diag K|, Ferrite's apply! semantics) and f is zeroed there (homogeneous Dirichlet).
function _constrain!(op::B200Operator, f, dh, nodes, components)
    ch = ConstraintHandler(dh)
    for d in components
        add!(ch, Dirichlet(:u, nodes, (x, t) -> 0.0, d))
    end
    close!(ch)
    update!(ch, 0.0)
    append!(op.prescribed, ch.prescribed_dofs)
    f[ch.prescribed_dofs] .= 0.0
    return ch
end
function apply_fixed_boundary!(op::B200Operator, f, dh, nodes)
    ch = _constrain!(op, f, dh, nodes, 1:3)
    println("Applied fixed boundary conditions to $(length(nodes)) nodes")
    return ch
end
function apply_sliding_boundary!(op::B200Operator, f, dh, nodes, fixed_dofs)
    ch = _constrain!(op, f, dh, nodes, fixed_dofs)
    println("Applied sliding boundary to $(length(nodes)) nodes, fixed DOFs: $fixed_dofs")
    return ch
end

"""
    assemble_stiffness_matrix_simp!(op::B200Operator, f, dh, cellvalues, material_model, density_data, args...)

Matrix-free method of the reference's assembly entry points (FiniteElementAnalysis.jl:229-243 six arguments, :637-671 nine):
`op` takes the place of K; the densities are uploaded, E_e and the Jacobi diagonal are formed on the device, f is zeroed
like `start_assemble(K, f)` does.  Material constants come from `set_material!` (default: OptimizationParameters defaults).
"""
function assemble_stiffness_matrix_simp!(op::B200Operator, f, dh, cellvalues, material_model, density_data, args...)
    ensure_mesh!(op)
    fill!(f, 0.0)
    pres = Int64.(unique(op.prescribed)) .- 1
    check(op.ctx, ccall((:esimp_set_dirichlet, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), op.ctx.ptr, pres, length(pres)), "esimp_set_dirichlet")
    check(op.ctx, ccall((:esimp_set_density, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), op.ctx.ptr, Vector{Float64}(density_data)), "esimp_set_density")
    return op
end

function Base.:*(op::B200Operator, u::Vector{Float64})
    y = similar(u)
    check(op.ctx, ccall((:esimp_matvec, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), op.ctx.ptr, u, y), "esimp_matvec")
    return y
end

"characteristic element size exactly as FilterCommon.jl:109-199 takes it: mean over the first <= 10 cells"
function characteristic_size(grid::Grid)
    n = min(10, getncells(grid))
    s = 0.0
    for i in 1:n
        cell = getcells(grid, i)
        x = [grid.nodes[j].x for j in cell.nodes]
        if cell isa Ferrite.Hexahedron
            s += cbrt(norm(x[2] - x[1]) * norm(x[4] - x[1]) * norm(x[5] - x[1]))
        else
            e = [norm(x[a] - x[b]) for (a, b) in ((1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4))]
            s += sum(e) / length(e)
        end
    end
    return s / n
end

function set_params!(ctx::Context, params::OptimizationParameters, R::Float64, structured::Bool; pcg_rtol, pcg_maxit, pcg_check_every, preconditioner)
    precond = preconditioner == :jacobi ? 0 : preconditioner == :gmg ? 1 : (structured ? 1 : 0)   # :auto
    prm = EsimpParams(params.E0, params.Emin, params.ν, params.p, params.volume_fraction, params.move_limit, params.damping, R,
                      params.filter_type == :density ? 1 : 0, pcg_maxit, pcg_rtol, pcg_check_every, 0, 1, precond)
    check(ctx, ccall((:esimp_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{EsimpParams}), ctx.ptr, prm), "esimp_set_params")
end

"Mesh + BCs + load + filter into the context.  Everything density independent is uploaded once."
function configure!(op::B200Operator, loads, boundary_conditions, params::OptimizationParameters; pcg_rtol, pcg_maxit, pcg_check_every, preconditioner)
    ctx, grid, dh = op.ctx, op.grid, op.dh
    sd = op.structured
    R = params.filter_radius * characteristic_size(grid)          # FilterCommon.jl:72-73
    set_params!(ctx, params, R, sd !== nothing; pcg_rtol, pcg_maxit, pcg_check_every, preconditioner)
    ensure_mesh!(op)
    if sd !== nothing
        # structured fast path: the library builds the stencil from (h, R); no KD-tree, no (n_cells, 8, 3) gather for the volumes
        check(ctx, ccall((:esimp_filter_build_structured, LIB), Cint, (Ptr{Cvoid},), ctx.ptr), "esimp_filter_build_structured")
    else
        volumes = calculate_element_volumes(grid, create_volume_quadrature(grid))
        fc = create_filter_cache(grid, params.filter_radius, volumes)
        rowptr = Int64[0; cumsum(length.(fc.neighbor_lists))]
        col = Int64.(reduce(vcat, fc.neighbor_lists)) .- 1
        centers = reduce(hcat, [collect(c) for c in fc.cell_centers])
        check(ctx, ccall((:esimp_filter_build_csr, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
                         ctx.ptr, rowptr, col, centers, fc.element_volumes, fc.filter_radius), "esimp_filter_build_csr")
    end
    bc_dofs = reduce(vcat, [ch.prescribed_dofs for ch in boundary_conditions]; init = Int[])
    prescribed = Int64.(unique(vcat(op.prescribed, bc_dofs))) .- 1
    check(ctx, ccall((:esimp_set_dirichlet, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), ctx.ptr, prescribed, length(prescribed)),
          "esimp_set_dirichlet")
    f = zeros(ndofs(dh))
    foreach(l -> apply_load_condition!(f, l), loads)      # density independent: assembled ONCE (reference: every iteration)
    check(ctx, ccall((:esimp_set_load, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, f), "esimp_set_load")
    return op
end

function get_fields(ctx::Context, n_cells::Int, n_dofs::Int)
    rho = zeros(n_cells); u = zeros(n_dofs)
    check(ctx, ccall((:esimp_get_physical_density, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, rho), "esimp_get_physical_density")
    check(ctx, ccall((:esimp_get_displacement, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, u), "esimp_get_displacement")
    return rho, u
end

"""
    stress_dict_b200(ctx, n_cells; chunk = 65536) -> Dict{Int,Vector{SymmetricTensor{2,3,Float64}}}

The `stress_field` of `calculate_stresses_simp` (FiniteElementAnalysis.jl:566-597) for the displacement and physical
densities the context holds: the type `OptimizationResult.stresses::Dict` and the unchanged PostProcessing
(`calculate_von_mises_stresses`, `calculate_element_energy`) expect.  Filled from `esimp_stress_tensors` chunk by chunk.
"""
function stress_dict_b200(ctx::Context, n_cells::Int; chunk::Int = 65536)
    nq = Int(ccall((:esimp_stress_nq, LIB), Cint, (Ptr{Cvoid},), ctx.ptr))
    out = Dict{Int,Vector{SymmetricTensor{2,3,Float64}}}()
    sizehint!(out, n_cells)
    sig = zeros(6, nq, min(chunk, n_cells))   # column-major view of the library's [cell][q][6], 6 = (11, 21, 31, 22, 32, 33)
    c0 = 0
    while c0 < n_cells
        n = min(chunk, n_cells - c0)
        check(ctx, ccall((:esimp_stress_tensors, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), ctx.ptr, c0, n, sig), "esimp_stress_tensors")
        for i in 1:n
            out[c0+i] = [SymmetricTensor{2,3,Float64}((sig[1, q, i], sig[2, q, i], sig[3, q, i], sig[4, q, i], sig[5, q, i], sig[6, q, i])) for q in 1:nq]
        end
        c0 += n
    end
    return out
end

"""
    stress_fields_b200(ctx) -> (von_mises, von_mises_avg, element_energy, max_von_mises, max_cell)

Cell arrays instead of the Dict, for meshes where 8 tensors per cell on the host are too much (33 M cells: 13 GB):
`von_mises` is the exported field (first Gauss point, PostProcessing.jl:204-222), `max_von_mises` / `max_cell` (1-based,
0 = none) are those of the Gauss-point average (FiniteElementAnalysis.jl:583-590).
"""
function stress_fields_b200(ctx::Context)
    n = Int(ccall((:esimp_n_cells, LIB), Int64, (Ptr{Cvoid},), ctx.ptr))
    vm, vma, ee = zeros(n), zeros(n), zeros(n)
    mx, cell = Ref{Cdouble}(0), Ref{Int64}(-1)
    check(ctx, ccall((:esimp_stress_fields, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Int64}),
                     ctx.ptr, vm, vma, ee, mx, cell), "esimp_stress_fields")
    return (von_mises = vm, von_mises_avg = vma, element_energy = ee, max_von_mises = mx[], max_cell = cell[] + 1)
end

"""
    simp_optimize(grid, dh, cellvalues, loads, boundary_conditions, params, acceleration_data = nothing;
                  operator = nothing, device = -1, pcg_rtol = 1e-8, pcg_maxit = 200000, pcg_check_every = 25,
                  preconditioner = :auto, stress_dict_max_cells = 2_000_000)

Same contract as the reference's `simp_optimize` (Optimization.jl:178-565): same arguments, same console lines, same logger
CSV / summary, same `tolerance_checkpoints` and `export_interval` exports, same `OptimizationResult`.  The loop body
(:273-379) is ONE ccall per iteration; density, displacement and sensitivities stay on the GPU and come back only when an
export branch or the final result needs them.  `acceleration_data` (:300-311) becomes one `esimp_set_body_force` call.
Above `stress_dict_max_cells` cells the result carries an empty `stresses` Dict (use `stress_fields_b200`).
"""
function simp_optimize(grid::Grid, dh::DofHandler, cellvalues, loads, boundary_conditions,
                       params::OptimizationParameters, acceleration_data = nothing;
                       operator::Union{Nothing,B200Operator} = nothing, device::Integer = -1, pcg_rtol::Float64 = 1e-8,
                       pcg_maxit::Integer = 200000, pcg_check_every::Integer = 25, preconditioner::Symbol = :auto,
                       stress_dict_max_cells::Integer = 2_000_000)
    println("Starting SIMP topology optimization (B200 backend)")
    logger = isempty(params.export_path) ? nothing : OptimizationLogger(params.export_path, params.task_name)
    op = operator === nothing ? B200Operator(grid, dh; device = device) : operator
    ctx = op.ctx
    configure!(op, loads, boundary_conditions, params; pcg_rtol, pcg_maxit = Int32(pcg_maxit), pcg_check_every = Int32(pcg_check_every), preconditioner)
    n_cells, n_dofs = getncells(grid), ndofs(dh)
    total_volume = op.structured !== nothing ? n_cells * op.structured[4] * op.structured[5] * op.structured[6] :
                   sum(calculate_element_volumes(grid, create_volume_quadrature(grid)))
    if acceleration_data !== nothing   # (acceleration_vector, base_density): f += rho_e * base * b * int N on every assembly
        acc, base = acceleration_data
        println("Variable density acceleration enabled: $(acc)")
        check(ctx, ccall((:esimp_set_body_force, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble), ctx.ptr,
                         Float64[acc[1], acc[2], acc[3]], Float64(base)), "esimp_set_body_force")
    end
    check(ctx, ccall((:esimp_init_design, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, C_NULL), "esimp_init_design")
    energy_history, volume_history = Float64[], Float64[]
    checkpoint_triggered = falses(length(params.tolerance_checkpoints))
    converged, iteration = false, 0
    stats = Ref{EsimpIterStats}()

    # an OptimizationResult of the state the device holds right now (export branches, :407-477)
    function snapshot(energy, volume, iter)
        rho, u = get_fields(ctx, n_cells, n_dofs)
        stresses = n_cells <= stress_dict_max_cells ? stress_dict_b200(ctx, n_cells) : Dict{Int,Vector{SymmetricTensor{2,3,Float64}}}()
        return OptimizationResult(rho, u, stresses, energy, volume, iter, false, copy(energy_history), copy(volume_history))
    end

    for iter in 1:params.max_iterations
        iteration = iter
        check(ctx, ccall((:esimp_iterate, LIB), Cint, (Ptr{Cvoid}, Ref{EsimpIterStats}), ctx.ptr, stats), "esimp_iterate")
        s = stats[]
        push!(energy_history, s.energy); push!(volume_history, s.volume)
        vf = s.volume / total_volume
        logger !== nothing && log_iteration!(logger, iter, s.energy, vf, s.change, s.lambda, s.grayness, s.max_u)     # :382-394
        @printf("Iter %4d | Energy: %.4e | Vol.Frac: %.4f | Change: %.4e\n", iter, s.energy, vf, s.change)            # :396-402
        if !isempty(params.tolerance_checkpoints) && !isempty(params.export_path)                                     # :407-445
            for (idx, cp) in enumerate(params.tolerance_checkpoints)
                if !checkpoint_triggered[idx] && s.change < cp
                    checkpoint_triggered[idx] = true
                    tol_str = @sprintf("%02d", round(Int, cp * 100))
                    println("Tolerance checkpoint $(cp) reached at iteration $(iter)")
                    cp_results_data = create_results_data(grid, dh, cellvalues, snapshot(s.energy, s.volume, iter))
                    PostProcessing.export_main_results(cp_results_data, joinpath(params.export_path, "final_results_$(tol_str)tol"))
                    println("Checkpoint exported: final_results_$(tol_str)tol.vtu")
                end
            end
        end
        if params.export_interval > 0 && iter % params.export_interval == 0 && !isempty(params.export_path)         # :448-477
            results_data = create_results_data(grid, dh, cellvalues, snapshot(s.energy, s.volume, iter))
            export_results_vtu(results_data, joinpath(params.export_path, "iter_$(lpad(iter, 4, '0'))"), include_history = false)
        end
        if s.change < params.tolerance                                                                                # :484-488
            converged = true
            println("Converged at iteration $(iter) with change $(s.change)")
            break
        end
    end
    # final analysis (:494-535): re-filter, re-assemble, re-solve on the device
    e, v, its = Ref{Cdouble}(0), Ref{Cdouble}(0), Ref{Cint}(0)
    check(ctx, ccall((:esimp_final_analysis, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cint}), ctx.ptr, e, v, its),
          "esimp_final_analysis")
    rho, u = get_fields(ctx, n_cells, n_dofs)
    sf = stress_fields_b200(ctx)   # :537-539 on the device state the final analysis left
    println("Maximum von Mises stress: $(sf.max_von_mises) at cell $(sf.max_cell)")
    stresses = n_cells <= stress_dict_max_cells ? stress_dict_b200(ctx, n_cells) : Dict{Int,Vector{SymmetricTensor{2,3,Float64}}}()
    if logger !== nothing
        write_summary(logger, e[], v[], converged)
        close_logger(logger)
    end
    println("Optimization completed")
    println("Final energy: $(e[])")
    println("Final volume fraction: $(v[] / total_volume)")
    return OptimizationResult(rho, u, stresses, e[], v[], iteration, converged, energy_history, volume_history)
end

# ---- N GPUs from this one Julia session (include/easysimp_b200.h, esimp_multi_*; SURVEY 8b) --------------------------------
# The reference is a single process (Optimization.jl:178).  An esimp_multi owns one context per device, the NCCL communicator
# and the x-slab partition; every call takes the same GLOBAL arrays as the one-GPU path and fans out to one host thread per GPU
# inside the library, so no MPI launcher is needed.  Structured generate_grid meshes only (imported meshes run as replicas).
mutable struct MultiContext
    ptr::Ptr{Cvoid}
    function MultiContext(devices::AbstractVector{<:Integer})
        r = Ref{Ptr{Cvoid}}(C_NULL)
        dev = Cint.(collect(devices))
        st = ccall((:esimp_multi_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), r, dev, length(dev))
        st < 0 && error("esimp_multi_create: " * unsafe_string(ccall((:esimp_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(r[])
        finalizer(c -> (c.ptr != C_NULL && ccall((:esimp_multi_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL), m)
        return m
    end
end
function check(m::MultiContext, st::Integer, what)
    st < 0 && error("$what failed ($st): " * unsafe_string(ccall((:esimp_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), m.ptr)))
    return st
end

"""
    simp_optimize_multi(grid, dh, cellvalues, loads, boundary_conditions, params, acceleration_data = nothing;
                        devices = 0:7, pcg_rtol = 1e-8, pcg_maxit = 200000)

`simp_optimize` (Optimization.jl:178-565) on several GPUs of one node: same arguments, console lines, logger and
`OptimizationResult`; the loop body is one `esimp_multi_iterate` per iteration.  The `stresses` Dict of the result is empty and
the per-iteration export branches are not wired (use the one-GPU entry point for those, or `esimp_multi_rank_ctx` + the stage calls).
"""
function simp_optimize_multi(grid::Grid, dh::DofHandler, cellvalues, loads, boundary_conditions, params::OptimizationParameters,
                             acceleration_data = nothing; devices = 0:7, pcg_rtol::Float64 = 1e-8, pcg_maxit::Integer = 200000)
    sd = structured_dims(grid)
    sd === nothing && error("simp_optimize_multi needs a generate_grid hexahedral mesh (imported meshes are not partitioned)")
    println("Starting SIMP topology optimization (B200 backend, $(length(devices)) GPUs)")
    logger = isempty(params.export_path) ? nothing : OptimizationLogger(params.export_path, params.task_name)
    m = MultiContext(devices)
    R = params.filter_radius * characteristic_size(grid)
    prm = EsimpParams(params.E0, params.Emin, params.ν, params.p, params.volume_fraction, params.move_limit, params.damping, R,
                      params.filter_type == :density ? 1 : 0, Int32(pcg_maxit), pcg_rtol, Int32(25), 0, 1, 1)   # multigrid-preconditioned CG
    check(m, ccall((:esimp_multi_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{EsimpParams}), m.ptr, prm), "esimp_multi_set_params")
    check(m, ccall((:esimp_multi_set_grid_structured, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cdouble, Cdouble, Ptr{Int64}),
                   m.ptr, Int(sd[1]), Int(sd[2]), Int(sd[3]), sd[4], sd[5], sd[6], dof_of_node(dh)), "esimp_multi_set_grid_structured")
    check(m, ccall((:esimp_multi_filter_build_structured, LIB), Cint, (Ptr{Cvoid},), m.ptr), "esimp_multi_filter_build_structured")
    prescribed = Int64.(unique(reduce(vcat, [ch.prescribed_dofs for ch in boundary_conditions]; init = Int[]))) .- 1
    check(m, ccall((:esimp_multi_set_dirichlet, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.ptr, prescribed, length(prescribed)),
          "esimp_multi_set_dirichlet")
    f = zeros(ndofs(dh))
    foreach(l -> apply_load_condition!(f, l), loads)
    check(m, ccall((:esimp_multi_set_load, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), m.ptr, f), "esimp_multi_set_load")
    if acceleration_data !== nothing
        acc, base = acceleration_data
        check(m, ccall((:esimp_multi_set_body_force, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble), m.ptr,
                       Float64[acc[1], acc[2], acc[3]], Float64(base)), "esimp_multi_set_body_force")
    end
    check(m, ccall((:esimp_multi_init_design, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), m.ptr, C_NULL), "esimp_multi_init_design")
    n_cells, n_dofs = getncells(grid), ndofs(dh)
    total_volume = n_cells * sd[4] * sd[5] * sd[6]
    energy_history, volume_history = Float64[], Float64[]
    converged, iteration = false, 0
    stats = Ref{EsimpIterStats}()
    for iter in 1:params.max_iterations
        iteration = iter
        check(m, ccall((:esimp_multi_iterate, LIB), Cint, (Ptr{Cvoid}, Ref{EsimpIterStats}), m.ptr, stats), "esimp_multi_iterate")
        s = stats[]
        push!(energy_history, s.energy); push!(volume_history, s.volume)
        vf = s.volume / total_volume
        logger !== nothing && log_iteration!(logger, iter, s.energy, vf, s.change, s.lambda, s.grayness, s.max_u)
        @printf("Iter %4d | Energy: %.4e | Vol.Frac: %.4f | Change: %.4e\n", iter, s.energy, vf, s.change)
        if s.change < params.tolerance
            converged = true
            println("Converged at iteration $(iter) with change $(s.change)")
            break
        end
    end
    e, v, its = Ref{Cdouble}(0), Ref{Cdouble}(0), Ref{Cint}(0)
    check(m, ccall((:esimp_multi_final_analysis, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cint}), m.ptr, e, v, its),
          "esimp_multi_final_analysis")
    rho = zeros(n_cells); u = zeros(n_dofs)
    check(m, ccall((:esimp_multi_get_physical_density, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), m.ptr, rho), "esimp_multi_get_physical_density")
    check(m, ccall((:esimp_multi_get_displacement, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), m.ptr, u), "esimp_multi_get_displacement")
    if logger !== nothing
        write_summary(logger, e[], v[], converged)
        close_logger(logger)
    end
    println("Optimization completed")
    println("Final energy: $(e[])")
    println("Final volume fraction: $(v[] / total_volume)")
    return OptimizationResult(rho, u, Dict{Int,Vector{SymmetricTensor{2,3,Float64}}}(), e[], v[], iteration, converged, energy_history, volume_history)
end

"the round-1 name of the entry point"
simp_optimize_b200(args...; kwargs...) = simp_optimize(args...; kwargs...)

export Context, B200Operator, setup_problem_b200, apply_fixed_boundary!, apply_sliding_boundary!, assemble_stiffness_matrix_simp!,
       simp_optimize, simp_optimize_b200, simp_optimize_multi, MultiContext, stress_fields_b200, stress_dict_b200

end # module
