# EasySIMPB200.jl -- the reference-side binding: how EasySIMP.jl reaches libeasysimp_b200.so through `ccall`.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain (probed: `which julia` -> none),
# so this file is the binding a maintainer adds to EasySIMP.jl, written against include/easysimp_b200.h and
# mirrored 1:1 by the ctypes binding that IS exercised (easysimp_b200/lib.py, tests/test_gpu_*.py).
#
# Drop-in scope (BASELINE.json north_star): `simp_optimize`, `OptimizationParameters`, `FilterCache` /
# `create_filter_cache` and `assemble_stiffness_matrix_simp!` keep their signatures; `MeshImport` and
# `PostProcessing` are untouched.  Usage inside EasySIMP.jl:
#
#     include("EasySIMPB200.jl"); using .EasySIMPB200
#     results = simp_optimize_b200(grid, dh, cellvalues, loads, boundary_conditions, params)
#
# or, to swap the backend globally, define in src/Optimization/Optimization.jl
#     simp_optimize(grid, dh, cv, loads, bcs, params::OptimizationParameters, acc = nothing) =
#         EasySIMPB200.simp_optimize_b200(grid, dh, cv, loads, bcs, params, acc)
module EasySIMPB200

using Ferrite
using ..FiniteElementAnalysis: apply_load_condition!, calculate_element_volumes, create_volume_quadrature,
                               calculate_stresses_simp, create_simp_material_model
using ..Optimization: OptimizationParameters, OptimizationResult, FilterCache, create_filter_cache

const LIB = get(ENV, "ESIMP_LIB", joinpath(@__DIR__, "..", "libeasysimp_b200.so"))

# esimp_params (include/easysimp_b200.h) -- field order and types must match the C struct
struct EsimpParams
    E0::Cdouble; Emin::Cdouble; nu::Cdouble; p::Cdouble
    volume_fraction::Cdouble; move_limit::Cdouble; damping::Cdouble; filter_radius::Cdouble
    filter_type::Int32; pcg_maxit::Int32
    pcg_rtol::Cdouble
    pcg_check_every::Int32; matvec_variant::Int32; warm_start::Int32; precond::Int32
end

# esimp_iter_stats
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
        finalizer(c -> (ccall((:esimp_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL), ctx)
        return ctx
    end
end

function check(ctx::Context, st::Integer, what)
    st < 0 && error("$what failed ($st): " * unsafe_string(ccall((:esimp_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.ptr)))
    return st
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
    nxn * nyn * nzn == length(x) && getncells(grid) == (nxn - 1) * (nyn - 1) * (nzn - 1) || return nothing
    h = (x[2][1] - x0[1], x[nxn+1][2] - x0[2], x[nxn*nyn+1][3] - x0[3])
    # verify node positions and the first cell's VTK vertex order
    for k in 0:nzn-1, j in 0:nyn-1, i in 0:nxn-1
        p = x[1+i+nxn*(j+nyn*k)]
        (abs(p[1] - x0[1] - i * h[1]) < 1e-9 * abs(h[1]) + 1e-12 && abs(p[2] - x0[2] - j * h[2]) < 1e-9 * abs(h[2]) + 1e-12 &&
         abs(p[3] - x0[3] - k * h[3]) < 1e-9 * abs(h[3]) + 1e-12) || return nothing
    end
    collect(getcells(grid, 1).nodes) == [1, 2, nxn + 2, nxn + 1, nxn * nyn + 1, nxn * nyn + 2, nxn * nyn + nxn + 2, nxn * nyn + nxn + 1] || return nothing
    return (nxn - 1, nyn - 1, nzn - 1, h...)
end

"dof_of_node (3 x n_nodes, 0-based) from the DofHandler: what get_node_dofs (FiniteElementAnalysis.jl:257-279) builds."
function dof_of_node(dh::DofHandler)
    grid = dh.grid
    out = fill(Int64(-1), 3, getnnodes(grid))
    for cell in CellIterator(dh)
        dofs = celldofs(cell)
        for (a, n) in enumerate(cell.nodes)
            out[1, n] < 0 && (out[:, n] .= dofs[3a-2:3a] .- 1)
        end
    end
    return out
end

"Mesh + BCs + load + filter into a context.  Everything density independent is uploaded once."
function configure!(ctx::Context, grid, dh, loads, boundary_conditions, params::OptimizationParameters, fc::FilterCache)
    sd = structured_dims(grid)
    # last field: CG preconditioner, ESIMP_PRECOND_GMG (1) on generate_grid grids, ESIMP_PRECOND_JACOBI (0) on imported meshes
    prm = EsimpParams(params.E0, params.Emin, params.ν, params.p, params.volume_fraction, params.move_limit, params.damping,
                      fc.filter_radius, params.filter_type == :density ? 1 : 0, 200000, 1e-8, 25, 0, 1, sd !== nothing ? 1 : 0)
    check(ctx, ccall((:esimp_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{EsimpParams}), ctx.ptr, prm), "esimp_set_params")
    if sd !== nothing
        don = dof_of_node(dh)
        check(ctx, ccall((:esimp_set_grid_structured, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cdouble, Cdouble, Ptr{Int64}),
                         ctx.ptr, sd[1], sd[2], sd[3], sd[4], sd[5], sd[6], don), "esimp_set_grid_structured")
        check(ctx, ccall((:esimp_filter_build_structured, LIB), Cint, (Ptr{Cvoid},), ctx.ptr), "esimp_filter_build_structured")
    else
        npe = length(getcells(grid, 1).nodes)
        xyz = reduce(hcat, [collect(n.x) for n in grid.nodes])                         # 3 x n_nodes
        conn = Int64.(reduce(hcat, [collect(c.nodes) for c in getcells(grid)])) .- 1   # npe x n_cells
        cd = Int64.(reduce(hcat, [celldofs(dh, i) for i in 1:getncells(grid)])) .- 1   # 3npe x n_cells
        check(ctx, ccall((:esimp_set_mesh_unstructured, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}),
                         ctx.ptr, npe, getncells(grid), getnnodes(grid), xyz, conn, cd), "esimp_set_mesh_unstructured")
        rowptr = Int64[0; cumsum(length.(fc.neighbor_lists))]
        col = Int64.(reduce(vcat, fc.neighbor_lists)) .- 1
        centers = reduce(hcat, [collect(c) for c in fc.cell_centers])
        check(ctx, ccall((:esimp_filter_build_csr, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
                         ctx.ptr, rowptr, col, centers, fc.element_volumes, fc.filter_radius), "esimp_filter_build_csr")
    end
    prescribed = Int64.(unique(reduce(vcat, [ch.prescribed_dofs for ch in boundary_conditions]; init = Int[]))) .- 1
    check(ctx, ccall((:esimp_set_dirichlet, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), ctx.ptr, prescribed, length(prescribed)),
          "esimp_set_dirichlet")
    f = zeros(ndofs(dh))
    foreach(l -> apply_load_condition!(f, l), loads)      # density independent: assembled ONCE (reference: every iteration)
    check(ctx, ccall((:esimp_set_load, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, f), "esimp_set_load")
    return ctx
end

"""
    simp_optimize_b200(grid, dh, cellvalues, loads, boundary_conditions, params, acceleration_data = nothing)

Same contract as `simp_optimize` (Optimization.jl:178-565).  The loop body (:273-379) is ONE ccall per iteration;
density, displacement and sensitivities stay on the GPU.  `acceleration_data` (:300-311) becomes one
`esimp_set_body_force` call; export / checkpoint branches (:407-477) read the von Mises / element-energy cell arrays from
`stress_fields_b200` instead of walking a Dict of tensors.
"""
function simp_optimize_b200(grid::Grid, dh::DofHandler, cellvalues, loads, boundary_conditions,
                            params::OptimizationParameters, acceleration_data = nothing; device::Integer = -1)
    volumes = calculate_element_volumes(grid, create_volume_quadrature(grid))
    fc = create_filter_cache(grid, params.filter_radius, volumes)
    ctx = configure!(Context(device), grid, dh, loads, boundary_conditions, params, fc)
    if acceleration_data !== nothing   # (acceleration_vector, base_density): f += rho_e * base * b * int N on every assembly
        acc, base = acceleration_data
        check(ctx, ccall((:esimp_set_body_force, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble), ctx.ptr,
                         Float64[acc[1], acc[2], acc[3]], Float64(base)), "esimp_set_body_force")
    end
    check(ctx, ccall((:esimp_init_design, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, C_NULL), "esimp_init_design")
    energy_history, volume_history = Float64[], Float64[]
    converged, iteration = false, 0
    stats = Ref{EsimpIterStats}()
    for iter in 1:params.max_iterations
        iteration = iter
        check(ctx, ccall((:esimp_iterate, LIB), Cint, (Ptr{Cvoid}, Ref{EsimpIterStats}), ctx.ptr, stats), "esimp_iterate")
        s = stats[]
        push!(energy_history, s.energy); push!(volume_history, s.volume)
        # logger / @printf / tolerance_checkpoints / export_interval exactly as Optimization.jl:382-477, using
        # s.energy, s.volume, s.change, s.lambda, s.grayness, s.max_u and (only when exporting)
        # esimp_get_physical_density / esimp_get_displacement / stress_fields_b200(ctx)
        if s.change < params.tolerance
            converged = true
            break
        end
    end
    e, v, its = Ref{Cdouble}(0), Ref{Cdouble}(0), Ref{Cint}(0)
    check(ctx, ccall((:esimp_final_analysis, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cint}), ctx.ptr, e, v, its),
          "esimp_final_analysis")
    rho = zeros(getncells(grid)); u = zeros(ndofs(dh))
    check(ctx, ccall((:esimp_get_physical_density, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, rho), "esimp_get_physical_density")
    check(ctx, ccall((:esimp_get_displacement, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.ptr, u), "esimp_get_displacement")
    stress_field = stress_fields_b200(ctx)   # :537-539 on the device state the final analysis left
    return OptimizationResult(rho, u, stress_field, e[], v[], iteration, converged, energy_history, volume_history)
end

"""
    stress_fields_b200(ctx) -> (von_mises, von_mises_avg, element_energy, max_von_mises, max_cell)

`calculate_stresses_simp` (FEA.jl:566-597) + `calculate_von_mises_stresses` / `calculate_element_energy`
(PostProcessing.jl:172-222) for the displacement and physical densities the context holds.  `von_mises` is the exported
field (first Gauss point), `max_von_mises` / `max_cell` (1-based, 0 = none) are those of the Gauss-point average.
`stress_tensors_b200(ctx, cells)` returns the Dict entries themselves for a range of cells.
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

function stress_tensors_b200(ctx::Context, cells::UnitRange{Int})
    nq = Int(ccall((:esimp_stress_nq, LIB), Cint, (Ptr{Cvoid},), ctx.ptr))
    sig = zeros(6, nq, length(cells))   # column-major view of the library's [cell][q][6]
    check(ctx, ccall((:esimp_stress_tensors, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}),
                     ctx.ptr, first(cells) - 1, length(cells), sig), "esimp_stress_tensors")
    return Dict(c => [SymmetricTensor{2,3,Float64}(Tuple(sig[:, q, i])) for q in 1:nq] for (i, c) in enumerate(cells))
end

"""
    assemble_stiffness_matrix_simp!(op::Context, f, dh, cellvalues, material_model, density_data, args...)

Matrix-free method of the reference's assembly entry points (FEA.jl:229, :637): `op` takes the place of the
`SparseMatrixCSC` K; the densities are uploaded and E_e + the Jacobi diagonal are formed on the device.
`K * u` becomes `esimp_matvec`.
"""
function assemble_stiffness_matrix_simp!(op::Context, f, dh, cellvalues, material_model, density_data, args...)
    fill!(f, 0.0)
    check(op, ccall((:esimp_set_density, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), op.ptr, density_data), "esimp_set_density")
    return op
end

function Base.:*(op::Context, u::Vector{Float64})
    y = similar(u)
    check(op, ccall((:esimp_matvec, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), op.ptr, u, y), "esimp_matvec")
    return y
end

export Context, simp_optimize_b200, assemble_stiffness_matrix_simp!, stress_fields_b200, stress_tensors_b200

end # module
