# make_julia_golden.jl -- reference-held parity vectors.  Run on a machine that has Julia and the reference's
# dependencies (none of which exist in the build image, which is why the repository's parity status is "unpinned"):
#
#     cd /path/to/EasySIMP.jl && julia --project=. /path/to/repo/tests/golden/make_julia_golden.jl /path/to/repo/tests/golden
#
# It runs the UNMODIFIED reference (no file of this repository is loaded) on the configurations the repository's own
# golden fixtures use -- the enabled smoke test of test/runtests.jl:16-106, BASELINE config 1 (README cantilever, 50
# iterations) and a short run of the 2x1x1 four-legs problem of test/Examples/05_3D_2x1x1_4Legs.jl:24-94 with the density
# filter -- and writes tests/golden/julia_<name>.json.  tests/test_julia_golden.py then holds BOTH the CPU oracle
# (-m "not gpu") and the CUDA path (-m gpu) against those numbers with the north_star tolerances (element sensitivities
# and filter 1e-10, per-iteration compliance 1e-6, final density L_inf 1e-3).  Dependency-free output (hand-written JSON).
using EasySIMP
using EasySIMP.FiniteElementAnalysis
using EasySIMP.Optimization
using Ferrite
using LinearAlgebra
using SparseArrays

outdir = length(ARGS) >= 1 ? ARGS[1] : @__DIR__

jnum(x::Real) = isfinite(x) ? repr(Float64(x)) : "null"
jarr(v) = "[" * join((jnum(x) for x in v), ",") * "]"
function write_json(path, pairs)
    open(path, "w") do io
        print(io, "{")
        for (i, (k, v)) in enumerate(pairs)
            i > 1 && print(io, ",")
            print(io, "\n \"", k, "\": ", v isa AbstractString ? "\"" * v * "\"" : v isa AbstractVector ? jarr(v) : jnum(v))
        end
        print(io, "\n}\n")
    end
    println("wrote ", path)
end

"first-iteration stage vectors with the reference's own functions (Optimization.jl:288-357 on the uniform start design)"
function first_iteration(grid, dh, cellvalues, loads, bcs, p::OptimizationParameters)
    n = getncells(grid)
    mm = create_simp_material_model(p.E0, p.ν, p.Emin, p.p)
    K = allocate_matrix(dh)
    f = zeros(ndofs(dh))
    rho = fill(p.volume_fraction, n)
    vols = calculate_element_volumes(grid, create_volume_quadrature(grid))
    fc = create_filter_cache(grid, p.filter_radius, vols)
    phys = copy(rho)
    if p.filter_type == :density
        Optimization.apply_density_filter_cached!(phys, fc, rho)
    end
    assemble_stiffness_matrix_simp!(K, f, dh, cellvalues, mm, phys)
    Optimization.apply_loads_and_bcs!(K, f, loads, bcs)
    u = cholesky(Symmetric(K, :L)) \ f
    energy = 0.5 * dot(u, K * u)
    dc = zeros(n)
    Optimization.calculate_sensitivities!(dc, grid, dh, cellvalues, phys, u, p.E0, p.Emin, p.ν, p.p)
    dcf = zeros(n)
    if p.filter_type == :density
        Optimization.apply_density_filter_chain_rule_cached!(dcf, fc, dc)
    else
        Optimization.apply_sensitivity_filter_cached!(dcf, fc, rho, dc)
    end
    return (energy = energy, max_u = maximum(abs.(u)), dc = dc, dcf = dcf, volumes = vols)
end

function cantilever(name, nel, right; E0, Emin, vf, r, iters, tol, ftype = :sensitivity)
    grid = generate_grid(Hexahedron, nel, Vec((0.0, 0.0, 0.0)), Vec(right))
    dh, cellvalues, K, f = setup_problem(grid)
    mm = create_simp_material_model(E0, 0.3, Emin, 3.0)
    fixed_nodes = select_nodes_by_plane(grid, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-3)
    force_nodes = select_nodes_by_circle(grid, [right[1], 0.0, right[3] / 2], [1.0, 0.0, 0.0], 1.0)
    assemble_stiffness_matrix_simp!(K, f, dh, cellvalues, mm, fill(vf, getncells(grid)))
    ch = apply_fixed_boundary!(K, f, dh, fixed_nodes)
    p = OptimizationParameters(E0 = E0, Emin = Emin, ν = 0.3, p = 3.0, volume_fraction = vf, max_iterations = iters, tolerance = tol,
                               filter_radius = r, filter_type = ftype, move_limit = 0.2, damping = 0.5)
    loads = [PointLoad(dh, collect(force_nodes), [0.0, -1.0, 0.0])]
    st = first_iteration(grid, dh, cellvalues, loads, [ch], p)
    res = simp_optimize(grid, dh, cellvalues, loads, [ch], p)
    write_json(joinpath(outdir, "julia_$(name).json"),
               ["problem" => "cantilever", "nel" => collect(nel), "right" => collect(right), "E0" => E0, "Emin" => Emin, "nu" => 0.3, "p" => 3.0,
                "volume_fraction" => vf, "filter_radius" => r, "filter_type" => String(ftype), "max_iterations" => iters, "tolerance" => tol,
                "iterations" => res.iterations, "energy_history" => res.energy_history, "volume_history" => res.volume_history,
                "final_energy" => res.energy, "final_volume" => res.volume, "final_density" => res.densities,
                "max_abs_u_final" => maximum(abs.(res.displacements)),
                "first_energy" => st.energy, "first_max_u" => st.max_u, "first_dc" => st.dc, "first_dc_filtered" => st.dcf,
                "n_force_nodes" => length(force_nodes), "n_fixed_nodes" => length(fixed_nodes)])
end

function four_legs(name; iters)
    # test/Examples/05_3D_2x1x1_4Legs.jl:24-94: 40x20x20 over 2x1x1, four corner patches fixed on x = 0, circular load on x = 2
    grid = generate_grid(Hexahedron, (40, 20, 20), Vec((0.0, 0.0, 0.0)), Vec((2.0, 1.0, 1.0)))
    dh, cellvalues, K, f = setup_problem(grid)
    mm = create_simp_material_model(1.0, 0.3, 1e-9, 3.0)
    on_x0 = select_nodes_by_plane(grid, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-6)
    inleg(v) = (v <= 0.3 + 1e-9) || (v >= 0.7 - 1e-9)
    fixed_nodes = Set(n for n in on_x0 if inleg(grid.nodes[n].x[2]) && inleg(grid.nodes[n].x[3]))
    force_nodes = select_nodes_by_circle(grid, [2.0, 0.5, 0.5], [1.0, 0.0, 0.0], 0.1)
    assemble_stiffness_matrix_simp!(K, f, dh, cellvalues, mm, fill(0.4, getncells(grid)))
    ch = apply_fixed_boundary!(K, f, dh, fixed_nodes)
    p = OptimizationParameters(E0 = 1.0, Emin = 1e-9, ν = 0.3, p = 3.0, volume_fraction = 0.4, max_iterations = iters, tolerance = 0.0,
                               filter_radius = 2.0, filter_type = :density, move_limit = 0.2, damping = 0.5)
    loads = [PointLoad(dh, collect(force_nodes), [0.0, 0.0, -1.0])]
    st = first_iteration(grid, dh, cellvalues, loads, [ch], p)
    res = simp_optimize(grid, dh, cellvalues, loads, [ch], p)
    write_json(joinpath(outdir, "julia_$(name).json"),
               ["problem" => "four_legs", "nel" => [40, 20, 20], "right" => [2.0, 1.0, 1.0], "E0" => 1.0, "Emin" => 1e-9, "nu" => 0.3, "p" => 3.0,
                "volume_fraction" => 0.4, "filter_radius" => 2.0, "filter_type" => "density", "max_iterations" => iters, "tolerance" => 0.0,
                "iterations" => res.iterations, "energy_history" => res.energy_history, "volume_history" => res.volume_history,
                "final_energy" => res.energy, "final_volume" => res.volume, "final_density" => res.densities,
                "max_abs_u_final" => maximum(abs.(res.displacements)),
                "first_energy" => st.energy, "first_max_u" => st.max_u, "first_dc" => st.dc, "first_dc_filtered" => st.dcf,
                "n_force_nodes" => length(force_nodes), "n_fixed_nodes" => length(fixed_nodes)])
end

cantilever("runtests_60x20x4", (60, 20, 4), (60.0, 20.0, 4.0); E0 = 200.0, Emin = 1e-6, vf = 0.4, r = 2.5, iters = 20, tol = 0.0)
cantilever("c1_cantilever_60x20x4", (60, 20, 4), (60.0, 20.0, 4.0); E0 = 1.0, Emin = 1e-9, vf = 0.5, r = 1.5, iters = 50, tol = 0.0)
four_legs("c2_4legs_40x20x20"; iters = 12)
