"""easysimp_b200 -- host-side mirror of EasySIMP.jl's SIMP hot-path interface over libeasysimp_b200.so."""
from .lib import Context, EsimpError, EsimpIterStats, EsimpParams, comm_unique_id, load_library  # noqa: F401
from .host import (  # noqa: F401
    ConstraintHandler, DofHandler, FilterCache, Grid, Hexahedron, OptimizationParameters, OptimizationResult, PointLoad,
    StiffnessOperator, Tetrahedron, apply_fixed_boundary_, apply_force_, apply_load_condition_, apply_sliding_boundary_,
    assemble_stiffness_matrix_simp_, calculate_cell_centers, calculate_element_volumes, create_filter_cache,
    create_simp_material_model, estimate_element_size, generate_grid, select_nodes_by_circle, select_nodes_by_plane,
    setup_problem, simp_optimize, StressField, apply_variable_density_volume_force_, calculate_stresses_simp, create_results_data,
    check_sensitivity_health,
)
from .partition import slab_partition  # noqa: F401
