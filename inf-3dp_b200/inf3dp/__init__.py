"""inf3dp -- B200-native (sm_100a) implementation of INF-3DP's data-parallel hot path.

Fused SIREN field queries (value / gradient / Hessian as forward-mode jets) and isocontour / field-intersection
extraction, behind the reference's own Python surface.  See DESIGN.md and INTEGRATION.md at the repo root.
"""
from . import _lib
from .siren import SingleBVPNet, MLPClassifier, set_engine, get_engine
from .contours import (extract_isocontours, get_fields_intersection_inter_slice, extract_isocontours_raw,
                       extract_isocontours_compact, contour_loops)
from .diffops import (gradient, hessian, hessian3d, jacobian, divergence, laplace, principal_curvatures,
                      min_max_curvs_and_vectors, unit_normals, unit_normals_rows)
from .query import field_querying, field_query_with_grads, HostPipeline
from .projection import batch_isosurface_projection, iter_project_contours_batch
from .volume import (grid_queries, grid_field, get_grid_fields, marching_cubes, sdf_to_mesh, gen_iso_layers,
                     get_fields_intersection_from_grids)
from .collision import (CollTVSDF, collision_evaluation, tvsdf_collision_forward, quat_collision_sweep,
                        quat_collision_sweep_unfused, transform_nozzle_pcd_with_quaternion)


def patch_reference(force=False):
    """Rebind the reference's hot-path classes in already-imported reference modules:
    `modules.SingleBVPNet` -> inf3dp.SingleBVPNet.  `diff_operators` is left untouched: its autograd calls work on
    the fused outputs.

    The drop-in is inference-only (no gradient path to the weights).  When the calling process has imported the
    reference's `training` module -- i.e. it is a training script -- rebinding the class would make the networks it is
    about to train untrainable, so this raises unless force=True (e.g. train_collision.py, which trains the quaternion
    network against FROZEN field networks: rebind by hand for those instead)."""
    import sys
    if not force and "training" in sys.modules and hasattr(sys.modules["training"], "train"):
        raise RuntimeError("inf3dp.patch_reference(): the reference's `training` module is imported -- this looks like a "
                           "training script, and inf3dp.SingleBVPNet is inference-only (weights get no gradients).  "
                           "Keep modules.SingleBVPNet for the networks being trained; pass force=True to rebind anyway.")
    m = sys.modules.get("modules")
    if m is not None:
        m.SingleBVPNet = SingleBVPNet
    return m is not None


__all__ = [
    "SingleBVPNet", "MLPClassifier", "set_engine", "get_engine", "extract_isocontours",
    "get_fields_intersection_inter_slice", "extract_isocontours_raw", "extract_isocontours_compact", "contour_loops",
    "gradient", "hessian", "hessian3d", "jacobian",
    "divergence", "laplace", "principal_curvatures", "min_max_curvs_and_vectors", "unit_normals", "unit_normals_rows", "field_querying",
    "field_query_with_grads", "HostPipeline", "patch_reference", "batch_isosurface_projection",
    "iter_project_contours_batch", "CollTVSDF", "collision_evaluation", "tvsdf_collision_forward", "quat_collision_sweep", "quat_collision_sweep_unfused",
    "transform_nozzle_pcd_with_quaternion", "grid_queries", "grid_field", "get_grid_fields", "marching_cubes",
    "sdf_to_mesh", "gen_iso_layers", "get_fields_intersection_from_grids",
]
