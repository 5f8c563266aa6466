"""Public names of the package (import as ``gadm_b200``)."""
from .projectors import BasicProjector, CudaProjector, DeferredProjection, ProjectionType, is_not_buffer  # noqa: F401
from .aggregation import (PackedMasks, aoi_attr_batched, bootstrap_statistic, loo_attr_batched, data_banzhaf, data_banzhaf_batched, data_shapley,  # noqa: F401
                          data_shapley_batched, evaluate_lds, group_reduce, lds_per_test_set, masks_from_remaining_idx,
                          spearman_matrix, lds_fit_sweep, convergence_metrics,
                          stable_rank, sym_pinv)
from .scoring import (LOCAL, TrakScorer, aggregate_by_class, col_mean_scaled, compute_dtrak_trak_scores,  # noqa: F401
                      compute_gradient_scores, gemm_tn, gradient_scores, group_and_rank, matvec_rows, row_norms,
                      trak_scores, transpose)
from .masks import (counterfactual_split, masks_from_seeds, remove_data_by_datamodel, remove_data_by_shapley,  # noqa: F401
                    remove_data_by_uniform)
from .datamodel import (RidgeCV, compute_datamodel_scores, datamodel, datamodel_ridge_batched,  # noqa: F401
                        ridge_cv_batched)
from .formats import (collect_data, journey_point_indices, load_lds_test_sets, read_behavior_db, run_traks,  # noqa: F401
                      save_lds_outputs, write_journey_group_csv)
from ._lib import GadmError, load_library  # noqa: F401

__all__ = [
    "BasicProjector", "CudaProjector", "DeferredProjection", "ProjectionType", "is_not_buffer",
    "PackedMasks", "aoi_attr_batched", "loo_attr_batched", "bootstrap_statistic", "data_banzhaf", "data_banzhaf_batched", "data_shapley",
    "data_shapley_batched", "evaluate_lds", "group_reduce", "lds_per_test_set", "masks_from_remaining_idx", "spearman_matrix", "stable_rank",
    "sym_pinv", "lds_fit_sweep", "convergence_metrics",
    "LOCAL", "matvec_rows", "TrakScorer", "aggregate_by_class", "col_mean_scaled", "compute_dtrak_trak_scores", "compute_gradient_scores",
    "gemm_tn", "gradient_scores", "group_and_rank", "row_norms", "trak_scores", "transpose",
    "counterfactual_split", "masks_from_seeds", "remove_data_by_datamodel", "remove_data_by_shapley",
    "remove_data_by_uniform",
    "RidgeCV", "datamodel_ridge_batched", "ridge_cv_batched", "datamodel", "compute_datamodel_scores",
    "collect_data", "journey_point_indices", "load_lds_test_sets", "read_behavior_db", "run_traks", "save_lds_outputs",
    "write_journey_group_csv",
    "GadmError", "load_library",
]
