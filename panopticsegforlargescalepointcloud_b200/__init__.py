"""B200-native (sm_100a) replacement for the sparse-conv + instance-clustering hot path of
prs-eth/PanopticSegForLargeScalePointCloud.

Sub-modules mirror the three native dependencies the reference calls:

  me        MinkowskiEngine-shaped namespace   (SparseTensor, MinkowskiConvolution, ...)
  tpk       torch_points_kernels-shaped ops    (ball_query, region_grow, instance_iou)
  hdbscan   hdbscan-shaped class               (HDBSCAN.fit_predict)

plus `models` (the PointGroup-style model assembled from them, torch_points3d BaseModel API) and
`parallel` helpers for the one-process-per-GPU data-parallel step.  Everything computes through
libpgs_b200.so (include/pgs_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
