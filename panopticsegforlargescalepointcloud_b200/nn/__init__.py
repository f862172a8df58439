"""Sparse-backend plugin modules for the reference's `torch_points3d.modules.SparseConv3d.nn` hook."""
