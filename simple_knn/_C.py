"""Stands where the reference's compiled module simple_knn._C stands (submodules/simple-knn/ext.cpp:15-17)."""
from luciddreamer_b200.simple_knn import distCUDA2  # noqa: F401
