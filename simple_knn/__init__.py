"""Import alias so reference code runs unchanged: `from simple_knn._C import distCUDA2` (scene/gaussian_model.py:19)
resolves to the B200-native implementation in luciddreamer_b200.simple_knn."""
