"""Drop-in for the reference's in-tree `simple-knn` extension (/root/reference/simple-knn, imported at
gs_renderer.py:14 as `from simple_knn._C import distCUDA2`), backed by libdgr_b200.so (dgr_dist_cuda2)."""
