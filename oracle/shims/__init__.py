"""Stub modules that let the UNMODIFIED reference files import in an image without
ase / torch_geometric / torch_scatter / torch_cluster.  Test infrastructure only."""
