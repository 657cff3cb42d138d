"""Form families: the closed sets of variational forms whose quadrature kernels
exist in the CUDA engine.  Each module mirrors the function names of the
reference example script that defines the same forms in UFL."""
