"""bnkit_b200 -- B200-native (sm_100a) per-alignment-column inference on the phylogenetic Bayesian
network behind GRASP ancestral sequence reconstruction: joint (max-product) and marginal
(sum-product) passes, behind the C ABI of include/bnkgpu.h.

    lib    ctypes binding of libbnkgpu.so (the product)
    phylo  host-side mirror of the reference's Java interface for this path
    synth  synthetic trees / alignments of the BASELINE.json shapes
    build  nvcc build of libbnkgpu.so (sm_100a, in-tree)
"""
__version__ = "0.1"
