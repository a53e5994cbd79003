"""pa_b200 — a B200-native engine for the PartitionedArrays.jl PSparseMatrix x PVector hot path.

The directory is named ``partitionedarrays.jl_b200``; import it as ``pa_b200`` (see /pa_b200.py shim)."""
from . import _capi, build, prange
from . import fem_example, hpcg
from ._capi import PA_CG_REFERENCE_OPS, PA_CG_TIMING, PA_SPMV_FUSED_EXCHANGE, PA_SPMV_INLINE_PEER_LOADS, PA_SPMV_OVERLAP, PA_SPMV_DEFAULT, PA_SPMV_EXPLICIT_EXCHANGE, PA_SPMV_SKIP_GHOST_REFRESH, PAError
from .hpcg import GaussSeidel, MgPreconditioner, cg_timings, hpcg_benchmark, pc_setup, ref_cg_pc_, report_results
from .gallery import build_p_matrix, compute_optimal_shape_xyz, fill_hash, laplacian_fdm, stencil_matrix
from .parrays import (consistent_matrix, rap, spmm, spmtm, CGResult, CUDAArray, ExchangeGraph, exchange, exchange_layout, PRange, PSparseMatrix, PVector, assemble_, consistent_, dot, mul_, mul_no_lat_, mul_transpose_, norm,
                      opt_cg_, spmv_, spmtv_, pfill, pones, psparse, pvector, pvector_from_global, pvector_from_triplets, pzeros, ref_cg_, uniform_partition,
                      variable_partition, with_cuda, with_cuda_multi)
from .prange import LocalIndices, local_range

__all__ = [n for n in dir() if not n.startswith("_")]
