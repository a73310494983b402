"""multigridcmt_b200 -- B200-native multigrid V-cycle eigensolver path of AndyMN/MultigridCMT.

Drop-in classes (same names and signatures as the reference's modules of the same name):

    from multigridcmt_b200 import MGCMTStencilMaker, MGCMTSolver, MGCMTProcessor

Everything numerical runs in libmgcmt_b200.so (hand-written sm_100a CUDA, C ABI in
include/mgcmt_b200.h); there is no CPU fallback.
"""
from .MGCMTProcessor import MGCMTProcessor
from .MGCMTSolver import MGCMTSolver, ZeroVector
from .MGCMTStencilMaker import MGCMTStencilMaker
from .operators import SeparableOperator, UnsupportedOperator

__all__ = ["MGCMTStencilMaker", "MGCMTSolver", "MGCMTProcessor", "SeparableOperator", "UnsupportedOperator", "ZeroVector"]
