"""B200-native QP-subproblem engine for SqpSolver.jl's ``external_optimizer`` slot.

Sub-packages
------------
csrc/   hand-written sm_100a CUDA kernels + the C-ABI (``libsqpqp.so``)
capi    ctypes binding of ``include/sqpqp.h`` (fails loudly if the .so is absent)
host/   host-side mirror of the reference's sub-optimizer / SQP-TR interface
nlp/    NLP evaluators used as workloads (toy problems, polar ACOPF)
julia/  the ``ccall`` shim a SqpSolver.jl maintainer would add (not runnable here)
"""
__version__ = "0.1.0"
