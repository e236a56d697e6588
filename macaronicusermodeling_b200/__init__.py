"""B200-native loopy-belief-propagation hot path of MacaronicUserModeling (LBP.py + array_utils).

Sub-modules
    LBP                         drop-in for the reference's LBP.py (FactorGraph / VariableNode / FactorNode / ...)
    array_utils.c_array_utils   drop-in for the reference's Cython helpers (imported as ``au`` by LBP.py:6)
    engine                      batched executor: many sentence graphs -> level-batched sm_100a kernels
    trainer                     batch_sgd_many / epoch loop / data-parallel theta all-reduce (train.py, train_mp.py)
    synth                       synthetic sentences and feature planes

All arithmetic runs in hand-written CUDA kernels behind the C ABI declared in include/mlbp.h
(``libmlbp.so``, built in-tree by ``__graft_entry__.build()``).  There is no CPU fallback: importing
``engine`` without the built library, or running it without a GPU, raises.
"""
__version__ = '0.1.0'
