"""oracle/shim/compressai -- a CPU restatement of the slice of CompressAI that sc2bench's
bottleneck path executes, shaped like the real package so the reference's own
`sc2bench/models/layer.py` can be imported on top of it to generate golden vectors.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: CompressAI (compressai>=1.2.3,
/root/reference/setup.py:28) is not vendored in the reference and not installable here; this
package restates its published behaviour, it is NOT CompressAI.
Product code never imports it; `sc2bench_b200` has its own CUDA-backed implementation.
"""
__version__ = '0+oracle-restatement'
