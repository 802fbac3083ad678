"""zktls_b200 -- B200-native (sm_100a) STARK proving backend for the `zktls prove -p r0` hot path.

The product is libzkb200.so (hand-written CUDA + the C-ABI in include/zkb200.h).  This package is the thin
host-side mirror used by the tests and the benchmark: `hal.B200Hal` (the risc0_zkp `Hal` operator surface),
`prover.SegmentProver` (Prover::commit_group / finalize + the segment driver) and `circuit` (circuit blobs).
"""
from ._lib import lib, ZkbError, SO_PATH  # noqa: F401
