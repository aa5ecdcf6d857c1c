"""omgsr_b200 — B200-native (sm_100a) implementation of OMGSR's one-step SR inference path.

Only what the hot path needs lives here: `csrc/` (hand-written CUDA + the C ABI), `capi` (ctypes binding),
`pack` (weight repacking into the kernel layouts) and the `engine_*` modules that drive the kernels behind
the reference's module surface (`infer/`).
"""
__version__ = "0.1.0"
