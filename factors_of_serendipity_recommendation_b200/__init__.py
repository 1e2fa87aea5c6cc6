"""B200-native LightGCN propagation + scoring engine: a drop-in for the PyTorch LightGCN path of
csjwj2023/factors-of-serendipity-recommendation (lightGCN/LightGCN-PyTorch-master/code).

Modules mirror the reference's: world, dataloader, model, utils, Procedure, register; ``_lgx`` binds
the C ABI of include/lgx.h (hand-written sm_100a CUDA in csrc/).  Importing the package never
touches the GPU; using it without the built library or off a B200 raises (no fallback)."""
__all__ = ["world", "dataloader", "model", "utils", "Procedure", "register", "synth"]
