"""modril_b200 -- B200-native (sm_100a) implementation of the FLOWRIL/MODRIL flow-matching density hot path.

Layout:
  csrc/            CUDA kernels + the C ABI (include/flowril_b200.h) -> lib/libflowril_b200.so
  _native.py       ctypes binding (no fallback: raises if the library is missing)
  flowril/         host-side mirror of the reference's modril/flowril package (networks, pretrain, flowril)
  bcf.py           BCF flow-matching policy (VectorNetwork, FlowPolicy, bc_step) -- SURVEY section 8 row f3
  ppo.py           PPO consumer (returns / GAE, DistActorCritic, ppo_update, PPO updater) -- row f4
  dist.py          row-sharded scoring / data-parallel training helpers (torch.distributed over NCCL)
"""
__version__ = "0.1.0"
