"""dnet_b200 -- the B200-native shard forward of dnet's pipelined ring.

Only what the north-star path needs: the C-ABI CUDA library (csrc/, lib/), and the
host-side mirror of the reference's plug-in seams for that path (ComputePolicy
registry, BaseRingModel operator API, WeightCache / pools / LayerManager,
ShardRuntime, ActivationCodec, the ring hop).  See DESIGN.md.
"""
__version__ = "0.1.0"
