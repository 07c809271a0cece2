"""Drop-in for .../network_architecture/NexToU_Encoder_Decoder.py."""
from nextou_b200.blocks import (OptInit, NexToU_Encoder, NexToU_Decoder, FFN, MRConv, GraphConv, DyGraphConv,  # noqa: F401
                                PoolDyGraphConv, Grapher, window_partition, window_reverse, SwinGrapher, PoolGrapher,
                                SwinGNNBlocks, PoolGNNBlocks)
