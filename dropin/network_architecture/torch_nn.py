"""Drop-in for .../network_architecture/torch_nn.py."""
from nextou_b200.layers import act_layer, norm_layer, MLP, BasicConv, batched_index_select  # noqa: F401
