"""Drop-in for .../network_architecture/pos_embed.py."""
from nextou_b200.pos_embed import (get_2d_relative_pos_embed, get_3d_relative_pos_embed, get_2d_sincos_pos_embed,  # noqa: F401
                                   get_3d_sincos_pos_embed, get_1d_sincos_pos_embed_from_grid)
