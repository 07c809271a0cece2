"""CPU: host-side logic of the product (module tree, hyper-parameter derivation, index maps, tables) and the
C-ABI surface.  No CUDA compute is called here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
from torch import nn

from oracle import torch_oracle as TO
from tests import helpers as H


def test_library_exports_every_symbol_declared_in_the_header():
    from nextou_b200 import _lib
    hdr = open(os.path.join(H.ROOT, "include", "nextou_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(nextou_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 10
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nextou_b200.h but not exported"
    assert lib.nextou_abi_version() == 1
    assert isinstance(_lib.launch_count(), int)


def test_missing_library_fails_loudly(monkeypatch):
    from nextou_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnextou_b200.so")
    with pytest.raises(_lib.NextouError):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    from nextou_b200 import ops
    from nextou_b200._lib import NextouError
    with pytest.raises(NextouError):
        ops.knn_graph(torch.randn(16, 8), 1, 16, k=4)
    with pytest.raises(NextouError):
        ops.bti_labels(torch.randn(1, 3, 4, 4))


@pytest.mark.parametrize("fname,cfg", [("model_mini3d_reference.npz", H.MINI3D), ("model_mini2d_reference.npz", H.MINI2D)],
                         ids=["mini3d", "mini2d"])
def test_state_dict_layout_matches_reference(fname, cfg):
    """Every parameter / buffer key and shape of the reference model exists in the product model (published
    checkpoints must load), and nothing extra is trainable."""
    npz = H.golden_model(fname)
    m = H.build_product(cfg)
    gold = H.golden_state_dict(npz)
    own = m.state_dict()
    for k, v in gold.items():
        assert k in own and own[k].shape == v.shape, k
    skip = lambda k: k.startswith("decoder.encoder.") or k.endswith(("relative_pos", "num_batches_tracked"))
    extra = [k for k in own if k not in gold and not skip(k) and ".all_modules." not in k]
    assert not extra, extra[:5]
    assert any(k.startswith("decoder.encoder.stages.0") for k in own)          # decoder stores the encoder (ED:212)
    assert any(".all_modules.0.weight" in k for k in own)                      # upstream alias keys
    H.load_golden_into(m, npz)


def test_full_3d_config_parameter_census():
    """3d_fullres_nextou: 1 101 state-dict entries, 67 023 694 parameters, 30 671 182 trainable."""
    m = H.build_product(H.FULL3D)
    assert len(m.state_dict()) == 1101
    params = list(m.parameters())
    assert sum(p.numel() for p in params) == 67023694
    assert sum(p.numel() for p in params if p.requires_grad) == 30671182
    enc = m.encoder
    ks = [(s[0][1].blocks[0][0].graph_conv.k, s[0][2].blocks[0][0].graph_conv.k) for s in list(enc.stages)[2:]]
    assert ks == [(14, 7), (28, 14), (32, 14), (32, 28)]
    assert [s[0][1].blocks[0][0].graph_conv.r for s in list(enc.stages)[2:]] == [4, 2, 1, 1]
    assert enc.stages[2][0][1].blocks[0][0].graph_conv.pool_size == [2, 2, 2]
    assert enc.stages[3][0][1].blocks[0][0].graph_conv.pool_size == [1, 1, 1]
    assert tuple(enc.stages[2][0][1].blocks[0][0].relative_pos.shape) == (1, 10752, 168)
    assert tuple(enc.stages[3][0][1].blocks[0][0].relative_pos.shape) == (1, 10752, 1344)
    assert tuple(enc.stages[2][0][2].blocks[0][0].relative_pos.shape) == (1, 168, 168)
    assert enc.stages[2][0][2].blocks[0][0].window_size == (4, 7, 6)
    assert enc.stages[2][0][2].blocks[0][0].shift_size == [2, 3, 3]


@pytest.mark.parametrize("cfg", [H.MINI3D, H.MINI2D, H.FULL3D], ids=["mini3d", "mini2d", "full3d"])
def test_hyperparameters_agree_with_oracle_plan(cfg):
    from nextou_b200.blocks import _k_schedule, _pool_size_for, _stage_shapes
    dim = len(cfg["patch"])
    plan = TO.derive_plan(cfg["patch"], cfg["strides"])
    shapes, _ = _stage_shapes(cfg["patch"], cfg["strides"], nn.Conv3d if dim == 3 else nn.Conv2d)
    assert [tuple(s) for s in shapes] == [tuple(s) for s in plan["shapes"]]
    k_list, _ = _k_schedule(shapes[-1], len(cfg["strides"]), dim)
    assert k_list == plan["k_list"]
    for s, st in enumerate(plan["stages"]):
        if st is not None:
            assert _pool_size_for(shapes[s], shapes[-1]) == st["pool_size"]


def test_relative_pos_tables_bit_identical_to_oracle():
    from nextou_b200.pos_embed import relative_pos_parameter
    for (C, n, m, dim) in [(24, 1536, 24, 3), (36, 24, 24, 3), (132, 168, 168, 3), (132, 343, 343, 3), (32, 256, 16, 2),
                           (16, 16, 16, 2)]:
        a = relative_pos_parameter(C, n, m, dim)
        b = TO.relative_pos_table(C, n, m, dim)
        assert a.requires_grad is False and tuple(a.shape) == (1, n, m)
        assert torch.equal(a.data, b)


@pytest.mark.parametrize("spatial,window,shift", [((8, 14, 12), (4, 7, 6), (2, 3, 3)), ((4, 6, 8), (2, 3, 4), (1, 1, 2)),
                                                  ((8, 8), (4, 4), (2, 2)), ((4, 7, 6), (4, 7, 6), (0, 0, 0))])
def test_shifted_window_row_map_equals_roll_plus_partition(spatial, window, shift):
    """The int32 row map reproduces torch.roll(-shift) + window_partition (ED:634-660, 784) as pure indexing."""
    from nextou_b200.blocks import shifted_window_row_map, window_partition, window_reverse
    B, C = 2, 3
    V = int(np.prod(spatial))
    x = torch.arange(B * C * V, dtype=torch.float32).reshape(B, C, *spatial)
    dims = tuple(range(2, 2 + len(spatial)))
    rolled = torch.roll(x, shifts=tuple(-s for s in shift), dims=dims) if max(shift) > 0 else x
    win = window_partition(rolled, window)                                   # (B*nW, C, *window)
    assert torch.equal(win, TO._windows(rolled, window))
    assert torch.equal(window_reverse(win, window, spatial), rolled)
    rows = shifted_window_row_map(B, spatial, window, shift, "cpu").long()
    tok = x.permute(0, *dims, 1).reshape(B * V, C)                           # natural channels-last rows
    n = int(np.prod(window))
    expect = win.reshape(-1, C, n).permute(0, 2, 1).reshape(-1, C)
    assert torch.equal(tok[rows], expect)
    assert sorted(rows.tolist()) == list(range(B * V))                       # a permutation


def test_interaction_tables_match_oracle():
    from nextou_b200.losses import BTI_Loss, TI_Loss
    from oracle.ref_shims import SYNAPSE_EXCLUSION, make_tensors
    exc = make_tensors(SYNAPSE_EXCLUSION)
    loss = BTI_Loss(dim=3, connectivity=26, inclusion=[], exclusion=exc, min_thick=1)
    assert loss.interaction_table() == tuple(TO.interaction_table([], exc))
    assert len(loss.interaction_list) == 12 and tuple(loss.kernel.shape) == (1, 1, 3, 3, 3)
    inc = make_tensors([[1, 2], [[3], [1, 2]]])
    l2 = BTI_Loss(dim=2, connectivity=4, inclusion=inc, exclusion=make_tensors([[1, 3]]))
    assert l2.interaction_table() == tuple(TO.interaction_table(inc, make_tensors([[1, 3]])))
    assert l2.kernel[0, 0].tolist() == [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    with pytest.raises(ValueError):
        TI_Loss(dim=3, connectivity=26, exclusion=[[torch.tensor([1, 2]), torch.tensor([3])]]).interaction_table()
    with pytest.raises(ValueError):
        BTI_Loss(dim=3, connectivity=8)


def test_error_conventions_match_reference():
    from nextou_b200.blocks import GraphConv, NexToU_Encoder
    from nextou_b200.layers import act_layer, norm_layer
    with pytest.raises(NotImplementedError):
        GraphConv(12, 24, conv="edge")                      # only 'mr' exists (ED:426-429)
    with pytest.raises(NotImplementedError):
        act_layer("swish")
    with pytest.raises(NotImplementedError):
        norm_layer("layer", 8, nn.Conv3d)
    with pytest.raises(ValueError):
        NexToU_Encoder(1, [32, 32], 5, 8, nn.Conv1d, 3, [[1]] + [[2]] * 4, 2)
    with pytest.raises(AssertionError):
        H.build_product(dict(H.MINI2D, kernels=[[3, 3]] * 4))


def test_fork_tokens_gradient_sum_cpu():
    """ops.fork_tokens is pure tensor plumbing (no kernel): the two aliases' gradients are summed by the package, in the
    padded token layout when both arrive in it, and the result equals autograd's own accumulation."""
    import torch
    from nextou_b200 import ops
    g = torch.Generator().manual_seed(0)
    rows, C = 64, 33
    x = torch.randn(rows, ops.pad8(C), generator=g)[:, :C].requires_grad_(True)
    w1 = torch.randn(rows, ops.pad8(C), generator=g)[:, :C]          # padded-pitch gradient
    w2 = torch.randn(rows, C, generator=g)                           # dense gradient
    a, b = ops.fork_tokens(x)
    seen = []
    h = x.register_hook(lambda t: seen.append(tuple(t.stride())))
    ((a * w1).sum() + (b * w2).sum()).backward()
    h.remove()
    assert torch.allclose(x.grad, w1 + w2)
    assert seen and seen[0][0] % 8 == 0 and seen[0][1] == 1          # summed into a channel-padded buffer
    x.grad = None
    a, b = ops.fork_tokens(x)
    ((a * w1).sum() + (b * w1).sum()).backward()                     # both padded: one pass over the physical rows
    assert torch.allclose(x.grad, 2 * w1)
    x.grad = None
    a, b = ops.fork_tokens(x)
    (a * w1).sum().backward()                                        # unused alias: its gradient is None, not zeros
    assert torch.equal(x.grad, w1)
    with torch.no_grad():
        p, q = ops.fork_tokens(x)
    assert p is x and q is x                                         # no autograd: no-op


def test_relative_pos_table_memo_and_disk_cache(tmp_path, monkeypatch):
    """In-process memo and NEXTOU_RELPOS_CACHE return the very bytes of a fresh computation (tables stay bit-identical)."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from nextou_b200 import pos_embed as PE
    def fresh(c, n, m, d):
        grid = int(n ** (1 / d))
        fn = PE.get_3d_relative_pos_embed if d == 3 else PE.get_2d_relative_pos_embed
        t = torch.from_numpy(np.float32(fn(c, grid))).unsqueeze(0).unsqueeze(1)
        return -F.interpolate(t, size=(n, m), mode="bicubic", align_corners=False).squeeze(1)
    PE._TABLE_MEMO.clear()
    monkeypatch.setenv("NEXTOU_RELPOS_CACHE", str(tmp_path))
    a = PE.relative_pos_parameter(12, 168, 168, 3)
    assert torch.equal(a.data, fresh(12, 168, 168, 3)) and not a.requires_grad
    assert (tmp_path / "relpos_c12_n168_m168_d3.pt").exists()
    b = PE.relative_pos_parameter(12, 168, 168, 3)                    # memo hit: equal values, independent storage
    assert torch.equal(a, b) and a.data_ptr() != b.data_ptr()
    PE._TABLE_MEMO.clear()
    c = PE.relative_pos_parameter(12, 168, 168, 3)                    # disk hit
    assert torch.equal(a, c)
    d2 = PE.relative_pos_parameter(8, 64, 16, 2)
    assert torch.equal(d2.data, fresh(8, 64, 16, 2))
    PE._TABLE_MEMO.clear()
