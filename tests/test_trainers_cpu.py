"""CPU: the nnU-Net trainer plug-ins (nextou_b200.trainers) build the reference's network / loss configuration."""
import types

import numpy as np
import pytest
import torch

from tests import helpers as H


def _managers(cfg):
    dim = len(cfg["patch"])
    cm = types.SimpleNamespace(conv_kernel_sizes=cfg["kernels"], pool_op_kernel_sizes=cfg["strides"], patch_size=list(cfg["patch"]),
                               UNet_base_num_features=cfg["feats"][0], unet_max_num_features=max(cfg["feats"]),
                               n_conv_per_stage_encoder=[2] * len(cfg["feats"]), n_conv_per_stage_decoder=[2] * (len(cfg["feats"]) - 1),
                               batch_dice=True)
    lm = types.SimpleNamespace(num_segmentation_heads=cfg["num_classes"], ignore_label=None)
    pm = types.SimpleNamespace(get_label_manager=lambda dataset_json: lm)
    return pm, cm


def test_build_network_architecture_matches_reference_layout():
    from nextou_b200.trainers import nnUNetTrainer_NexToU
    cfg = dict(H.MINI3D, feats=(6, 12, 24, 48, 48, 48))   # base 6, doubled per stage, capped at 48 (TR:78-79)
    pm, cm = _managers(cfg)
    torch.manual_seed(0)
    net = nnUNetTrainer_NexToU.build_network_architecture(pm, {}, cm, 1, True)
    ref = H.build_product(cfg)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    assert net.decoder.deep_supervision is True
    w = net.encoder.stages[0][0].convs[0].conv
    assert float(w.bias.abs().max()) == 0.0                      # InitWeights_He: zero bias (TR:88)
    assert isinstance(net.encoder.stages[0][0].convs[0].norm, torch.nn.BatchNorm3d)
    assert net.encoder.stages[0][0].convs[0].nonlin.negative_slope == 0.01


@pytest.mark.parametrize("name,n_inter,compound", [("nnUNetTrainer_NexToU_BTI_Synapse", 12, "DC_and_CE_and_BTI_Loss"),
                                                    ("nnUNetTrainer_NexToU_BTI_ICA_NoMirroring", 17, "DC_and_CE_and_BTI_Loss"),
                                                    ("nnUNetTrainer_NexToU_BTI_RAVIR", 1, "DC_and_CE_and_BTI_Loss"),
                                                    ("nnUNetTrainer_NexToU_TI", 10, "DC_and_CE_and_TI_Loss"),
                                                    ("nnUNetTrainer_NexToU_TI_NoMirroring", 10, "DC_and_CE_and_TI_Loss")])
def test_build_loss_configuration(name, n_inter, compound):
    from nextou_b200 import trainers
    pm, cm = _managers(H.MINI3D)
    tr = getattr(trainers, name)(plans_manager=pm, configuration_manager=cm, dataset_json={"labels": {"bg": 0, "a": 1, "b": 2, "c": 3, "d": 4, "e": 5}},
                                 device="cpu")
    loss = tr._build_loss()
    assert type(loss).__name__ == "DeepSupervisionWrapper"
    w = np.array(loss.weight_factors)
    assert np.allclose(w, np.array([1, 0.5, 0.25, 0.125, 0]) / 1.875)          # SYN:23-27
    inner = loss.loss
    assert type(inner).__name__ == compound
    assert inner.weight_ti == 1e-6 and inner.ti.connectivity == 26 and inner.ti.dim == 3
    assert len(inner.ti.interaction_list) == n_inter
    ma, mc, inc = inner.ti.interaction_table()
    assert not any(inc) and all(a & c == 0 for a, c in zip(ma, mc))               # exclusion sets are disjoint
    if "NoMirroring" in name:
        _, _, _, mirror = tr.configure_rotation_dummyDA_mirroring_and_inital_patch_size()
        assert mirror is None and tr.inference_allowed_mirroring_axes is None


def test_dropin_overlay_reexports_reference_module_paths():
    import importlib.util
    import os
    root = os.path.join(H.ROOT, "dropin")
    expected = {"network_architecture/NexToU.py": ["NexToU"],
                "network_architecture/NexToU_Encoder_Decoder.py": ["NexToU_Encoder", "NexToU_Decoder", "SwinGrapher", "PoolGrapher", "FFN", "MRConv", "OptInit"],
                "network_architecture/torch_edge.py": ["DenseDilatedKnnGraph", "dense_knn_matrix", "xy_dense_knn_matrix"],
                "network_architecture/torch_nn.py": ["BasicConv", "batched_index_select", "act_layer", "norm_layer"],
                "network_architecture/pos_embed.py": ["get_3d_relative_pos_embed", "get_2d_relative_pos_embed"],
                "loss/bti_loss.py": ["BTI_Loss"], "loss/ti_loss.py": ["TI_Loss"],
                "loss/compound_bti_loss.py": ["DC_and_CE_and_BTI_Loss"], "loss/compound_ti_loss.py": ["DC_and_CE_and_TI_Loss"],
                "nnUNetTrainer/nnUNetTrainer_NexToU.py": ["nnUNetTrainer_NexToU"],
                "nnUNetTrainer/nnUNetTrainer_NexToU_BTI_Synapse.py": ["nnUNetTrainer_NexToU_BTI_Synapse"]}
    for rel, names in expected.items():
        spec = importlib.util.spec_from_file_location("dropin_" + rel.replace("/", "_")[:-3], os.path.join(root, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for n in names:
            assert hasattr(mod, n), (rel, n)
