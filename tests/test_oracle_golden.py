"""CPU: the oracle (oracle/) against the committed golden fixtures that the UNMODIFIED reference produced
(oracle/make_golden.py).  These pin the checker; the -m gpu tests then compare the CUDA path with the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import torch_oracle as TO
from tests import helpers as H
from oracle.make_golden import BTI_CASES, bti_case, bti_interactions


@pytest.fixture(scope="module")
def knn_gold():
    return np.load(os.path.join(H.GOLDEN, "knn_reference.npz"))


@pytest.mark.parametrize("case", H.KNN_CASES, ids=[c[0] for c in H.KNN_CASES])
def test_c_oracle_knn_matches_reference_golden(case, knn_gold):
    """Tie-aware equality: neighbour SETS must be equal except where the reference's own fp32 distance gap at the
    k-th / (k+1)-th boundary is below 1e-5 (MKL vs sequential-FMA accumulation order)."""
    name, B, N, M, C, k, d, rp = case
    x, y, relpos = H.knn_inputs(case)
    ref = knn_gold[name].astype(np.int64)
    mine = c_oracle.knn_graph(x.numpy(), None if y is None else y.numpy(), None if relpos is None else relpos[0].numpy(),
                              k, d)
    assert mine.shape == ref.shape == (B, N, k)
    x4 = x.permute(0, 2, 1).unsqueeze(-1)
    y4 = None if y is None else y.permute(0, 2, 1).unsqueeze(-1)
    dist = TO.knn_distances(x4, y4, relpos)                      # fp64
    bad = 0
    for b in range(B):
        rows = np.nonzero((mine[b] != ref[b]).any(-1))[0]
        for i in rows:
            dm = dist[b, i, torch.from_numpy(mine[b, i])]
            dr = dist[b, i, torch.from_numpy(ref[b, i])]
            # positions may swap / differ only between candidates that are (nearly) equidistant
            if not torch.allclose(dm, dr, rtol=0, atol=1e-5):
                bad += 1
    assert bad == 0
    frac = float((mine != ref).any(-1).mean())
    assert frac < 0.02, f"{frac:.4f} of rows differ from the reference even tie-aware"


def test_c_oracle_knn_order_is_ascending_distance_lowest_index_first():
    x = torch.zeros(1, 8, 4)
    x[0, :, 0] = torch.tensor([0., 1., 1., 2., 2., 2., 3., 3.])   # duplicates -> exact ties
    idx = c_oracle.knn_graph(x.numpy(), None, None, 4, 1, normalize=False)
    assert idx[0, 0].tolist() == [0, 1, 2, 3]
    assert idx[0, 3].tolist() == [3, 4, 5, 1]
    assert idx[0, 7].tolist() == [6, 7, 3, 4]


@pytest.fixture(scope="module")
def bti_gold():
    return np.load(os.path.join(H.GOLDEN, "bti_reference.npz"))


@pytest.mark.parametrize("case", BTI_CASES, ids=[c[0] for c in BTI_CASES])
def test_bti_oracles_match_reference_golden(case, bti_gold):
    name, shape, nc, seed, conn, thick, kind = case
    logits, target = bti_case(shape, nc, seed)
    inc, exc = bti_interactions(kind, nc)
    dim = len(shape) - 1
    ma, mc, flags = TO.interaction_table(inc, exc)
    labels = logits.argmax(1)
    ref_crit = np.unpackbits(bti_gold[f"{name}.bti.crit"])[: labels.numel()].reshape(labels.shape)
    # torch restatement: bit exact map, loss to fp64 round-off
    crit_t = TO.bti_critical_map(labels, ma, mc, flags, dim, conn, thick)
    assert np.array_equal(crit_t.numpy().astype(np.uint8), ref_crit)
    loss = TO.bti_loss(logits, target, inc, exc, dim, conn, thick)
    assert abs(loss.item() - float(bti_gold[f"{name}.bti.loss"])) <= 1e-10 * max(1.0, abs(loss.item()))
    # plain-C restatement: bit exact map
    crit_c = c_oracle.bti_critical(labels.numpy().astype(np.uint8), ma, mc, flags, conn, thick)
    assert np.array_equal(crit_c, ref_crit)
    assert 0.001 < ref_crit.mean() < 0.9, "fixture should have a non-trivial critical region"
    if f"{name}.ti.crit" in bti_gold.files:  # TI twin (scalar labels) is the same map
        ti_crit = np.unpackbits(bti_gold[f"{name}.ti.crit"])[: labels.numel()].reshape(labels.shape)
        assert np.array_equal(ti_crit, ref_crit)


@pytest.mark.parametrize("fname,cfg", [("model_mini3d_reference.npz", H.MINI3D), ("model_mini2d_reference.npz", H.MINI2D)],
                         ids=["mini3d", "mini2d"])
def test_torch_oracle_model_matches_reference_golden(fname, cfg):
    """Whole-network forward + backward of the functional restatement, teacher-forced with the reference's own
    neighbour lists (a randomly initialised NexToU is chaotic in them, see oracle/torch_oracle.py::ReplayKnn)."""
    npz = H.golden_model(fname)
    sd = H.golden_state_dict(npz)
    dim = len(cfg["patch"])
    plan = TO.derive_plan(cfg["patch"], cfg["strides"])
    # relative_pos tables are not stored: rebuild them and attach under the reference's key names
    for s in range(plan["gnn_from"], len(cfg["feats"])):
        st = plan["stages"][s]
        C = cfg["feats"][s]
        n_pool = int(np.prod(st["shape"])) // int(np.prod(st["pool_size"]))
        n_win = int(np.prod(st["window"]))
        prefixes = [f"encoder.stages.{s}.0"]
        j = len(cfg["feats"]) - 2 - s
        if j >= 0:
            prefixes.append(f"decoder.stages.{j}")
        for p in prefixes:
            sd[f"{p}.1.blocks.0.0.relative_pos"] = TO.relative_pos_table(C, n_pool, n_pool // st["r"] ** dim, dim)
            sd[f"{p}.2.blocks.0.0.relative_pos"] = TO.relative_pos_table(C, n_win, n_win, dim)
    for k in sd:
        if sd[k].dtype.is_floating_point:
            sd[k] = sd[k].clone().requires_grad_(not k.endswith(("running_mean", "running_var", "relative_pos")))
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    replay = TO.ReplayKnn(H.golden_knn_list(npz), tol=1e-4)
    outs = TO.nextou_forward(sd, x, cfg["patch"], cfg["strides"], training=True, knn=replay)
    assert replay.pos == len(replay.recorded)
    for i, o in enumerate(outs):
        ref = torch.from_numpy(npz[f"out/{i}"])
        got = o.detach() if o.numel() <= 70000 else o.detach().reshape(-1)[::97]
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, rtol=1e-3, atol=2e-3), (i, (got - ref).abs().max())
    loss = sum(o.float().mean() for o in outs)
    assert abs(loss.item() - float(npz["loss"])) < 1e-3
    loss.backward()
    checked = 0
    for k in npz.files:
        if k.startswith("grad/"):
            gref = torch.from_numpy(npz[k])
            gmine = sd[k[5:]].grad.reshape(-1)[:256]
            scale = gref.abs().max().item() + 1e-8
            assert (gmine - gref).abs().max().item() <= 2e-2 * scale + 1e-6, (k, (gmine - gref).abs().max(), scale)
            checked += 1
    assert checked > 20


def test_torch_oracle_eval_forward_matches_reference_golden():
    """Inference path of the oracle (eval-mode BatchNorm with running statistics, deep supervision off) against the
    UNMODIFIED reference in eval mode (tests/golden/model_mini3d_reference_eval.npz, generated by
    oracle/make_golden.py::gen_model_eval): this pins the checker the GPU inference tests compare with."""
    cfg = H.MINI3D
    base = H.golden_model("model_mini3d_reference.npz")
    ev = H.golden_model("model_mini3d_reference_eval.npz")
    sd = H.golden_state_dict(base)
    n_stats = 0
    for k in ev.files:
        if k.startswith("sd/"):
            assert k[3:] in sd, k
            sd[k[3:]] = torch.from_numpy(ev[k])            # running statistics after the momentum-1 forward
            n_stats += 1
    assert n_stats > 100
    dim = len(cfg["patch"])
    plan = TO.derive_plan(cfg["patch"], cfg["strides"])
    for s in range(plan["gnn_from"], len(cfg["feats"])):
        st = plan["stages"][s]
        C = cfg["feats"][s]
        n_pool = int(np.prod(st["shape"])) // int(np.prod(st["pool_size"]))
        n_win = int(np.prod(st["window"]))
        prefixes = [f"encoder.stages.{s}.0"]
        j = len(cfg["feats"]) - 2 - s
        if j >= 0:
            prefixes.append(f"decoder.stages.{j}")
        for p in prefixes:
            sd[f"{p}.1.blocks.0.0.relative_pos"] = TO.relative_pos_table(C, n_pool, n_pool // st["r"] ** dim, dim)
            sd[f"{p}.2.blocks.0.0.relative_pos"] = TO.relative_pos_table(C, n_win, n_win, dim)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 1, *cfg["patch"], generator=g)
    replay = TO.ReplayKnn(H.golden_knn_list(ev), tol=1e-4)
    with torch.no_grad():
        y = TO.nextou_forward(sd, x, cfg["patch"], cfg["strides"], deep_supervision=False, training=False, knn=replay)
    assert replay.pos == len(replay.recorded) == 14
    assert tuple(y.shape) == tuple(int(v) for v in ev["out_shape"])
    ref = torch.from_numpy(ev["out/0"])
    got = y.reshape(-1)[::97]
    # same bound as the train-mode fixture: teacher-forced graphs, the max-unpool positions may still flip on 1e-7
    assert torch.allclose(got, ref, rtol=1e-3, atol=2e-3), (got - ref).abs().max()
