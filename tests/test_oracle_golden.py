"""CPU: the oracle against the committed golden vectors, which were produced by the REFERENCE's own Python
(tests/golden/make_golden.py: model/resunet.py, lib/eval.py, lib/colocation_trainer.py imported in the build container)."""
import os

import numpy as np
import torch

import oracle.me_cpu as OME
from oracle import gcl_loss as oloss
from oracle import matching as omatch
from gcl_b200.resunet import make_models
from helpers import numpy_seeded_weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_resunet_restatement_matches_reference_model_output():
  g = np.load(os.path.join(GOLD, "resunet_bn2c.npz"))
  C = torch.from_numpy(g["coords"])
  model = make_models(OME)["ResUNetBN2C"](1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  numpy_seeded_weights(model, seed=int(g["weight_seed"])).eval()
  with torch.no_grad():
    out = model(OME.SparseTensor(torch.ones(len(C), 1), coordinates=C)).F
  ref = torch.from_numpy(g["feats_out"])
  assert out.shape == ref.shape
  assert (out - ref).abs().max() < 2e-5          # same graph, same operators; only BLAS summation order may differ


def test_nn_matches_reference_find_nn_gpu():
  g = np.load(os.path.join(GOLD, "nn.npz"))
  F0, F1 = torch.from_numpy(g["F0"]), torch.from_numpy(g["F1"])
  idx, d = omatch.find_nn(F0, F1, nn_max_n=250, return_distance=True)
  assert np.array_equal(idx.numpy(), g["idx"]) and np.allclose(d.numpy(), g["dist"], atol=1e-7)
  idx2, d2 = omatch.find_nn(F0, F1, return_distance=True, dist_type="L2")
  assert np.array_equal(idx2.numpy(), g["idx_l2"]) and np.allclose(d2.numpy(), g["dist_l2"], atol=1e-6)
  pairs, _, _ = omatch.mutual_nn(F0, F1)
  assert np.array_equal(pairs, g["mutual_pairs"])


def test_gcl_loss_matches_reference_trainer():
  g = np.load(os.path.join(GOLD, "gcl_loss.npz"))
  for name, square, finest in (("finest_sq", True, True), ("finest_l2", False, True), ("location", False, False)):
    F = torch.from_numpy(g["F"]).clone().requires_grad_(True)
    np.random.seed(5)
    sel = oloss.draw_selections(len(g["group"]), len(F), 256, 512)
    pos, fin, neg = oloss.group_contrastive_loss(F, g["group"], g["index"], g["index_hash"], g["finest_flag"], *sel,
                                                 square_loss=square, with_finest=finest)
    (1.0 * pos + 0.5 * fin + 2.0 * neg).backward()
    want = g[name + "_losses"]
    assert np.allclose([pos.item(), fin.item(), neg.item()], want, rtol=1e-5, atol=1e-7), name
    assert torch.allclose(F.grad, torch.from_numpy(g[name + "_grad"]), atol=1e-7, rtol=1e-4), name


def test_exhaustive_hash_symmetry():
  h = oloss.exhaustive_hash([[3, 7, 9], [1, 2]], 100)
  assert sorted(h.tolist()) == sorted([3 * 100 + 7, 3 * 100 + 9, 7 * 100 + 9, 1 * 100 + 2])
  assert oloss.neg_hash([7], [3], 100)[0] == oloss.neg_hash([3], [7], 100)[0] == 307


def _group_cases():
  g = np.load(os.path.join(GOLD, "groups.npz"))
  for c in range(4):
    nbs, Ts, j = [], [], 0
    while f"c{c}_nb{j}" in g:
      nbs.append(g[f"c{c}_nb{j}"]); Ts.append(g[f"c{c}_T{j}"]); j += 1
    K = int(g[f"c{c}_K"])
    yield c, g[f"c{c}_centre"], nbs, Ts, float(g[f"c{c}_radius"]), (None if K < 0 else K), g


def test_colocation_groups_match_reference_function():
  """oracle/groups.py against the fixture produced by the reference's own get_matching_indices_colocation /
  get_matching_indices (util/pointcloud.py:53-132) -- see tests/golden/make_golden_groups.py for what is pinned"""
  from oracle import groups as og
  for c, centre, nbs, Ts, radius, K, g in _group_cases():
    grp, idx, fin = og.colocation_groups(centre, nbs, Ts, radius, K)
    assert np.array_equal(grp, g[f"c{c}_group"]) and np.array_equal(idx, g[f"c{c}_index"]) and np.array_equal(fin, g[f"c{c}_finest"])
    assert fin.sum() == len(grp) and grp.sum() == len(idx)
    if c == 0:
      assert np.array_equal(og.matching_indices(nbs[0], centre, Ts[0], radius, K), g[f"c{c}_pairs"])


def test_sc2pcr_restatement_matches_reference_matcher():
  """oracle/sc2pcr.py (the next row to be built) against the transforms the reference's own Matcher.SC2_PCR
  produced on CPU (tests/golden/make_golden_sc2pcr.py), and against the ground-truth motion of the synthetic correspondences"""
  from oracle import sc2pcr as osc
  g = np.load(os.path.join(GOLD, "sc2pcr.npz"))
  for c in range(3):
    T = osc.sc2_pcr(torch.from_numpy(g[f"c{c}_src"])[None], torch.from_numpy(g[f"c{c}_tgt"])[None])[0].numpy()
    assert np.abs(T - g[f"c{c}_trans"]).max() < 1e-4
    assert np.abs(T - g[f"c{c}_gt"]).max() < 1e-2


def test_pair_hashes_vs_reference_golden():
  """a17: oracle restatement of util/misc.py:29-40 against keys produced by the reference's own functions
  (tests/golden/make_golden_pair_hash.py)"""
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pair_hash.npz"))
  for c in range(3):
    grp, idx, M = g[f"c{c}_group"], g[f"c{c}_index"], int(g[f"c{c}_M"])
    split = np.split(idx, np.cumsum(grp)[:-1])
    assert np.array_equal(oloss.exhaustive_hash(split, M), g[f"c{c}_keys"])
  assert np.array_equal(oloss.neg_hash(g["neg_i1"], g["neg_i2"], int(g["neg_M"])), g["neg_keys"])
  loss = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gcl_loss.npz"))
  assert np.array_equal(loss["index_hash"], g["c2_keys"])


def test_hardest_contrastive_oracle_vs_reference_golden():
  """f4: oracle/metrics.hardest_contrastive against values + gradients produced by the reference's own
  HardestContrastiveLossTrainer.contrastive_hardest_negative_loss (tests/golden/make_golden_metrics.py)"""
  from oracle import metrics as omet
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hardest_loss.npz"))
  a, b = torch.from_numpy(g["F0"]).requires_grad_(True), torch.from_numpy(g["F1"]).requires_grad_(True)
  pos, neg = omet.hardest_contrastive(a, b, g["pairs"], g["sel0"], g["sel1"], g["pos_sel"])
  (pos + neg).backward()
  assert abs(float(pos) - g["losses"][0]) < 1e-6 and abs(float(neg) - g["losses"][1]) < 1e-6
  assert np.abs(a.grad.numpy() - g["g0"]).max() < 1e-7 and np.abs(b.grad.numpy() - g["g1"]).max() < 1e-7


def test_metric_formulas_known_answers():
  from oracle import metrics as omet
  T = torch.eye(4)
  E = torch.eye(4)
  ang = np.deg2rad(3.0)
  E[:3, :3] = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
  E[:3, 3] = torch.tensor([0.3, 0.4, 0.0])
  rte, rre = omet.rte_rre(E, T)
  assert abs(rte - 0.5) < 1e-6 and abs(np.rad2deg(rre) - 3.0) < 1e-3
  x = torch.randn(100, 3)
  assert omet.evaluate_hit_ratio(x, x @ E[:3, :3].t() + E[:3, 3], E, 0.1) == 1.0


CIRCLE_VARIANTS = {"sq_block": (True, True, False), "sq_open": (True, False, False), "l2_block": (False, True, False),
                   "l2_open": (False, False, False), "sq_pair": (True, True, True)}


def test_circle_loss_oracle_vs_reference_golden():
  """f4, the circle-loss head: oracle/gcl_loss.circle_loss against values + gradients produced by the reference's own
  FinestContrastiveLossTrainer.location_circle_loss (tests/golden/make_golden_circle.py), five settings"""
  g = np.load(os.path.join(GOLD, "circle_loss.npz"))
  for name, (square, block, pair) in CIRCLE_VARIANTS.items():
    F = torch.from_numpy(g["F"]).clone().requires_grad_(True)
    np.random.seed(5)
    pos, fin, neg = oloss.circle_loss(F, g["group"], g["index"], g["finest_flag"], g["points"], g["batch_lengths"].tolist(),
                                      max_pos_cluster=256, square_loss=square, block_finest_gradient=block,
                                      use_pair_group_positive_loss=pair)
    (1.0 * pos + 0.5 * fin + 2.0 * neg).backward()
    assert np.allclose([pos.item(), fin.item(), neg.item()], g[name + "_losses"], rtol=1e-5, atol=1e-7), name
    assert torch.allclose(F.grad, torch.from_numpy(g[name + "_grad"]), atol=1e-7, rtol=1e-4), name

