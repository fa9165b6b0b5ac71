"""ORACLE (test infrastructure, NOT product code): GCL group-wise contrastive losses on CPU.

Restates /root/reference/lib/colocation_trainer.py:
  finest_contrastive_loss    :430-535   (script default: square_loss=True, block_finest_gradient=False,
                                         use_pair_group_positive_loss=False, use_hard_negative=True)
  location_contrastive_loss  :734-809   (selected when finest_weight == 0, :425-428)
  location_circle_loss       :538-681   (the circle-loss head; golden tests/golden/circle_loss.npz from the reference's own method,
                                         tests/golden/make_golden_circle.py)
and the pair-hash helpers /root/reference/util/misc.py:29-40 (_exhaustive_hash, _neg_hash).

The host-side random selections (pos_sel :456-459, sel_hn1/sel_hn2 :506-507) are drawn from `rng` with the
same calls in the same order as the reference, so a seeded run reproduces the reference's selection.
Pinned against the reference functions imported in the build container (tests/test_reference_live.py;
golden vectors tests/golden/gcl_loss.npz, pair_hash.npz; tests/test_oracle_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def exhaustive_hash(index_split, M):
  out = []
  for idx in index_split:
    idx = np.asarray(idx, dtype=np.int64)
    for i in range(len(idx) - 1):
      a, b = idx[i], idx[i + 1:]
      out.append(np.minimum(a + b * M, a * M + b))
  return np.concatenate(out) if out else np.zeros(0, np.int64)


def neg_hash(i1, i2, M):
  i1, i2 = np.asarray(i1, np.int64), np.asarray(i2, np.int64)
  return np.minimum(i1 * M + i2, i1 + i2 * M)


def draw_selections(n_groups, n_rows, max_pos_cluster, max_hn_samples, rng=np.random):
  """Same np.random call sequence as colocation_trainer.py:456-459 and :506-507."""
  if n_groups > max_pos_cluster:
    pos_sel = rng.choice(n_groups, max_pos_cluster, replace=False)
  else:
    pos_sel = np.arange(n_groups)
  sel_hn1 = rng.choice(n_rows, min(n_rows, max_hn_samples), replace=False)
  sel_hn2 = rng.choice(n_rows, min(n_rows, max_hn_samples), replace=False)
  return pos_sel, sel_hn1, sel_hn2


def group_contrastive_loss(F_out, group, index, index_hash, finest_flag, pos_sel, sel_hn1, sel_hn2,
                           pos_thresh=0.1, neg_thresh=1.4, finest_thresh=0.2, square_loss=True,
                           with_finest=True):
  """Returns (pos_loss, finest_loss, neg_loss) as 0-d tensors attached to F_out's graph."""
  N = len(F_out)
  group = [int(g) for g in group]
  starts = np.concatenate([[0], np.cumsum(group)])
  index = torch.as_tensor(index, dtype=torch.int64)
  finest_flag = torch.as_tensor(finest_flag, dtype=torch.bool)
  pos_loss = F_out.new_zeros(())
  finest_loss = F_out.new_zeros(())
  for g in pos_sel:
    s, e = starts[g], starts[g + 1]
    fs = F_out[index[s:e]]
    mu = fs.mean(0)
    d2 = (mu - fs).pow(2).sum(-1)
    if square_loss:
      pos_loss = pos_loss + F.relu(d2.mean() - pos_thresh)
    else:
      pos_loss = pos_loss + F.relu(torch.sqrt(d2 + 1e-7).mean() - pos_thresh)
    if with_finest:
      ff = fs[finest_flag[s:e]][0]
      e2 = (mu - ff).pow(2).sum()
      finest_loss = finest_loss + (F.relu(e2 - finest_thresh) if square_loss
                                   else F.relu(torch.sqrt(e2 + 1e-7) - finest_thresh))
  pos_loss, finest_loss = pos_loss / len(pos_sel), finest_loss / len(pos_sel)

  sub1, sub2 = F_out[torch.as_tensor(sel_hn1)], F_out[torch.as_tensor(sel_hn2)]
  D = torch.sqrt((sub1.unsqueeze(1) - sub2.unsqueeze(0)).pow(2).sum(2) + 1e-7)
  Dmin, Dind = D.min(1)
  closest = np.asarray(sel_hn2)[Dind.numpy()]
  mask_self = np.asarray(sel_hn1) != closest
  mask = ~np.isin(neg_hash(sel_hn1, closest, N), np.asarray(index_hash))
  neg = F.relu(neg_thresh - Dmin[torch.from_numpy(mask & mask_self)]).pow(2)
  return pos_loss, finest_loss, neg.mean()


def circle_loss(F_out, group, index, finest_flag, points, batch_lengths, max_pos_cluster=256, pos_thresh=0.1, neg_thresh=1.4,
                finest_thresh=0.2, square_loss=True, block_finest_gradient=True, use_pair_group_positive_loss=False, log_scale=16,
                safe_radius=0.75, rng=np.random):
  """lib/colocation_trainer.py:538-681 (`location_circle_loss`): the circle-loss head.  One loop iteration per selected group,
  the reference's RNG calls in the reference's order (:562 group selection, :598 the positive pair of a group).
  Returns (pos_loss, finest_loss, neg_loss) as torch scalars attached to F_out."""
  group = np.asarray(group, np.int64)
  index = torch.as_tensor(index, dtype=torch.int64)
  flag = torch.as_tensor(finest_flag, dtype=torch.bool)
  starts = np.concatenate([[0], np.cumsum(group)])
  G = len(group)
  pos_sel = np.sort(rng.choice(G, max_pos_cluster, replace=False)) if G > max_pos_cluster else np.arange(G)   # :561-564
  S, C = len(pos_sel), F_out.shape[1]
  ends = np.cumsum(np.asarray(batch_lengths, np.float64))            # :573-578 running ends of the batch items
  counts = np.zeros(len(batch_lengths), np.int64)
  coords = torch.zeros((S, 3), dtype=torch.float32)
  means = torch.zeros((S, C), dtype=torch.float32)
  pos_total, fin_total = 0, 0

  def gap(a, b, thr):                                               # squared or eps-guarded L2 distance minus a threshold
    d2 = (a - b).pow(2).sum(-1)
    return (d2 if square_loss else torch.sqrt(d2 + 1e-7)) - thr

  def soft_lse(x):                                                  # :614-619 / :636-641: self-weighted log-sum-exp, softplus
    w = torch.clamp(x, min=0).detach()
    return F.softplus(torch.logsumexp(log_scale * x * w, dim=-1)) / log_scale

  for n, g in enumerate(pos_sel):
    members = index[starts[g]:starts[g + 1]]
    fl = flag[starts[g]:starts[g + 1]]
    fs = F_out[members]
    coords[n] = torch.as_tensor(points[int(members[0])], dtype=torch.float32)          # :587
    mean = fs.mean(0)
    means[n] = mean                                                                     # :589
    counts[int(np.sum(int(members[0]) > ends))] += 1                                    # :592-594
    if use_pair_group_positive_loss:                                                    # :595-607
      a, b = rng.choice(len(fs), 2, replace=False)
      pos_total = pos_total + F.softplus(gap(fs[a], fs[b], pos_thresh))
    else:                                                                               # :608-619
      pos_total = pos_total + soft_lse(gap(mean, fs, pos_thresh / 2))
    anchor = fs[fl][0]                                                                  # first finest member (:624 / :631)
    if block_finest_gradient:
      fin_total = fin_total + soft_lse(gap(fs[~fl], anchor.detach(), finest_thresh))
    else:
      fin_total = fin_total + soft_lse(gap(fs, anchor, finest_thresh))
  pos_loss, fin_loss = pos_total / S, fin_total / S
  same_item = torch.zeros((S, S), dtype=torch.bool)                                     # :647-652 diagonal blocks by COUNT
  s0 = 0
  for c in counts:
    same_item[s0:s0 + c, s0:s0 + c] = True
    s0 += int(c)

  def sqdist(x, normalised):                                                            # util/misc.py:7-26
    d = -2 * x @ x.T
    d = d + 2 if normalised else d + (x ** 2).sum(-1)[:, None] + (x ** 2).sum(-1)[None, :]
    return torch.clamp(d, min=1e-12)

  cd = torch.sqrt(sqdist(coords, False))
  fd = torch.sqrt(sqdist(means, True))
  neg_mask = (cd > safe_radius) & same_item                                             # :666
  has_neg = (neg_mask.sum(-1) > 0)
  w = torch.clamp(neg_thresh - (fd + 1e5 * (~neg_mask).float()), min=0).detach()        # :669-671
  rows = F.softplus(torch.logsumexp(log_scale * (neg_thresh - fd) * w, dim=-1)) / log_scale
  return pos_loss, fin_loss, rows[has_neg].mean()

