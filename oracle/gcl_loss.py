"""ORACLE (test infrastructure, NOT product code): GCL group-wise contrastive losses on CPU.

Restates /root/reference/lib/colocation_trainer.py:
  finest_contrastive_loss    :430-535   (script default: square_loss=True, block_finest_gradient=False,
                                         use_pair_group_positive_loss=False, use_hard_negative=True)
  location_contrastive_loss  :734-809   (selected when finest_weight == 0, :425-428)
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
