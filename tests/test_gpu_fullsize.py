"""GPU tests at BASELINE.json's full sizes, where the CPU oracle is too slow to be the checker: size-independent
properties of the domain (idempotence, symmetry of kernel maps, strided<->transposed duality, linearity and
permutation-equivariance of the convolution, optimality of the returned neighbours) plus cross-checks between the two
independent CUDA implementations (exact-fp32 CUDA-core kernels vs tcgen05 tensor-core kernels).

  config 1/2  KITTI-shape 64-beam scans (~130k points, 0.3 m voxels)
  config 3    nuScenes-shape 32-beam scans, batch of 8 pairs = 16 clouds
  config 5    voxel-size sweep 0.1-0.5 m on a dense synthetic surface (up to ~1M voxels)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def G():
  from gcl_b200 import MinkowskiEngine as ME, engine, matching, ops, synth
  from gcl_b200.resunet import make_models

  class NS:
    pass

  ns = NS()
  ns.ME, ns.ops, ns.engine, ns.matching, ns.synth, ns.make_models = ME, ops, engine, matching, synth, make_models
  return ns


@pytest.fixture(scope="module")
def kitti_pair(G):
  x0, x1, T = G.synth.scan_pair(scene_seed=5, pair_seed=5)
  return x0, x1, T


def _check_voxelize_properties(G, xyz_np, ptr, voxel):
  xyz = torch.from_numpy(xyz_np).to(DEV)
  cm, umap, inv = G.ops.voxelize(xyz, voxel, ptr, return_inverse=True)
  V = cm.n
  # (1) unique_map strictly ascending first-occurrence indices
  assert bool((umap[1:] > umap[:-1]).all())
  # (2) the kept point of every voxel lies in it, and every point maps to the voxel of its floor()ed coordinate
  # NB: computed on the CPU like the reference's loaders do (IEEE fp32 division).  torch's CUDA `tensor / python_float`
  # multiplies by the reciprocal instead and differs on voxel boundaries; the kernel follows the reference (CPU) semantics.
  disc = torch.floor(torch.from_numpy(xyz_np) / voxel).int().to(DEV)
  assert torch.equal(cm.coords[:, 1:], disc[umap])
  assert torch.equal(cm.coords[inv.long()][:, 1:], disc)
  # (3) first occurrence: no earlier point falls into the same voxel  <=>  umap[v] == min{i : inv[i] == v}
  first = torch.full((V,), xyz.shape[0], dtype=torch.int64, device=DEV)
  first.scatter_reduce_(0, inv.long(), torch.arange(xyz.shape[0], device=DEV), reduce="amin")
  assert torch.equal(first, umap)
  # (4) rows are unique (count equals torch.unique's) and the table finds every row at its own index
  b = torch.searchsorted(ptr.to(DEV)[1:].contiguous(), torch.arange(xyz.shape[0], device=DEV), right=True).int()
  full = torch.cat([b[:, None], disc], 1)
  assert V == torch.unique(full, dim=0).shape[0]
  assert torch.equal(G.ops.hash_query(cm, cm.coords), torch.arange(V, dtype=torch.int32, device=DEV))
  # (5) idempotence: voxelising the voxel centres returns the same map in the same order
  centres = (cm.coords[:, 1:].float() + 0.5) * voxel
  starts = torch.searchsorted(cm.coords[:, 0].contiguous(), torch.arange(ptr.numel(), dtype=torch.int32, device=DEV))
  cm2, umap2 = G.ops.voxelize(centres, voxel, starts.cpu().long())
  assert torch.equal(cm2.coords, cm.coords) and torch.equal(umap2, torch.arange(V, device=DEV))
  return cm


def _check_maps_properties(G, cm1):
  cm2 = G.ops.stride_map(cm1, 2)
  cm4 = G.ops.stride_map(cm2, 2)
  for cm, parent in ((cm2, cm1), (cm4, cm2)):
    s = cm.tensor_stride
    assert bool((cm.coords[:, 1:] % s == 0).all())                               # lattice
    assert cm.n == torch.unique(cm.coords, dim=0).shape[0]                       # unique
    par = parent.coords.clone()
    par[:, 1:] = torch.div(par[:, 1:], s, rounding_mode="floor") * s
    assert bool((G.ops.hash_query(cm, par) >= 0).all())                          # covers every parent row
  for cm in (cm1, cm2):
    nbr = G.ops.kernel_map(cm, cm, 3)
    n = cm.n
    ar = torch.arange(n, dtype=torch.int32, device=DEV)
    assert torch.equal(nbr[:, 13], ar)                                           # centre offset = identity
    # symmetry: nbr[o,k] = i  <=>  nbr[i,26-k] = o
    o, k = torch.nonzero(nbr >= 0, as_tuple=True)
    i = nbr[o, k].long()
    assert torch.equal(nbr[i, 26 - k].long(), o)
    # geometric meaning: coord[i] == coord[o] + off_k * stride
    kk = k
    off = torch.stack([kk % 3 - 1, (kk // 3) % 3 - 1, kk // 9 - 1], 1).int() * cm.tensor_stride
    assert torch.equal(cm.coords[i][:, 1:], cm.coords[o][:, 1:] + off)
    assert torch.equal(cm.coords[i][:, 0], cm.coords[o][:, 0])                   # never across clouds
  down = G.ops.kernel_map(cm1, cm2, 3)
  up = G.ops.kernel_map(cm2, cm1, 3, transposed=True)
  c, k = torch.nonzero(down >= 0, as_tuple=True)
  f = down[c, k].long()
  assert torch.equal(up[f, k].long(), c)                                         # strided <-> transposed duality
  assert int((down >= 0).sum()) == int((up >= 0).sum())
  assert bool(((down >= 0).sum(1) >= 1).all()) and bool(((up >= 0).sum(1) >= 1).all())
  return cm2, down, up


def test_kitti_scan_voxelize_and_maps(G, kitti_pair):
  x0, x1, _ = kitti_pair
  xyz = np.concatenate([x0, x1])
  ptr = torch.tensor([0, len(x0), len(x0) + len(x1)])
  cm1 = _check_voxelize_properties(G, xyz, ptr, 0.3)
  assert 20000 < cm1.n < 80000
  _check_maps_properties(G, cm1)


def test_nuscenes_batch16_voxelize_and_maps(G):
  clouds = [G.synth.cast(G.synth.Scene(20 + i // 2), G.synth.NUSCENES, (3.0 * (i % 2), 0.0, 0.02 * i), seed=i) for i in range(16)]
  xyz = np.concatenate(clouds)
  ptr = torch.tensor(np.cumsum([0] + [len(c) for c in clouds]))
  cm1 = _check_voxelize_properties(G, xyz, ptr, 0.3)
  assert int(cm1.coords[:, 0].max()) == 15
  _check_maps_properties(G, cm1)


@pytest.mark.parametrize("voxel", [0.1, 0.2, 0.5])
def test_voxel_sweep_dense_surface(G, voxel):
  pts = G.synth.dense_surface(6_000_000 if voxel == 0.1 else 2_000_000, seed=3)
  ptr = torch.tensor([0, len(pts)])
  cm1 = _check_voxelize_properties(G, pts, ptr, voxel)
  if voxel == 0.1:
    assert cm1.n > 900_000            # the ~1M-voxel stress case of BASELINE config 5
  cm2, down, up = _check_maps_properties(G, cm1)
  # conv at this size: tensor-core kernel vs exact-fp32 kernel, plus linearity of the operator
  torch.manual_seed(0)
  n = cm1.n
  nbr = G.ops.kernel_map(cm1, cm1, 3)
  x, y = torch.randn(n, 32, device=DEV), torch.randn(n, 32, device=DEV)
  W = torch.randn(27, 32, 64, device=DEV) / 30
  ref = G.ops.spconv_fwd(x, W, nbr, n, algo=1)
  srt, perm, mask = G.ops.kernel_map_sort(nbr)
  Wt = G.ops.weights_to_tc(W)
  got = G.ops.spconv_fwd(x, Wt, srt, n, algo=2, row_perm=perm, tile_mask=mask)
  assert ((got - ref).norm() / ref.norm()).item() < 1e-3
  lin = G.ops.spconv_fwd(2.0 * x - 3.0 * y, W, nbr, n, algo=1)
  ref_y = G.ops.spconv_fwd(y, W, nbr, n, algo=1)
  assert ((lin - (2.0 * ref - 3.0 * ref_y)).abs().max() / ref.abs().max()).item() < 1e-5
  # strided conv then transposed conv keep row counts / duality at this size
  z = G.ops.spconv_fwd(x, torch.randn(27, 32, 32, device=DEV) / 30, down, cm2.n, algo=1)
  u = G.ops.spconv_fwd(z, torch.randn(27, 32, 32, device=DEV) / 30, up, n, algo=1)
  assert z.shape == (cm2.n, 32) and u.shape == (n, 32) and bool(torch.isfinite(u).all())


def test_kitti_pair_network_tc_vs_fp32_and_equivariance(G, kitti_pair):
  import bench
  x0, x1, _ = kitti_pair
  model = bench.seeded_model(G.ME)
  xyz = torch.from_numpy(np.concatenate([x0, x1])).to(DEV)
  ptr = torch.tensor([0, len(x0), len(x0) + len(x1)])
  eng_tc = G.engine.ResUNetEngine(model, device=DEV)
  eng_32 = G.engine.ResUNetEngine(model, device=DEV, algo=1)
  f_tc, cm, umap = eng_tc.extract(xyz, 0.3, ptr)
  f_32, cm_b, _ = eng_32.extract(xyz, 0.3, ptr)
  assert torch.equal(cm.coords, cm_b.coords)
  rel = ((f_tc - f_32).norm() / f_32.norm()).item()
  assert rel < 1e-3, rel                                   # north_star tolerance at full KITTI size
  assert (f_tc.norm(dim=1) - 1).abs().max().item() < 1e-5
  # determinism: bit-identical run to run (no atomics on the feature path)
  f_again, _, _ = eng_tc.extract(xyz, 0.3, ptr)
  assert torch.equal(f_tc, f_again)
  # permutation equivariance of the whole network on the voxel rows
  perm = torch.randperm(cm.n, device=DEV)
  cm_p = G.ops.hash_build(cm.coords[perm].contiguous())
  f_p = eng_tc.forward(cm_p, torch.ones(cm.n, 1, device=DEV))
  assert ((f_p - f_tc[perm]).norm() / f_tc.norm()).item() < 1e-3
  # batch separation: the two clouds do not influence each other
  cm_a, _ = G.ops.voxelize(xyz[:len(x0)].contiguous(), 0.3)
  f_a = eng_tc.forward(cm_a, torch.ones(cm_a.n, 1, device=DEV))
  assert ((f_a - f_tc[:cm_a.n]).norm() / f_a.norm()).item() < 1e-3


def test_full_size_nn_optimality_and_mutual(G):
  g = torch.Generator(device="cpu").manual_seed(4)
  A = torch.nn.functional.normalize(torch.randn(20000, 32, generator=g), dim=1).to(DEV)
  B = torch.nn.functional.normalize(torch.randn(19000, 32, generator=g), dim=1).to(DEV)
  B[:5000] = torch.nn.functional.normalize(A[:5000] + 0.05 * torch.randn(5000, 32, device=DEV), dim=1)
  res = {}
  for algo in (1, 2):
    idx01, d01, idx10, d10, *_ = G.ops.nn_search(A, B, both=True, algo=algo)
    res[algo] = (idx01, d01, idx10, d10)
    # optimality against random candidates and exactness of the reported distance
    j = torch.randint(0, B.shape[0], (A.shape[0], 64), device=DEV)
    d_rand = (A[:, None, :] - B[j]).pow(2).sum(-1)
    assert bool((d01[:, None] <= d_rand + 1e-6).all())
    assert torch.allclose(d01, (A - B[idx01]).pow(2).sum(1), atol=2e-6)
    assert torch.allclose(d10, (B - A[idx10]).pow(2).sum(1), atol=2e-6)
  same = (res[1][0] == res[2][0]).float().mean().item()
  assert same > 0.9995                                        # fp32 tiles vs tensor-core split GEMM
  pairs = G.matching.mutual_nn(A, B)
  assert len(pairs) > 4000 and bool((np.diff(pairs[:, 0]) > 0).all())
  i, jj = torch.from_numpy(pairs[:, 0]).to(DEV), torch.from_numpy(pairs[:, 1]).to(DEV)
  assert torch.equal(res[2][0][i], jj) and torch.equal(res[2][2][jj], i)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configurations against the CPU ORACLE directly (round-1 verdict: the full-size tests only compared the CUDA
# path with itself).  The oracle takes ~1 s per KITTI-shape scan, so these are cheap.  Bars (north_star): coordinates and
# unique maps bit-exact, features <= 1e-3 relative in Frobenius norm AND per row (descriptors are unit vectors, so the
# per-row relative error is the norm of the row difference).
# ------------------------------------------------------------------------------------------------------------------
def _oracle_forward(G, model_name, clouds, voxel=0.3, seed=0):
  import bench
  import oracle.me_cpu as OME
  cs, sels = [], []
  for x in clouds:
    t = torch.from_numpy(x)
    _, sel = OME.utils.sparse_quantize(t / voxel, return_index=True)
    cs.append(torch.floor(t[sel] / voxel).int())
    sels.append(sel)
  C, F = OME.utils.sparse_collate(cs, [torch.ones(len(c), 1) for c in cs])
  torch.manual_seed(seed)
  om = G.make_models(OME)[model_name](**bench.MODEL)
  g = torch.Generator().manual_seed(seed + 1)
  for mod in om.modules():
    if isinstance(mod, torch.nn.BatchNorm1d):
      mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
      mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
      mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
      mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
  om.eval()
  torch.set_num_threads(max(1, (__import__("os").cpu_count() or 1)))
  with torch.no_grad():
    ref = om(OME.SparseTensor(F, coordinates=C)).F
  off = np.cumsum([0] + [len(x) for x in clouds])
  umap = torch.cat([s + int(off[i]) for i, s in enumerate(sels)])
  return om, C, umap, ref


def _check_engine_vs_oracle(G, model_name, clouds, half=True):
  om, C, umap_ref, ref = _oracle_forward(G, model_name, clouds)
  eng = G.engine.ResUNetEngine(om, device=DEV, half=half)
  xyz = torch.from_numpy(np.concatenate(clouds)).to(DEV)
  ptr = torch.tensor(np.cumsum([0] + [len(c) for c in clouds]))
  feats, cm, umap = eng.extract(xyz, 0.3, ptr)
  assert torch.equal(cm.coords.cpu(), C)                      # voxel coordinates, row order included: bit-exact
  assert torch.equal(umap.cpu(), umap_ref)                    # unique_map: bit-exact
  d = feats.cpu() - ref
  fro = (d.norm() / ref.norm()).item()
  worst = (d.norm(dim=1) / ref.norm(dim=1)).max().item()
  print(f"[{model_name} half={half}] V={len(ref)} frobenius {fro:.2e} worst row {worst:.2e}")
  assert fro < 1e-3, fro
  assert worst < 1e-3, worst
  eng.check_range()                                            # no fp16 saturation / underflow flagged
  return fro, worst


def test_kitti_pair_fp16_engine_vs_oracle(G, kitti_pair):
  """BASELINE config 2 at full size: one LoKITTI-style pair (2 x ~130k points) through the fp16 tcgen05 engine vs the
  fp32 CPU oracle"""
  x0, x1, _ = kitti_pair
  _check_engine_vs_oracle(G, "ResUNetBN2C", [x0, x1])


def test_kitti_pair_tf32_engine_vs_oracle(G, kitti_pair):
  """same with fp32 activation storage (kind::tf32 operands)"""
  x0, x1, _ = kitti_pair
  _check_engine_vs_oracle(G, "ResUNetBN2C", [x0[:60000], x1[:60000]], half=False)


def test_nuscenes_batch16_engine_vs_oracle(G):
  """BASELINE config 3: 16 nuScenes-shape clouds (8 pairs) in one batch vs the oracle"""
  clouds = [G.synth.cast(G.synth.Scene(20 + i // 2), G.synth.NUSCENES, (3.0 * (i % 2), 0.0, 0.02 * i), seed=i) for i in range(16)]
  _check_engine_vs_oracle(G, "ResUNetBN2C", clouds)


def test_resunet_fatbn_engine_vs_oracle(G):
  """the variant GCL's scripts train / evaluate (scripts/train_gcl_kitti.sh:13, generalization_ETH/evaluate.py:227):
  128-wide decoder, 160-channel conv1_tr input"""
  clouds = [G.synth.cast(G.synth.Scene(40 + i), G.synth.NUSCENES, (2.0 * i, 0.0, 0.0), seed=i) for i in range(2)]
  _check_engine_vs_oracle(G, "ResUNetFatBN", clouds)
