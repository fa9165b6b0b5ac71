"""GPU parity tests: libgclb200 (through the C ABI / the public Python surface) against the CPU oracle on the
same seeded inputs.  Bit-exact for integer work (voxel indices, hash lookups, kernel maps, NN indices away from
ties); features within 1e-3 relative (north_star tolerance; fp32 path is held to 1e-5)."""
import os

import numpy as np
import pytest
import torch

import oracle.me_cpu as OME
from oracle import gcl_loss as oloss
from oracle import matching as omatch

pytestmark = pytest.mark.gpu

FEAT_TOL = 1e-3      # north_star: output features within 1e-3 relative (tensor-core path, fp32 accumulate)
FP32_TOL = 2e-5      # exact-fp32 kernels only differ from the oracle by summation order


@pytest.fixture(scope="module")
def G():
  import gcl_b200
  from gcl_b200 import MinkowskiEngine as ME, engine, loss, matching, ops, synth
  from gcl_b200.resunet import make_models

  class NS:
    pass

  ns = NS()
  ns.ME, ns.ops, ns.engine, ns.matching, ns.loss, ns.synth, ns.make_models = ME, ops, engine, matching, loss, synth, make_models
  ns.dev = torch.device("cuda:0")
  return ns


def _rel(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _random_cloud(seed, n=4000, extent=15.0):
  rng = np.random.RandomState(seed)
  pts = np.concatenate([rng.uniform(-extent, extent, (n, 2)), rng.normal(0, 0.4, (n, 1))], 1)
  return pts.astype(np.float32)


def _oracle_voxelize(clouds, voxel):
  cs, sels, base = [], [], 0
  for x in clouds:
    t = torch.from_numpy(x)
    _, sel = OME.utils.sparse_quantize(t / voxel, return_index=True)
    cs.append(torch.floor(t[sel] / voxel).int())
    sels.append(sel + base)
    base += len(x)
  C, _ = OME.utils.sparse_collate(cs, [torch.ones(len(c), 1) for c in cs])
  return C, torch.cat(sels)


# ----------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("voxel", [0.3, 0.1, 0.05])
def test_voxelize_bit_exact(G, voxel):
  clouds = [_random_cloud(1), _random_cloud(2, 2500), np.zeros((0, 3), np.float32), _random_cloud(3, 100)]
  clouds[0][:50] = clouds[0][50:100]            # exact duplicates
  clouds[1][:, 0] -= 40.0                       # negative coordinates
  C_ref, sel_ref = _oracle_voxelize(clouds, voxel)
  xyz = torch.from_numpy(np.concatenate(clouds)).to(G.dev)
  ptr = torch.tensor(np.cumsum([0] + [len(c) for c in clouds]), dtype=torch.int64)
  cm, umap, inv = G.ops.voxelize(xyz, voxel, ptr, return_inverse=True)
  assert torch.equal(umap.cpu(), sel_ref)
  assert torch.equal(cm.coords.cpu(), C_ref)
  # inverse map: every point lands on the row of its voxel
  assert torch.equal(cm.coords[inv.long()][:, 1:].cpu(), torch.floor(xyz.cpu() / voxel).int())
  # the voxeliser's table is usable as the coordinate hash
  assert torch.equal(G.ops.hash_query(cm, cm.coords).cpu(), torch.arange(cm.n, dtype=torch.int32))


def test_voxelize_fp32_division_matches_torch(G):
  # coordinates chosen to sit on voxel boundaries where x/v vs x*(1/v) differ
  v = 0.3
  k = np.arange(-2000, 2000, dtype=np.float32)
  x = np.stack([k * np.float32(v), k * np.float32(v) + 1e-6, -k * np.float32(v)], 1).astype(np.float32)
  C_ref, sel_ref = _oracle_voxelize([x], v)
  cm, umap = G.ops.voxelize(torch.from_numpy(x).to(G.dev), v)
  assert torch.equal(umap.cpu(), sel_ref) and torch.equal(cm.coords.cpu(), C_ref)


def test_voxelize_range_error_is_loud(G):
  from gcl_b200._lib import GclbError
  x = torch.tensor([[0.0, 0.0, 0.0], [1e9, 0.0, 0.0]], device=G.dev)
  with pytest.raises(GclbError):
    G.ops.voxelize(x, 0.3)
  with pytest.raises(GclbError):
    G.ops.voxelize(torch.tensor([[float("nan"), 0.0, 0.0]], device=G.dev), 0.3)


def test_sparse_quantize_api(G):
  pts = torch.from_numpy(_random_cloud(5, 3000))
  for inp in (pts / 0.3, (pts / 0.3).double().numpy(), torch.floor(pts / 0.3).int()):
    ref = OME.utils.sparse_quantize(inp, return_index=True, return_inverse=True)
    got = G.ME.utils.sparse_quantize(inp, return_index=True, return_inverse=True)
    for r, g in zip(ref, got):
      assert type(r) is type(g)
      assert np.array_equal(np.asarray(r), np.asarray(g))
  um = G.ME.utils.sparse_quantize(pts / 0.3, return_maps_only=True)
  assert torch.equal(um, OME.utils.sparse_quantize(pts / 0.3, return_maps_only=True))


def test_hash_build_query_and_duplicates(G):
  from gcl_b200._lib import GclbError
  C_ref, _ = _oracle_voxelize([_random_cloud(7)], 0.2)
  cm = G.ops.hash_build(C_ref.to(G.dev))
  q = torch.cat([C_ref[::3], C_ref[:50] + torch.tensor([0, 1000, 0, 0], dtype=torch.int32)])
  want = torch.cat([torch.arange(0, len(C_ref), 3), torch.full((50,), -1)]).int()
  assert torch.equal(G.ops.hash_query(cm, q.to(G.dev)).cpu(), want)
  with pytest.raises(GclbError):
    G.ops.hash_build(torch.cat([C_ref, C_ref[:1]]).to(G.dev))
  empty = G.ops.hash_build(torch.zeros((0, 4), dtype=torch.int32, device=G.dev))
  assert G.ops.hash_query(empty, q.to(G.dev)).eq(-1).all()


# ----------------------------------------------------------------------------------------------- K2
def test_stride_maps_and_kernel_maps_bit_exact(G):
  clouds = [_random_cloud(11, 6000), _random_cloud(12, 3000)]
  clouds[1][:, :2] -= 25.0
  C_ref, _ = _oracle_voxelize(clouds, 0.3)
  cm1 = G.ops.hash_build(C_ref.to(G.dev))
  ref = {1: C_ref.numpy()}
  cms = {1: cm1}
  for s in (2, 4, 8):
    ref[s] = OME.stride_coords(ref[s // 2], s)
    cms[s], parent = G.ops.stride_map(cms[s // 2], 2, return_parent_rows=True)
    assert cms[s].tensor_stride == s
    assert np.array_equal(cms[s].coords.cpu().numpy(), ref[s])         # same rows in the same (canonical) order
    lk = OME._Lookup(ref[s])
    par = ref[s // 2].astype(np.int64).copy()
    par[:, 1:] = np.floor_divide(par[:, 1:], s) * s
    assert np.array_equal(parent.cpu().numpy(), lk.query(par))
  for s in (1, 2, 4, 8):
    for ks in ((3, 5) if s == 1 else (3,)):
      want = OME.build_neighbor_table(ref[s], ref[s], OME.kernel_offsets(ks, s))
      got, cnt = G.ops.kernel_map(cms[s], cms[s], ks, count_pairs=True)
      assert np.array_equal(got.cpu().numpy(), want)
      assert np.array_equal(cnt.cpu().numpy(), (want >= 0).sum(0))
  for s in (1, 2, 4):
    down = OME.build_neighbor_table(ref[s], ref[2 * s], OME.kernel_offsets(3, s))
    up = OME.build_neighbor_table(ref[2 * s], ref[s], -OME.kernel_offsets(3, s))
    assert np.array_equal(G.ops.kernel_map(cms[s], cms[2 * s], 3).cpu().numpy(), down)
    assert np.array_equal(G.ops.kernel_map(cms[2 * s], cms[s], 3, transposed=True).cpu().numpy(), up)
    # pair lists (ME layout), canonical order
    i, o, off = G.ops.kernel_map_pairs(G.ops.kernel_map(cms[s], cms[2 * s], 3))
    pairs = OME.neighbor_table_to_pairs(down)
    off = off.tolist()
    for k, (pi, po) in enumerate(pairs):
      assert np.array_equal(i[off[k]:off[k + 1]].cpu().numpy(), pi)
      assert np.array_equal(o[off[k]:off[k + 1]].cpu().numpy(), po)


# ----------------------------------------------------------------------------------------------- K3
CONV_SHAPES = [(1, 32, 5), (32, 32, 3), (32, 64, 3), (64, 64, 3), (128, 128, 3), (256, 256, 3), (3, 16, 3), (20, 24, 3)]


@pytest.mark.parametrize("cin,cout,ks", CONV_SHAPES)
def test_spconv_fwd_vs_oracle(G, cin, cout, ks):
  torch.manual_seed(cin * 1000 + cout)
  C_ref, _ = _oracle_voxelize([_random_cloud(21, 3000, 8.0)], 0.3)
  n = len(C_ref)
  nbr = OME.build_neighbor_table(C_ref.numpy(), C_ref.numpy(), OME.kernel_offsets(ks, 1))
  x = torch.randn(n, cin)
  W = torch.randn(ks ** 3, cin, cout) / np.sqrt(cin * ks ** 3)
  ref = OME.sparse_conv_reference(x, W, nbr, n)
  nbr_d = torch.from_numpy(nbr).int().to(G.dev)
  got = G.ops.spconv_fwd(x.to(G.dev), W.to(G.dev), nbr_d, n, algo=1)
  assert _rel(got, ref) < FP32_TOL
  if G.ops.tc_supported(cin, 0, cout, ks ** 3):   # tcgen05 kind::tf32 path, tensor-core weight layout
    Wt = G.ops.weights_to_tc(W.to(G.dev))
    got = G.ops.spconv_fwd(x.to(G.dev), Wt, nbr_d, n, algo=2)
    assert _rel(got, ref) < FEAT_TOL
    assert (got.cpu() - ref).abs().max() < 5e-3 * ref.abs().max()
  # fused epilogue: scale/shift, residual, relu
  sc, sh, res = torch.rand(cout) + 0.5, torch.randn(cout), torch.randn(n, cout)
  ref2 = torch.relu(ref * sc + sh + res)
  got2 = G.ops.spconv_fwd(x.to(G.dev), W.to(G.dev), nbr_d, n, scale=sc.to(G.dev), shift=sh.to(G.dev),
                          residual=res.to(G.dev), relu=True, algo=1)
  assert _rel(got2, ref2) < FP32_TOL


@pytest.mark.parametrize("n,cin,cout", [(128, 32, 32), (1000, 64, 64), (5000, 96, 64), (333, 256, 256), (77, 128, 128)])
def test_tc_dense_pointwise_gemm(G, n, cin, cout):
  """K == 1 (identity map) through the tcgen05 kernel is a dense GEMM: pins descriptors / swizzle / TMEM readout."""
  torch.manual_seed(n)
  x, W = torch.randn(n, cin), torch.randn(1, cin, cout) / np.sqrt(cin)
  ref = x @ W[0]
  got = G.ops.spconv_fwd(x.to(G.dev), G.ops.weights_to_tc(W.to(G.dev)), None, n, algo=2)
  assert _rel(got, ref) < FEAT_TOL
  # exact when the inputs are tf32-representable (integers): proves the only error source is operand rounding
  xi, Wi = torch.randint(-4, 5, (n, cin)).float(), torch.randint(-4, 5, (1, cin, cout)).float()
  got = G.ops.spconv_fwd(xi.to(G.dev), G.ops.weights_to_tc(Wi.to(G.dev)), None, n, algo=2)
  assert torch.equal(got.cpu(), xi @ Wi[0])


def test_tc_persistent_many_tiles(G):
  """more 128-row tiles than SMs (several tiles per persistent CTA), for short (K = 1, 2-3 stages per tile) and long
  (K = 27) tiles: exercises ring wrap-around, accumulator double-buffering and the deferred-arrival drain"""
  torch.manual_seed(14)
  n = 148 * 128 * 5 + 77
  for cin, cout in ((64, 32), (96, 64), (32, 32)):
    x = torch.randint(-3, 4, (n, cin)).float()
    W = torch.randint(-3, 4, (1, cin, cout)).float()
    got = G.ops.spconv_fwd(x.to(G.dev), G.ops.weights_to_tc(W.to(G.dev)), None, n, algo=2)
    assert torch.equal(got.cpu(), x @ W[0])
  cloud = np.concatenate([_random_cloud(70 + i, 30000, 60.0) + np.array([200.0 * i, 0, 0], np.float32) for i in range(3)])
  cm, _ = G.ops.voxelize(torch.from_numpy(cloud).to(G.dev), 0.3)
  assert cm.n > 148 * 128 * 3
  nbr = G.ops.kernel_map(cm, cm, 3)
  x = torch.randint(-2, 3, (cm.n, 64), device=G.dev).float()
  W = torch.randint(-2, 3, (27, 64, 64), device=G.dev).float()
  ref = G.ops.spconv_fwd(x, W, nbr, cm.n, algo=1)           # exact-fp32 kernel; integer data => both are exact
  srt, perm, mask = G.ops.kernel_map_sort(nbr)
  got = G.ops.spconv_fwd(x, G.ops.weights_to_tc(W), srt, cm.n, algo=2, row_perm=perm, tile_mask=mask)
  assert torch.equal(got, ref)
  res = torch.randint(-2, 3, (cm.n, 64), device=G.dev).float()
  got = G.ops.spconv_fwd(x, G.ops.weights_to_tc(W), srt, cm.n, algo=2, row_perm=perm, tile_mask=mask, residual=res, relu=True)
  assert torch.equal(got, torch.relu(ref + res))


def test_tc_fused_l2_normalise_tail(G):
  torch.manual_seed(8)
  n = 3001
  a, b = torch.randn(n, 64), torch.randn(n, 32)
  W1, W2, bias = torch.randn(96, 64) / 10, torch.randn(64, 32) / 8, torch.randn(32)
  y = torch.relu(torch.cat([a, b], 1) @ W1) @ W2 + bias
  ref = y / y.norm(dim=1, keepdim=True)
  h = G.ops.spconv_fwd(a.to(G.dev), G.ops.weights_to_tc(W1.to(G.dev)), None, n, in1=b.to(G.dev), relu=True, algo=2)
  got = G.ops.spconv_fwd(h, G.ops.weights_to_tc(W2.to(G.dev)), None, n, shift=bias.to(G.dev), normalize=True, algo=2)
  assert _rel(got, ref) < FEAT_TOL
  assert (got.norm(dim=1) - 1).abs().max() < 1e-5


def test_kmap_row_bucketing_is_a_pure_reordering(G):
  torch.manual_seed(12)
  C_ref, _ = _oracle_voxelize([_random_cloud(24, 9000, 12.0), _random_cloud(25, 500, 3.0)], 0.3)
  cm = G.ops.hash_build(C_ref.to(G.dev))
  cm2 = G.ops.stride_map(cm, 2)
  for args in ((cm, cm, 3, 1, False), (cm2, cm, 3, 1, True), (cm, cm2, 3, 1, False)):
    nbr, keys = G.ops.kernel_map(*args, with_keys=True)
    srt, perm, mask = G.ops.kernel_map_sort(nbr)
    none, perm2, mask2 = G.ops.kernel_map_sort(nbr, keys, copy=False)      # keys + row masks straight from the build
    assert none is None and torch.equal(perm, perm2) and torch.equal(mask, mask2)
    assert torch.equal(keys[1], ((nbr >= 0).int() << torch.arange(27, device=G.dev, dtype=torch.int32)).sum(1).int())
    n = nbr.shape[0]
    assert torch.equal(torch.sort(perm.long()).values.cpu(), torch.arange(n))
    assert torch.equal(srt, nbr[perm.long()])
    # keys are non-decreasing and the sort is stable inside a bucket
    v = (srt >= 0).cpu().numpy()
    k = np.arange(27)
    ix, iy, iz = k % 3, (k // 3) % 3, k // 9
    key = sum((v[:, sel].any(1).astype(np.int64) << b) for b, sel in enumerate([ix == 0, ix == 2, iy == 0, iy == 2, iz == 0, iz == 2]))
    assert (np.diff(key) >= 0).all()
    p = perm.cpu().numpy()
    same = np.diff(key) == 0
    assert (np.diff(p)[same] > 0).all()
    # the convolution result does not depend on the row order
    x, W = torch.randn(int(srt.max()) + 1, 64, device=G.dev), torch.randn(27, 64, 64, device=G.dev) / 40
    Wt = G.ops.weights_to_tc(W)
    a = G.ops.spconv_fwd(x, Wt, nbr, n, algo=2)
    b = G.ops.spconv_fwd(x, Wt, srt, n, algo=2, row_perm=perm)
    c = G.ops.spconv_fwd(x, Wt, srt, n, algo=2, row_perm=perm, tile_mask=mask)
    d = G.ops.spconv_fwd(x, Wt, nbr, n, algo=2, row_perm=perm, tile_mask=mask, nbr_is_sorted=False)   # table read through perm
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, d)
    want_mask = [int(sum(1 << k for k in range(27) if v[t:t + 128, k].any())) for t in range(0, n, 128)]
    assert mask.tolist() == want_mask


def test_tma_gather4_primitive(G):
  """the TMA tile::gather4 load the conv kernel is built on: 4 arbitrary rows x 32 channels land as consecutive
  SWIZZLE_128B rows (16-byte chunk j of row r at r*128 + ((j ^ r) << 4)); out-of-range rows are zero-filled"""
  import ctypes
  from gcl_b200 import _lib
  lib = _lib.load()
  n, c = 777, 96
  X = torch.arange(n * c, dtype=torch.float32, device=G.dev).reshape(n, c)
  for rows, col in (([5, 17, 700, 3], 32), ([0, -1, 9999, 776], 64)):
    out = torch.full((256,), -1.0, device=G.dev)
    r = (ctypes.c_int32 * 4)(*rows)
    assert lib.gclb_debug_tma_gather4(X.data_ptr(), n, c, 1, col, ctypes.cast(r, ctypes.c_void_p), out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    o = out.cpu().reshape(8, 8, 4)          # [smem row, 16-byte chunk, float]
    for i, row in enumerate(rows):
      for j in range(8):
        want = (X[row, col + 4 * j: col + 4 * j + 4].cpu() if 0 <= row < n else torch.zeros(4))
        assert torch.equal(o[i, j ^ i], want)
    assert bool((o[4:] == -777.0).all())     # nothing beyond the four rows is written


def test_tc_two_source_epilogue(G):
  torch.manual_seed(6)
  C_ref, _ = _oracle_voxelize([_random_cloud(23, 5000, 9.0)], 0.3)
  n = len(C_ref)
  nbr = OME.build_neighbor_table(C_ref.numpy(), C_ref.numpy(), OME.kernel_offsets(3, 1))
  a, b = torch.randn(n, 64), torch.randn(n, 64)
  W = torch.randn(27, 128, 64) / 60
  sc, sh, res = torch.rand(64) + 0.5, torch.randn(64), torch.randn(n, 64)
  ref = torch.relu(OME.sparse_conv_reference(torch.cat([a, b], 1), W, nbr, n) * sc + sh + res)
  got = G.ops.spconv_fwd(a.to(G.dev), G.ops.weights_to_tc(W.to(G.dev)), torch.from_numpy(nbr).int().to(G.dev), n,
                         in1=b.to(G.dev), scale=sc.to(G.dev), shift=sh.to(G.dev), residual=res.to(G.dev), relu=True, algo=2)
  assert _rel(got, ref) < FEAT_TOL


@pytest.mark.parametrize("cin,cout", [(64, 64), (128, 64), (256, 256), (64, 32), (128, 128), (32, 32), (32, 64), (96, 64), (64, 128)])
def test_tc_fp16_activations(G, cin, cout):
  """kind::f16 path: fp16 activations / residual / weight image, fp32 accumulation, fp32 or fp16 output.
  (a) small integers are exact in fp16: bit-exact against integer arithmetic (pins the 64-channel row layout, the fp16
  weight image, the residual decode and both output stores);  (b) random data against the oracle convolution of the
  fp16-rounded operands (products of two fp16 values are exact in fp32, so only the summation order differs)."""
  torch.manual_seed(cin + cout)
  C_ref, _ = _oracle_voxelize([_random_cloud(24, 4000, 9.0)], 0.3)
  n = len(C_ref)
  nbr = OME.build_neighbor_table(C_ref.numpy(), C_ref.numpy(), OME.kernel_offsets(3, 1))
  nbr_d = torch.from_numpy(nbr).int().to(G.dev)
  srt, perm, mask = G.ops.kernel_map_sort(nbr_d)
  # two-source gather for the wide cases; (96: 64 + 32) and (64 -> 128: 32 + 32) take the 32-channel-row layout like (32, *)
  c0 = cin // 2 if (cin >= 128 or cout == 128) else (64 if cin == 96 else cin)
  xi = torch.randint(-2, 3, (n, cin)).float()
  Wi = torch.randint(-2, 3, (27, cin, cout)).float()
  ri = torch.randint(-2, 3, (n, cout)).float()
  ref = torch.relu(OME.sparse_conv_reference(xi, Wi, nbr, n) + ri)
  Wh = G.ops.weights_to_tc(Wi.to(G.dev), half=True, c0=c0)
  a = xi[:, :c0].contiguous().half().to(G.dev)
  b = xi[:, c0:].contiguous().half().to(G.dev) if c0 < cin else None
  for out_dtype in (torch.float32, torch.float16):
    got = G.ops.spconv_fwd(a, Wh, srt, n, in1=b, residual=ri.half().to(G.dev), relu=True, algo=2, row_perm=perm,
                           tile_mask=mask, out_dtype=out_dtype)
    assert got.dtype == out_dtype and torch.equal(got.float().cpu(), ref)       # |values| < 2048: exact in fp16 too
  x = torch.randn(n, cin)
  W = torch.randn(27, cin, cout) / np.sqrt(27 * cin)
  sc, sh = torch.rand(cout) + 0.5, torch.randn(cout)
  ref = OME.sparse_conv_reference(x.half().float(), W.half().float(), nbr, n) * sc + sh
  got = G.ops.spconv_fwd(x.half().to(G.dev), G.ops.weights_to_tc(W.to(G.dev), half=True), nbr_d, n, scale=sc.to(G.dev),
                         shift=sh.to(G.dev), algo=2, out_dtype=torch.float32)
  assert _rel(got, ref) < 1e-5
  # against the UNROUNDED operands: the storage precision itself, same bar as the tf32 path
  assert _rel(got, OME.sparse_conv_reference(x, W, nbr, n) * sc + sh) < FEAT_TOL
  # fp32 -> fp16 converting layer (tf32 kernel writing fp16) and the K == 1 dense path with fp16 operands
  g32 = G.ops.spconv_fwd(x.to(G.dev), G.ops.weights_to_tc(W.to(G.dev)), nbr_d, n, algo=2)
  g16 = G.ops.spconv_fwd(x.to(G.dev), G.ops.weights_to_tc(W.to(G.dev)), nbr_d, n, algo=2, out_dtype=torch.float16)
  assert g16.dtype == torch.float16 and torch.equal(g16, g32.half())
  W1 = torch.randint(-3, 4, (1, cin, cout)).float()
  got = G.ops.spconv_fwd(xi.half().to(G.dev), G.ops.weights_to_tc(W1.to(G.dev), half=True), None, n, algo=2,
                         out_dtype=torch.float32)
  assert torch.equal(got.cpu(), xi @ W1[0])


def test_spconv_two_source_and_pointwise(G):
  torch.manual_seed(5)
  C_ref, _ = _oracle_voxelize([_random_cloud(22, 2000, 6.0)], 0.3)
  n = len(C_ref)
  nbr = OME.build_neighbor_table(C_ref.numpy(), C_ref.numpy(), OME.kernel_offsets(3, 1))
  a, b = torch.randn(n, 64), torch.randn(n, 32)
  W = torch.randn(27, 96, 64) / 50
  ref = OME.sparse_conv_reference(torch.cat([a, b], 1), W, nbr, n)
  got = G.ops.spconv_fwd(a.to(G.dev), W.to(G.dev), torch.from_numpy(nbr).int().to(G.dev), n, in1=b.to(G.dev), algo=1)
  assert _rel(got, ref) < FP32_TOL
  W1, W2, bias = torch.randn(96, 64) / 10, torch.randn(64, 32) / 8, torch.randn(32)
  y = torch.relu(torch.cat([a, b], 1) @ W1) @ W2 + bias
  ref_t = y / y.norm(dim=1, keepdim=True)
  got_t = G.ops.pointwise_tail(a.to(G.dev), b.to(G.dev), W1.to(G.dev), W2.to(G.dev), bias.to(G.dev))
  assert _rel(got_t, ref_t) < FP32_TOL
  got_mm = G.ops.spconv_fwd(a.to(G.dev), W1[:64].contiguous().to(G.dev), None, n, algo=1)
  assert _rel(got_mm, a @ W1[:64]) < FP32_TOL


@pytest.mark.parametrize("ks", [3, 5, 7])
def test_probe_conv_emits_k3_table(G, ks):
  """the inner probes of the fused-probe convolution ARE the stride-1 3x3x3 kernel map: table, row keys, row masks and
  key histogram must equal what the separate gclb_kmap_build pass produces, bit for bit (and the conv output is unchanged)"""
  torch.manual_seed(33)
  C_ref, _ = _oracle_voxelize([_random_cloud(29, 4000, 9.0), _random_cloud(30, 1500, 5.0)], 0.3)
  cm = G.ops.hash_build(C_ref.to(G.dev))
  x = torch.randn(cm.n, 1, device=G.dev)
  W = torch.randn(ks ** 3, 1, 32, device=G.dev)
  ref_out = G.ops.spconv_fwd_probe(x, W, cm, ks)
  out, (nbr, (keys, masks, hist)) = G.ops.spconv_fwd_probe(x, W, cm, ks, emit_k3=True)
  assert torch.equal(out, ref_out)
  nbr_ref, (k_ref, m_ref, h_ref) = G.ops.kernel_map(cm, cm, 3, with_keys=True)
  assert torch.equal(nbr, nbr_ref) and torch.equal(keys, k_ref) and torch.equal(masks, m_ref) and torch.equal(hist, h_ref)
  onbr = OME.build_neighbor_table(C_ref.numpy(), C_ref.numpy(), OME.kernel_offsets(3, 1))
  assert np.array_equal(nbr.cpu().numpy(), onbr)


@pytest.mark.parametrize("cin,cout,ks", [(1, 32, 5), (3, 16, 3), (4, 64, 3)])
def test_spconv_probe_fused_kernel_map(G, cin, cout, ks):
  torch.manual_seed(31)
  C_ref, _ = _oracle_voxelize([_random_cloud(27, 3000, 8.0), _random_cloud(28, 2000, 6.0)], 0.3)
  n = len(C_ref)
  nbr = OME.build_neighbor_table(C_ref.numpy(), C_ref.numpy(), OME.kernel_offsets(ks, 1))
  x, W = torch.randn(n, cin), torch.randn(ks ** 3, cin, cout) / np.sqrt(cin * ks ** 3)
  sc, sh = torch.rand(cout) + 0.5, torch.randn(cout)
  ref = OME.sparse_conv_reference(x, W, nbr, n) * sc + sh
  cm = G.ops.hash_build(C_ref.to(G.dev))
  got = G.ops.spconv_fwd_probe(x.to(G.dev), W.to(G.dev), cm, ks, scale=sc.to(G.dev), shift=sh.to(G.dev))
  assert _rel(got, ref) < FP32_TOL


def _seed_bn(model):
  g = torch.Generator().manual_seed(123)
  for m in model.modules():
    if isinstance(m, torch.nn.BatchNorm1d):
      m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
      m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
      m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
      m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


@pytest.fixture(scope="module")
def net(G):
  """oracle model + its output on a 2-cloud batch; the CUDA model shares the state_dict."""
  torch.manual_seed(0)
  clouds = [G.synth.cast(G.synth.Scene(3, n_boxes=25, extent=35.0), dict(G.synth.NUSCENES, azimuth_steps=300)),
            _random_cloud(31, 3000, 10.0)]
  C_ref, _ = _oracle_voxelize(clouds, 0.3)
  F_in = torch.ones(len(C_ref), 1)
  om = G.make_models(OME)["ResUNetBN2C"](1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  _seed_bn(om)
  om.eval()
  with torch.no_grad():
    ref = om(OME.SparseTensor(F_in, coordinates=C_ref)).F
  return om, clouds, C_ref, F_in, ref


def test_resunet_module_path_vs_oracle(G, net):
  om, clouds, C_ref, F_in, ref = net
  gm = G.make_models(G.ME)["ResUNetBN2C"](1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  gm.load_state_dict(om.state_dict())
  gm = gm.to(G.dev).eval()
  with torch.no_grad():
    st = G.ME.SparseTensor(F_in.to(G.dev), coordinates=C_ref.to(G.dev))
    out = gm(st)
  assert out.coordinate_map_key == st.coordinate_map_key
  assert st.coordinate_manager.stats == {"kmap_builds": 11, "stride_builds": 3}
  assert _rel(out.F, ref) < FEAT_TOL
  # the same modules with the exact-fp32 kernels everywhere
  G.ME.set_inference_conv_algo("fp32")
  try:
    with torch.no_grad():
      out32 = gm(G.ME.SparseTensor(F_in.to(G.dev), coordinates=C_ref.to(G.dev)))
  finally:
    G.ME.set_inference_conv_algo("auto")
  assert _rel(out32.F, ref) < 1e-5
  assert not torch.equal(out32.F, out.F)      # i.e. the default really went through the tensor-core kernels


def test_resunet_engine_vs_oracle(G, net):
  om, clouds, C_ref, F_in, ref = net
  eng = G.engine.ResUNetEngine(om, device=G.dev)
  xyz = torch.from_numpy(np.concatenate(clouds)).to(G.dev)
  ptr = torch.tensor([0, len(clouds[0]), len(clouds[0]) + len(clouds[1])])
  feats, cm, umap = eng.extract(xyz, 0.3, ptr)
  assert torch.equal(cm.coords.cpu(), C_ref)
  rel = _rel(feats, ref)
  worst = (feats.cpu() - ref).norm(dim=1).max().item()
  assert rel < FEAT_TOL and worst < 5e-3, (rel, worst)
  # permutation equivariance: shuffling the input rows permutes the output rows
  perm = torch.randperm(len(C_ref))
  cm_p = G.ops.hash_build(C_ref[perm].to(G.dev))
  feats_p = eng.forward(cm_p, F_in.to(G.dev))
  assert _rel(feats_p, ref[perm]) < FEAT_TOL


def test_fp16_activation_range_is_monitored(G, net):
  """fp16 activation storage (engine default) against magnitudes a real checkpoint could produce: BN scales pushing the
  activations to ~1e5 (beyond fp16's 65504) or down to ~1e-7 (fp16 subnormals).  The reference computes in fp32 and has
  no such limit, so the engine must either stay within tolerance or report it loudly -- never saturate silently.  The
  fp32-storage engine (kind::tf32) handles the same weights within tolerance."""
  import copy
  from gcl_b200._lib import GclbError
  om, clouds, C_ref, F_in, ref = net
  xyz = torch.from_numpy(np.concatenate(clouds)).to(G.dev)
  ptr = torch.tensor([0, len(clouds[0]), len(clouds[0]) + len(clouds[1])])
  eng = G.engine.ResUNetEngine(om, device=G.dev)
  eng.extract(xyz, 0.3, ptr)
  eng.check_range()                                   # the seeded model is comfortably inside the fp16 range
  for factor, what in ((3e5, "fp16 range"), (1e-7, "subnormal")):
    m = copy.deepcopy(om)
    with torch.no_grad():            # scale the FIRST layer's BN output: block1 and its residual path see the scaled tensor
      m.norm1.bn.weight.mul_(factor)
      m.norm1.bn.bias.mul_(factor)
    m.eval()
    with torch.no_grad():
      want = m(OME.SparseTensor(F_in, coordinates=C_ref)).F
    assert bool(torch.isfinite(want).all())
    e16 = G.engine.ResUNetEngine(m, device=G.dev)
    f16, _, _ = e16.extract(xyz, 0.3, ptr)
    ok = _rel(f16, want) < FEAT_TOL
    flagged = False
    try:
      e16.check_range()
    except GclbError as ex:
      flagged = what in str(ex)
    assert ok or flagged, f"factor {factor}: fp16 engine is off by {_rel(f16, want):.2e} and did not flag it"
    assert flagged, f"factor {factor}: activations of magnitude ~{factor} must set the status word"
    e32 = G.engine.ResUNetEngine(m, device=G.dev, half=False)       # the remedy named in the error message
    f32, _, _ = e32.extract(xyz, 0.3, ptr)
    assert _rel(f32, want) < FEAT_TOL
    e32.check_range()


@pytest.mark.parametrize("c,relu,affine,momentum", [(32, False, True, 0.05), (64, True, True, 0.05), (256, True, True, 0.1),
                                                   (128, False, False, 0.02), (96, True, True, None)])
def test_batchnorm_train_kernels_vs_torch(G, c, relu, affine, momentum):
  """a7: MinkowskiBatchNorm in training mode on libgclb200's own kernels (csrc/bn.cu) against torch.nn.BatchNorm1d on the CPU
  (what the oracle's MinkowskiBatchNorm wraps, model/common.py:6): output, running statistics (unbiased variance, momentum,
  cumulative average for momentum=None), num_batches_tracked and the gradients w.r.t. input, gamma and beta -- with and
  without the fused ReLU.  fp32 arithmetic on both sides: 1e-5 relative."""
  torch.manual_seed(c)
  n = 5000 + c
  x = torch.randn(n, c) * 2.0 + torch.randn(c)
  gy = torch.randn(n, c)
  ref = torch.nn.BatchNorm1d(c, eps=1e-5, momentum=momentum, affine=affine)
  mod = G.ME.MinkowskiBatchNorm(c, eps=1e-5, momentum=momentum, affine=affine).to(G.dev)
  if affine:
    with torch.no_grad():
      ref.weight.uniform_(0.5, 1.5); ref.bias.normal_(0, 0.3)
    mod.bn.load_state_dict(ref.state_dict())
  coords = torch.cat([torch.zeros(n, 1, dtype=torch.int32), torch.arange(n, dtype=torch.int32)[:, None], torch.zeros(n, 2, dtype=torch.int32)], 1)
  for step in range(2):                      # two steps: the running statistics accumulate
    xo = x.clone().requires_grad_(True)
    yo = ref(xo)
    yo = torch.relu(yo) if relu else yo
    yo.backward(gy)
    xg = x.clone().to(G.dev).requires_grad_(True)
    st = G.ME.SparseTensor(xg, coordinates=coords.to(G.dev)) if step == 0 else G.ME.SparseTensor(xg, coordinate_map_key=key, coordinate_manager=mgr)
    key, mgr = st.coordinate_map_key, st.coordinate_manager
    out = mod(st, relu=relu) if relu else mod(st)
    out.F.backward(gy.to(G.dev))
    assert _rel(out.F, yo) < 1e-5
    assert _rel(xg.grad, xo.grad) < 2e-5
    if affine:
      assert _rel(mod.bn.weight.grad, ref.weight.grad) < 2e-5 and _rel(mod.bn.bias.grad, ref.bias.grad) < 2e-5
      mod.bn.weight.grad = None; mod.bn.bias.grad = None; ref.weight.grad = None; ref.bias.grad = None
    assert _rel(mod.bn.running_mean, ref.running_mean) < 1e-5 and _rel(mod.bn.running_var, ref.running_var) < 1e-5
    assert int(mod.bn.num_batches_tracked) == int(ref.num_batches_tracked) == step + 1
  # eval mode afterwards uses the accumulated statistics (folded affine kernel)
  mod.eval(); ref.eval()
  with torch.no_grad():
    ye = mod(G.ME.SparseTensor(x.to(G.dev), coordinate_map_key=key, coordinate_manager=mgr)).F
  assert _rel(ye, ref(x)) < 1e-5


@pytest.mark.parametrize("mode,w,tol", [("fp32", (16, 32, 8), 1e-4), ("tf32", (32, 64, 32), 3e-3)])
def test_training_step_grads_vs_oracle(G, mode, w, tol):
  """conv dgrad / wgrad (stride-1, strided, transposed, 1x1) and train-mode BN through a small U-shaped net; in 'tf32'
  mode forward and dgrad run on the tcgen05 kernel (channel widths chosen so that every layer but the first qualifies)."""
  torch.manual_seed(3)
  C_ref, _ = _oracle_voxelize([_random_cloud(41, 1500, 6.0), _random_cloud(42, 1500, 6.0)], 0.3)
  F_in = torch.randn(len(C_ref), 3)
  c_a, c_b, c_o = w

  def build(ME):
    torch.manual_seed(9)
    return torch.nn.ModuleDict(dict(
        c1=ME.MinkowskiConvolution(3, c_a, kernel_size=3, stride=1, dimension=3), n1=ME.MinkowskiBatchNorm(c_a),
        c2=ME.MinkowskiConvolution(c_a, c_b, kernel_size=3, stride=2, dimension=3), n2=ME.MinkowskiBatchNorm(c_b),
        c3=ME.MinkowskiConvolution(c_b, c_b, kernel_size=3, stride=1, dimension=3),
        t2=ME.MinkowskiConvolutionTranspose(c_b, c_a, kernel_size=3, stride=2, dimension=3),
        f=ME.MinkowskiConvolution(2 * c_a, c_o, kernel_size=1, stride=1, bias=True, dimension=3)))

  def run(ME, net, Fi, Ci):
    MEF = ME.MinkowskiFunctional
    x = ME.SparseTensor(Fi, coordinates=Ci)
    a = MEF.relu(net["n1"](net["c1"](x)))
    b = MEF.relu(net["n2"](net["c2"](a)))
    b2 = net["c3"](b)
    b2 += b
    u = net["t2"](MEF.relu(b2))
    return net["f"](ME.cat(u, a)).F

  on = build(OME)
  gn = build(G.ME)
  gn.load_state_dict(on.state_dict())
  gn = gn.to(G.dev)
  Fo = F_in.clone().requires_grad_(True)
  Fg = F_in.clone().to(G.dev).requires_grad_(True)
  yo = run(OME, on, Fo, C_ref)
  G.ME.set_training_conv_algo(mode)
  try:
    yg = run(G.ME, gn, Fg, C_ref.to(G.dev))
    assert _rel(yg.detach(), yo.detach()) < tol
    wgt = torch.randn_like(yo)
    (yo * wgt).sum().backward()
    (yg * wgt.to(G.dev)).sum().backward()
  finally:
    G.ME.set_training_conv_algo("fp32")
  assert _rel(Fg.grad, Fo.grad) < tol
  # weight gradients pass through the train-mode BatchNorm backward (mean subtraction => cancellation), which amplifies
  # the tf32 operand rounding of the upstream dgrad: ~1e-2 there, as usual for TF32 training; fp32 mode stays at 1e-4
  ptol = tol if mode == "fp32" else 2e-2
  for (k, po), (_, pg) in zip(on.named_parameters(), gn.named_parameters()):
    assert _rel(pg.grad, po.grad) < ptol, k
  for (k, bo), (_, bg) in zip(on.named_buffers(), gn.named_buffers()):
    assert _rel(bg.float(), bo.float()) < (1e-5 if mode == "fp32" else tol), k


@pytest.mark.gpu
@pytest.mark.parametrize("K,cin,cout,flip", [(27, 32, 64, True), (27, 64, 64, False), (8, 128, 64, False), (1, 96, 32, False)])
def test_dgrad_weight_image_matches_flip_transpose_image(G, K, cin, cout, flip):
  """gclb_weights_to_tc_dgrad == gclb_weights_to_tc of the flipped / transposed weights (bit-exact: same rounding, same layout)"""
  if cin % 8 or cout % 32:
    pytest.skip("shape outside the tensor-core image")
  torch.manual_seed(K + cin)
  W = torch.randn(K, cin, cout, device=G.dev)
  want = G.ops.weights_to_tc((W.flip(0) if flip else W).transpose(1, 2).contiguous())
  got = G.ops.weights_to_tc_dgrad(W, flip)
  assert got.shape == want.shape and torch.equal(got, want)


@pytest.mark.parametrize("cin,cout", [(32, 32), (64, 64), (32, 64), (128, 128), (256, 256), (256, 64), (64, 128), (128, 256),
                                      (96, 64), (192, 32)])
def test_wgrad_tc_vs_fp32(G, cin, cout):
  """tcgen05 weight gradient (MN-major tf32 operands gathered by TMA) against the exact-fp32 CUDA-core kernel (itself
  checked against oracle autograd above): stride-1 table, strided (down) table, and the K == 1 identity case."""
  C_ref, _ = _oracle_voxelize([_random_cloud(51, 2500, 7.0), _random_cloud(52, 2300, 7.0)], 0.3)
  cm = G.ops.hash_build(C_ref.to(G.dev))
  g = torch.Generator().manual_seed(cin * 1000 + cout)
  cm2 = G.ops.stride_map(cm, 2)
  for name, cin_map, cout_map in [("s1", cm, cm), ("down", cm, cm2)]:
    nbr, keys = G.ops.kernel_map(cin_map, cout_map, 3, with_keys=True)
    srt, perm, mask = G.ops.kernel_map_sort(nbr, keys)
    x = torch.randn(cin_map.n, cin, generator=g).to(G.dev)
    go = torch.randn(cout_map.n, cout, generator=g).to(G.dev)
    ref = G.ops.spconv_wgrad(x, go, nbr, 27)
    got = G.ops.spconv_wgrad_tc(x, go, srt, 27, row_perm=perm, tile_mask=mask)
    assert _rel(got, ref) < 3e-3, (name, _rel(got, ref))
    got2 = G.ops.spconv_wgrad_tc(x, go, srt, 27, row_perm=perm, tile_mask=None)     # no masks: every tile, every offset
    assert _rel(got2, ref) < 3e-3, (name, "nomask", _rel(got2, ref))
  x = torch.randn(cm.n, cin, generator=g).to(G.dev)
  go = torch.randn(cm.n, cout, generator=g).to(G.dev)
  ref = (x.double().T @ go.double()).float()[None]
  got = G.ops.spconv_wgrad_tc(x, go, None, 1)
  assert _rel(got, ref) < 3e-3, ("mm", _rel(got, ref))


# ----------------------------------------------------------------------------------------------- K4
def _unit(n, c, seed):
  x = torch.randn(n, c, generator=torch.Generator().manual_seed(seed))
  return x / x.norm(dim=1, keepdim=True)


@pytest.mark.parametrize("n,m,c", [(5000, 5000, 32), (777, 1301, 32), (1, 5, 32), (300, 200, 16), (130, 257, 7)])
def test_nn_vs_oracle(G, n, m, c):
  A, B = _unit(n, c, 1), _unit(m, c, 2)
  ref_i, ref_d = omatch.find_nn(A, B, nn_max_n=500, return_distance=True)
  got_i, got_d = G.matching.find_nn_gpu(A.to(G.dev), B.to(G.dev), nn_max_n=500, return_distance=True)
  assert got_i.dtype == torch.int64 and got_i.device.type == "cpu" and got_d.shape == (n, 1)
  assert torch.allclose(got_d, ref_d, atol=1e-6)
  bad = got_i != ref_i
  if bad.any():   # documented exception: near-ties (second-best within fp32 summation noise of the best)
    margin = omatch.nn_margin(A, B)
    assert (margin[bad] < 1e-6).all()
  assert bad.float().mean() < 1e-3


@pytest.mark.parametrize("algo", [1, 2])
@pytest.mark.parametrize("scale", [1.0, 37.5])
def test_nn_both_kernels_both_directions(G, algo, scale):
  """fp32 CUDA-core tiles (algo 1) and the tcgen05 3-way-split distance GEMM (algo 2) against the oracle; un-normalised
  features (scale) make sure nothing assumes unit norms; a planted exact duplicate checks the first-index tie rule."""
  A, B = _unit(3000, 32, 5) * scale, _unit(4100, 32, 6) * scale
  B[1234] = B[77]                      # exact tie between candidates 77 and 1234 -> 77 must win
  A[5] = B[77]
  idx01, d01, idx10, d10, *_ = G.ops.nn_search(A.to(G.dev), B.to(G.dev), both=True, algo=algo)
  r01, rd01 = omatch.find_nn(A, B, nn_max_n=500, return_distance=True)
  r10, rd10 = omatch.find_nn(B, A, nn_max_n=500, return_distance=True)
  assert idx01[5].item() == 77 and d01[5].item() <= 1e-6 * scale * scale
  for got, ref, gd, rd, X, Y in ((idx01, r01, d01, rd01, A, B), (idx10, r10, d10, rd10, B, A)):
    assert torch.allclose(gd.cpu(), rd[:, 0], atol=2e-6 * scale * scale)
    bad = got.cpu() != ref
    if bad.any():
      assert (omatch.nn_margin(X, Y)[bad] < 2e-6 * scale * scale).all()
    assert bad.float().mean() < 2e-3


def test_mutual_nn_and_batched_segments(G):
  sizes = [(900, 1100), (0, 50), (640, 0), (1500, 1300)]
  As = [_unit(n, 32, 10 + i) for i, (n, _) in enumerate(sizes)]
  Bs = [torch.cat([a[: min(len(a), m) // 2] + 0.01 * _unit(min(len(a), m) // 2, 32, 50 + i),
                   _unit(m - min(len(a), m) // 2, 32, 90 + i)]) if m else torch.zeros(0, 32)
        for i, (a, (_, m)) in enumerate(zip(As, sizes))]
  a_ptr = np.cumsum([0] + [len(a) for a in As]).tolist()
  b_ptr = np.cumsum([0] + [len(b) for b in Bs]).tolist()
  pairs, pair_ptr, idx01, idx10 = G.matching.mutual_nn_device(torch.cat(As).to(G.dev), torch.cat(Bs).to(G.dev), a_ptr, b_ptr)
  pp = pair_ptr.tolist()
  for s, (a, b) in enumerate(zip(As, Bs)):
    got = pairs[pp[s]:pp[s + 1]].cpu().numpy()
    if len(a) == 0 or len(b) == 0:
      assert len(got) == 0
      continue
    want, nn01, nn10 = omatch.mutual_nn(a, b)
    assert np.array_equal(idx01[a_ptr[s]:a_ptr[s + 1]].cpu().numpy(), nn01)
    assert np.array_equal(idx10[b_ptr[s]:b_ptr[s + 1]].cpu().numpy(), nn10)
    assert np.array_equal(got, want)
  single = G.matching.mutual_nn(As[0].numpy(), Bs[0].numpy())
  assert np.array_equal(single, omatch.mutual_nn(As[0], Bs[0])[0])


def test_match_pair_vs_oracle(G):
  """a14: Matcher.match_pair (scripts/SC2_PCR/SC2_PCR.py:276-302) -- the literal `sqrt(2 - 2 F0 F1^T + 1e-6)` arg-min of
  the oracle against the fused K4 kernel on unit-norm descriptors; rows may differ only where best and second-best dot
  products are closer than 1e-6 (documented near-ties)"""
  rng = np.random.RandomState(9)
  n, m = 3000, 3500
  F0 = torch.nn.functional.normalize(torch.from_numpy(rng.randn(n, 32).astype(np.float32)), dim=1)
  F1 = torch.nn.functional.normalize(torch.from_numpy(rng.randn(m, 32).astype(np.float32)), dim=1)
  F1[:1000] = torch.nn.functional.normalize(F0[:1000] + 0.05 * torch.from_numpy(rng.randn(1000, 32).astype(np.float32)), dim=1)
  x0 = torch.from_numpy(rng.uniform(-30, 30, (1, n, 3)).astype(np.float32))
  x1 = torch.from_numpy(rng.uniform(-30, 30, (1, m, 3)).astype(np.float32))
  s_ref, t_ref, idx_ref = omatch.match_pair(x0, x1, F0[None], F1[None])
  s, t = G.matching.match_pair(x0.to(G.dev), x1.to(G.dev), F0[None].to(G.dev), F1[None].to(G.dev))
  assert s.shape == (1, n, 3) and t.shape == (1, n, 3) and torch.equal(s.cpu(), s_ref)
  same = (t.cpu() == t_ref).all(dim=2)[0]
  dots = F0 @ F1.T
  top2 = torch.topk(dots, 2, dim=1).values
  near_tie = (top2[:, 0] - top2[:, 1]) < 1e-6
  assert bool((same | near_tie).all()) and same.float().mean().item() > 0.999
  # sub-sampled variant: same RNG calls as the reference (np.random.choice WITH replacement, src then tgt)
  r = np.random.RandomState(4)
  si, ti = r.choice(n, 500), r.choice(m, 500)
  s_ref2, t_ref2, _ = omatch.match_pair(x0[:, si], x1[:, ti], F0[None][:, si], F1[None][:, ti])
  s2, t2 = G.matching.match_pair(x0.to(G.dev), x1.to(G.dev), F0[None].to(G.dev), F1[None].to(G.dev), num_node=500,
                                 rng=np.random.RandomState(4))
  assert torch.equal(s2.cpu(), s_ref2) and (t2.cpu() == t_ref2).all(dim=2).float().mean().item() > 0.99


def test_subsample_is_a_sample_without_replacement(G):
  clouds = [_random_cloud(61, 9000, 20.0), _random_cloud(62, 300, 3.0), _random_cloud(63, 7000, 18.0), _random_cloud(64, 6000, 18.0)]
  xyz = torch.from_numpy(np.concatenate(clouds)).to(G.dev)
  ptr = torch.tensor(np.cumsum([0] + [len(c) for c in clouds]))
  cm, _ = G.ops.voxelize(xyz, 0.3, ptr)
  counts = torch.bincount(cm.coords[:, 0].long(), minlength=4).tolist()
  starts = np.cumsum([0] + counts)
  S = 2000
  cloud_rows, sel_ptr, sel, cap = G.ops.subsample(cm, 4, S, groups=2, seed=5)
  assert cloud_rows.tolist() == starts.tolist() and cap == S
  sp = sel_ptr.cpu().numpy()
  picked = []
  for c in range(4):
    g, sgm = c % 2, c // 2
    rows = sel[g, sp[g, sgm]:sp[g, sgm + 1]].cpu().numpy()
    assert len(rows) == min(counts[c], S)
    assert len(np.unique(rows)) == len(rows)                              # without replacement
    assert rows.min() >= starts[c] and rows.max() < starts[c + 1]         # inside its own cloud
    if counts[c] <= S:
      assert np.array_equal(rows, np.arange(starts[c], starts[c + 1]))    # small clouds are kept whole, in order
    else:
      picked.append((rows - starts[c]) / counts[c])
  # pseudo-random: roughly uniform over the cloud, and a different seed gives a different sample
  for u in picked:
    assert abs(u.mean() - 0.5) < 0.03 and abs(np.mean(u < 0.25) - 0.25) < 0.04
  _, _, sel2, _ = G.ops.subsample(cm, 4, S, groups=2, seed=6)
  assert not torch.equal(sel, sel2)


def test_pair_matcher_pipeline_vs_oracle(G, net):
  """the public one-call API: correspondences equal the oracle's mutual NN on the same features / same subsample"""
  om, clouds, C_ref, F_in, ref = net
  from gcl_b200.pipeline import PairMatcher
  pm = PairMatcher(om, voxel=0.3, subsample=1500, device=G.dev, seed=3)
  xyz = torch.from_numpy(np.concatenate(clouds))
  ptr = torch.tensor([0, len(clouds[0]), len(clouds[0]) + len(clouds[1])])
  out = pm.match(xyz.pin_memory(), ptr)
  assert out["n_voxels_total"] == len(C_ref) and torch.equal(out["coords"].cpu(), C_ref)
  a, b = out["a_ptr"].tolist(), out["b_ptr"].tolist()
  r0, r1 = out["sel0"][a[0]:a[1]].cpu(), out["sel1"][b[0]:b[1]].cpu()
  assert len(r0) == 1500 and len(r1) == 1500
  F = out["feats"].cpu()
  want, nn01, nn10 = omatch.mutual_nn(F[r0], F[r1])
  k = int(out["pair_ptr"][-1])
  got = out["pairs"][:k].cpu().numpy()
  assert np.array_equal(got, want)
  assert _rel(F, ref) < FEAT_TOL


def test_match_many_pipelined_equals_sequential(G, net):
  """two batches in flight on alternating streams give exactly the results of sequential match() calls"""
  om, clouds, C_ref, F_in, ref = net
  from gcl_b200.pipeline import PairMatcher
  xyz = torch.from_numpy(np.concatenate(clouds)).pin_memory()
  ptr = torch.tensor([0, len(clouds[0]), len(clouds[0]) + len(clouds[1])])
  batches = [(xyz, ptr), ((xyz + 0.07).pin_memory(), ptr), ((xyz - 0.11).pin_memory(), ptr), (xyz, ptr)]
  seq = PairMatcher(om, voxel=0.3, subsample=1000, device=G.dev, seed=9)
  want = [seq.match(x, p) for x, p in batches]
  torch.cuda.synchronize()
  pip = PairMatcher(om, voxel=0.3, subsample=1000, device=G.dev, seed=9)
  got = list(pip.match_many(iter(batches), depth=2))
  assert len(got) == len(want)
  for a, b in zip(got, want):
    assert a["n_voxels_total"] == b["n_voxels_total"]
    assert torch.equal(a["feats"], b["feats"]) and torch.equal(a["sel0"], b["sel0"])
    k = int(b["pair_ptr"][-1])
    assert torch.equal(a["pair_ptr"], b["pair_ptr"]) and torch.equal(a["pairs"][:k], b["pairs"][:k])


# ----------------------------------------------------------------------------------------------- K5
def _loss_inputs(seed, N=6000, G_=900, C=32):
  rng = np.random.RandomState(seed)
  F = _unit(N, C, seed)
  sizes = rng.randint(2, 8, G_)
  index = np.concatenate([rng.choice(N, s, replace=False) for s in sizes]).astype(np.int64)
  starts = np.concatenate([[0], np.cumsum(sizes)])
  flag = np.zeros(len(index), bool)
  flag[starts[:-1] + np.array([rng.randint(0, s) for s in sizes])] = True
  # make positives somewhat close so both active and inactive hinge branches occur
  for g in range(G_):
    members = index[starts[g]:starts[g + 1]]
    F[members] = F[members[0]] + 0.25 * rng.rand() * torch.randn(len(members), C, generator=torch.Generator().manual_seed(g))
  F = F / F.norm(dim=1, keepdim=True)
  ih = oloss.exhaustive_hash([index[starts[g]:starts[g + 1]] for g in range(G_)], N)
  return F, sizes, index, flag, ih


@pytest.mark.parametrize("square,finest", [(True, True), (False, True), (False, False)])
def test_group_loss_fwd_bwd_vs_oracle(G, square, finest):
  F, sizes, index, flag, ih = _loss_inputs(7)
  N = len(F)
  rng = np.random.RandomState(0)
  sel = oloss.draw_selections(len(sizes), N, 512, 1024, rng)
  Fo = F.clone().requires_grad_(True)
  po, fo, no = oloss.group_contrastive_loss(Fo, sizes, index, ih, flag, *sel, square_loss=square, with_finest=finest)
  w = (1.0, 0.7 if finest else 0.0, 1.3)
  (w[0] * po + w[1] * fo + w[2] * no).backward()
  crit = G.loss.GroupContrastiveLoss(square_loss=square, pos_weight=w[0], finest_weight=w[1], neg_weight=w[2],
                                     rng=np.random.RandomState(0))
  Fg = F.clone().to(G.dev).requires_grad_(True)
  fn = crit.finest_contrastive_loss if finest else crit.location_contrastive_loss
  if not finest:
    crit.square_loss = False
  pg, fg, ng = fn(Fg, torch.from_numpy(sizes), torch.from_numpy(index), ih, torch.from_numpy(flag),
                  max_pos_cluster=512, max_hn_samples=1024)
  (w[0] * pg + w[1] * fg + w[2] * ng).backward()
  assert abs(pg.item() - po.item()) < 1e-5 * max(1, abs(po.item()))
  assert abs(fg.item() - fo.item()) < 1e-5 * max(1, abs(fo.item()))
  assert abs(ng.item() - no.item()) < 1e-5 * max(1, abs(no.item()))
  assert po.item() > 0 and no.item() > 0
  assert _rel(Fg.grad, Fo.grad) < 1e-4


# ----------------------------------------------------------------------------------------------- SURVEY 8f #2
def _load_group_cases():
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "groups.npz"))
  for c in range(4):
    nbs, Ts, j = [], [], 0
    while f"c{c}_nb{j}" in g:
      nbs.append(g[f"c{c}_nb{j}"]); Ts.append(g[f"c{c}_T{j}"]); j += 1
    K = int(g[f"c{c}_K"])
    yield c, g[f"c{c}_centre"], nbs, Ts, float(g[f"c{c}_radius"]), (None if K < 0 else K), g


def test_colocation_groups_vs_reference_fixture(G):
  """voxel-hash radius search + group assembly on the GPU against the fixture generated by the reference's own
  get_matching_indices_colocation (util/pointcloud.py:69-132): group sizes, member indices (order included) and the
  finest-neighbour flags are integer outputs and must match exactly (K = 5, K = None, radius = 1.0 / 1.5 / 2.0 voxels)"""
  from gcl_b200 import groups as gg
  for c, centre, nbs, Ts, radius, K, g in _load_group_cases():
    grp, idx, fin = gg.colocation_groups(torch.from_numpy(centre).to(G.dev), [torch.from_numpy(x) for x in nbs], Ts, 0.3, radius, K)
    assert grp.dtype == torch.int64 and idx.dtype == torch.int64 and fin.dtype == torch.bool
    assert np.array_equal(grp.cpu().numpy(), g[f"c{c}_group"]), c
    assert np.array_equal(idx.cpu().numpy(), g[f"c{c}_index"]), c
    assert np.array_equal(fin.cpu().numpy(), g[f"c{c}_finest"]), c
  # reference return convention + the loss consumes it: torch.split(index, group) (lib/colocation_trainer.py:440-470)
  c, centre, nbs, Ts, radius, K, g = next(_load_group_cases())
  lg, li, lf, ld = gg.get_matching_indices_colocation(torch.from_numpy(centre).to(G.dev), [torch.from_numpy(x) for x in nbs], Ts,
                                                      radius, 0.3, K=K)
  assert lg == g["c0_group"].tolist() and li == g["c0_index"].tolist() and ld == [] and sum(lf) == len(lg)


def test_colocation_groups_edge_cases(G):
  from gcl_b200 import groups as gg
  from gcl_b200._lib import GclbError
  from oracle import groups as og
  rng = np.random.RandomState(5)
  # (a) a neighbour cloud far away: every centre is dropped -> empty outputs
  centre = (rng.randint(-20, 20, (300, 3)) * 0.3 + 0.15).astype(np.float32)
  centre = np.unique(centre, axis=0)
  far = centre + np.float32(500.0)
  grp, idx, fin = gg.colocation_groups(torch.from_numpy(centre).to(G.dev), [torch.from_numpy(far)], [np.eye(4)], 0.3, 0.45, 5)
  assert grp.numel() == 0 and idx.numel() == 0 and fin.numel() == 0
  # (b) lattice points: many exactly tied distances -> smallest index first, same as the oracle; negative coordinates;
  #     a rotated + translated neighbour cloud
  yaw = 0.4
  T = np.eye(4); T[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]; T[:3, 3] = [1.0, -2.0, 0.3]
  nb = ((centre.astype(np.float64) - T[:3, 3]) @ T[:3, :3]).astype(np.float32)          # T^-1 centre
  import oracle.me_cpu as OME2
  _, sel = OME2.utils.sparse_quantize(torch.from_numpy(nb) / 0.3, return_index=True)
  nb = nb[sel.numpy()]
  # second neighbour cloud: the same lattice seen from a sensor 1.1 cm away (a cloud identical to the centre cloud would make
  # the finest-neighbour rule compare a float32 norm with the float64 norm of the SAME point: a pure rounding coin toss)
  shift = np.array([0.011, -0.007, 0.003])
  nb2 = (centre.astype(np.float64) + shift).astype(np.float32)
  T2 = np.eye(4); T2[:3, 3] = -shift
  for K in (3, None):
    eg, ei, ef = og.colocation_groups(centre, [nb, nb2], [T, T2], 0.45, K)
    grp, idx, fin = gg.colocation_groups(torch.from_numpy(centre).to(G.dev), [torch.from_numpy(nb), torch.from_numpy(nb2)],
                                         [T, T2], 0.3, 0.45, K)
    assert np.array_equal(grp.cpu().numpy(), eg) and np.array_equal(idx.cpu().numpy(), ei) and np.array_equal(fin.cpu().numpy(), ef)
  # (c) two points in one voxel is not a loader cloud: refused loudly
  dup = np.concatenate([centre, centre[:1] + np.float32(0.01)])
  with pytest.raises(GclbError):
    gg.colocation_groups(torch.from_numpy(dup).to(G.dev), [torch.from_numpy(nb)], [T], 0.3, 0.45, 5)


def test_exhaustive_hash_vs_reference_order(G):
  """a17: device _exhaustive_hash against (1) keys produced by the reference's own util/misc.py:29-36
  (tests/golden/pair_hash.npz: same keys, same order) and (2) the ORACLE restatement on ragged / size-1 / empty group lists"""
  from gcl_b200 import groups as gg
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pair_hash.npz"))
  to_d = lambda a: torch.from_numpy(np.asarray(a, np.int64)).to(G.dev)
  for c in range(3):
    got = gg.exhaustive_hash(to_d(g[f"c{c}_group"]), to_d(g[f"c{c}_index"]), int(g[f"c{c}_M"]))
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), g[f"c{c}_keys"]), c
  rng = np.random.RandomState(3)
  for grp, idx, M in [(np.array([1, 4, 2, 1, 7]), np.arange(15)[::-1].copy(), 50),
                      (np.array([3, 1, 1, 9, 2]), rng.randint(0, 10 ** 6, 16), 10 ** 6),
                      (np.zeros(0, np.int64), np.zeros(0, np.int64), 10)]:
    split = np.split(idx, np.cumsum(grp)[:-1]) if len(grp) else []
    want = oloss.exhaustive_hash(split, M)
    got = gg.exhaustive_hash(to_d(grp), to_d(idx), M)
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)


def test_group_loss_honours_upstream_gradients(G):
  """the reference trainer's exact sequence (lib/colocation_trainer.py:874-879): in-place `/= iter_size` on the three
  returned losses, weighted sum, backward -- with iter_size = 2 and weights that differ from the constructor's; then a
  backward through ONE term only.  dL/dF must follow autograd's upstream gradients (round-1 bug: they were ignored)."""
  F, sizes, index, flag, ih = _loss_inputs(11)
  N = len(F)
  sel = oloss.draw_selections(len(sizes), N, 512, 1024, np.random.RandomState(0))
  iter_size, w = 2, (0.3, 1.7, 0.9)

  def run(loss_fn, Fx):
    pos, fin, neg = loss_fn(Fx)
    pos /= iter_size          # in place, like the trainer
    fin /= iter_size
    neg /= iter_size
    loss = w[0] * pos + w[1] * fin + w[2] * neg
    loss.backward()
    return float(loss.detach())

  Fo = F.clone().requires_grad_(True)
  lo = run(lambda x: oloss.group_contrastive_loss(x, sizes, index, ih, flag, *sel, square_loss=True, with_finest=True), Fo)
  crit = G.loss.GroupContrastiveLoss(square_loss=True)          # constructor weights (1, 1, 1) must be irrelevant
  Fg = F.clone().to(G.dev).requires_grad_(True)
  lg = run(lambda x: crit.finest_contrastive_loss(x, torch.from_numpy(sizes), torch.from_numpy(index), ih,
                                                  torch.from_numpy(flag), selections=sel), Fg)
  assert abs(lg - lo) < 1e-5 * max(1, abs(lo))
  assert _rel(Fg.grad, Fo.grad) < 1e-4
  # one term only, scaled: the other two receive no upstream gradient at all
  Fo2, Fg2 = F.clone().requires_grad_(True), F.clone().to(G.dev).requires_grad_(True)
  (3.0 * oloss.group_contrastive_loss(Fo2, sizes, index, ih, flag, *sel, square_loss=True, with_finest=True)[2]).backward()
  (3.0 * crit.finest_contrastive_loss(Fg2, torch.from_numpy(sizes), torch.from_numpy(index), ih, torch.from_numpy(flag),
                                      selections=sel)[2]).backward()
  assert _rel(Fg2.grad, Fo2.grad) < 1e-4
  # a selected group without a finest member: the reference raises IndexError; the product raises instead of reading
  # out of bounds on the device
  bad = flag.copy()
  bad[:sizes[0]] = False
  from gcl_b200._lib import GclbError
  with pytest.raises(GclbError):
    crit.finest_contrastive_loss(Fg2.detach(), torch.from_numpy(sizes), torch.from_numpy(index), ih, torch.from_numpy(bad),
                                 selections=(np.arange(len(sizes)), sel[1], sel[2]))


def test_ingest_velodyne_bin_files(G, tmp_path):
  """f3: KITTI .bin reader -> pinned buffer -> one H2D -> gclb_ingest_points, against the reference's
  `np.fromfile(fname, dtype=np.float32).reshape(-1, 4)[:, :3]` (lib/complement_data_loader.py:358-361): bit-identical without
  augmentation; with the loaders' random rotation + scale (:65-70, :753-781, float32 numpy) within 1e-6 relative (BLAS may fuse
  the multiply-adds), and the voxelisation downstream equals the oracle's on the ingested points"""
  from gcl_b200 import ingest
  rng = np.random.RandomState(3)
  paths, clouds = [], []
  for i, n in enumerate((12345, 1, 40000)):
    xyz = (rng.randn(n, 3) * [30, 30, 2]).astype(np.float32)
    p = str(tmp_path / f"{i:06d}.bin")
    ingest.write_velodyne_bin(p, xyz, rng.rand(n).astype(np.float32))
    paths.append(p); clouds.append(xyz)
  reader = ingest.ScanReader(capacity_points=60000)
  rec, ptr = reader.read(paths)
  assert rec.is_pinned() and ptr.tolist() == [0, 12345, 12346, 52346]
  want = np.concatenate([np.fromfile(p, dtype=np.float32).reshape(-1, 4)[:, :3] for p in paths])
  got = ingest.points_to_device(rec, ptr, G.dev)
  assert got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), want)
  cm, umap = G.ops.voxelize(got, 0.3, ptr)
  C_ref, sel_ref = _oracle_voxelize([want[ptr[i]:ptr[i + 1]] for i in range(3)], 0.3)
  assert torch.equal(cm.coords.cpu(), C_ref) and torch.equal(umap.cpu(), sel_ref)
  # augmentation path
  Ts, scales, ref = [], [], []
  for i in range(3):
    a = rng.uniform(-np.pi / 4, np.pi / 4)
    T = np.eye(4); T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]; T[:3, 3] = rng.randn(3)
    sc = 0.8 + 0.4 * rng.rand()
    t32 = T.astype(np.float32)
    pts = clouds[i] @ t32[:3, :3].T + t32[:3, 3]              # apply_transform (:65-70)
    ref.append(np.float32(sc) * pts); Ts.append(T); scales.append(sc)
  got = ingest.points_to_device(rec, ptr, G.dev, transforms=Ts, scales=scales).cpu().numpy()
  ref = np.concatenate(ref)
  assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max() + 1e-6
  # a truncated file is refused loudly
  with open(paths[0], "ab") as f:
    f.write(b"\x00\x00")
  with pytest.raises(Exception):
    reader.read(paths)
