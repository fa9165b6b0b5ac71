"""Known-answer tests that pin the oracle's conv semantics WITHOUT MinkowskiEngine:
on a fully occupied box a sparse conv must equal torch's dense conv3d (cross-correlation, zero padding),
with W_torch[co,ci,kz,ky,kx] = W_me[k,ci,co], k = kx + K*ky + K^2*kz (Appendix A5)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle.me_cpu as ME


def _full_grid(n, batch=1, step=1, origin=0):
  g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3) * step + origin
  return np.concatenate([np.concatenate([np.full((len(g), 1), b), g], 1) for b in range(batch)]).astype(np.int32)


def _to_dense(coords, feats, n, step=1, origin=0, batch=1):
  d = torch.zeros(batch, feats.shape[1], n, n, n)   # [B,C,z,y,x]
  c = (torch.as_tensor(coords[:, 1:]).long() - origin) // step
  d[torch.as_tensor(coords[:, 0]).long(), :, c[:, 2], c[:, 1], c[:, 0]] = feats
  return d


def _w_dense(W, K):
  return W.reshape(K, K, K, W.shape[1], W.shape[2]).permute(4, 3, 0, 1, 2).contiguous()  # [co,ci,kz,ky,kx]


@pytest.mark.parametrize("K", [3, 5])
def test_stride1_equals_conv3d(K):
  torch.manual_seed(0)
  n, cin, cout = 7, 3, 4
  coords = _full_grid(n, batch=2, origin=-3)      # negative coordinates on purpose
  feats = torch.randn(len(coords), cin)
  conv = ME.MinkowskiConvolution(cin, cout, kernel_size=K, stride=1, dimension=3)
  y = conv(ME.SparseTensor(feats, coordinates=torch.from_numpy(coords)))
  ref = F.conv3d(_to_dense(coords, feats, n, origin=-3, batch=2), _w_dense(conv.kernel.detach(), K), padding=K // 2)
  got = _to_dense(coords, y.F.detach(), n, origin=-3, batch=2)
  assert torch.allclose(got, ref, atol=1e-5)


def test_stride2_equals_conv3d_even_origin():
  torch.manual_seed(1)
  n, cin, cout = 8, 2, 5
  coords = _full_grid(n)
  feats = torch.randn(len(coords), cin)
  conv = ME.MinkowskiConvolution(cin, cout, kernel_size=3, stride=2, dimension=3)
  y = conv(ME.SparseTensor(feats, coordinates=torch.from_numpy(coords)))
  assert y.tensor_stride == [2, 2, 2] and len(y) == (n // 2) ** 3
  # out[c] = sum_off in[c + off], off in {-1,0,1}: dense conv3d stride 2 padding 1
  ref = F.conv3d(_to_dense(coords, feats, n), _w_dense(conv.kernel.detach(), 3), stride=2, padding=1)
  got = _to_dense(y.C.numpy(), y.F.detach(), n // 2, step=2)
  assert torch.allclose(got, ref, atol=1e-5)


def test_transposed_equals_conv_transpose3d():
  torch.manual_seed(2)
  n, cin, cout = 8, 3, 2
  coords = _full_grid(n)
  x = ME.SparseTensor(torch.randn(len(coords), 4), coordinates=torch.from_numpy(coords))
  down = ME.MinkowskiConvolution(4, cin, kernel_size=3, stride=2, dimension=3)
  up = ME.MinkowskiConvolutionTranspose(cin, cout, kernel_size=3, stride=2, dimension=3)
  z = down(x)
  y = up(z)
  assert y.coordinate_map_key == x.coordinate_map_key and len(y) == len(x)
  zd = _to_dense(z.C.numpy(), z.F.detach(), n // 2, step=2)
  # out[f] += in[c] W[k] for f = c + off_k: conv_transpose3d with un-flipped kernel, padding 1, output_padding 1
  Wt = up.kernel.detach().reshape(3, 3, 3, cin, cout).permute(3, 4, 0, 1, 2).contiguous()  # [ci,co,kz,ky,kx]
  ref = F.conv_transpose3d(zd, Wt, stride=2, padding=1, output_padding=1)
  got = _to_dense(coords, y.F.detach(), n)
  assert torch.allclose(got, ref, atol=1e-5)


def test_strided_map_floor_for_negatives_and_order():
  c = np.array([[0, -1, -1, -1], [0, 0, 0, 0], [0, -2, -1, 3], [0, 1, 1, 1], [1, -1, -1, -1]], np.int32)
  s = ME.stride_coords(c, 2)
  assert s.tolist() == [[0, -2, -2, -2], [0, 0, 0, 0], [0, -2, -2, 2], [1, -2, -2, -2]]


def test_sparse_quantize_first_occurrence():
  pts = torch.tensor([[0.2, 0.2, 0.2], [1.7, 0.1, 0.0], [0.9, 0.9, 0.9], [-0.1, 0.0, 0.0], [1.2, 0.5, 0.3]])
  c, idx = ME.utils.sparse_quantize(pts, return_index=True)
  assert idx.tolist() == [0, 1, 3]
  assert c.tolist() == [[0, 0, 0], [1, 0, 0], [-1, 0, 0]]
  um, inv = ME.utils.sparse_quantize(pts, return_maps_only=True, return_inverse=True)
  assert inv.tolist() == [0, 1, 0, 2, 1]


def test_cat_and_iadd_require_same_map():
  a = ME.SparseTensor(torch.ones(2, 1), coordinates=torch.tensor([[0, 0, 0, 0], [0, 1, 0, 0]], dtype=torch.int32))
  b = ME.SparseTensor(torch.ones(2, 1), coordinates=torch.tensor([[0, 0, 0, 0], [0, 1, 0, 0]], dtype=torch.int32))
  with pytest.raises(ValueError):
    ME.cat(a, b)
  with pytest.raises(ValueError):
    a += b
  c = ME.MinkowskiFunctional.relu(a)
  assert ME.cat(a, c).F.shape == (2, 2)
