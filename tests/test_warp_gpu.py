"""GPU parity of the crop warp + normalise kernel (rsg_warp_affine) through the C ABI: every byte of the warped crop
and every bit of the normalised model input equal the CPU oracle and the reference-executed golden fixtures
(cv2.warpAffine + torchvision ToTensor/Normalize)."""
import os

import numpy as np
import pytest
import torch

from oracle import warp_oracle
from rsgnet_b200 import synth
from rsgnet_b200.utils import transforms

pytestmark = pytest.mark.gpu


def test_warp_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'warp_cases.npz'), allow_pickle=False)
    imgs, c, s = synth.images(int(g['n']), seed=int(g['seed']))
    for i, img in enumerate(imgs):
        rot, size = float(g['rots'][i]), tuple(int(v) for v in g['sizes'][i])
        out = transforms.crop(img, c[i], s[i], size, rot)                        # reference signature, NumPy in/out
        assert out.dtype == np.uint8 and out.shape == (size[1], size[0], 3)
        assert np.array_equal(out, g[f'crop{i}']), i
        if f'input{i}' in g.files:
            m = transforms.get_affine_transform(c[i], s[i], rot, size)
            x, u8 = transforms.warp_crops([img], m[None], size, color_rgb=True, normalize=True, return_u8=True)
            assert np.array_equal(x[0].cpu().numpy(), g[f'input{i}'])             # bit-exact fp32
            assert np.array_equal(u8[0].cpu().numpy(), g[f'crop{i}'][:, :, ::-1]) # RGB order of the same bytes


@pytest.mark.parametrize('size', [(192, 256), (288, 384), (50, 37)])
def test_warp_batch_vs_oracle(size):
    """A batch whose crops share source images (several people per photo), hang over every border, rotate, and
    whose width is not a multiple of the CTA tile."""
    imgs, _, _ = synth.images(6, seed=21)
    rs = np.random.RandomState(5)
    n = 40
    idx = rs.randint(0, len(imgs), n)
    centers = np.stack([[rs.uniform(-0.2, 1.2) * imgs[j].shape[1], rs.uniform(-0.2, 1.2) * imgs[j].shape[0]] for j in idx]).astype(np.float32)
    sc = rs.uniform(0.1, 4.0, n).astype(np.float32)
    scales = np.stack([sc, sc * 1.25], axis=1)
    rots = np.where(rs.uniform(size=n) < 0.5, 0.0, rs.uniform(-80, 80, n))
    mats = transforms.affine_matrices(centers, scales, rots, size)
    dev = [torch.from_numpy(im).cuda() for im in imgs]                            # sources already resident
    x, u8 = transforms.warp_crops(dev, mats, size, image_index=idx, color_rgb=True, normalize=True, return_u8=True)
    x, u8 = x.cpu().numpy(), u8.cpu().numpy()
    lut = warp_oracle.normalize_lut()
    for i in range(n):
        ref = warp_oracle.warp_affine_u8(imgs[idx[i]], mats[i], size)[:, :, ::-1]
        assert np.array_equal(u8[i], ref), i
        assert np.array_equal(x[i], np.stack([lut[k][ref[:, :, k]] for k in range(3)])), i


def test_warp_edge_cases():
    rs = np.random.RandomState(2)
    # identity on an integer grid: fx = fy = 0 everywhere -> OpenCV's (32767, 0, 0, 1) weight set
    img = rs.randint(0, 256, (20, 30, 3)).astype(np.uint8)
    eye = np.array([[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]])
    out = transforms.warp_crops([img], eye, (30, 20), normalize=False, return_u8=True)[0].cpu().numpy()
    assert np.array_equal(out, warp_oracle.warp_affine_u8(img, eye[0], (30, 20)))
    # 1x1 and 2x2 sources, strong magnification, everything near a border
    for shp in ((1, 1, 3), (2, 2, 3), (1, 7, 3)):
        small = rs.randint(0, 256, shp).astype(np.uint8)
        m = np.array([[[9.5, 0.3, 4.0], [-0.2, 7.25, 3.0]]])
        out = transforms.warp_crops([small], m, (33, 17), normalize=False, return_u8=True)[0].cpu().numpy()
        assert np.array_equal(out, warp_oracle.warp_affine_u8(small, m[0], (33, 17)))
    # a crop entirely outside its image is all border (0) -> normalised zeros of the table
    m = transforms.affine_matrices(np.array([[5000.0, 5000.0]], np.float32), np.array([[1.0, 1.25]], np.float32), 0, (48, 64))
    x = transforms.warp_crops([img], m, (48, 64))
    lut = warp_oracle.normalize_lut()
    assert np.array_equal(x[0, :, 0, 0].cpu().numpy(), lut[:, 0]) and float(x[0, 0].std()) == 0.0
    # empty batch
    assert transforms.warp_crops([img], np.zeros((0, 2, 3)), (48, 64)).shape == (0, 3, 64, 48)
    with pytest.raises(AssertionError):
        transforms.warp_crops([img.astype(np.float32)], eye, (30, 20))


def test_warp_feeds_the_model_input_layout():
    """The f32 output is the NCHW tensor the model's stem kernel consumes (same layout as synth.crops)."""
    imgs, c, s = synth.images(3, seed=3)
    mats = transforms.affine_matrices(c, s, 0, (192, 256))
    x = transforms.warp_crops(imgs, mats, (192, 256), color_rgb=True)
    assert x.shape == (3, 3, 256, 192) and x.dtype == torch.float32 and x.is_contiguous()
    ref = np.stack([warp_oracle.crop_input(imgs[i], c[i], s[i], (192, 256)) for i in range(3)])
    assert np.array_equal(x.cpu().numpy(), ref)


def test_pipeline_from_photos_equals_pipeline_from_crops():
    """CropPipeline.infer_images (warp kernel writing the model's input buffer) gives exactly what feeding the
    oracle-made crops to the same pipeline gives."""
    from rsgnet_b200.pipeline import CropPipeline
    from tests.gpu_util import build
    cfg, net, _ = build('tiny_hrnet', 3)
    imgs, c, s = synth.images(4, seed=9)
    size = (net.spec.image_w, net.spec.image_h)
    pipe = CropPipeline(net, cfg, 4, use_graph=False)
    p1, m1 = pipe.infer_images(imgs, c, s, color_rgb=True)
    p1, m1 = p1.cpu().numpy(), m1.cpu().numpy()
    x = np.stack([warp_oracle.crop_input(imgs[i], c[i], s[i], size) for i in range(4)])
    p2, m2 = pipe(x, c, s)
    assert np.array_equal(p1, p2) and np.array_equal(m1, m2)


def test_warp_full_step_properties():
    """BASELINE step size (256 crops of 256x192 from 64 photos): size-independent properties instead of the oracle on
    everything -- an integer translation reproduces the source window byte for byte (OpenCV's (32767, 0, 0, 1) weight set
    rounds back to the centre tap), windows hanging over the border are zero there, and the normalised output is the
    table applied to those bytes; a sample of general crops is checked against the oracle."""
    rs = np.random.RandomState(4)
    imgs = [rs.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(8)]
    dev = [torch.from_numpy(imgs[i % 8]).cuda().clone() for i in range(64)]
    n, size = 256, (192, 256)
    idx = np.arange(n) % 64
    tx = rs.randint(-100, 500, n)
    ty = rs.randint(-100, 300, n)
    mats = np.zeros((n, 2, 3))
    mats[:, 0, 0] = mats[:, 1, 1] = 1.0
    mats[:, 0, 2] = -tx                                     # dst(x, y) = src(x + tx, y + ty)
    mats[:, 1, 2] = -ty
    x, u8 = transforms.warp_crops(dev, mats, size, image_index=idx, normalize=True, return_u8=True)
    u8 = u8.cpu().numpy()
    lut = torch.from_numpy(warp_oracle.normalize_lut()).cuda()
    for k in range(3):
        assert torch.equal(x[:, k], lut[k][torch.from_numpy(u8[..., k]).cuda().long()])
    for i in range(n):
        src = imgs[idx[i] % 8]
        ref = np.zeros((256, 192, 3), np.uint8)
        y0, y1 = max(0, -ty[i]), min(256, 480 - ty[i])
        x0, x1 = max(0, -tx[i]), min(192, 640 - tx[i])
        if y1 > y0 and x1 > x0:
            ref[y0:y1, x0:x1] = src[y0 + ty[i]:y1 + ty[i], x0 + tx[i]:x1 + tx[i]]
        assert np.array_equal(u8[i], ref), i
    # general crops at the same batch size: a sample against the oracle
    centers = np.stack([rs.uniform(0, 640, n), rs.uniform(0, 480, n)], axis=1).astype(np.float32)
    sc = rs.uniform(0.3, 3.0, n).astype(np.float32)
    m2 = transforms.affine_matrices(centers, np.stack([sc, sc * 1.25], axis=1), 0, size)
    u2 = transforms.warp_crops(dev, m2, size, image_index=idx, normalize=False, return_u8=True).cpu().numpy()
    for i in range(0, n, 37):
        assert np.array_equal(u2[i], warp_oracle.warp_affine_u8(imgs[idx[i] % 8], m2[i], size)), i
