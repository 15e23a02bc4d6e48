"""GPU parity: CLIP preprocessing kernels vs the real thing -- PIL's bicubic resize + torchvision
CenterCrop / ToTensor / Normalize run on the CPU (the transform the reference's DataLoader applies).
Integer resampling and single-rounding float ops: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# (96,96) .. (97,101) take the fused shared-memory kernel (odd byte counts: every image at a different 16-byte
# phase), the rest the general two-pass kernels
SIZES = [(96, 96), (32, 32), (100, 150), (97, 101), (28, 28), (224, 224), (375, 500), (500, 375), (600, 300), (225, 224), (1000, 40)]


def _cpu_transform(arr):
    from PIL import Image

    from lossyless_b200.clip_vit import clip_preprocess

    return clip_preprocess(224)(Image.fromarray(arr))


@pytest.mark.parametrize("h,w", SIZES)
def test_matches_pil_torchvision(cuda_device, h, w):
    from lossyless_b200.preprocess import GpuClipPreprocess

    rng = np.random.default_rng(h * 1000 + w)
    imgs = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    imgs[1] = 255  # saturation: bicubic overshoot must clamp exactly like Pillow's clip8
    imgs[2, ::2] = 0
    ref = torch.stack([_cpu_transform(a) for a in imgs])
    pre = GpuClipPreprocess(224)
    x = torch.from_numpy(imgs).to(cuda_device)
    out32 = pre(x, dtype=torch.float32)
    assert torch.equal(out32.cpu(), ref)
    out16 = pre(x)  # default fp16: what compress_dataset feeds the encoder (x.to(device).half())
    assert out16.dtype == torch.float16 and torch.equal(out16.cpu(), ref.half())
    out_chw = pre(x.permute(0, 3, 1, 2).contiguous(), chw=True, dtype=torch.float32)
    assert torch.equal(out_chw.cpu(), ref)
    # the general two-pass kernels give the same bits as whatever was chosen above
    from lossyless_b200 import _native as nat

    nat.call("ll_clip_preprocess_set_impl", 1)
    try:
        assert torch.equal(pre(x, dtype=torch.float32).cpu(), ref) and torch.equal(pre(x).cpu(), ref.half())
        assert torch.equal(pre(x, dtype=torch.bfloat16), out32.bfloat16())
    finally:
        nat.call("ll_clip_preprocess_set_impl", 0)
    assert torch.equal(pre(x, dtype=torch.bfloat16), out32.bfloat16())


@pytest.mark.parametrize("side", [96, 224])
def test_compress_dataset_from_raw_pixels(cuda_device, tmp_path, side):
    """STL10-shaped uint8 images (BASELINE config 2) and CLIP-sized ones (bench.py's e2e feed: Resize / CenterCrop are
    the identity, ToTensor / Normalize run in clip_normalize_u8_kernel): coding the raw pixels with GPU preprocessing
    gives the same container file, byte for byte, as coding the CPU-transformed tensors."""
    import hubconf

    comp, transform = hubconf.clip_compressor_b005(device="cuda")
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 256, (50, side, side, 3), dtype=np.uint8)

    class Raw(torch.utils.data.Dataset):
        def __len__(self):
            return len(raw)

        def __getitem__(self, i):
            return torch.from_numpy(raw[i]), i % 10

    class Transformed(Raw):
        def __getitem__(self, i):
            from PIL import Image

            return transform(Image.fromarray(raw[i])), i % 10

    kw = dict(batch_size=16, num_workers=0)
    comp.compress_dataset(Raw(), tmp_path / "raw.bin", label_file=tmp_path / "y1.npy", kwargs_dataloader=kw, is_info=False)
    comp.compress_dataset(Transformed(), tmp_path / "cpu.bin", label_file=tmp_path / "y2.npy", kwargs_dataloader=kw, is_info=False)
    assert (tmp_path / "raw.bin").read_bytes() == (tmp_path / "cpu.bin").read_bytes()
    assert np.array_equal(np.load(tmp_path / "y1.npy"), np.load(tmp_path / "y2.npy"))
