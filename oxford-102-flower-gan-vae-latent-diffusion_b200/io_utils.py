"""Host-side plumbing either side of the sampling path: the two checkpoint conventions of the reference and
an image-grid writer that replaces its matplotlib figures.  No arithmetic of the hot path lives here."""
import os
import struct
import zlib

import torch


def parse_epoch_from_filename(path):
    """v2:1354-1358: `conditional_diffusion_epoch_<N>.pt` -> N.  Raises IndexError / ValueError exactly where the
    reference's own split / int would (its caller catches both and restarts from epoch 0)."""
    filename = os.path.basename(path)
    return int(filename.split("epoch_")[1].split(".pt")[0])


def load_unet_checkpoint(unet, checkpoint_path, device=None):
    """The resume branch of main() (v2:1352-1364): returns the start epoch (0 when the file is missing or its name
    carries no epoch), loading the state_dict with strict=False as the reference does."""
    if not checkpoint_path or not os.path.exists(checkpoint_path):
        return 0
    try:
        start_epoch = parse_epoch_from_filename(checkpoint_path)
    except (IndexError, ValueError):
        return 0
    unet.load_state_dict(torch.load(checkpoint_path, map_location=device), strict=False)
    return start_epoch


def _png_chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path, rgb):
    """rgb: (H, W, 3) uint8 tensor / array -> an 8-bit RGB PNG (no third-party encoder needed)."""
    rgb = torch.as_tensor(rgb).to(torch.uint8).cpu().contiguous()
    h, w, c = rgb.shape
    if c != 3:
        raise ValueError("write_png expects (H, W, 3)")
    raw = b"".join(b"\x00" + rgb[y].numpy().tobytes() for y in range(h))
    png = (b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
           + _png_chunk(b"IDAT", zlib.compress(raw, 6)) + _png_chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(png)


def to_uint8(images):
    """(N, 3, H, W) float images in [0, 1] (what decode / the pixel sampler return) -> (N, H, W, 3) uint8, the conversion
    matplotlib's imshow applies to the reference's `samples[i].cpu().permute(1, 2, 0)` (v2:873-874)."""
    x = images.detach().float().clamp(0, 1).mul(255.0).round().to(torch.uint8)
    return x.permute(0, 2, 3, 1).contiguous()


def save_image_grid(images, path, cols=None, pad=2):
    """The row / grid figures of generate_class_samples (v2:870-881) and generate_samples_grid (v4:204-224) as one PNG."""
    x = to_uint8(images).cpu()
    n, h, w, _ = x.shape
    cols = cols or n
    rows = (n + cols - 1) // cols
    canvas = torch.full((rows * h + (rows + 1) * pad, cols * w + (cols + 1) * pad, 3), 255, dtype=torch.uint8)
    for i in range(n):
        r, c = divmod(i, cols)
        y0, x0 = pad + r * (h + pad), pad + c * (w + pad)
        canvas[y0:y0 + h, x0:x0 + w] = x[i]
    write_png(path, canvas)
    return path
