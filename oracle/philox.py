"""Philox4x32-10 + Box-Muller: the specification of the in-kernel noise stream.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws its noise
from torch's global generator (v2:589,595), which a custom kernel cannot
replay; the product instead defines its stream by this spec, and parity tests
feed the SAME numbers to the reference through explicit noise tensors.

Stream definition (shared with csrc/philox.cuh):
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (quad, sample & 0xffffffff, step, sample >> 32)
      quad   = element_index // 4 inside one latent row
      sample = GLOBAL sample index (so sharding over GPUs never changes a draw)
      step   = t for the noise added by the update at timestep t (v2:589),
               n_steps for the initial x_T draw (v2:595)
  The four 32-bit outputs r0..r3 become four normals:
      u(r)  = ((r >> 8) + 0.5) * 2**-24                in (0, 1)
      z0,z1 = sqrt(-2 ln u(r0)) * (cos, sin)(2 pi u(r1))
      z2,z3 = sqrt(-2 ln u(r2)) * (cos, sin)(2 pi u(r3))
  element 4*quad + j  <-  z_j.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcastable). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _u01(r):
    return ((r >> np.uint32(8)).astype(np.float64) + 0.5) * (2.0 ** -24)


def normal_rows(seed, sample_offset, n_samples, step, dim=256):
    """(n_samples, dim) float32 standard normals for global samples
    [sample_offset, sample_offset + n_samples) at `step`."""
    assert dim % 4 == 0
    seed = int(seed)
    quads = np.arange(dim // 4, dtype=np.uint64)[None, :]
    samples = (np.arange(n_samples, dtype=np.uint64) + np.uint64(sample_offset))[:, None]
    r0, r1, r2, r3 = philox4x32_10(quads, samples & MASK, np.uint64(step), samples >> np.uint64(32),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((n_samples, dim // 4, 4), dtype=np.float64)
    for j, (ra, rb) in enumerate(((r0, r1), (r2, r3))):
        rad = np.sqrt(-2.0 * np.log(_u01(ra)))
        ang = 2.0 * np.pi * _u01(rb)
        out[:, :, 2 * j] = rad * np.cos(ang)
        out[:, :, 2 * j + 1] = rad * np.sin(ang)
    return out.reshape(n_samples, dim).astype(np.float32)
