"""Build csrc/*.cu into libldm_b200.so (sm_100a only) with nvcc; no torch headers involved.

    python oxford-102-flower-gan-vae-latent-diffusion_b200/build.py [--force] [--verbose]

The library is built in-tree so that it travels with the repository snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libldm_b200.so")
SOURCES = ["api.cu", "chain.cu", "v3loop.cu", "rowwise.cu", "gemm_f32.cu", "gemm_tc.cu", "conv_tc.cu", "pack.cu", "decoder.cu", "decoder_norm.cu", "pixel.cu", "ublock.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]
if os.environ.get("LDM_CHAIN_PREINIT"):        # chain.cu: additive terms written into the accumulator ahead of the MMAs (A/B builds)
    NVCC_FLAGS.append("-DLDM_CHAIN_PREINIT=" + os.environ["LDM_CHAIN_PREINIT"])
if os.environ.get("LDM_CHAIN_TRACE"):          # per-CTA clock stamps inside chain_kernel (tools/chain_sweep.py)
    NVCC_FLAGS.append("-DLDM_CHAIN_TRACE")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ldm_b200.h")]
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest(deps)
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
