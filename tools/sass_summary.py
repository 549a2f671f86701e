"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.cp -> UTCCP; legacy mma.sync would show as HMMA).

    python tools/sass_summary.py > profiles/sass_summary.txt

Reads the in-tree libldm_b200.so with cuobjdump (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oxford-102-flower-gan-vae-latent-diffusion_b200", "libldm_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "HMMA", "FFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, i = {}, [], None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i] if i < len(names) else m.group(1)
            i += 1
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)[:70]
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print("libldm_b200.so: architectures %s; %d kernels" % (archs, len(order)))
    print("%-70s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
    for name in sorted(order, key=lambda n: -counts[n]["UTCHMMA"]):
        c = counts[name]
        print("%-70s %7d " % (name, c["_total"]) + " ".join("%7s" % (c[k] or ".") for k in KEYS))
    tc = [n for n in order if counts[n]["UTCHMMA"]]
    print("\n%d kernels issue tcgen05.mma (UTCHMMA); %d use TMA tensor loads (UTMALDG); %d read TMEM (LDTM); HMMA (legacy mma.sync) kernels: %d"
          % (len(tc), sum(1 for n in order if counts[n]["UTMALDG"]), sum(1 for n in order if counts[n]["LDTM"]),
             sum(1 for n in order if counts[n]["HMMA"])))


if __name__ == "__main__":
    sys.exit(main())
