"""Per-layer table of one reverse step of the v4 path from an ncu launch list (gpu__time_duration.sum CSV)."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
L = [(r[ki][:44], r[gi], float(r[vi].replace(',', '')) / 1e3) for r in rows[hdr + 1:] if len(r) > vi]
names = ["conv_in", "c1b", "down1", "c2a", "c2b", "down2", "c3a", "c3b", "b0", "b2", "up1", "c4a", "c4b", "up2", "c5a", "c5b", "out", "update"]
mm = [7, 151, 134, 151, 151, 134, 151, 151, 302, 302, 134, 302, 151, 134, 302, 151, 7, 0]
tot = 0
for n, (k, g, us), m in zip(names, L[-18:], mm):
    tot += us
    print(f"{n:7s} {k:44s} grid={g:16s} {us:8.1f} us  {m * B * 2 / us:8.1f} TFLOP/s")
print("sum", tot)
