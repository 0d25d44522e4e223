"""Format an `ncu --metrics gpu__time_duration.sum --csv` log as the launch list kept under profiles/:
   python tools/launch_list.py raw.csv > launch_list.txt"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
iK, iB, iG, iV = hdr.index("Kernel Name"), hdr.index("Block Size"), hdr.index("Grid Size"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    k = r[iK].split("(")[0][:70]
    tot[k] += float(r[iV].replace(",", "")) / 1e6
    cnt[k] += 1
T = sum(tot.values())
print("%-70s %6s %14s %7s" % ("kernel", "count", "total_ms", "share"))
for k, v in tot.most_common():
    print("%-70s %6d %14.3f %6.2f%%" % (k, cnt[k], v, 100 * v / T))
print("\n# launch by launch")
for i, r in enumerate(rows):
    print("%3d %-70s grid %-18s block %-16s %10.3f ms" % (i, r[iK].split("(")[0][:70], r[iG], r[iB], float(r[iV].replace(",", "")) / 1e6))
