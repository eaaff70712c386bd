"""Summarise an `ncu --page source --csv --print-source sass` dump into regions of equal execution count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc = hdr.index("Source"); ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed")
data = []
for r in rows[2:]:
    if len(r) < 10 or r[0] == "Kernel Name": break
    if r[0] == "Address": continue
    data.append((r[isrc].strip(), int(r[ie]), int(r[it])))
regions = []; start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[start][1]) > 0.03 * max(data[start][1], 1):
        regions.append((start, i - 1, i - start, sum(d[1] for d in data[start:i]), sum(d[2] for d in data[start:i]))); start = i
tot = sum(r[3] for r in regions); tth = sum(r[4] for r in regions)
print("instructions %d  warp-inst %.4e  thread-inst %.4e  avg threads %.2f" % (len(data), tot, tth, tth / tot))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
for s, e, n, ex, th in regions:
    if ex / tot > thr:
        print("instr %4d-%4d n=%3d exec/inst=%8.3fM share=%5.1f%% avgthreads=%4.1f  first: %s" % (s, e, n, ex / n / 1e6, 100 * ex / tot, th / max(ex, 1), data[s][0][:60]))
if len(sys.argv) > 3:
    a, b = map(int, sys.argv[3].split("-"))
    for i in range(a, b + 1):
        print("%4d %8.3fM %5.1f %s" % (i, data[i][1] / 1e6, data[i][2] / max(data[i][1], 1), data[i][0][:90]))
