"""Dev helper: group an `ncu --page source --csv` export into runs of SASS lines with the same execution count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = float(sys.argv[2]) if len(sys.argv) > 2 else 1048576.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr, data = rows[1], rows[2:]
iS, iI, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
print("kernel:", rows[0][1][:90])
tot = sum(int(r[iI]) for r in data); ts = sum(int(r[iN]) for r in data)
print("warp instructions %d = %.1f per row; samples %d" % (tot, tot / N, ts))
groups = []
for k, r in enumerate(data):
    n, s = int(r[iI]), int(r[iN])
    if groups and abs(groups[-1]['n'] - n) <= 0.03 * max(n, 1):
        g = groups[-1]; g['cnt'] += 1; g['tot'] += n; g['samp'] += s; g['end'] = k
    else:
        groups.append(dict(n=n, cnt=1, tot=n, samp=s, start=k, end=k))
for g in groups:
    if g['tot'] / N > thr:
        print("%4d-%4d lines=%3d per-line/row=%6.2f total/row=%7.2f samples=%5d (%4.1f%%) %s" % (
            g['start'], g['end'], g['cnt'], g['n'] / N, g['tot'] / N, g['samp'], 100.0 * g['samp'] / ts, data[g['start']][iS].strip()[:44]))
if len(sys.argv) > 4:
    a, b = map(int, sys.argv[4].split('-'))
    for k in range(a, b + 1):
        r = data[k]
        print(k, r[iS].strip()[:80].ljust(80), "%7.2f" % (int(r[iI]) / N), r[iN])
