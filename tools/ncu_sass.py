"""Per-SASS-instruction executed counts and stall samples of one kernel of an .ncu-rep (source page):
    python tools/ncu_sass.py rep.ncu-rep <kernel-substring> [min_exec]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
out, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = r[1]; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = {h: i for i, h in enumerate(r)}; continue
    if cur and pat in cur and hdr and len(r) > 5:
        out.append((r[hdr['Address']], r[hdr['Source']].strip(), int(r[hdr['# Samples']] or 0), int(r[hdr['Instructions Executed']] or 0),
                    float(r[hdr['Avg. Threads Executed']] or 0), int(r[hdr['L1 Wavefronts Shared']] or 0)))
tot = sum(o[3] for o in out); ts = sum(o[2] for o in out)
print("total inst", tot, "samples", ts)
base = int(out[0][0], 16)
for a, s, smp, ex, thr, wf in out:
    print("%5x %-70s ex=%9d (%4.1f%%) smp=%5d (%4.1f%%) thr=%4.1f wf=%d" % (int(a, 16) - base, s[:70], ex, 100.0 * ex / tot, smp, 100.0 * smp / max(ts, 1), thr, wf))
