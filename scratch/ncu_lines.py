"""Executed-instruction and stall-sample shares per source line of one ncu report: python scratch/ncu_lines.py rep [top]"""
import csv, collections, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass,cuda'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; hdr = None; agg = collections.Counter(); samp = collections.Counter(); lines = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if hdr and r[0].isdigit():
        try: n = int(r[ie]); s = int(r[ss])
        except ValueError: continue
        agg[(cur, int(r[0]))] += n; samp[(cur, int(r[0]))] += s; lines[(cur, int(r[0]))] = r[1].strip()[:100]
tot = sum(agg.values()); ts = max(sum(samp.values()), 1)
print('total warp instructions', tot, 'stall samples', ts)
for (f, l), n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print('%5.1f%% instr %5.1f%% samples  %s:%d  %s' % (100 * n / tot, 100 * samp[(f, l)] / ts, f, l, lines[(f, l)]))
