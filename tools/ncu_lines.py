"""Aggregate an ncu source-page CSV by source line: executed warp instructions per (cell, group) task and share of
stall samples.  usage: ncu_lines.py report.ncu-rep tasks [top]"""
import collections, csv, subprocess, sys
rep, tasks = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ''])
ops = collections.Counter(); stalls = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == 'Line No': hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 10: continue
    if r[0].isdigit():
        try:
            key = (cur, int(r[0])); agg[key][0] += int(r[7]); agg[key][1] += int(r[6]); agg[key][2] = r[1].strip()[:100]
        except ValueError: pass
    elif r[0] == '' and r[2].startswith('0x'):
        try: n = int(r[7])
        except ValueError: continue
        t = r[3].split()
        op = t[1] if t and t[0].startswith('@') else (t[0] if t else '?')
        ops[op.split('.')[0]] += n
        for h, i in idx.items():
            if h.startswith('stall_') and 'Not Issued' not in h:
                try: stalls[h] += int(r[i])
                except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f'total warp-instructions per task: {tot / tasks:.0f}')
byfile = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v[0]
print({k: round(v / tasks) for k, v in byfile.most_common()})
st = sum(stalls.values())
print('stalls:', ', '.join(f'{k[6:]} {100 * v / st:.1f}%' for k, v in stalls.most_common(9)))
print('opcodes per task:', ', '.join(f'{k} {v / tasks:.0f}' for k, v in ops.most_common(22)))
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f'{f:13s}:{l:4d} {v[0] / tasks:7.1f}/task  samples {100 * v[1] / tots:4.1f}%  {v[2]}')
