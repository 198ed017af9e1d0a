"""Static SASS instruction counts per source line of one kernel: python scratch/sass_lines.py file.cubin <mangled-name-prefix> [top]"""
import re, collections, subprocess, sys
txt = subprocess.run(['nvdisasm', '-g', '-c', sys.argv[1]], stdout=subprocess.PIPE, text=True).stdout
lines = txt.split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + sys.argv[2])][0]
end = [i for i, l in enumerate(lines) if i > start and l.startswith('//---------------------')]
end = end[0] if end else len(lines)
cur = None; cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for line in lines[start:end]:
    mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if mm:
        cur = (mm.group(1).split('/')[-1], int(mm.group(2))); continue
    mi = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if mi and cur:
        cnt[cur] += 1; ops[cur][mi.group(2)] += 1
print('total', sum(cnt.values()))
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 15]:
    print(v, k, dict(ops[k].most_common(6)))
