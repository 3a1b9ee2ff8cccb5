"""Per-region instruction breakdown of a kernel from an ncu report (source page): buckets the SASS by
position, prints warp-level and thread-level instruction counts, samples and average active lanes."""
import csv, subprocess, sys

def main(path, pattern, bucket=40, which=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pattern],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], []
    for r in rows:
        if r and r[0] == 'Kernel Name':
            if cur: blocks.append(cur)
            cur = []
        else:
            cur.append(r)
    blocks.append(cur)
    b = blocks[which]
    h = b[0]
    ia, isrc, isamp, iinst, ithr = (h.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed'))
    seen, data = set(), []
    for r in b[1:]:
        if len(r) > isamp and r[isamp].isdigit() and r[ia] not in seen:
            seen.add(r[ia]); data.append((int(r[isamp]), int(r[iinst]), int(r[ithr]), r[isrc].strip()))
    tw, tt, ts = sum(d[1] for d in data), sum(d[2] for d in data), sum(d[0] for d in data)
    print(f"{len(data)} SASS instrs, {tw/1e6:.1f} M warp-inst, {tt/1e9:.2f} G thread-inst, avg lanes {tt/max(tw,1):.1f}, {ts} samples")
    for i in range(0, len(data), bucket):
        ch = data[i:i + bucket]
        w, t, s = sum(d[1] for d in ch), sum(d[2] for d in ch), sum(d[0] for d in ch)
        if w * 200 < tw: continue
        print(f"{i:5d}-{i+len(ch)-1:5d} warp {w/1e6:8.2f}M ({100*w/tw:4.1f}%) lanes {t/max(w,1):5.1f} samples {100*s/max(ts,1):4.1f}%  {ch[0][3][:50]}")

if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, int(sys.argv[4]) if len(sys.argv) > 4 else 0)
