"""Hot SASS instructions of each kernel in an ncu report: ncu -i rep --page source --csv | this script [topN]."""
import csv, sys
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(sys.argv[1])))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1][:90]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        ci = {h: k for k, h in enumerate(hdr)}
        s_all, s_ex = ci['# Samples'], ci['Instructions Executed']
        tot = sum(int(r[s_all] or 0) for r in body)
        totx = sum(int(r[s_ex] or 0) for r in body)
        print('==', name, 'samples', tot, 'inst', totx)
        # running regions: print every instruction with >= 1% samples, with index
        for k, r in enumerate(body):
            n = int(r[s_all] or 0)
            if n >= max(1, tot // 100 * (100 // top if top < 100 else 1)) or n * 100 >= tot * 1.5:
                stalls = {h[6:]: int(r[ci[h]] or 0) for h in hdr if h.startswith('stall_') and '(' not in h}
                best = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
                print('%5d %5.1f%% x%-9s %-70s %s' % (k, 100.0 * n / tot, r[s_ex], r[ci['Source']][:70], best))
        i = j
    else:
        i += 1
