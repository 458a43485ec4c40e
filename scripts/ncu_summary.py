"""Summarise an ncu report (.ncu-rep) into a compact per-launch table (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
cols = [('Kernel Name', 'kernel', 44), ('gpu__time_duration.sum', 'us', 8), ('dram__bytes_read.sum', 'rdMB', 8), ('dram__bytes_write.sum', 'wrMB', 8),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 6),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 7), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 6),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6), ('smsp__inst_executed.sum', 'Minst', 8),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf', 9), ('launch__registers_per_thread', 'regs', 5),
        ('launch__grid_size', 'grid', 7), ('lts__t_sectors_srcunit_tex_op_read.sum', 'l2rdMB', 8)]
idx = [hdr.index(c[0]) if c[0] in hdr else -1 for c in cols]
units = rows[1]
print(' '.join(('%-' + str(c[2]) + 's') % c[1] for c in cols))
tot = 0.0
for r in rows[2:]:
    out = []
    for (name, short, w), i in zip(cols, idx):
        v = r[i] if i >= 0 else '-'
        if short == 'kernel':
            v = v.replace('spb200::', '').replace('void ', '')[:w]
        elif short in ('Minst',):
            v = '%.2f' % (float(v.replace(',', '')) / 1e6)
        elif short == 'l2rdMB' and i >= 0:
            v = '%.1f' % (float(v.replace(',', '')) * 32 / 1e6)
        elif short in ('rdMB', 'wrMB') and i >= 0:
            f = float(v.replace(',', ''))
            u = units[i]
            f = f * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
            v = '%.1f' % f
        elif short == 'us' and i >= 0:
            f = float(v.replace(',', ''))
            u = units[i]
            f = f * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
            tot += f
            v = '%.1f' % f
        else:
            try:
                v = '%.1f' % float(v.replace(',', ''))
            except Exception:
                pass
        out.append(('%-' + str(w) + 's') % v)
    print(' '.join(out))
print('total %.1f us' % tot)
