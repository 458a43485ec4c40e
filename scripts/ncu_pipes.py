"""Pipe utilisation per launch from an ncu --set full report: tensor pipe, the shared-memory data pipe split into the
tensor core's operand fetches (l1tex__data_pipe_tc_wavefronts_mem_shared) and the LSU's loads / stores
(l1tex__data_pipe_lsu_wavefronts_mem_shared), L1 / L2 / DRAM throughput, issue slots.  The two are separate data pipes (they do not add up:
l1tex% is the larger of them); the operand fetch is the bound the residual-block kernels sit on (4.0).  Usage: python scripts/ncu_pipes.py prof.ncu-rep"""
import csv, io, subprocess, sys

COLS = [('gpu__time_duration.sum', 'us', 1), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor%', 1),
        ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem_tc%', 1),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem_lsu%', 1),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex%', 1), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 1),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 1),
        ('sm__mio_inst_issued.avg.pct_of_peak_sustained_elapsed', 'mio%', 1)]


def main():
    raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = rows[0]
    idx = {n: head.index(n) for n, _, _ in COLS if n in head}
    kn = head.index('Kernel Name')
    print('%-46s' % 'kernel' + ''.join('%10s' % s for _, s, _ in COLS))
    for r in rows[2:]:
        if len(r) <= kn:
            continue
        out = '%-46s' % r[kn].replace('void ', '').replace('spb200::', '')[:45]
        for n, _, f in COLS:
            try:
                out += '%10.1f' % (float(r[idx[n]].replace(',', '')) * f)
            except Exception:
                out += '%10s' % '-'
        print(out)


if __name__ == '__main__':
    main()
