"""SASS evidence per hot kernel of libspb200.so (runs where cuobjdump is installed; no GPU needed):
    python scripts/sass_excerpts.py > profiles/r02_sass_excerpts.txt
For every kernel: instruction counts of the Blackwell-specific mnemonics (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier ops, ACQBULK / PREEXIT = the
griddepcontrol pair of programmatic dependent launch) and the first lines around a tcgen05.mma."""
import os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, 'feature-point-cnn_b200', 'libspb200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', out)[1:]
WANT = ('halo_tc_kernel', 'block_tc_kernel', 'stem_planes_kernel', 'stem_tc_kernel', 'stem_wide_kernel', 'match_tc', 'nms_round0_kernel<4, true>',
        'nms_finish_kernel', 'sample_desc128_kernel', 'planes_kernel')
MNEMONICS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'SYNCS', 'ACQBULK', 'PREEXIT', 'HMMA', 'LDG', 'STG', 'LDS', 'STS']
demangle = subprocess.run(['c++filt'], input='\n'.join(f.split('\n', 1)[0].strip() for f in funcs), capture_output=True, text=True).stdout.splitlines()
print('# cuobjdump -sass feature-point-cnn_b200/libspb200.so: mnemonic counts per kernel (sm_100a)\n')
seen = set()
for name, f in zip(demangle, funcs):
    short = re.sub(r'\(.*', '', name).replace('spb200::', '').replace('void ', '')
    if not any(w in short for w in WANT) or short in seen or '__nv_bfloat16' in short:
        continue
    seen.add(short)
    body = f.split('\n', 1)[1]
    lines = [l for l in body.splitlines() if re.search(r'/\*[0-9a-f]{4}\*/', l)]
    counts = {m: sum(1 for l in lines if re.search(r'\b' + m + r'\b', l) or (' ' + m + '.') in l or (' ' + m + ' ') in l) for m in MNEMONICS}
    print('## %s' % short)
    print('   ' + '  '.join('%s %d' % (m, c) for m, c in counts.items() if c) + '  (of %d instructions)' % len(lines))
    for i, l in enumerate(lines):
        if 'UTCHMMA' in l:
            for x in lines[max(0, i - 3):i + 4]:
                print('      ' + re.sub(r'\s+/\* 0x[0-9a-f]+ \*/\s*$', '', x).strip())
            break
    print()
