"""SASS evidence for profiles/: per-kernel instruction counts of the signature mnemonics (tcgen05 = UTCHMMA / UTCBAR / LDTM,
TMA = UTMALDG / UBLKCP, FP64 tensor = DMMA ...) and the first occurrence of each, from cuobjdump -sass of the built library."""
import collections
import re
import subprocess

out = subprocess.run(['cuobjdump', '-sass', 'wisecondor_b200/libwisecondor_b200.so'], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', out)
want = ['wc_dist_topk_tc_kernelILi0ELb0', 'wc_dist_topk_tc_kernelILi1ELb0', 'wc_dist_topk_tc_kernelILi2ELb0', 'wc_fin_rescore_kernelILi4ELi2',
        'wc_fin_select_kernel', 'wc_fin_rank_kernel', 'wc_prepare_f16_kernel', 'wc_dist_topk_kernelILb1', 'wc_dist_topk_f16_kernelILb1',
        'wc_zscore_kernel', 'wc_segment_kernelILb0']
keys = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'DMMA', 'HMMA', 'LDSM', 'SHF.L.W', 'F2F', 'FADD', 'FFMA', 'FSETP', 'DADD', 'DMUL',
        'LDS', 'STS', 'LDG', 'STG', 'ATOMG', 'RED', 'SHFL', 'BAR.SYNC', 'USETMAXREG', 'R2UR']
print('# SASS of wisecondor_b200/libwisecondor_b200.so (cuobjdump -sass, sm_100a): instruction counts per kernel, first occurrences.')
print('# Regenerate: python tools/sass_summary.py > profiles/sass_r03.txt\n')
for f in funcs[1:]:
    name = f.split('\n', 1)[0]
    if not any(w in name for w in want):
        continue
    ins = [l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4,6}\*/', l)]
    cnt = collections.Counter()
    for l in ins:
        m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m:
            for k in keys:
                if m.group(1).startswith(k):
                    cnt[k] += 1
    print('== %s' % name[:160])
    print('   instructions %d | ' % len(ins) + ', '.join('%s %d' % (k, cnt[k]) for k in keys if cnt[k]))
    shown = set()
    for l in ins:
        for k in ['UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'DMMA', 'HMMA', 'USETMAXREG', 'UTCBAR']:
            if k in l and k not in shown:
                shown.add(k)
                print('   ' + re.sub(r'\s+/\* 0x[0-9a-f]+ \*/', '', l).strip())
    print()
