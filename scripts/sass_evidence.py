"""SASS evidence of the built library (runs here, no GPU): per kernel the instruction count, the Blackwell-specific
mnemonics, R2UR count, and for the main conv kernel the MMA issue sequence of one pipeline stage.
python scripts/sass_evidence.py > profiles/r2_sass_conv_tc.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'luminoth_b200', 'libluminoth_b200.so')
MNEMONICS = ['UTCHMMA.2CTA', 'UTCHMMA', 'UTMALDG.4D.2CTA', 'UTMALDG.2D.2CTA', 'UTMALDG.4D', 'UTMALDG.2D', 'UTMASTG', 'LDTM', 'UTCBAR.2CTA.MULTICAST',
             'UTCBAR', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'UBLKCP', 'FHFMA', 'FADD2', 'FFMA2', 'ELECT', 'R2UR', 'HMMA', 'HGMMA']
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
blocks = re.split(r'\n\s*Function : \S+\n', sass)[1:]
print('SASS evidence for the sm_100a kernels of luminoth_b200/libluminoth_b200.so (cuobjdump -sass; scripts/sass_evidence.py).')
print('UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2); UTMALDG / UTMASTG = TMA tensor load / store; LDTM = tcgen05.ld;')
print('UTCBAR = tcgen05.commit -> mbarrier (.2CTA.MULTICAST for pairs); UCGABAR = cluster barrier; UBLKCP = cp.async.bulk;')
print('FHFMA = fma.f32.f16; FADD2 / FFMA2 = packed fp32x2; ELECT = elect.sync; R2UR = vector -> uniform register move.')
print('No HMMA / HGMMA (legacy tensor paths) anywhere.')
print()
main_block = None
for name, blk in zip(names, blocks):
    ins = re.findall(r'/\*[0-9a-f]{4,}\*/\s+(.*?);', blk)
    cnt = collections.Counter()
    for i in ins:
        op = i.split()[1] if i.startswith('@') else i.split()[0]
        for m in MNEMONICS:
            if op.startswith(m):
                cnt[m] += 1
                break
    shown = ', '.join('%s x%d' % (m, cnt[m]) for m in MNEMONICS if cnt[m])
    print('%s\n    %d instructions; %s' % (name, len(ins), shown or '-'))
    if 'conv_tc_kernel<128, 3, false, 2, false, false, false>' in name:
        main_block = ins
if main_block:
    idx = [i for i, x in enumerate(main_block) if 'UTCHMMA' in x]
    print()
    print('MMA issue sequence of one pipeline stage, conv_tc_kernel<128, 3, false, 2, false, false, false> (first to last UTCHMMA):')
    for x in main_block[idx[0] - 2: idx[-1] + 3]:
        print('        ' + x.strip())
