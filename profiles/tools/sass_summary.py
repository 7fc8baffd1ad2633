"""Regenerates profiles/r2_sass.md: static SASS opcode counts per kernel of the in-tree libdct3d.so (cuobjdump -sass)."""
import collections, re, subprocess
so = '3ddctvideoencoding_b200/libdct3d.so'
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern).replace('void ', '').replace('dct3d::', '')
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        counts[kern][m.group(1)] += 1
ops = ['UTMALDG', 'UTMASTG', 'SYNCS', 'FFMA', 'FADD', 'FMUL', 'DFMA', 'DADD', 'DMUL', 'F2IP', 'PRMT', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'BAR',
       'ATOMG', 'REDG', 'HMMA', 'UTCHMMA', 'UTCQMMA']
lines = ['# Round 2 SASS opcode summary (libdct3d.so, sm_100a)', '',
         'Static instruction counts per kernel from `cuobjdump -sass 3ddctvideoencoding_b200/libdct3d.so` (regenerate: `python profiles/tools/sass_summary.py`).',
         '`UTMALDG` = TMA tensor loads (the warp-private unit loads of `encode_kernel`), `UTMASTG` = TMA tensor stores (the pixel tile of '
         '`reconstruct_coo_kernel<., TAIL_TMA>`, one per unit, DESIGN.md 3.5), `SYNCS` = mbarrier operations; there is '
         'no tensor-core instruction (`HMMA`/`UTC*MMA`): the north star excludes tensor cores for this path.', '',
         '| kernel | total | ' + ' | '.join(ops) + ' |', '|---|---|' + '---|' * len(ops)]
for k, c in counts.items():
    lines.append(f'| {k} | {sum(c.values())} | ' + ' | '.join(str(c.get(o, 0)) for o in ops) + ' |')
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
lines += ['', f"Whole library: UTMALDG {tot['UTMALDG']}, UTMASTG {tot['UTMASTG']}, SYNCS {tot['SYNCS']}, FFMA {tot['FFMA']}, DFMA {tot['DFMA']}, "
              f"tensor-core instructions {tot['HMMA'] + tot['UTCHMMA'] + tot['UTCQMMA']}."]
open('profiles/r2_sass.md', 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[-2:]))
