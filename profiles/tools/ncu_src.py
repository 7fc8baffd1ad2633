import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
iS=hdr.index('Source'); iN=hdr.index('# Samples'); iI=hdr.index('Instructions Executed')
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data=[r for r in rows[2:] if len(r)>max(iN,iI) and r[iN].strip().isdigit() and r[iI].strip().isdigit()]
tot_s=sum(int(r[iN] or 0) for r in data); tot_i=sum(int(r[iI] or 0) for r in data)
print('total samples',tot_s,'total warp instr',tot_i, 'n sass', len(data))
# group by opcode
op=collections.Counter(); ops=collections.Counter()
for r in data:
    o=r[iS].split()[0] if not r[iS].startswith('@') else r[iS].split()[1]
    op[o.split('.')[0]]+=int(r[iI] or 0); ops[o.split('.')[0]]+=int(r[iN] or 0)
print('--- instr mix (warp instr executed) / samples')
for k,v in op.most_common(28): print(f'{k:12s} {v:12d} {v/tot_i:6.1%}   samples {ops[k]/tot_s:6.1%}')
print('--- top sampled SASS lines')
top=sorted(data,key=lambda r:-int(r[iN] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]
for r in top:
    st={s:int(r[hdr.index(s)] or 0) for s in stalls}
    main=sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(f"{int(r[iN]):7d} {int(r[iN])/tot_s:6.1%} ex={r[iI]:>9s} {r[iS][:70]:70s} {main}")
