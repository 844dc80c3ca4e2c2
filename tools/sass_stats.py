import csv, sys, collections, re
rows=list(csv.reader(open(sys.argv[1])))
# find first kernel block header
i=0
while rows[i][0]!='Address': i+=1
hdr=rows[i]; 
ia=hdr.index('Address'); isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); ismp=hdr.index('# Samples')
ops=collections.Counter(); smp=collections.Counter(); tot=0; tots=0
body=[]
for r in rows[i+1:]:
    if len(r)<=iex or r[0]=='Address' or r[0]=='Kernel Name': break
    try: ex=int(r[iex]); s=int(r[ismp])
    except: continue
    src=r[isrc].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
    op=m.group(2) if m else src
    base=op.split('.')[0]
    if base in('LDS','STS','LDG','STG','F2F','MUFU','SHFL','BAR'): base=op if base in ('MUFU','F2F') else base
    ops[base]+=ex; smp[base]+=s; tot+=ex; tots+=s
    body.append((r[ia],src,ex,s))
frames=float(sys.argv[2]) if len(sys.argv)>2 else 48000
print('total warp-inst',tot,'per frame',tot/frames,'lane-inst/sample',tot/frames*32/2048)
for k,v in ops.most_common(45):
    print(f'{k:22s} {v/frames:9.1f}/frame  {100*v/tot:5.1f}%   samples {100*smp[k]/max(tots,1):5.1f}%')
if len(sys.argv)>3:
    top=sorted(body,key=lambda x:-x[3])[:int(sys.argv[3])]
    for a,s,e,sm in top: print(a,sm,e,s[:90])
