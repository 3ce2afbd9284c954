#!/bin/bash
# usage: scripts/ncu_hot.sh <report.ncu-rep> <kernel regex> [top N] [launch index]  -- hottest CUDA source lines by warp-stall samples
rep=$1; k=$2; n=${3:-25}; skip=${4:-}
ncu -i "$rep" --page source --csv --print-source cuda,sass --kernel-name "regex:$k" ${skip:+--launch-skip $skip --launch-count 1} 2>/dev/null > /tmp/_src.csv
python - "$n" <<'PY'
import csv, sys
n=int(sys.argv[1])
rows=list(csv.reader(open('/tmp/_src.csv')))
hdr=None; data=[]; tot=0; fname=''
for r in rows:
    if len(r)>=2 and r[0]=='File Path': fname=r[1].split('/')[-1]
    if len(r)>3 and r[0]=='Line No':
        hdr=r; ix={}
        for i,h in enumerate(hdr): ix.setdefault(h,i)
        continue
    if hdr is None or len(r)<len(hdr) or not r[0].strip().isdigit(): continue
    try: s=int(r[ix['# Samples']])
    except: continue
    tot+=s; data.append((s,fname,r,ix))
data.sort(key=lambda x:-x[0])
print("total samples",tot)
for s,f,r,ix in data[:n]:
    stalls=[h for h in ix if h.startswith('stall_') and 'Not Issued' not in h]
    top=sorted(((int(r[ix[h]] or 0),h) for h in stalls),reverse=True)[:3]
    print(f"{100*s/max(tot,1):5.1f}% {f}:{r[0]:>4} {r[1].strip()[:100]:100s} | "+", ".join(f"{h[6:]}={v}" for v,h in top if v))
PY
